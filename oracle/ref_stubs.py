"""TEST INFRASTRUCTURE — not product code.

Makes the UNMODIFIED reference (`/root/reference/models/*.py`, `utils/dataset.py`,
`training/loss.py`) importable in this container, where `torch_geometric`, `lightning`,
`meshkernel`, ... are not installed, by registering small `sys.modules` stubs first
(SURVEY.md §8c / Appendix A).  The only stub that carries arithmetic is
`torch_geometric.utils.scatter`, which restates the native-torch path of PyG 2.4.0
(`requirements.txt:25` pins torch_geometric==2.4.0; its source is not under `/root/reference`):
``sum`` = ``zeros.scatter_add_``, ``mean`` = that divided by ``count.clamp(min=1)``.

`/root/reference` exists only in the build container.  Nothing that runs on the GPU box
(`-m gpu` tests, `smoke()`, `bench.py`) may import this module; it is used by
`oracle/gen_golden.py` and by the CPU tests that pin `oracle/swe_oracle.py` against the real
reference (those tests skip when the directory is absent).
"""
from __future__ import annotations

import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MSWE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "gnn.py"))


def pyg_scatter(src, index, dim=0, dim_size=None, reduce="sum"):
    """PyG 2.4.0 `torch_geometric.utils.scatter`, native-torch branch (call sites
    `models/gnn.py:254,256,437`)."""
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[0] = dim_size
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    out = src.new_zeros(shape).scatter_add_(0, idx, src)
    if reduce in ("sum", "add"):
        return out
    if reduce != "mean":
        raise NotImplementedError(reduce)
    cnt = src.new_zeros(dim_size).scatter_add_(0, index, src.new_ones(src.shape[0])).clamp(min=1)
    return out / cnt.view(-1, *([1] * (src.dim() - 1)))


class StubData:
    """Attribute bag standing in for `torch_geometric.data.Data` inside the reference code."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def keys(self):
        return list(self.__dict__)

    def clone(self):
        out = self.__class__()
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.clone() if torch.is_tensor(v) else v
        return out


class StubBatch(StubData):
    pass


_installed = False


def install():
    """Register the stubs and put the reference root on sys.path.  Idempotent."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _NoConv(torch.nn.Module):
        def __init__(self, *a, **k):
            raise NotImplementedError("PyG conv baselines are not part of the hot path")

    class _MultiscaleMesh:
        pass

    class _Lightning:
        class LightningModule(torch.nn.Module):
            def log(self, *a, **k):
                pass

        class LightningDataModule:
            pass

    mod("torch_geometric")
    mod("torch_geometric.nn", ChebConv=_NoConv, TAGConv=_NoConv, GATConv=_NoConv)
    mod("torch_geometric.utils", scatter=pyg_scatter, to_undirected=None)
    mod("torch_geometric.data", Data=StubData, Batch=StubBatch, DataLoader=None, Dataset=object)
    mod("torch_geometric.data.batch", Batch=StubBatch)
    mod("torch_geometric.loader", DataLoader=None)
    mod("database")
    mod("database.graph_creation", MultiscaleMesh=_MultiscaleMesh, rotate_mesh=None)
    # Reference modules named `models`, `utils`, `training` must win over anything else.
    for k in [k for k in sys.modules if k.split(".")[0] in ("models", "utils", "training")]:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def load_reference():
    """Returns a namespace with the reference classes / helpers used to pin the oracle."""
    install()
    import importlib
    gnn = importlib.import_module("models.gnn")
    models = importlib.import_module("models.models")
    dataset = importlib.import_module("utils.dataset")
    loss = importlib.import_module("training.loss")
    return types.SimpleNamespace(
        GNN=gnn.GNN, MSGNN=gnn.MSGNN, SWEGNN=gnn.SWEGNN,
        BaseFloodModel=models.BaseFloodModel, make_mlp=models.make_mlp,
        activation_functions=models.activation_functions,
        apply_boundary_condition=dataset.apply_boundary_condition,
        use_prediction=dataset.use_prediction, create_scale_mask=dataset.create_scale_mask,
        loss_function=loss.loss_function, Data=StubData, Batch=StubBatch)


def to_stub(data, batch: bool = False):
    """Convert one of our `Data` objects into the stub type the reference code sees."""
    cls = StubBatch if batch else StubData
    return cls(**{k: getattr(data, k) for k in data.keys()})


def load_checkpoint_state_dict(name: str):
    """`results/Pareto_front/models/<name>.h5` → model state_dict (Lightning prefix `model.`
    stripped)."""
    path = os.path.join(REFERENCE_ROOT, "results", "Pareto_front", "models", name + ".h5")
    ck = torch.load(path, map_location="cpu", weights_only=False)
    return {k[len("model."):]: v for k, v in ck["state_dict"].items() if k.startswith("model.")}
