"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported under oracle/ref_stubs.py) on seeded synthetic meshes.

    python oracle/gen_golden.py            # only works where /root/reference exists

Every fixture stores the reference's outputs (and, for the trained checkpoint, the weights, which
are reference artefacts that cannot travel to the GPU box otherwise).  Inputs are NOT stored: they
are regenerated from the recorded generator arguments by `mswe_gnn_b200.utils.synthetic`, and
random-init weights from the recorded constructor arguments + seed (the constructors consume the
RNG in the reference's order; `weights_sha` pins that).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_stubs  # noqa: E402

import mswe_gnn_b200  # noqa: E402,F401
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh, make_tri_mesh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sd_sha(sd) -> str:
    h = hashlib.sha256()
    for k in sd:
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def ref_rollout(R, model, data, steps):
    """The reference's rollout_test loop body (training/train.py:87-93) driven with the
    reference's own helpers and model."""
    t = ref_stubs.to_stub(data).clone()
    nd = model.previous_t * model.NUM_WATER_VARS
    preds = []
    with torch.no_grad():
        for s in range(steps):
            t.x[:, -nd:] = R.apply_boundary_condition(t.x[:, -nd:], t.BC[:, :, s], t.node_BC, type_BC=t.type_BC)
            p = model(t)
            t.x = R.use_prediction(t.x, p, model.previous_t)
            preds.append(p)
    return torch.stack(preds, -1)


def main():
    R = ref_stubs.load_reference()
    os.makedirs(OUT, exist_ok=True)
    cfg = yaml.safe_load(open(os.path.join(ref_stubs.REFERENCE_ROOT, "config.yaml")))["models"]
    cfg.pop("model_type")
    cases = []

    def dump(name, meta, **arrays):
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta), **arrays)
        cases.append(name)
        print(name, {k: v.shape for k, v in arrays.items()})

    # 1. default config.yaml model (K4/F64, random init seed 666), cfg1 mesh, forward + 3-step rollout
    mesh = dict(nx=32, ny=24, num_scales=4, previous_t=3, rollout_steps=3, wet="random", seed=0)
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=4, previous_t=3, **cfg)
    m = R.MSGNN(**ctor)
    d = make_tri_mesh(**mesh)
    with torch.no_grad():
        fwd = m(ref_stubs.to_stub(d))
    roll = ref_rollout(R, m, d, 3)
    dump("msgnn_k4f64_cfg1", dict(model="MSGNN", ctor=ctor, mesh=mesh, weights_sha=sd_sha(m.state_dict()),
                                  n_params=sum(p.numel() for p in m.parameters())),
         forward=fwd.numpy(), rollout=roll.numpy())

    # 2. irregular inter-scale edges (in-degree 0/2, ghost links), small F=16 model, K list
    mesh = dict(nx=16, ny=8, num_scales=3, previous_t=2, rollout_steps=2, wet="random", seed=3,
                link_ghosts=True, orphan_every=7, extra_parent_every=5)
    ctor = dict(num_node_features=6, num_edge_features=1, num_scales=3, previous_t=2, hid_features=16,
                mlp_layers=2, K=[2, 1, 3], seed=7, learned_residuals="all", with_WL=False)
    m = R.MSGNN(**ctor)
    d = make_tri_mesh(**mesh)
    with torch.no_grad():
        fwd = m(ref_stubs.to_stub(d))
    dump("msgnn_k213f16_irregular", dict(model="MSGNN", ctor=ctor, mesh=mesh, weights_sha=sd_sha(m.state_dict()),
                                         n_params=sum(p.numel() for p in m.parameters())),
         forward=fwd.numpy(), rollout=ref_rollout(R, m, d, 2).numpy())

    # 3. trained checkpoint K4_F32 on a dry bed with inflow: physically meaningful 8-step rollout
    sd = ref_stubs.load_checkpoint_state_dict("K4_F32")
    mesh = dict(nx=32, ny=24, num_scales=4, previous_t=3, rollout_steps=8, wet="dry", inflow=0.3, seed=0)
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=4, previous_t=3, **{**cfg, "hid_features": 32})
    m = R.MSGNN(**ctor)
    print(m.load_state_dict(sd))
    d = make_tri_mesh(**mesh)
    roll = ref_rollout(R, m, d, 8)
    m64 = R.MSGNN(**ctor).double()
    m64.load_state_dict({k: v.double() for k, v in sd.items()})
    d64 = make_tri_mesh(**mesh)
    for k in ("x", "edge_attr", "BC"):
        setattr(d64, k, getattr(d64, k).double())
    roll64 = ref_rollout(R, m64, d64, 8)
    weights = {"w::" + k: v.numpy() for k, v in sd.items()}
    dump("msgnn_k4f32_trained_drybed", dict(model="MSGNN", ctor=ctor, mesh=mesh, weights_sha=sd_sha(sd),
                                            n_params=sum(p.numel() for p in m.parameters()),
                                            checkpoint="results/Pareto_front/models/K4_F32.h5"),
         rollout=roll.numpy(), rollout_fp64=roll64.numpy(), **weights)

    # 4. single-scale SWE-GNN (GNN), config.yaml hyper-parameters, K=3
    gcfg = {k: v for k, v in cfg.items() if k not in ("learned_pooling", "skip_connections")}
    gcfg["K"] = 3
    mesh = dict(nx=24, ny=16, previous_t=3, rollout_steps=2, wet="random", seed=5)
    ctor = dict(num_node_features=8, num_edge_features=1, previous_t=3, n_GNN_layers=2, **gcfg)
    m = R.GNN(**ctor)
    d = make_single_scale_mesh(**mesh)
    with torch.no_grad():
        fwd = m(ref_stubs.to_stub(d))
    dump("gnn_k3f64_single", dict(model="GNN", ctor=ctor, mesh=mesh, weights_sha=sd_sha(m.state_dict()),
                                  n_params=sum(p.numel() for p in m.parameters())),
         forward=fwd.numpy(), rollout=ref_rollout(R, m, d, 2).numpy())

    # 5. SWEGNN operator alone: with and without gradient / filter / edge features
    torch.manual_seed(11)
    d = make_single_scale_mesh(8, 8, seed=9)
    n, e = d.x.shape[0], d.edge_index.shape[1]
    xs, xd, ea = torch.randn(n, 16), torch.randn(n, 16) * (torch.rand(n, 1) < 0.5), torch.randn(e, 16)
    arrays = dict(x_s=xs.numpy(), x_d=xd.numpy(), edge_attr=ea.numpy(), edge_index=d.edge_index.numpy())
    metas = []
    for i, kw in enumerate([dict(edge_features=16, K=3), dict(edge_features=0, K=1, with_filter_matrix=False,
                                                               with_gradient=False),
                            dict(edge_features=16, K=2, upwind_mode=True), dict(edge_features=16, K=2, normalize=False)]):
        torch.manual_seed(100 + i)
        op = R.SWEGNN(16, 16, n_layers=2, activation="prelu", bias=True, **kw)
        with torch.no_grad():
            out = op(xs, xd, d.edge_index, ea if kw["edge_features"] else None)
        arrays[f"out{i}"] = out.numpy()
        metas.append(dict(kw=kw, seed=100 + i, weights_sha=sd_sha(op.state_dict())))
    dump("swegnn_operator", dict(model="SWEGNN", variants=metas), **arrays)

    # published parameter counts (results/Pareto_front/overview_{MSGNN,GNN}.csv) as known answers
    json.dump(dict(MSGNN={"2,16": 47981, "3,32": 196781, "4,32": 203949, "4,64": 811309, "5,64": 839981},
                   GNN={"16,10": 16068, "32,10": 63348, "64,18": 317140}),
              open(os.path.join(OUT, "param_counts.json"), "w"), indent=1)
    print("wrote", cases)


if __name__ == "__main__":
    main()
