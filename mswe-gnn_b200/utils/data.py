"""Minimal stand-ins for ``torch_geometric.data.Data`` / ``Batch``.

The models only use attribute access on the graph object (reference ``models/gnn.py:267-277``),
so any object with the right attributes works, including a real PyG ``Data``.  PyG is not a
dependency of this package; these two classes give the tests, the benchmark and the rollout loop
something to hold the fields listed in SURVEY.md Appendix C, and ``Batch.from_data_list`` restates
the PyG collation rules the reference's ``adapt_batch_training`` relies on
(``training/train.py:14-65``): attributes whose name contains ``index`` are concatenated along
the last dimension and shifted by the running node count, every other tensor attribute is
concatenated along dim 0 un-shifted, ``ptr`` holds the per-graph node offsets.
"""
from __future__ import annotations

import copy
from typing import Iterable, List

import torch


class Data:
    """Attribute bag with the handful of methods the reference calls on a PyG ``Data``
    (``keys()`` at ``training/train.py:22``, ``clone()`` at ``training/train.py:17,80``)."""

    def __init__(self, **fields):
        for k, v in fields.items():
            setattr(self, k, v)

    def keys(self) -> List[str]:
        return [k for k in self.__dict__ if not k.startswith("_")]

    def __contains__(self, key: str) -> bool:
        return key in self.__dict__

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.clone() if torch.is_tensor(v) else copy.deepcopy(v)
        return out

    def to(self, device, non_blocking: bool = False):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v
        return out

    @property
    def num_nodes(self) -> int:
        return int(self.x.shape[0])

    def __repr__(self):
        parts = []
        for k in self.keys():
            v = getattr(self, k)
            parts.append(f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={v!r}")
        return f"{self.__class__.__name__}({', '.join(parts)})"


class Batch(Data):
    """Several graphs stacked the way PyG's ``Batch.from_data_list`` stacks them."""

    @classmethod
    def from_data_list(cls, graphs: Iterable[Data]) -> "Batch":
        graphs = list(graphs)
        assert len(graphs) > 0
        out = cls()
        offsets = [0]
        for g in graphs:
            offsets.append(offsets[-1] + g.num_nodes)
        out.ptr = torch.tensor(offsets, dtype=torch.long)
        out.num_graphs = len(graphs)
        out._graphs = graphs
        out.batch = torch.cat([torch.full((g.num_nodes,), i, dtype=torch.long)
                               for i, g in enumerate(graphs)])
        for key in graphs[0].keys():
            vals = [getattr(g, key) for g in graphs]
            if torch.is_tensor(vals[0]) and vals[0].dim() > 0:
                if "index" in key:
                    vals = [v + off for v, off in zip(vals, offsets[:-1])]
                    setattr(out, key, torch.cat(vals, dim=-1))
                else:
                    setattr(out, key, torch.cat(vals, dim=0))
            elif torch.is_tensor(vals[0]):
                setattr(out, key, torch.stack(vals))
            elif isinstance(vals[0], (int, float)):
                setattr(out, key, torch.tensor(vals))
            else:
                setattr(out, key, vals)
        return out

    def keys(self) -> List[str]:
        return [k for k in self.__dict__ if not k.startswith("_")]

    def __getitem__(self, i: int) -> Data:
        return self._graphs[i]

    def clone(self):
        out = super().clone()
        out.__dict__["_graphs"] = self._graphs
        return out

    def to(self, device, non_blocking: bool = False):
        out = super().to(device, non_blocking)
        out.__dict__["_graphs"] = self._graphs
        return out
