"""Synthetic structured triangular meshes ``tri(nx, ny)`` shaped like the reference's PyG graphs.

The Zenodo datasets the reference trains on are not available offline, so the tests and the
benchmark use this generator.  It follows the conventions of the reference's mesh → graph
conversion (``database/graph_creation.py``) so that the models see inputs of
the same form they would get from ``convert_mesh_to_pyg`` (``graph_creation.py:1483-1582``):

* graph nodes are mesh faces, edges are the dual graph (``graph_creation.py:63-66``);
* scales are stacked fine → coarse with global node ids (``graph_creation.py:887-893,1527-1530``);
* per scale the undirected dual edges come first in (row, col) order, followed by one directed
  ghost → boundary-face edge (``graph_creation.py:1259-1263,1398-1400``); the ghost cell is the
  last face of its scale (``graph_creation.py:1391-1394``);
* inter-scale edges are (coarse, fine) pairs from centroid containment, ordered by coarse id then
  fine id (``graph_creation.py:422-436,925-927``);
* ``node_BC`` lists only the finest ghost cell (``graph_creation.py:1577``).

Geometry: an ``nx × ny`` grid of unit squares, each cut by its (0,0)-(1,1) diagonal into a
lower-right triangle (t=0) and an upper-left triangle (t=1); face id ``2*(j*nx+i)+t``.  Level
``s+1`` halves ``nx`` and ``ny``; a coarse triangle contains the centroids of exactly four finer
triangles.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .data import Data


def tri_level_sizes(nx: int, ny: int, num_scales: int):
    """(N_s, E_s, I_s) per level, SURVEY.md §8 formulae."""
    out = []
    for s in range(num_scales):
        a, b = nx >> s, ny >> s
        out.append((2 * a * b + 1, 2 * (3 * a * b - a - b) + 1, 2 * a * b))
    return out


def _dual_edges(nx: int, ny: int) -> np.ndarray:
    """Directed dual edges of one level in (row, col) order, local ids, shape [2, E]."""
    i = np.arange(nx, dtype=np.int64)[None, :].repeat(ny, 0)
    j = np.arange(ny, dtype=np.int64)[:, None].repeat(nx, 1)
    sq = j * nx + i
    f0, f1 = 2 * sq, 2 * sq + 1
    big = np.iinfo(np.int64).max
    # t=0: diagonal ↔ same square t=1; bottom ↔ (i, j-1) t=1; right ↔ (i+1, j) t=1
    n0 = np.stack([f1,
                   np.where(j > 0, 2 * (sq - nx) + 1, big),
                   np.where(i < nx - 1, 2 * (sq + 1) + 1, big)], -1)
    # t=1: diagonal ↔ same square t=0; top ↔ (i, j+1) t=0; left ↔ (i-1, j) t=0
    n1 = np.stack([f0,
                   np.where(j < ny - 1, 2 * (sq + nx), big),
                   np.where(i > 0, 2 * (sq - 1), big)], -1)
    nbr = np.stack([n0, n1], 2).reshape(-1, 3)          # [2*nx*ny, 3] in face-id order
    nbr.sort(axis=1)
    row = np.arange(2 * nx * ny, dtype=np.int64)[:, None].repeat(3, 1)
    keep = nbr != big
    return np.stack([row[keep], nbr[keep]])


def _containment(nx: int, ny: int) -> np.ndarray:
    """(coarse, fine) pairs between a level of ``nx × ny`` squares and the next coarser one,
    local ids, ordered by coarse then fine, shape [2, 2*nx*ny]."""
    i = np.arange(nx, dtype=np.int64)[None, :].repeat(ny, 0)
    j = np.arange(ny, dtype=np.int64)[:, None].repeat(nx, 1)
    a, b = i & 1, j & 1
    csq = (j >> 1) * (nx >> 1) + (i >> 1)
    fine0 = (2 * (j * nx + i)).ravel()
    # centroid of fine t=0 is (a+2/3, b+1/3)/2, of fine t=1 is (a+1/3, b+2/3)/2; coarse t=0 is y<x
    par0 = np.where(b <= a, 0, 1).ravel()               # fine t=0: lower iff b < a + 1/3
    par1 = np.where(b < a, 0, 1).ravel()                # fine t=1: lower iff b + 1/3 < a
    coarse = np.concatenate([2 * csq.ravel() + par0, 2 * csq.ravel() + par1])
    fine = np.concatenate([fine0, fine0 + 1])
    order = np.lexsort((fine, coarse))
    return np.stack([coarse[order], fine[order]])


def make_tri_mesh(nx: int, ny: int, num_scales: int = 4, previous_t: int = 3, rollout_steps: int = 1,
                  wet: str = "random", wet_fraction: float = 0.3, inflow: float = 0.3,
                  seed: int = 0, link_ghosts: bool = False, orphan_every: int = 0,
                  extra_parent_every: int = 0, dtype=torch.float32, with_y: bool = True) -> Data:
    """Multi-scale graph of ``tri(nx, ny)`` with ``num_scales`` levels.

    wet='random' : 30 % wet nodes with h,|q| ~ U(0,1) in every window slot (stress fixture).
    wet='dry'    : dry bed; the only forcing is the discharge boundary condition ``inflow``.
    link_ghosts  : add ghost(coarse) → ghost(fine) inter-scale edges.
    orphan_every / extra_parent_every : drop / duplicate every n-th inter-scale edge so that
        un-pool in-degree 0 and 2 and pool fan-in 0 occur (Appendix D-5).  0 = off.
    """
    assert nx % (1 << (num_scales - 1)) == 0 and ny % (1 << (num_scales - 1)) == 0, \
        "nx, ny must be divisible by 2**(num_scales-1)"
    sizes = tri_level_sizes(nx, ny, num_scales)
    node_ptr = np.concatenate([[0], np.cumsum([s[0] for s in sizes])]).astype(np.int64)
    edges, intra = [], []
    for s in range(num_scales):
        a, b = nx >> s, ny >> s
        e = _dual_edges(a, b) + node_ptr[s]
        ghost = node_ptr[s + 1] - 1
        e = np.concatenate([e, np.array([[ghost], [node_ptr[s]]], dtype=np.int64)], 1)
        edges.append(e)
        if s < num_scales - 1:
            c = _containment(a, b)
            c[0] += node_ptr[s + 1]
            c[1] += node_ptr[s]
            if orphan_every:
                keep = np.ones(c.shape[1], dtype=bool)
                keep[orphan_every - 1::orphan_every] = False
                c = c[:, keep]
            if extra_parent_every:
                # give every n-th fine face a second parent (the next coarse face, wrapping)
                sel = c[:, extra_parent_every - 1::extra_parent_every].copy()
                nc = sizes[s + 1][0] - 1
                sel[0] = node_ptr[s + 1] + (sel[0] - node_ptr[s + 1] + 1) % nc
                c = np.concatenate([c, sel], 1)
                c = c[:, np.lexsort((c[1], c[0]))]
            if link_ghosts:
                c = np.concatenate([c, np.array([[node_ptr[s + 2] - 1], [node_ptr[s + 1] - 1]],
                                                dtype=np.int64)], 1)
            intra.append(c)
    edge_ptr = np.concatenate([[0], np.cumsum([e.shape[1] for e in edges])]).astype(np.int64)
    intra_edge_ptr = np.concatenate([[0], np.cumsum([c.shape[1] for c in intra])]).astype(np.int64)
    N, E = int(node_ptr[-1]), int(edge_ptr[-1])

    g0 = torch.Generator().manual_seed(seed)          # mesh / static values
    g1 = torch.Generator().manual_seed(seed + 1)      # dynamic values
    area = torch.randn(N, generator=g0, dtype=dtype)
    dem = torch.rand(N, generator=g0, dtype=dtype) * 2.0
    edge_attr = torch.randn(E, 1, generator=g0, dtype=dtype)
    dyn = torch.zeros(N, 2 * previous_t, dtype=dtype)
    if wet == "random":
        is_wet = torch.rand(N, generator=g1) < wet_fraction
        dyn = torch.rand(N, 2 * previous_t, generator=g1, dtype=dtype) * is_wet[:, None].to(dtype)
    elif wet != "dry":
        raise ValueError("wet must be 'random' or 'dry'")
    x = torch.cat([area[:, None], dem[:, None], dyn], 1)
    node_BC = torch.tensor([int(node_ptr[1]) - 1], dtype=torch.long)
    BC = torch.full((1, previous_t, rollout_steps + 1), float(inflow), dtype=dtype)
    # targets (only the training step reads them; the rollout reads just their last dimension)
    y = torch.rand(N, 2, rollout_steps, generator=g1, dtype=dtype) if with_y else torch.empty(0, 2, rollout_steps, dtype=dtype)

    data = Data(
        x=x,
        edge_index=torch.from_numpy(np.concatenate(edges, 1)),
        edge_attr=edge_attr,
        node_ptr=torch.from_numpy(node_ptr),
        edge_ptr=torch.from_numpy(edge_ptr),
        intra_mesh_edge_index=torch.from_numpy(np.concatenate(intra, 1)) if intra
        else torch.zeros(2, 0, dtype=torch.long),
        intra_edge_ptr=torch.from_numpy(intra_edge_ptr),
        node_BC=node_BC, BC=BC, type_BC=2, previous_t=previous_t, temporal_res=120, y=y,
    )
    return data


def make_single_scale_mesh(nx: int, ny: int, previous_t: int = 3, rollout_steps: int = 1,
                           wet: str = "random", wet_fraction: float = 0.3, inflow: float = 0.3,
                           seed: int = 0, dtype=torch.float32) -> Data:
    """Single-scale graph for the SWE-GNN (``GNN``) model: no ``*_ptr`` fields, so the
    reference's ``adapt_batch_training`` takes its single-scale branch (``train.py:22-28``)."""
    d = make_tri_mesh(nx, ny, 1, previous_t, rollout_steps, wet, wet_fraction, inflow, seed, dtype=dtype)
    for k in ("node_ptr", "edge_ptr", "intra_mesh_edge_index", "intra_edge_ptr"):
        delattr(d, k)
    return d


CONFIGS = {
    # name: (nx, ny, num_scales)  — SURVEY.md §8(d)
    "cfg1": (32, 24, 4),        # 2,044 nodes
    "cfg2": (160, 160, 1),      # 51,201 nodes per graph, batch 8 (single-scale GNN)
    "cfg3": (712, 712, 4),      # 1,346,574 nodes
    "cfg4": (2832, 2832, 4),    # 21,303,724 nodes
    "cfg5": (224, 224, 4),      # 133,284 nodes per simulation
}
