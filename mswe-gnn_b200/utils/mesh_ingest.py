"""Mesh → multi-scale graph ingest: what the reference's ``database/graph_creation.py`` does with
``meshkernel`` / NetCDF inputs (``MultiscaleMesh.stack_meshes`` 866-909, ``get_intra_edges`` 911-931,
``connect_coarse_to_fine_mesh`` 422-436, ghost cells 1391-1400, ``convert_mesh_to_pyg`` 1483-1582), restated for a neutral
input — per level the node coordinates and the face → node table of an unstructured mesh (``.npz``; the D-Hydro readers are
not available offline) — with the integer artefacts the hot path consumes as output:

* graph nodes = mesh faces, edges = the dual graph, both directions, in (row, col) order;
* one ghost face per level (appended last) with a directed ghost → boundary-face edge; the boundary face of a level is the
  one whose boundary edge midpoint is closest to the given inflow location (``interpolate_BC_location_multiscale``);
* levels stacked fine → coarse with global ids (``node_ptr``, ``edge_ptr``);
* inter-scale edges (coarse, fine) wherever a fine face centre lies inside a coarse face, ordered by coarse then fine
  (``intra_mesh_edge_index``, ``intra_edge_ptr``) — not necessarily a tree on real meshes (Appendix D-5);
* ``node_BC`` = the finest ghost face only (``graph_creation.py:1577``);
* optionally a locality-preserving renumbering of the faces of every level (Morton order of the face centres), which keeps
  the gathers of the kernels local and makes the partitioner's contiguous blocks of coarse faces compact.

Everything is vectorised NumPy (sorting / unique / uniform-grid binning for the containment test): the ingest runs once per
mesh on the host, like the reference's.  ``tests/test_mesh_ingest.py`` feeds the structured ``tri(nx, ny)`` geometry through
this general path and gets ``utils.synthetic.make_tri_mesh``'s arrays back bit for bit.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .data import Data


def face_centres(node_xy: np.ndarray, face_nodes: np.ndarray) -> np.ndarray:
    """Mean of the face's vertices (``get_barycenter``, graph_creation.py:438-454); -1 pads faces with fewer nodes."""
    valid = face_nodes >= 0
    xy = node_xy[np.where(valid, face_nodes, 0)] * valid[..., None]
    return xy.sum(1) / valid.sum(1, keepdims=True)


def dual_graph(face_nodes: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Directed dual edges [2, E] in (row, col) order, and the boundary edges: (face id, [node a, node b]) of every mesh
    edge that belongs to one face only."""
    F, K = face_nodes.shape
    valid = face_nodes >= 0
    n_per = valid.sum(1)
    nxt = np.where(np.arange(K)[None, :] + 1 < n_per[:, None], np.roll(face_nodes, -1, 1), face_nodes[:, :1])
    a, b = face_nodes[valid], nxt[valid]
    face = np.repeat(np.arange(F, dtype=np.int64), K).reshape(F, K)[valid]
    lo, hi = np.minimum(a, b).astype(np.int64), np.maximum(a, b).astype(np.int64)
    key = lo * (int(face_nodes.max()) + 1) + hi
    order = np.argsort(key, kind="stable")
    key_s, face_s = key[order], face[order]
    same = key_s[1:] == key_s[:-1]
    if (same[1:] & same[:-1]).any():
        raise ValueError("a mesh edge is shared by more than two faces")
    f0, f1 = face_s[:-1][same], face_s[1:][same]
    row, col = np.concatenate([f0, f1]), np.concatenate([f1, f0])
    o = np.lexsort((col, row))
    interior = np.zeros(key_s.shape[0], dtype=bool)
    interior[:-1] |= same
    interior[1:] |= same
    bsel = order[~interior]
    return np.stack([row[o], col[o]]), face[bsel], np.stack([a[bsel], b[bsel]], 1).astype(np.int64)


def containment_edges(coarse_xy: np.ndarray, coarse_faces: np.ndarray, fine_centres: np.ndarray) -> np.ndarray:
    """(coarse, fine) pairs with the fine face centre strictly inside the (convex or not) coarse polygon — the even-odd
    rule of ``matplotlib.path.Path.contains_points`` (graph_creation.py:430-433) — ordered by coarse then fine.  Candidates
    come from a uniform grid over the coarse faces' bounding boxes instead of the reference's all-pairs test."""
    Fc, K = coarse_faces.shape
    valid = coarse_faces >= 0
    P = coarse_xy[np.where(valid, coarse_faces, coarse_faces[:, :1])]           # [Fc, K, 2], padded with the first vertex
    lo, hi = P.min(1), P.max(1)
    g = max(1, int(np.sqrt(Fc)))
    origin, span = lo.min(0), np.maximum(hi.max(0) - lo.min(0), 1e-30)
    cell = lambda xy: np.clip(((xy - origin) / span * g).astype(np.int64), 0, g - 1)
    c_lo, c_hi = cell(lo), cell(hi)
    f_cell = cell(fine_centres)
    # bucket the coarse faces by every grid cell their bounding box touches
    nx_ = c_hi[:, 0] - c_lo[:, 0] + 1
    ny_ = c_hi[:, 1] - c_lo[:, 1] + 1
    cnt = nx_ * ny_
    face_rep = np.repeat(np.arange(Fc, dtype=np.int64), cnt)
    k = np.arange(cnt.sum(), dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    cx = c_lo[face_rep, 0] + k % nx_[face_rep]
    cy = c_lo[face_rep, 1] + k // nx_[face_rep]
    bucket = cx * g + cy
    order = np.argsort(bucket, kind="stable")
    bucket_s, face_s = bucket[order], face_rep[order]
    starts = np.searchsorted(bucket_s, np.arange(g * g))
    ends = np.searchsorted(bucket_s, np.arange(g * g), side="right")
    fb = f_cell[:, 0] * g + f_cell[:, 1]
    n_cand = ends[fb] - starts[fb]
    fine_rep = np.repeat(np.arange(fine_centres.shape[0], dtype=np.int64), n_cand)
    kk = np.arange(n_cand.sum(), dtype=np.int64) - np.repeat(np.cumsum(n_cand) - n_cand, n_cand)
    coarse_c = face_s[starts[fb][fine_rep] + kk]
    # even-odd crossing test of every candidate pair
    pt = fine_centres[fine_rep]
    poly = P[coarse_c]
    x0, y0 = poly[:, :, 0], poly[:, :, 1]
    x1, y1 = np.roll(x0, -1, 1), np.roll(y0, -1, 1)
    px, py = pt[:, :1], pt[:, 1:2]
    crosses = ((y0 > py) != (y1 > py)) & (px < (x1 - x0) * (py - y0) / np.where(y1 != y0, y1 - y0, 1.0) + x0)
    inside = (crosses.sum(1) % 2) == 1
    c, f = coarse_c[inside], fine_rep[inside]
    o = np.lexsort((f, c))
    return np.stack([c[o], f[o]])


def morton_order(xy: np.ndarray, bits: int = 16) -> np.ndarray:
    """Permutation that sorts points along the Z-order curve of their (quantised) coordinates."""
    lo, span = xy.min(0), np.maximum(xy.max(0) - xy.min(0), 1e-30)
    q = np.minimum(((xy - lo) / span * (1 << bits)).astype(np.uint64), (1 << bits) - 1)

    def spread(v):
        v = (v | (v << 16)) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << 8)) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << 4)) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << 2)) & np.uint64(0x3333333333333333)
        v = (v | (v << 1)) & np.uint64(0x5555555555555555)
        return v
    return np.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)), kind="stable")


def build_multiscale_graph(levels: Sequence[Dict[str, np.ndarray]], inflow_xy: Sequence[float], previous_t: int = 2,
                           rollout_steps: int = 1, inflow: float = 0.0, renumber: bool = False, type_BC: int = 2,
                           dtype=torch.float32) -> Data:
    """levels: fine → coarse, each ``{"node_xy": [n, 2] float, "face_nodes": [F, K] int (-1 padded)}``; inflow_xy: location of
    the inflow boundary (its closest boundary edge of every level gets that level's ghost face).  Returns a ``Data`` with the
    fields of SURVEY.md Appendix C; node features: [face area, 0 (DEM), zeros(2 previous_t)]; edge feature: centre distance."""
    S = len(levels)
    node_ptr, edge_ptr, intra_ptr = [0], [0], [0]
    edges, intras, areas, dists, centres_all = [], [], [], [], []
    prepared = []
    for lv in levels:
        xy, fn = np.asarray(lv["node_xy"], dtype=np.float64), np.asarray(lv["face_nodes"], dtype=np.int64)
        ctr = face_centres(xy, fn)
        if renumber:
            perm = morton_order(ctr)
            fn, ctr = fn[perm], ctr[perm]
        prepared.append((xy, fn, ctr))
    for s, (xy, fn, ctr) in enumerate(prepared):
        F = fn.shape[0]
        ei, b_face, b_nodes = dual_graph(fn)
        if b_face.size == 0:
            raise ValueError(f"level {s}: the mesh has no boundary edge")
        mid = xy[b_nodes].mean(1)
        bc_face = int(b_face[np.argmin(((mid - np.asarray(inflow_xy, dtype=np.float64)) ** 2).sum(1))])
        ghost = F                                                            # appended as the last face of the level
        e = np.concatenate([ei, np.array([[ghost], [bc_face]], dtype=np.int64)], 1) + node_ptr[-1]
        edges.append(e)
        # shoelace area; the ghost face mirrors its boundary face
        valid = fn >= 0
        P = xy[np.where(valid, fn, fn[:, :1])]
        x0, y0 = P[:, :, 0], P[:, :, 1]
        area = 0.5 * np.abs((x0 * np.roll(y0, -1, 1) - np.roll(x0, -1, 1) * y0).sum(1))
        areas.append(np.concatenate([area, area[bc_face:bc_face + 1]]))
        gctr = 2 * mid[np.argmin(((mid - np.asarray(inflow_xy, dtype=np.float64)) ** 2).sum(1))] - ctr[bc_face]
        ctr_g = np.concatenate([ctr, gctr[None]])
        centres_all.append(ctr_g)
        loc = e - node_ptr[-1]
        dists.append(np.sqrt(((ctr_g[loc[0]] - ctr_g[loc[1]]) ** 2).sum(1)))
        node_ptr.append(node_ptr[-1] + F + 1)
        edge_ptr.append(edge_ptr[-1] + e.shape[1])
    for s in range(S - 1):
        xy_c, fn_c, _ = prepared[s + 1]
        c = containment_edges(xy_c, fn_c, prepared[s][2])
        c[0] += node_ptr[s + 1]
        c[1] += node_ptr[s]
        intras.append(c)
        intra_ptr.append(intra_ptr[-1] + c.shape[1])
    N = node_ptr[-1]
    x = torch.zeros(N, 2 + 2 * previous_t, dtype=dtype)
    x[:, 0] = torch.from_numpy(np.concatenate(areas)).to(dtype)
    return Data(
        x=x, edge_index=torch.from_numpy(np.concatenate(edges, 1)),
        edge_attr=torch.from_numpy(np.concatenate(dists)).to(dtype)[:, None],
        node_ptr=torch.tensor(node_ptr, dtype=torch.long), edge_ptr=torch.tensor(edge_ptr, dtype=torch.long),
        intra_mesh_edge_index=torch.from_numpy(np.concatenate(intras, 1)) if intras else torch.zeros(2, 0, dtype=torch.long),
        intra_edge_ptr=torch.tensor(intra_ptr, dtype=torch.long),
        node_BC=torch.tensor([node_ptr[1] - 1], dtype=torch.long),
        BC=torch.full((1, previous_t, rollout_steps + 1), float(inflow), dtype=dtype), type_BC=type_BC, previous_t=previous_t,
        temporal_res=120, y=torch.empty(0, 2, rollout_steps, dtype=dtype),
        pos=torch.from_numpy(np.concatenate(centres_all)).to(dtype))


def save_levels(path: str, levels: Sequence[Dict[str, np.ndarray]]):
    """The neutral on-disk format: ``node_xy_<s>`` / ``face_nodes_<s>`` per level, fine → coarse."""
    np.savez_compressed(path, n_levels=len(levels), **{f"{k}_{s}": lv[k] for s, lv in enumerate(levels) for k in ("node_xy", "face_nodes")})


def load_levels(path: str) -> List[Dict[str, np.ndarray]]:
    z = np.load(path)
    return [{k: z[f"{k}_{s}"] for k in ("node_xy", "face_nodes")} for s in range(int(z["n_levels"]))]


def structured_tri_levels(nx: int, ny: int, num_scales: int) -> List[Dict[str, np.ndarray]]:
    """Geometry of the synthetic ``tri(nx, ny)`` family (utils/synthetic.py) as plain mesh levels: unit squares (2^s at
    level s) cut by their (0,0)-(1,1) diagonal; face 2 (j nx + i) + t, t = 0 lower-right, t = 1 upper-left."""
    out = []
    for s in range(num_scales):
        a, b, h = nx >> s, ny >> s, float(1 << s)
        jj, ii = np.meshgrid(np.arange(b + 1), np.arange(a + 1), indexing="ij")
        node_xy = np.stack([ii.ravel() * h, jj.ravel() * h], 1)
        nid = lambda i, j: j * (a + 1) + i
        j, i = np.meshgrid(np.arange(b), np.arange(a), indexing="ij")
        i, j = i.ravel(), j.ravel()
        t0 = np.stack([nid(i, j), nid(i + 1, j), nid(i + 1, j + 1)], 1)
        t1 = np.stack([nid(i, j), nid(i + 1, j + 1), nid(i, j + 1)], 1)
        out.append({"node_xy": node_xy, "face_nodes": np.stack([t0, t1], 1).reshape(-1, 3).astype(np.int64)})
    return out
