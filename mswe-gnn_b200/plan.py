"""Graph plan: the integer artefacts the kernels need, derived once per mesh topology.

The reference walks ``edge_index`` in its given (row-sorted) order and aggregates with
``scatter_add_`` (``models/gnn.py:437-438``); its GPU path uses atomics.  Here every edge set is
re-ordered ONCE into a *stable* destination-CSR (``swe_csr_build``), so aggregation is a
sequential in-segment sum in the original edge order: deterministic, atomics-free and bit-equal
to the CPU ``scatter_add_`` order.  Meshes are static across a rollout, so plans are cached.

Plan order of nodes: scale-major, graph-minor.  For a single graph that is the identity; for an
adapted batch (``node_ptr`` of shape ``[G, S+1]``, reference ``training/train.py:48-60``) the
nodes of each scale are made contiguous, which lets every kernel work on a plain row range
instead of the reference's ``(mask == i)`` selections (``models/gnn.py:307,322,331``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import lib


@dataclass
class EdgeSet:
    """Stable destination-CSR of one edge set (all ids int32, plan order)."""
    rowptr: torch.Tensor          # [n_dst + 1]
    src: torch.Tensor             # [E]
    dst: torch.Tensor             # [E]  non-decreasing
    eid: torch.Tensor             # [E]  original edge index (relative to the edge-set slice)
    n_edges: int
    dst_lo: int
    n_dst: int
    src_lo: int
    src_hi: int
    # transposed view (grouped by src, stable in CSR position) — built lazily for the backward pass
    t_rowptr: Optional[torch.Tensor] = None
    t_pos: Optional[torch.Tensor] = None
    # largest number of edges of 4 consecutive destinations (the s-ring hop stages 12 per such block; -1: unknown)
    max_block4: int = -1


def _build_edge_set(row, col, node_map, dst_lo, n_dst, src_lo, src_hi) -> EdgeSet:
    rowptr, src, dst, eid = lib.csr_build(row.contiguous(), col.contiguous(), node_map, dst_lo, n_dst, src_lo, src_hi)
    es = EdgeSet(rowptr, src, dst, eid, int(row.numel()), dst_lo, n_dst, src_lo, src_hi)
    if n_dst > 0:
        # blocks as the s-ring hop cuts them: 4 destinations starting at multiples of 4 (one host read per edge set, once
        # per topology)
        idx = torch.arange(0, n_dst + 4, 4, device=rowptr.device).clamp_(max=n_dst)
        rp = rowptr[idx.long()]
        es.max_block4 = int((rp[1:] - rp[:-1]).max()) if rp.numel() > 1 else 0
    else:
        es.max_block4 = 0
    return es


@dataclass
class GraphPlan:
    n_nodes: int
    num_scales: int
    scale_lo: List[int]
    scale_n: List[int]
    perm: Optional[torch.Tensor]          # int32 [N] plan -> original, None = identity
    inv: Optional[torch.Tensor]           # int32 [N] original -> plan
    edges: List[EdgeSet]                  # per scale
    pool: List[EdgeSet] = field(default_factory=list)      # level j: keyed by coarse (scale j+1), src = fine
    unpool: List[EdgeSet] = field(default_factory=list)    # level j: keyed by fine (scale j), src = coarse
    edge_slices: List[Tuple[int, int]] = field(default_factory=list)   # [lo, hi) of each scale in edge_index
    key: tuple = ()
    # partitioned meshes (parallel.partition_graph): per scale the first scale_owned[s] rows are owned by this rank, the
    # rest are halo copies; every edge set then ends in owned rows only (None: all rows are owned)
    scale_owned: Optional[List[int]] = None

    @property
    def n_edges_total(self) -> int:
        return sum(e.n_edges for e in self.edges)

    @property
    def max_edges(self) -> int:
        m = max([e.n_edges for e in self.edges] + [1])
        if self.unpool:
            m = max(m, max(e.n_edges for e in self.unpool))
        return m


def _topology_key(graph, multiscale: bool) -> tuple:
    ei = graph.edge_index
    key = [ei.data_ptr(), tuple(ei.shape), ei._version, int(graph.x.shape[0])]
    if multiscale:
        for name in ("node_ptr", "edge_ptr", "intra_mesh_edge_index", "intra_edge_ptr"):
            t = getattr(graph, name)
            key += [t.data_ptr(), tuple(t.shape), t._version]
    return tuple(key)


def build_plan(graph, num_scales: int, multiscale: bool) -> GraphPlan:
    """Derive the plan of ``graph`` (a PyG-like ``Data`` / adapted ``Batch``)."""
    ei = graph.edge_index
    if not ei.is_cuda:
        raise RuntimeError("graph tensors must live on a CUDA device (no CPU fallback)")
    dev = ei.device
    N = int(graph.x.shape[0])
    if ei.dtype != torch.int64 or ei.dim() != 2 or ei.shape[0] != 2:
        raise ValueError("edge_index must be an int64 tensor of shape [2, E]")
    if ei.numel() and (int(ei.min()) < 0 or int(ei.max()) >= N):
        raise ValueError("edge_index refers to nodes outside [0, num_nodes)")
    key = _topology_key(graph, multiscale)

    if not multiscale:
        owned = getattr(graph, "n_owned", None)
        n_dst = N if owned is None else int(owned[0])
        es = _build_edge_set(ei[0], ei[1], None, 0, n_dst, 0, N)
        return GraphPlan(N, 1, [0], [N], None, None, [es], edge_slices=[(0, int(ei.shape[1]))], key=key,
                         scale_owned=None if owned is None else [n_dst])

    S = num_scales
    node_ptr = graph.node_ptr.detach().to("cpu", torch.int64)
    edge_ptr = [int(v) for v in graph.edge_ptr.detach().to("cpu").tolist()]
    intra_ptr = [int(v) for v in graph.intra_edge_ptr.detach().to("cpu").tolist()]
    ie = graph.intra_mesh_edge_index
    if node_ptr.shape[-1] != S + 1:
        raise ValueError(f"node_ptr has {node_ptr.shape[-1]} entries per graph, expected num_scales+1 = {S + 1}")
    if len(edge_ptr) != S + 1 or len(intra_ptr) != S:
        raise ValueError("edge_ptr must have num_scales+1 entries and intra_edge_ptr num_scales entries")
    if ie.numel() and (int(ie.min()) < 0 or int(ie.max()) >= N):
        raise ValueError("intra_mesh_edge_index refers to nodes outside [0, num_nodes)")
    ptr2 = node_ptr.reshape(-1, S + 1)
    G = ptr2.shape[0]
    if int(ptr2[-1, -1]) != N or int(ptr2[0, 0]) != 0 or bool((ptr2[:, 1:] < ptr2[:, :-1]).any()):
        raise ValueError("node_ptr is not a cumulative partition of the nodes")

    perm = inv = None
    if G == 1:
        scale_lo = [int(v) for v in ptr2[0, :-1]]
        scale_n = [int(ptr2[0, s + 1] - ptr2[0, s]) for s in range(S)]
    else:
        pieces = [torch.arange(int(ptr2[g, s]), int(ptr2[g, s + 1]), dtype=torch.int64)
                  for s in range(S) for g in range(G)]
        perm64 = torch.cat(pieces)
        if perm64.numel() != N:
            raise ValueError("node_ptr ranges do not cover every node exactly once")
        scale_n = [int((ptr2[:, s + 1] - ptr2[:, s]).sum()) for s in range(S)]
        scale_lo = [0]
        for s in range(S - 1):
            scale_lo.append(scale_lo[-1] + scale_n[s])
        inv64 = torch.empty(N, dtype=torch.int64)
        inv64[perm64] = torch.arange(N, dtype=torch.int64)
        perm = perm64.to(torch.int32).to(dev)
        inv = inv64.to(torch.int32).to(dev)

    # A rank's part of a partitioned mesh: destinations are the OWNED rows only (the halo rows behind them are written by
    # their owners' exchanges, never by a local kernel); inter-scale edges whose destination is a halo row are dropped
    owned = getattr(graph, "n_owned", None)
    if owned is not None:
        owned = [int(v) for v in owned]
        if G != 1 or len(owned) != S or any(o < 0 or o > n for o, n in zip(owned, scale_n)):
            raise ValueError("n_owned must give, for an un-batched graph, the number of owned rows of every scale")
    dst_n = owned if owned is not None else scale_n

    edges, pool, unpool, slices = [], [], [], []
    for s in range(S):
        lo, hi = edge_ptr[s], edge_ptr[s + 1]
        slices.append((lo, hi))
        try:
            edges.append(_build_edge_set(ei[0, lo:hi], ei[1, lo:hi], inv, scale_lo[s], dst_n[s],
                                         scale_lo[s], scale_lo[s] + scale_n[s]))
        except ValueError as e:
            raise ValueError(f"scale {s}: edges must connect nodes of the same scale"
                             f"{' and end in owned rows' if owned is not None else ''}: {e}") from e
    for j in range(S - 1):
        lo, hi = intra_ptr[j], intra_ptr[j + 1]
        coarse, fine = ie[0, lo:hi], ie[1, lo:hi]
        f_lo, f_n, c_lo, c_n = scale_lo[j], scale_n[j], scale_lo[j + 1], scale_n[j + 1]
        pc, pf, uc, uf = coarse, fine, coarse, fine
        if owned is not None:
            keep_p = coarse < c_lo + dst_n[j + 1]                    # pooling: destination = coarse row
            keep_u = fine < f_lo + dst_n[j]                          # un-pooling: destination = fine row
            pc, pf, uc, uf = coarse[keep_p], fine[keep_p], coarse[keep_u], fine[keep_u]
        try:
            pool.append(_build_edge_set(pf, pc, inv, c_lo, dst_n[j + 1], f_lo, f_lo + f_n))
            unpool.append(_build_edge_set(uc, uf, inv, f_lo, dst_n[j], c_lo, c_lo + c_n))
        except ValueError as e:
            raise ValueError(f"inter-scale level {j}: row 0 must hold scale-{j + 1} (coarse) and row 1 "
                             f"scale-{j} (fine) nodes: {e}") from e
    return GraphPlan(N, S, scale_lo, scale_n, perm, inv, edges, pool, unpool, slices, key, scale_owned=owned)


class PlanCache:
    """Small per-model cache keyed on the identity of the topology tensors."""

    def __init__(self, capacity: int = 4):
        self.capacity = capacity
        self._items: Dict[tuple, GraphPlan] = {}

    def get(self, graph, num_scales: int, multiscale: bool) -> GraphPlan:
        key = _topology_key(graph, multiscale)
        hit = self._items.get(key)
        if hit is None:
            plan = build_plan(graph, num_scales, multiscale)
            if len(self._items) >= self.capacity:
                self._items.pop(next(iter(self._items)))
            # the key is built from data_ptr()s: keep the tensors alive for as long as the entry lives, otherwise
            # the allocator may hand the same address to a DIFFERENT graph of the same shape and the key would lie
            keep = [graph.edge_index] + ([getattr(graph, n) for n in ("node_ptr", "edge_ptr", "intra_mesh_edge_index",
                                                                       "intra_edge_ptr")] if multiscale else [])
            plan._key_tensors = keep
            self._items[key] = plan
            return plan
        return hit

    def clear(self):
        self._items.clear()
