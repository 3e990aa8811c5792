"""TEST INFRASTRUCTURE — numpy restatement of the integer plan artefacts.  NOT product code.

Pins `swe_csr_build` (stable destination-CSR) bit for bit.  The property that matters is the one
the reference's aggregation order implies: CPU `Tensor.scatter_add_` (PyG `scatter`,
/root/reference/models/gnn.py:437-438, :256) adds the contributions of a destination in the order
the edges appear in `edge_index`; a STABLE sort by destination keeps exactly that order inside
every CSR segment.
"""
from __future__ import annotations

import numpy as np


def stable_dst_csr(row, col, node_map, dst_lo, n_dst, by_row=False):
    """Returns (rowptr[n_dst+1], src[E], dst[E], eid[E]) as int32 arrays in plan ids."""
    row = np.asarray(row, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    if node_map is not None:
        node_map = np.asarray(node_map, dtype=np.int64)
        row, col = node_map[row], node_map[col]
    key, oth = (row, col) if by_row else (col, row)
    eid = np.argsort(key, kind="stable")
    counts = np.bincount(key - dst_lo, minlength=n_dst)
    rowptr = np.concatenate([[0], np.cumsum(counts)])
    return (rowptr.astype(np.int32), oth[eid].astype(np.int32), key[eid].astype(np.int32), eid.astype(np.int32))


def stable_dst_csr_loops(row, col, dst_lo, n_dst):
    """Pure-Python version for tiny cases (independent of numpy's sort)."""
    buckets = [[] for _ in range(n_dst)]
    for e, (r, c) in enumerate(zip(row, col)):
        buckets[c - dst_lo].append((e, r))
    rowptr, src, dst, eid = [0], [], [], []
    for i, b in enumerate(buckets):
        for e, r in b:
            eid.append(e)
            src.append(r)
            dst.append(dst_lo + i)
        rowptr.append(len(eid))
    return rowptr, src, dst, eid


def batch_permutation(node_ptr_2d):
    """plan -> original node permutation for an adapted batch: scale-major, graph-minor
    (node_ptr [G, S+1] cumulative, /root/reference/training/train.py:48-60)."""
    ptr = np.asarray(node_ptr_2d, dtype=np.int64)
    G, S1 = ptr.shape
    perm = np.concatenate([np.arange(ptr[g, s], ptr[g, s + 1]) for s in range(S1 - 1) for g in range(G)])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    return perm.astype(np.int32), inv.astype(np.int32)
