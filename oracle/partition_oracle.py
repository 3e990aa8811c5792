"""TEST INFRASTRUCTURE — plain-Python-loop restatement of the graph partitioner
(`mswe-gnn_b200/parallel.py`).  The reference has no graph partitioning, so
there is no reference code to cite: this file states the RULES independently with loops and dicts, and
`tests/test_partition.py` checks the vectorised NumPy product code against it bit for bit.

Rules
  owner(coarsest node i of nc)  = min(i * P // nc, P - 1)
  owner(finer node)             = owner of its first parent in intra_mesh_edge_index order
  owner(parent-less node)       = owner of the first edge neighbour that has one (as source first), else 0
  local nodes of rank r, scale s = owned (ascending global id) then halo (by owner, then global id)
  halo(r, s)  = sources of scale-s edges ending in an r-owned node, children (scale s) of r-owned
                parents and parents (scale s) of r-owned children that r does not own
  local edges = edges ending in an r-owned node, global order kept
"""
from __future__ import annotations


def owner_map_loops(node_ptr, edge_index, intra, intra_ptr, world):
    S = len(node_ptr) - 1
    N = int(node_ptr[-1])
    owner = [-1] * N
    lo, hi = int(node_ptr[S - 1]), int(node_ptr[S])
    nc = hi - lo
    for i in range(nc):
        owner[lo + i] = min(i * world // max(nc, 1), world - 1)
    for j in range(S - 2, -1, -1):
        seen = set()
        for e in range(int(intra_ptr[j]), int(intra_ptr[j + 1])):
            c, f = int(intra[0][e]), int(intra[1][e])
            if f not in seen:
                seen.add(f)
                owner[f] = owner[c]
    E = len(edge_index[0])
    for _ in range(4):
        if all(o >= 0 for o in owner):
            break
        snap = list(owner)
        done = set()
        for e in range(E):
            r, c = int(edge_index[0][e]), int(edge_index[1][e])
            if snap[r] < 0 and snap[c] >= 0 and r not in done:
                owner[r] = snap[c]; done.add(r)
        snap = list(owner)
        done = set()
        for e in range(E):
            r, c = int(edge_index[0][e]), int(edge_index[1][e])
            if snap[c] < 0 and snap[r] >= 0 and c not in done:
                owner[c] = snap[r]; done.add(c)
    return [o if o >= 0 else 0 for o in owner]


def local_sets_loops(node_ptr, edge_index, edge_ptr, intra, intra_ptr, owner, rank):
    """Returns (local_to_global list, per-scale halo lists, per-scale {peer: send global ids},
    local edge list [(src_g, dst_g)] per scale)."""
    S = len(node_ptr) - 1
    halo = [set() for _ in range(S)]
    send = [dict() for _ in range(S)]

    def want(scale, src, dst):
        if owner[src] != owner[dst]:
            if owner[dst] == rank:
                halo[scale].add(src)
            if owner[src] == rank:
                send[scale].setdefault(owner[dst], set()).add(src)

    edges = []
    for s in range(S):
        loc = []
        for e in range(int(edge_ptr[s]), int(edge_ptr[s + 1])):
            r, c = int(edge_index[0][e]), int(edge_index[1][e])
            want(s, r, c)
            if owner[c] == rank:
                loc.append((r, c))
        edges.append(loc)
    for j in range(S - 1):
        for e in range(int(intra_ptr[j]), int(intra_ptr[j + 1])):
            c, f = int(intra[0][e]), int(intra[1][e])
            want(j, f, c)
            want(j + 1, c, f)
    l2g, halos = [], []
    for s in range(S):
        owned = [n for n in range(int(node_ptr[s]), int(node_ptr[s + 1])) if owner[n] == rank]
        h = sorted(halo[s], key=lambda n: (owner[n], n))
        l2g += owned + h
        halos.append(h)
    sends = [{q: sorted(v) for q, v in d.items()} for d in send]
    return l2g, halos, sends, edges
