"""Multi-GPU modes of the hot path (one process per GPU, ``torch.distributed``).

The reference has no explicit distributed code (only Lightning-implicit DDP), so
this module is new functionality with two modes (``BASELINE.json`` north_star):

* **Partitioned large-mesh rollout** — the graph is cut by blocks of coarsest-level cells; every
  finer node inherits the owner of its (first) parent, so pooling / un-pooling are local; every
  level is sharded (no replicated level, hence no all-gather).  A rank's local graph holds, per
  scale, its owned nodes followed by the halo nodes (sources of edges that end in an owned node),
  grouped by owner, and exactly the edges that end in an owned node IN GLOBAL EDGE ORDER — so each
  in-segment sum runs over the same edges in the same order as on one GPU and owned rows are
  bit-identical to the single-GPU result.  Halo rows are refreshed by neighbour exchange: once
  before each SWEGNN call (its input x_d) and after every hop but the last; the node inputs ``x``
  (8 floats) once per rollout step.  Receives land directly in the halo row range (contiguous
  per peer); sends are packed by ``swe_pack_rows``.
* **Data-parallel training** — independent simulations per rank, one all-reduce of the flat fp32
  gradient per optimizer step (what Lightning DDP does for the reference, ``main.py:107``).

Integer artefacts (owner map, local numbering, send / receive lists) are checked bit-exactly against
the loop restatement in ``oracle/partition_oracle.py``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import new_static_token, static_inputs
from .utils.data import Data


# ------------------------------------------------------------------------------------------------
# partitioning (host, NumPy, int64)
# ------------------------------------------------------------------------------------------------
def owner_map(node_ptr: np.ndarray, edge_index: np.ndarray, edge_ptr: np.ndarray, intra: np.ndarray,
              intra_ptr: np.ndarray, world: int) -> np.ndarray:
    """Owner rank of every node.  Coarsest scale: contiguous index blocks; finer scales: owner of
    the FIRST parent in ``intra_mesh_edge_index`` order; nodes without a parent (ghost cells,
    orphans): owner of the first edge neighbour that has one, else rank 0."""
    S = len(node_ptr) - 1
    N = int(node_ptr[-1])
    owner = np.full(N, -1, dtype=np.int64)
    lo, hi = int(node_ptr[S - 1]), int(node_ptr[S])
    nc = hi - lo
    owner[lo:hi] = np.minimum((np.arange(nc, dtype=np.int64) * world) // max(nc, 1), world - 1)
    for j in range(S - 2, -1, -1):
        a, b = int(intra_ptr[j]), int(intra_ptr[j + 1])
        coarse, fine = intra[0, a:b], intra[1, a:b]
        uf, first = np.unique(fine, return_index=True)
        owner[uf] = owner[coarse[first]]
    for _ in range(4):                      # parent-less nodes: take the owner of an edge neighbour
        todo = owner < 0
        if not todo.any():
            break
        row, col = edge_index[0], edge_index[1]
        m = todo[row] & (owner[col] >= 0)
        if m.any():
            ur, first = np.unique(row[m], return_index=True)
            owner[ur] = owner[col[m][first]]
        todo = owner < 0
        m = todo[col] & (owner[row] >= 0)
        if m.any():
            uc, first = np.unique(col[m], return_index=True)
            owner[uc] = owner[row[m][first]]
    owner[owner < 0] = 0
    return owner


@dataclass
class LocalPartition:
    """Everything one rank needs: its local graph and the exchange lists, per scale."""
    rank: int
    world: int
    num_scales: int
    graph: Data                                  # local multiscale graph (CPU tensors)
    local_to_global: np.ndarray                  # [N_local]
    n_owned: List[int]                           # per scale
    n_halo: List[int]
    scale_lo: List[int]                          # first local row of each scale
    # per scale: {peer: int64 local rows to send (sorted by global id)}, {peer: (first local row, count)}
    send: List[Dict[int, np.ndarray]] = field(default_factory=list)
    recv: List[Dict[int, Tuple[int, int]]] = field(default_factory=list)
    inter_cross: List[bool] = field(default_factory=list)   # inter-scale level j has edges crossing ranks
    owned_rows: Optional[np.ndarray] = None      # local rows that are owned (all scales)
    owned_global: Optional[np.ndarray] = None    # their global ids


def partition_graph(graph, world: int, rank: int, owner: Optional[np.ndarray] = None) -> LocalPartition:
    """Cut a multiscale graph (fields of SURVEY.md Appendix C, CPU tensors) for ``rank`` of ``world``.
    A single-scale graph (no ``node_ptr``) is treated as one scale."""
    x = graph.x
    N = int(x.shape[0])
    ei = graph.edge_index.numpy()
    multiscale = hasattr(graph, "node_ptr") and getattr(graph, "node_ptr") is not None
    if multiscale:
        node_ptr = graph.node_ptr.numpy().astype(np.int64).reshape(-1)
        edge_ptr = graph.edge_ptr.numpy().astype(np.int64)
        intra = graph.intra_mesh_edge_index.numpy()
        intra_ptr = graph.intra_edge_ptr.numpy().astype(np.int64)
    else:
        node_ptr = np.array([0, N], dtype=np.int64)
        edge_ptr = np.array([0, ei.shape[1]], dtype=np.int64)
        intra = np.zeros((2, 0), dtype=np.int64)
        intra_ptr = np.array([0], dtype=np.int64)
    S = len(node_ptr) - 1
    if owner is None:
        owner = owner_map(node_ptr, ei, edge_ptr, intra, intra_ptr, world)
    r = rank

    halo_sets: List[List[np.ndarray]] = [[] for _ in range(S)]          # global ids wanted as halo, per scale
    send_sets: List[Dict[int, List[np.ndarray]]] = [dict() for _ in range(S)]

    def want(scale: int, src: np.ndarray, dst: np.ndarray):
        """edges src -> dst whose value at src is consumed at dst: halo / send bookkeeping."""
        os_, od = owner[src], owner[dst]
        cross = os_ != od
        mine = cross & (od == r)
        if mine.any():
            halo_sets[scale].append(src[mine])
        out = cross & (os_ == r)
        if out.any():
            peers = od[out]
            nodes = src[out]
            for q in np.unique(peers):
                send_sets[scale].setdefault(int(q), []).append(nodes[peers == q])
        return bool(cross.any())

    for s in range(S):
        a, b = int(edge_ptr[s]), int(edge_ptr[s + 1])
        want(s, ei[0, a:b], ei[1, a:b])
    inter_cross = []
    for j in range(S - 1):
        a, b = int(intra_ptr[j]), int(intra_ptr[j + 1])
        coarse, fine = intra[0, a:b], intra[1, a:b]
        c1 = want(j, fine, coarse)          # pooling reads fine rows at the coarse owner
        c2 = want(j + 1, coarse, fine)      # un-pooling reads coarse rows at the fine owner
        inter_cross.append(c1 or c2)

    g2l = np.full(N, -1, dtype=np.int64)
    l2g_parts, n_owned, n_halo, scale_lo = [], [], [], []
    recv: List[Dict[int, Tuple[int, int]]] = []
    cursor = 0
    for s in range(S):
        lo, hi = int(node_ptr[s]), int(node_ptr[s + 1])
        ids = np.arange(lo, hi, dtype=np.int64)
        owned = ids[owner[lo:hi] == r]
        halo = np.unique(np.concatenate(halo_sets[s])) if halo_sets[s] else np.zeros(0, dtype=np.int64)
        halo = halo[np.lexsort((halo, owner[halo]))]                     # by owner, then global id
        scale_lo.append(cursor)
        g2l[owned] = cursor + np.arange(owned.size)
        g2l[halo] = cursor + owned.size + np.arange(halo.size)
        rmap: Dict[int, Tuple[int, int]] = {}
        if halo.size:
            ho = owner[halo]
            for q in np.unique(ho):
                idx = np.nonzero(ho == q)[0]
                rmap[int(q)] = (cursor + owned.size + int(idx[0]), int(idx.size))
        recv.append(rmap)
        l2g_parts += [owned, halo]
        n_owned.append(int(owned.size)); n_halo.append(int(halo.size))
        cursor += owned.size + halo.size
    l2g = np.concatenate(l2g_parts) if l2g_parts else np.zeros(0, dtype=np.int64)
    send: List[Dict[int, np.ndarray]] = []
    for s in range(S):
        send.append({q: g2l[np.unique(np.concatenate(v))] for q, v in sorted(send_sets[s].items())})

    # ---- local graph: edges that end in an owned node, in global order; inter-scale edges with an owned end
    e_parts, ea_parts, e_counts = [], [], []
    for s in range(S):
        a, b = int(edge_ptr[s]), int(edge_ptr[s + 1])
        keep = np.nonzero(owner[ei[1, a:b]] == r)[0] + a
        e_parts.append(g2l[ei[:, keep]])
        ea_parts.append(keep)
        e_counts.append(keep.size)
    i_parts, i_counts = [], []
    for j in range(S - 1):
        a, b = int(intra_ptr[j]), int(intra_ptr[j + 1])
        keep = np.nonzero((owner[intra[0, a:b]] == r) | (owner[intra[1, a:b]] == r))[0] + a
        i_parts.append(g2l[intra[:, keep]])
        i_counts.append(keep.size)
    e_keep = np.concatenate(ea_parts) if ea_parts else np.zeros(0, dtype=np.int64)
    local = Data(
        x=x[torch.from_numpy(l2g)].contiguous(),
        edge_index=torch.from_numpy(np.concatenate(e_parts, 1) if e_parts else np.zeros((2, 0), dtype=np.int64)).contiguous(),
        edge_attr=graph.edge_attr[torch.from_numpy(e_keep)].contiguous(),
    )
    if multiscale:
        local.node_ptr = torch.tensor(scale_lo + [cursor], dtype=torch.long)
        local.edge_ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(e_counts)]).astype(np.int64))
        local.intra_mesh_edge_index = torch.from_numpy(np.concatenate(i_parts, 1) if i_parts
                                                       else np.zeros((2, 0), dtype=np.int64)).contiguous()
        local.intra_edge_ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(i_counts)]).astype(np.int64))
    if hasattr(graph, "node_BC"):
        nbc = graph.node_BC.numpy().astype(np.int64)
        present = np.nonzero(g2l[nbc] >= 0)[0]
        local.node_BC = torch.from_numpy(g2l[nbc[present]])
        local.BC = graph.BC[torch.from_numpy(present)].contiguous()
        local.type_BC = graph.type_BC
    if hasattr(graph, "y") and graph.y is not None and graph.y.shape[0] == N:
        local.y = graph.y[torch.from_numpy(l2g)].contiguous()
    for k in ("previous_t", "temporal_res"):
        if hasattr(graph, k):
            setattr(local, k, getattr(graph, k))
    local.n_owned = list(n_owned)            # plan.build_plan: every edge set ends in owned rows only
    owned_rows = np.concatenate([scale_lo[s] + np.arange(n_owned[s]) for s in range(S)]) if S else np.zeros(0, np.int64)
    return LocalPartition(rank, world, S, local, l2g, n_owned, n_halo, scale_lo, send, recv, inter_cross,
                          owned_rows.astype(np.int64), l2g[owned_rows.astype(np.int64)])


# ------------------------------------------------------------------------------------------------
# halo exchange
# ------------------------------------------------------------------------------------------------
class HaloExchanger:
    """Neighbour exchange of halo rows for one rank.

    transport='nccl'   : ``dist.batch_isend_irecv`` on device tensors (NVLink / NVSwitch);
    transport='staged' : the same messages through pinned host buffers (gloo) — used by the tests
                         that run two ranks on one GPU or on CPU tensors."""

    def __init__(self, part: LocalPartition, device, transport: str = "nccl", group=None):
        import torch.distributed as dist
        self.dist = dist
        self.part, self.device, self.transport, self.group = part, torch.device(device), transport, group
        self.send_idx = [{q: torch.from_numpy(v.astype(np.int32)).to(self.device) for q, v in d.items()} for d in part.send]
        self.recv = part.recv
        self._bufs: Dict[Tuple[int, int, int], torch.Tensor] = {}
        self.n_exchanges = 0
        self.bytes_sent = 0

    def _send_buf(self, scale: int, q: int, n: int, width: int) -> torch.Tensor:
        key = (scale, q, width)
        b = self._bufs.get(key)
        if b is None:
            b = torch.empty(max(n, 1), width, dtype=torch.float32, device=self.device)
            self._bufs[key] = b
        return b

    def exchange(self, arr: torch.Tensor, scale: int):
        """Refresh the halo rows of scale ``scale`` in ``arr`` ([N_local, width] fp32, local row order)."""
        sends, recvs = self.send_idx[scale], self.recv[scale]
        if not sends and not recvs:
            return
        dist = self.dist
        width = arr.shape[1]
        ops, staged = [], []
        for q, idx in sends.items():
            n = int(idx.numel())
            buf = self._send_buf(scale, q, n, width)
            if arr.is_cuda:
                from . import lib
                lib.pack_rows(arr, idx, n, buf)
            else:
                torch.index_select(arr, 0, idx.long(), out=buf[:n])
            self.bytes_sent += n * width * 4
            if self.transport == "nccl":
                ops.append(dist.P2POp(dist.isend, buf[:n], q, group=self.group))
            else:
                ops.append(dist.P2POp(dist.isend, buf[:n].cpu(), q, group=self.group))
        for q, (row, n) in recvs.items():
            if self.transport == "nccl":
                ops.append(dist.P2POp(dist.irecv, arr[row:row + n], q, group=self.group))
            else:
                tmp = torch.empty(n, width, dtype=torch.float32)
                staged.append((row, n, tmp))
                ops.append(dist.P2POp(dist.irecv, tmp, q, group=self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for row, n, tmp in staged:
            arr[row:row + n].copy_(tmp)
        self.n_exchanges += 1


class _RawCuda:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2,
                                         "strides": None}


class PeerHalo:
    """Halo exchange over peer memory (NVLink / NVSwitch) — the production transport of the partitioned rollout.

    Every node array that takes part in an exchange (the model's per-forward arrays and the node inputs x) lives in ONE
    cudaMalloc'ed arena per rank, exported through CUDA IPC and mapped by every other rank.  ``exchange`` is a single
    kernel (``swe_halo_exchange``): it stores this rank's boundary rows straight into the halo rows of its neighbours'
    copies of the array, signals them through a sequence-numbered flag and waits for theirs.  No pack buffers, no NCCL, no
    host synchronisation: the step is a fixed kernel sequence and is captured in a CUDA graph.

    ``sync='host'`` (tests that run several ranks on ONE GPU, where kernels of different processes must not wait for one
    another): push kernel, device synchronise + process-group barrier, then the wait kernel (which finds its flags set).
    """
    owned_only = True           # local kernels write owned rows only; the halo rows belong to the neighbours' stores

    def __init__(self, part: LocalPartition, device, group=None, n_arrays: int = 8, width: int = 64, x_cols: int = 8,
                 sync: str = "device"):
        import torch.distributed as dist
        from . import lib
        self.lib, self.dist, self.group = lib, dist, group
        self.part, self.device, self.sync = part, torch.device(device), sync
        self.rank, self.world = part.rank, part.world
        n_local = int(part.local_to_global.shape[0])
        al = lambda b: (b + 255) // 256 * 256
        self.size = al(n_local * x_cols * 4) + n_arrays * al(n_local * width * 4) + 4096
        with torch.cuda.device(self.device):
            self.base, handle = lib.ipc_alloc(self.size)
        self.cursor = 0
        self.offsets: Dict[str, int] = {}
        self._names: Dict[int, str] = {}
        self._keep: List[torch.Tensor] = []
        self.flags = self._carve("_flags", (self.world + 8,), torch.int32)       # [q]: latest exchange of neighbour q; seq; done
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        self.peer_base = {}
        with torch.cuda.device(self.device):
            for q in range(self.world):
                if q != self.rank:
                    self.peer_base[q] = lib.ipc_open(handles[q])
        self.send_idx = [{q: torch.from_numpy(v.astype(np.int32)).to(self.device) for q, v in d.items()} for d in part.send]
        # symmetric neighbour sets per scale: every exchange signals and awaits the same peers on both sides
        self.peers = [sorted(set(part.send[s]) | set(part.recv[s])) for s in range(part.num_scales)]
        self._remote = None
        self.n_exchanges = 0
        self.bytes_sent = 0
        self.closed = False

    # ---- arena -------------------------------------------------------------------------------------------------
    def _carve(self, name, shape, dtype):
        nbytes = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        off = (self.cursor + 255) // 256 * 256
        if off + nbytes > self.size:
            raise RuntimeError(f"peer arena of {self.size} bytes is full (array '{name}')")
        self.cursor = off + nbytes
        typestr = {torch.float32: "<f4", torch.int32: "<i4"}[dtype]
        t = torch.as_tensor(_RawCuda(self.base + off, shape, typestr), device=self.device)
        self.offsets[name] = off
        self._names[t.data_ptr()] = name
        self._keep.append(t)
        return t

    def alloc(self, name: str, rows: int, cols: int) -> torch.Tensor:
        """A [rows, cols] fp32 array inside the arena (the model's workspace hook and the runner's x)."""
        return self._carve(name, (rows, cols), torch.float32)

    def finalize(self):
        """After every exchanged array has been allocated: learn where the neighbours put theirs."""
        meta = [None] * self.world
        mine = (dict(self.offsets), [{int(q): (int(r0), int(n)) for q, (r0, n) in d.items()} for d in self.part.recv])
        self.dist.all_gather_object(meta, mine, group=self.group)
        self._remote = meta

    # ---- exchange ----------------------------------------------------------------------------------------------
    def exchange(self, arr: torch.Tensor, scale: int):
        peers = self.peers[scale]
        if not peers:
            return
        name = self._names.get(arr.data_ptr())
        if name is None:
            raise RuntimeError("PeerHalo.exchange: the array does not live in the peer arena")
        if self._remote is None:
            self.finalize()
        width = int(arr.shape[1])
        fl_off = self.offsets["_flags"]
        idx_ptrs, n_send, rrows, rflags, lflags = [], [], [], [], []
        for q in peers:
            idx = self.send_idx[scale].get(q)
            n = 0 if idx is None else int(idx.numel())
            offs_q, recv_q = self._remote[q]
            row0 = recv_q[scale].get(self.rank, (0, 0))[0]
            if n and recv_q[scale].get(self.rank, (0, 0))[1] != n:
                raise RuntimeError("halo maps of the ranks disagree")
            idx_ptrs.append(idx.data_ptr() if n else 0)
            n_send.append(n)
            rrows.append(self.peer_base[q] + offs_q[name] + row0 * width * 4 if n else 0)
            rflags.append(self.peer_base[q] + offs_q["_flags"] + 4 * self.rank)
            lflags.append(self.base + fl_off + 4 * q)
            self.bytes_sent += n * width * 4
        seq_ptr = self.base + fl_off + 4 * self.world
        done_ptr = seq_ptr + 4
        if self.sync == "device":
            self.lib.halo_exchange(arr, idx_ptrs, n_send, rrows, rflags, lflags, seq_ptr, done_ptr, True, True)
        else:
            self.lib.halo_exchange(arr, idx_ptrs, n_send, rrows, rflags, lflags, seq_ptr, done_ptr, True, False)
            torch.cuda.synchronize(self.device)
            self.dist.barrier(group=self.group)
            self.lib.halo_exchange(arr, idx_ptrs, n_send, rrows, rflags, lflags, seq_ptr, done_ptr, False, True)
        self.n_exchanges += 1

    def close(self):
        """Unmap the neighbours' arenas and free this rank's (collective: nobody may still be storing into it)."""
        if self.closed:
            return
        self.closed = True
        torch.cuda.synchronize(self.device)
        self.dist.barrier(group=self.group)
        with torch.cuda.device(self.device):
            for p_ in self.peer_base.values():
                self.lib.ipc_close(p_)
            self.dist.barrier(group=self.group)
            self._keep.clear()
            self.lib.ipc_free(self.base)


class PartitionedRollout:
    """Autoregressive rollout of one large mesh cut over the ranks of ``group``.

    Same loop as ``RolloutRunner`` (``training/train.py:67-95``) run eagerly on the local graph,
    with the halo refreshes inserted by the model's launch sequence.  ``preds`` holds the local rows
    (owned + halo); ``owned_predictions()`` returns the owned rows with their global ids."""

    def __init__(self, model, graph_cpu, n_steps: int, device, transport: str = "nccl", group=None,
                 part: Optional[LocalPartition] = None):
        import torch.distributed as dist
        from . import lib
        from .utils.dataset import NUM_WATER_VARS, check_type_BC
        self.lib = lib
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.part = part if part is not None else partition_graph(graph_cpu, world, rank)
        self.model, self.T = model, int(n_steps)
        self.graph = self.part.graph.to(device, non_blocking=True)
        self.peer = transport in ("peer", "peer-hostsync")
        if self.peer:
            self.halo = PeerHalo(self.part, device, group, n_arrays=len(model._WS_NAMES) + 1, width=model._FP,
                                 x_cols=int(self.graph.x.shape[1]), sync="device" if transport == "peer" else "host")
        else:
            self.halo = HaloExchanger(self.part, device, transport, group)
        model._check_input(self.graph)
        multiscale = model.type_model == "MSGNN"
        self.plan = model._plans.get(self.graph, getattr(model, "num_scales", 1), multiscale)
        N = self.plan.n_nodes
        if self.peer:
            # node arrays of the forward and the node inputs live in the IPC arena; the neighbours learn their offsets
            model._ws_alloc = self.halo.alloc
            try:
                model._ws.pop((self.plan.key, self.plan.n_nodes), None)
                model._workspace(self.plan, model._WS_NAMES)
            finally:
                model._ws_alloc = None
            self.x = self.halo.alloc("x", N, int(self.graph.x.shape[1]))
            self.x.copy_(self.graph.x)
            self.halo.finalize()
        else:
            self.x = self.graph.x.detach().clone().contiguous()
        self.preds = torch.empty(self.T, N, NUM_WATER_VARS, dtype=torch.float32, device=device)
        self.step = torch.zeros(1, dtype=torch.int32, device=device)
        self.type_BC = int(self.graph.type_BC)
        check_type_BC(self.type_BC)
        self.node_BC = self.graph.node_BC.to(device, torch.int64).contiguous()
        self.bc = self.graph.BC.to(device, torch.float32).contiguous()
        self.n_static_raw = self.x.shape[1] - model.previous_t * NUM_WATER_VARS
        self.launches_per_step = 0
        self.done = 0                             # host mirror of the device step counter (bounds check in run())
        self._token = new_static_token()
        # the peer transport's step is a fixed kernel sequence without host synchronisation: captured and replayed
        self.use_cuda_graph = transport == "peer" and self.T > 2
        self._graph = None

    def close(self):
        """Release the peer arena (collective).  The model's cached workspace of this plan pointed into it."""
        if not self.peer:
            return
        if not self.halo.closed:
            self._graph = None
            self.model._ws.pop((self.plan.key, self.plan.n_nodes), None)
            self.model._plans.clear()
            self.halo.close()

    def _one_step(self):
        lib, m = self.lib, self.model
        c0 = lib.launch_count
        x0, b0 = self.halo.n_exchanges, self.halo.bytes_sent
        if self.node_BC.numel():
            lib.apply_bc(self.x, self.n_static_raw, m.previous_t, self.type_BC, self.node_BC, self.bc, self.step)
        with static_inputs(self._token, xs_static=not m.with_WL):         # static columns, mesh part and weights are constant over the rollout
            m._launch(self.plan, self.graph, self.x, self.preds, step_ptr=self.step,
                      pred_stride=self.preds.shape[1] * 2, x_next=self.x, halo=self.halo)
        lib.step_advance(self.step)
        # the window shift wrote garbage into the halo rows of x: refresh them from their owners
        for s in range(self.part.num_scales):
            self.halo.exchange(self.x, s)
        self.launches_per_step = lib.launch_count - c0
        self.exchanges_per_step = self.halo.n_exchanges - x0
        self.halo_bytes_per_step = self.halo.bytes_sent - b0

    def reset(self):
        """Collective for the peer transport: nobody may still be storing x rows of the previous rollout."""
        if self.peer:
            torch.cuda.synchronize()
            self.halo.dist.barrier(group=self.halo.group)
        self.x.copy_(self.graph.x)
        self.step.zero_()
        self.done = 0
        if self.peer:
            torch.cuda.synchronize()
            self.halo.dist.barrier(group=self.halo.group)

    def rebind(self, graph_cpu):
        """Another simulation on the SAME partitioned mesh (new node inputs, boundary series, edge attributes — e.g. the
        next sample of an ensemble, given as this rank's local graph on the host): the values are copied into the buffers
        the captured step reads; plan, peer arena and captured graph are reused.  Collective."""
        if self.peer:
            torch.cuda.synchronize()
            self.halo.dist.barrier(group=self.halo.group)
        self.graph.x.copy_(graph_cpu.x, non_blocking=True)
        self.graph.edge_attr.copy_(graph_cpu.edge_attr, non_blocking=True)      # (in place: bumps the version the encoder cache checks)
        self.bc.copy_(graph_cpu.BC, non_blocking=True)
        self.x.copy_(self.graph.x)
        self.step.zero_()
        self.done = 0
        ws = self.model._ws.get((self.plan.key, self.plan.n_nodes))
        if ws is not None and ws.get("_plan") is self.plan:
            self.model._encoded_edges(self.plan, self.graph, ws)
        else:
            self._graph = None
        if self.peer:
            torch.cuda.synchronize()
            self.halo.dist.barrier(group=self.halo.group)

    def _stream_out(self, out_host, step: int):
        """Device -> host copy of the OWNED rows of prediction slot `step` (one contiguous range per scale) on a side
        stream, ordered after the kernels issued so far: it overlaps the next step (training/train.py:RolloutRunner)."""
        main = torch.cuda.current_stream()
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.preds.device)
        ev = torch.cuda.Event()
        ev.record(main)
        self._copy_stream.wait_event(ev)
        with torch.cuda.stream(self._copy_stream):
            off = 0
            for s_ in range(self.part.num_scales):
                lo, cnt = int(self.part.scale_lo[s_]), int(self.part.n_owned[s_])
                out_host[step, off:off + cnt].copy_(self.preds[step, lo:lo + cnt], non_blocking=True)
                off += cnt

    def run(self, n_steps: Optional[int] = None, out_host: Optional[torch.Tensor] = None):
        """out_host: optional pinned host tensor ``[T, n_owned_rows, 2]`` receiving the owned rows (the order of
        ``part.owned_rows``) of every step while the next one runs."""
        n = self.T - self.done if n_steps is None else int(n_steps)
        if n < 0 or self.done + n > self.T:
            raise ValueError(f"rollout of {self.T} steps: {self.done} done, {n} more requested (call reset() first)")
        if out_host is not None:
            want = (self.T, int(len(self.part.owned_rows)), int(self.preds.shape[-1]))
            if not (out_host.is_pinned() and out_host.dtype == torch.float32 and out_host.is_contiguous()
                    and tuple(out_host.shape) == want):
                raise ValueError(f"out_host must be a pinned contiguous float32 tensor of shape {want}")
        done = 0
        if self.use_cuda_graph and self._graph is None and n > 1:
            self._one_step()                     # eager: lazy packing / allocation happen here
            done = 1
            if out_host is not None:
                self._stream_out(out_host, self.done)
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                self._one_step()
            self._graph = g
        for j in range(done, n):
            if self._graph is not None:
                self._graph.replay()
            else:
                self._one_step()
            if out_host is not None:
                self._stream_out(out_host, self.done + j)
        if out_host is not None:
            torch.cuda.current_stream().wait_stream(self._copy_stream)
        self.done += n
        return self.preds

    def owned_predictions(self):
        rows = torch.from_numpy(self.part.owned_rows).to(self.preds.device)
        return self.preds[:, rows], self.part.owned_global


# ------------------------------------------------------------------------------------------------
# data-parallel training
# ------------------------------------------------------------------------------------------------
def allreduce_gradients(params: Sequence[torch.Tensor], group=None, average: bool = True):
    """One all-reduce of the flat fp32 gradient (811,309 floats = 3.25 MB for the default model)."""
    import torch.distributed as dist
    # the list must not depend on the data: a rank whose batch never touched a parameter (grad is None) still takes
    # part with zeros, otherwise the flat buffers differ in length between ranks and the collective hangs or mixes
    ps = [p for p in params if p.requires_grad]
    if not ps:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in ps:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel() * 4


def shard_simulations(n_sims: int, world: int, rank: int) -> List[int]:
    """Indices of the simulations rank ``rank`` trains on (round-robin, every index exactly once)."""
    return list(range(rank, n_sims, world))
