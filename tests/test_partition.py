"""Host-side multi-GPU logic on CPU: partition artefacts bit-exact against the loop oracle, halo exchange and
gradient all-reduce over gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mswe_gnn_b200  # noqa: F401
from mswe_gnn_b200.parallel import HaloExchanger, allreduce_gradients, owner_map, partition_graph, shard_simulations
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh, make_tri_mesh
from oracle import partition_oracle as PO


def _fields(g):
    return (g.node_ptr.numpy(), g.edge_index.numpy(), g.edge_ptr.numpy(), g.intra_mesh_edge_index.numpy(),
            g.intra_edge_ptr.numpy())


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("mesh_kw", [dict(), dict(orphan_every=5, extra_parent_every=3), dict(link_ghosts=True)])
def test_partition_bit_exact_vs_loop_oracle(world, mesh_kw):
    g = make_tri_mesh(8, 8, 3, seed=1, **mesh_kw)
    node_ptr, ei, edge_ptr, intra, intra_ptr = _fields(g)
    own = owner_map(node_ptr, ei, edge_ptr, intra, intra_ptr, world)
    assert own.tolist() == PO.owner_map_loops(node_ptr, ei, intra, intra_ptr, world)
    covered = np.zeros(g.x.shape[0], dtype=int)
    parts = [partition_graph(g, world, r) for r in range(world)]
    for r, p in enumerate(parts):
        l2g, halos, sends, edges = PO.local_sets_loops(node_ptr, ei, edge_ptr, intra, intra_ptr, own.tolist(), r)
        assert p.local_to_global.tolist() == l2g
        assert p.n_halo == [len(h) for h in halos]
        covered[p.owned_global] += 1
        for s in range(p.num_scales):
            assert {q: p.local_to_global[v].tolist() for q, v in p.send[s].items()} == sends[s]
            # local edges: the global edges that end in an owned node, in global order
            lo, hi = int(p.graph.edge_ptr[s]), int(p.graph.edge_ptr[s + 1])
            le = p.graph.edge_index[:, lo:hi].numpy()
            assert [(int(p.local_to_global[a]), int(p.local_to_global[b])) for a, b in le.T] == edges[s]
            # receive ranges are contiguous, ordered by peer, and sit after the owned rows
            for q, (row, n) in p.recv[s].items():
                assert row >= p.scale_lo[s] + p.n_owned[s]
                ids = p.local_to_global[row:row + n]
                assert (own[ids] == q).all() and (np.diff(ids) > 0).all()
    assert (covered == 1).all()                       # every node owned by exactly one rank
    # what r receives from q is exactly what q sends to r, in the same order
    for r, p in enumerate(parts):
        for s in range(p.num_scales):
            for q, (row, n) in p.recv[s].items():
                sent = parts[q].local_to_global[parts[q].send[s][r]]
                assert sent.tolist() == p.local_to_global[row:row + n].tolist()


def test_partition_single_scale_and_world_one():
    g = make_single_scale_mesh(10, 6, seed=2)
    p = partition_graph(g, 1, 0)
    assert p.n_halo == [0] and p.local_to_global.tolist() == list(range(g.x.shape[0]))
    assert torch.equal(p.graph.edge_index, g.edge_index)
    parts = [partition_graph(g, 2, r) for r in range(2)]
    assert sum(q.n_owned[0] for q in parts) == g.x.shape[0]
    assert all(q.n_halo[0] > 0 for q in parts)


def test_shard_simulations_covers_every_index_once():
    for n, w in [(10, 4), (3, 8), (16, 2)]:
        got = sorted(i for r in range(w) for i in shard_simulations(n, w, r))
        assert got == list(range(n))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = make_tri_mesh(8, 8, 3, seed=3, extra_parent_every=4)
        p = partition_graph(g, world, rank)
        ex = HaloExchanger(p, "cpu", transport="staged")
        n_local = p.local_to_global.size
        ok = True
        for width in (8, 16):
            arr = torch.full((n_local, width), -1.0)
            owned = torch.from_numpy(p.owned_rows)
            arr[owned] = torch.from_numpy(p.owned_global).float()[:, None] + torch.arange(width).float()[None, :] / 100
            for s in range(p.num_scales):
                ex.exchange(arr, s)
            want = torch.from_numpy(p.local_to_global).float()[:, None] + torch.arange(width).float()[None, :] / 100
            ok &= bool(torch.equal(arr, want))
        # data-parallel gradient all-reduce: mean over ranks of rank-dependent gradients
        ps = [torch.nn.Parameter(torch.zeros(3, 2)), torch.nn.Parameter(torch.zeros(5))]
        for i, q in enumerate(ps):
            q.grad = torch.full_like(q, float(rank + 1 + i))
        nbytes = allreduce_gradients(ps)
        mean0 = sum(r + 1 for r in range(world)) / world
        ok &= bool(torch.allclose(ps[0].grad, torch.full((3, 2), mean0))) and bool(torch.allclose(ps[1].grad, torch.full((5,), mean0 + 1)))
        ok &= nbytes == 11 * 4
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_halo_exchange_and_gradient_allreduce_gloo_world2():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
