"""Partitioned large-mesh rollout: P ranks (processes) share the test GPU and exchange halos through the staged
(gloo) transport; the owned rows of every rank must equal the single-GPU rollout BIT FOR BIT (same edges in the
same order per destination).  The NCCL transport is exercised by `bench.py --gpus N --workload cfg4`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

CTOR = dict(num_node_features=8, num_edge_features=1, num_scales=3, previous_t=3, hid_features=64, mlp_layers=3,
            seed=21, learned_residuals=True, mlp_activation="prelu", gnn_activation="tanh", with_WL=True, K=3)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, steps, mesh_kw, out, transport="staged"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mswe_gnn_b200  # noqa: F401
        from mswe_gnn_b200.models.gnn import MSGNN
        from mswe_gnn_b200.parallel import PartitionedRollout
        from mswe_gnn_b200.training.train import rollout_test
        from mswe_gnn_b200.utils.synthetic import make_tri_mesh
        dev = torch.device("cuda", 0)
        torch.cuda.set_device(dev)
        model = MSGNN(**CTOR).to(dev)
        g = make_tri_mesh(24, 16, 3, rollout_steps=steps, seed=5, **mesh_kw)
        pr = PartitionedRollout(model, g, steps, dev, transport=transport)
        # the owned rows of every step also stream to a pinned host buffer while the next step runs (run(out_host=...))
        host = torch.full((steps, len(pr.part.owned_rows), 2), float("nan")).pin_memory()
        pr.run(out_host=host)
        torch.cuda.synchronize()
        preds, gids = pr.owned_predictions()
        assert torch.equal(host, preds.cpu())
        out[rank] = (preds.cpu().numpy(), gids, pr.halo.n_exchanges)
        pr.close()
        if rank == 0:
            with torch.no_grad():
                ref = rollout_test(model, g.to(dev), use_cuda_graph=False)      # [N, 2, T]
            out["ref"] = ref.permute(2, 0, 1).cpu().numpy()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["staged", "peer-hostsync"])
@pytest.mark.parametrize("world,mesh_kw", [(2, dict()), (3, dict()), (2, dict(extra_parent_every=4, orphan_every=7))])
def test_partitioned_rollout_equals_single_gpu_bit_exact(world, mesh_kw, transport):
    """'staged': the halo rows travel through pinned host buffers (gloo); 'peer-hostsync': the production transport —
    boundary rows stored straight into the neighbours' IPC-mapped arrays by swe_halo_exchange — with the cross-rank
    wait done on the host, because kernels of several processes sharing ONE GPU must not wait for one another (the
    device-side flag wait is exercised on real multi-GPU boxes by bench.py's parity check)."""
    steps = 3
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), steps, mesh_kw, out, transport), nprocs=world, join=True)
        ref = out["ref"]
        seen = np.zeros(ref.shape[1], dtype=int)
        for r in range(world):
            preds, gids, n_ex = out[r]
            seen[gids] += 1
            assert n_ex > 0
            assert np.array_equal(preds, ref[:, gids]), f"rank {r}: max diff {np.abs(preds - ref[:, gids]).max()}"
        assert (seen == 1).all()
