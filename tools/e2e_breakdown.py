"""Where the end-to-end rollout_test(model, host_graph) time goes (cfg3, K steps)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mswe_gnn_b200  # noqa
from mswe_gnn_b200.models.gnn import MSGNN
from mswe_gnn_b200.training.train import RolloutRunner
from mswe_gnn_b200.utils.synthetic import make_tri_mesh
K = int(os.environ.get("K", 48))
dev = torch.device("cuda", 0)
ctor = dict(num_node_features=8, num_edge_features=1, num_scales=4, previous_t=3, hid_features=64, mlp_layers=3, seed=666,
            learned_residuals=True, mlp_activation="prelu", gnn_activation="tanh", with_WL=True, K=4)
model = MSGNN(**ctor).to(dev)
host = make_tri_mesh(712, 712, 4, rollout_steps=K, with_y=False)
for k in host.keys():
    v = getattr(host, k)
    if torch.is_tensor(v): setattr(host, k, v.pin_memory())
out_host = torch.empty(K, host.x.shape[0], 2).pin_memory()
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = sync()
    g = host.to(dev, non_blocking=True); t1 = sync()
    temp = g.clone(); t2 = sync()
    runner = RolloutRunner(model, temp, K, None); t3 = sync()
    runner._one_step(); t4 = sync()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        runner._one_step()
    t5 = sync()
    for _ in range(K - 1): gr.replay()
    t6 = sync()
    out_host.copy_(runner.preds, non_blocking=True); t7 = sync()
    print(f"rep {rep}: h2d {1e3*(t1-t0):.1f}  clone {1e3*(t2-t1):.1f}  runner+plan {1e3*(t3-t2):.1f}  eager step {1e3*(t4-t3):.1f}  "
          f"capture {1e3*(t5-t4):.1f}  {K-1} replays {1e3*(t6-t5):.1f} ({1e3*(t6-t5)/(K-1):.2f}/step)  d2h {1e3*(t7-t6):.1f}  total {1e3*(t7-t0):.1f} ms")
    del runner, gr, temp, g
