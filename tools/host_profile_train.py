"""Host-side (Python) profile of the training step: where the CPU time of one step goes (GPU work is asynchronous;
the profile is taken without synchronising inside the step, so it shows the launch path only)."""
import sys, os, cProfile, pstats, time, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import mswe_gnn_b200  # noqa
from mswe_gnn_b200 import lib
from mswe_gnn_b200.training.train import training_step

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg5-train"
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
lib.load()
kind, ctor, model, batch_host, R, _ = bench._train_setup(wl, 0, dev)
batch = batch_host.to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=3e-3, weight_decay=0.0, fused=True)

def step():
    opt.zero_grad(set_to_none=True)
    loss = training_step(model, batch, R, only_where_water=True, velocity_scaler=7.0)
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    return loss

for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t1 = time.perf_counter()          # host time to ISSUE 5 steps (may include back-pressure from the launch queue)
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue time per step {1e3 * (t1 - t0) / 5:.1f} ms; with drain {1e3 * (t2 - t0) / 5:.1f} ms; launches/step {lib.launch_count}")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:6000])
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:5000])

s = io.StringIO()
pstats.Stats(pr, stream=s).print_callers("method 'to' of")
print(s.getvalue()[:4000])
