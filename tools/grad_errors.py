"""Per-parameter gradient error of the training step against the fp64 oracle at a given mesh size, for the tensor-core
and the exact-fp32 training paths (MSWE_TRAIN_GEMM=tc|ffma).  Usage: python tools/grad_errors.py [nx ny] (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch

import mswe_gnn_b200  # noqa: F401
from helpers import REF_CONFIG_MODELS, rel_l2
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh
import test_gpu_backward as TB

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (160, 160)
gc = {k: v for k, v in REF_CONFIG_MODELS.items() if k not in ("learned_pooling", "skip_connections")}
ctor = dict(num_node_features=8, num_edge_features=1, previous_t=3, n_GNN_layers=2, **gc)
data = make_single_scale_mesh(nx, ny, rollout_steps=1, seed=5)
errs = {}
orig = TB._check_grads


def collect(ours, ref64, ref32, floor=2e-4, mult=20.0):
    out = []
    for k, g64 in ref64.items():
        if g64 is None or float(g64.norm()) == 0.0:
            continue
        out.append((rel_l2(ours[k].cpu(), g64), rel_l2(ref32[k], g64), k, tuple(g64.shape)))
    out.sort(reverse=True)
    for e, y, k, shp in out[:12]:
        print(f"   {e:.3e}  (fp32 oracle {y:.3e})  {k} {shp}")
    return out[0]


TB._check_grads = collect
for mode in ("tc", "ffma"):
    os.environ["MSWE_TRAIN_GEMM"] = mode
    print(f"== MSWE_TRAIN_GEMM={mode}, tri({nx},{ny})")
    TB._train_compare("GNN", ctor, data.clone(), 1)
