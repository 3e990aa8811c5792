import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_backward as T
from helpers import REF_CONFIG_MODELS, rel_l2
from mswe_gnn_b200.utils.synthetic import make_tri_mesh

def run(parts):
    os.environ["MSWE_TRAIN_TC_PARTS"] = parts
    errs = {}
    orig = T._check_grads
    def chk(ours, ref64, ref32, floor=2e-4, mult=20.0):
        for k, g64 in ref64.items():
            if g64 is None or float(g64.norm()) == 0: continue
            errs[k] = (rel_l2(ours[k].cpu(), g64), rel_l2(ref32[k], g64))
        return ("", 0)
    T._check_grads = chk
    ctor = dict(num_node_features=8, num_edge_features=1, num_scales=4, previous_t=3, **REF_CONFIG_MODELS)
    data = make_tri_mesh(16, 16, 4, rollout_steps=1, seed=5)
    T._train_compare("MSGNN", ctor, data, 1)
    T._check_grads = orig
    return errs

res = {p: run(p) for p in ["none", "fwd", "dx,dw", "fwd,dx,dw"]}
keys = sorted(res["none"], key=lambda k: -res["fwd,dx,dw"][k][0])[:14]
print("%-46s" % "key", *["%10s" % p for p in res], "      yard")
for k in keys:
    print("%-46s" % k[-46:], *["%10.2e" % res[p][k][0] for p in res], "%10.2e" % res["none"][k][1])
