import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import mswe_gnn_b200.autograd as A
from mswe_gnn_b200.models.gnn import SWEGNN
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh
DEV = "cuda"
torch.manual_seed(3)
F = 64
d = make_single_scale_mesh(40, 30, seed=2)
n, e = d.x.shape[0], d.edge_index.shape[1]
xs = torch.randn(n, F); xd = torch.randn(n, F); xd[torch.rand(n) < 0.4] = 0.0
ea = torch.randn(e, F)
op = SWEGNN(F, F, n_layers=3, activation="prelu", bias=True, edge_features=F, K=2, normalize=True, with_filter_matrix=True,
            with_gradient=True).to(DEV)
calls = []
orig = A.swegnn_forward_train
def wrap(*a, **k):
    c = orig(*a, **k); calls.append(c); return c
A.swegnn_forward_train = wrap
outs = {}
for parts in ["none", "fwd"]:
    os.environ["MSWE_TRAIN_TC_PARTS"] = parts
    xs_g, xd_g = xs.to(DEV).requires_grad_(True), xd.to(DEV).requires_grad_(True)
    ea_g = ea.to(DEV).requires_grad_(True)
    out = op(xs_g, xd_g, d.edge_index.to(DEV), ea_g)
    outs[parts] = calls[-1]
a, b = outs["none"], outs["fwd"]
print("E", e, "N", n)
for i in range(3):
    pa, pb = a.pres[i], b.pres[i]
    diff = (pa - pb).abs()
    print("pre", i, tuple(pa.shape), "rel", (diff.norm() / pa.norm()).item(), "max abs", diff.max().item(),
          "rows with big err", int((diff.max(1).values > 1e-3).sum()))
    bad = (diff.max(1).values > 1e-3).nonzero().flatten()[:10].tolist()
    print("   bad rows", bad)
diff = (a.s - b.s).abs()
print("s rel", (diff.norm() / a.s.norm()).item(), "max", diff.max().item())
