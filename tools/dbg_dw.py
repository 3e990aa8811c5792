import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mswe_gnn_b200  # noqa
from mswe_gnn_b200 import lib
DEV = "cuda:0"
L = lib.load()

def run(delta, X, n=128):
    R = delta.shape[0]
    rows = lib.make_rows([(X, None, X.shape[1], X.shape[1], 0, None)])
    grid = lib.mlp_layer_bwd_dw_tc_grid(R)
    part = torch.zeros(grid * n * X.shape[1], device=DEV)
    lib.mlp_layer_bwd_dw_tc(delta, R, n, rows, part)
    torch.cuda.synchronize()
    return part.view(grid, n, X.shape[1]).sum(0)

def onehot(R, W, r, c):
    t = torch.zeros(R, W, device=DEV); t[r, c] = 1.0; return t

print("same row experiments (expect D[m0,q1]=1)")
for (r, m0, q1) in [(0, 0, 0), (0, 5, 9), (0, 37, 70), (1, 5, 9), (8, 5, 9), (9, 37, 70), (31, 100, 127)]:
    D = run(onehot(32, 128, r, m0), onehot(32, 128, r, q1))
    nz = D.nonzero().tolist()
    print((r, m0, q1), "->", [(a, b, D[a, b].item()) for a, b in nz][:6])
print("row mismatch experiments: delta row 0, X row r1")
for r1 in range(0, 32):
    D = run(onehot(32, 128, 0, 3), onehot(32, 128, r1, 4))
    nz = D.nonzero().tolist()
    if nz: print("r1", r1, nz[:4])
print("delta row r0, X row 0")
for r0 in range(0, 32):
    D = run(onehot(32, 128, r0, 3), onehot(32, 128, 0, 4))
    nz = D.nonzero().tolist()
    if nz: print("r0", r0, nz[:4])
# full-row tests
d = torch.randn(32, 128, device=DEV); x = torch.randn(32, 128, device=DEV)
D = run(d, x); ref = d.double().t() @ x.double()
print("random 32 rows rel err", ((D.double() - ref).norm() / ref.norm()).item())

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
E = 1_228_800
delta = torch.randn(E, 128, device=DEV); pre = torch.randn(E, 128, device=DEV)
rows1 = lib.make_rows([(pre, None, 128, 128, 0, None)])
g = lib.mlp_layer_bwd_dw_tc_grid(E)
part = torch.empty(g * 128 * 256, device=DEV)
for dbg in (0, 1, 2, 3):
    L.swe_train_tc_set_debug(dbg)
    t = timeit(lambda: lib.mlp_layer_bwd_dw_tc(delta, E, 128, rows1, part))
    print("debug", dbg, f"{t:.3f} ms")
L.swe_train_tc_set_debug(0)
