"""Time the tensor-core backward GEMMs against the CUDA-core ones on edge-MLP shapes (E rows)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mswe_gnn_b200  # noqa: F401
from mswe_gnn_b200 import lib

DEV = "cuda:0"
E = int(sys.argv[1]) if len(sys.argv) > 1 else 1_228_800


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


n_nodes = E // 3
delta = torch.randn(E, 128, device=DEV)
delta64 = torch.randn(E, 64, device=DEV)
pre = torch.randn(E, 128, device=DEV)
xs = torch.randn(n_nodes, 64, device=DEV)
src = torch.randint(0, n_nodes, (E,), device=DEV, dtype=torch.int32)
dst = torch.randint(0, n_nodes, (E,), device=DEV, dtype=torch.int32)
w = torch.randn(128, 320, device=DEV) * 0.1
w2 = torch.randn(128, 128, device=DEV) * 0.1
w3 = torch.randn(64, 128, device=DEV) * 0.1
dx = torch.empty(E, 128, device=DEV)
dxa, dxb = torch.empty(E, 64, device=DEV), torch.empty(E, 64, device=DEV)

# dx
t = timeit(lambda: lib.mlp_layer_bwd_dx_tc(delta, E, 128, w2, 128, 0, 128, 128, dx, False))
print(f"dx_tc  n=128 ko=128        {t:.3f} ms  {2*E*128*128/t/1e9:.1f} TFLOP/s")
t = timeit(lambda: lib.mlp_layer_bwd_dx_tc(delta, E, 128, w, 320, 0, 64, 64, dxa, False))
print(f"dx_tc  n=128 ko=64         {t:.3f} ms  {2*E*128*64/t/1e9:.1f} TFLOP/s")
t = timeit(lambda: lib.mlp_layer_bwd_dx_tc(delta, E, 128, w, 320, 0, 128, 128, dxa, False, dxb, False, 64))
print(f"dx_tc  n=128 ko=2x64       {t:.3f} ms  {2*E*128*128/t/1e9:.1f} TFLOP/s")
t = timeit(lambda: lib.mlp_layer_bwd_dx_tc(delta64, E, 64, w3, 128, 0, 128, 128, dx, False))
print(f"dx_tc  n=64  ko=128        {t:.3f} ms  {2*E*64*128/t/1e9:.1f} TFLOP/s")
grid = lib.mlp_layer_bwd_dx_grid(E)
t = timeit(lambda: lib.mlp_layer_bwd_dx(delta, None, 0, None, E, 128, w2, 128, 0, 128, 128, dx, False, False, None))
print(f"dx_ffma n=128 ko=128       {t:.3f} ms  {2*E*128*128/t/1e9:.1f} TFLOP/s")

# dw
g = lib.mlp_layer_bwd_dw_tc_grid(E)
part = torch.empty(g * 128 * 256, device=DEV)
rows4 = lib.make_rows([(xs, src, 64, 64, 0, None), (xs, dst, 64, 64, 0, None), (xs, src, 64, 64, 0, None), (xs, dst, 64, 64, 0, None)])
t = timeit(lambda: lib.mlp_layer_bwd_dw_tc(delta, E, 128, rows4, part))
print(f"dw_tc  n=128 X=4x64 gather {t:.3f} ms  {2*E*128*256/t/1e9:.1f} TFLOP/s")
rows1 = lib.make_rows([(pre, None, 128, 128, 1, None)])
t = timeit(lambda: lib.mlp_layer_bwd_dw_tc(delta, E, 128, rows1, part))
print(f"dw_tc  n=128 X=act(pre)128 {t:.3f} ms  {2*E*128*128/t/1e9:.1f} TFLOP/s")
t = timeit(lambda: lib.mlp_layer_bwd_dw_tc(delta64, E, 64, rows1, part))
print(f"dw_tc  n=64  X=act(pre)128 {t:.3f} ms  {2*E*64*128/t/1e9:.1f} TFLOP/s")
rows64 = lib.make_rows([(xs, src, 64, 64, 0, None)])
t = timeit(lambda: lib.mlp_layer_bwd_dw_tc(delta, E, 128, rows64, part))
print(f"dw_tc  n=128 X=64 gather   {t:.3f} ms  {2*E*128*64/t/1e9:.1f} TFLOP/s")
gf = lib.mlp_layer_bwd_dw_grid(E)
partf = torch.empty(gf * 128 * 128, device=DEV)
t = timeit(lambda: lib.mlp_layer_bwd_dw(delta, E, 128, rows1, 128, partf))
print(f"dw_ffma n=128 X=act(pre)128 {t:.3f} ms  {2*E*128*128/t/1e9:.1f} TFLOP/s")
rows64p = lib.make_rows([(pre[:, :64].contiguous(), None, 64, 64, 1, None)])
t = timeit(lambda: lib.mlp_layer_bwd_dw_tc(delta64, E, 64, rows64p, part))
print(f"dw_tc  n=64  X=act(pre)64  {t:.3f} ms  {2*E*64*64/t/1e9:.1f} TFLOP/s")
partf64 = torch.empty(gf * 64 * 64, device=DEV)
t = timeit(lambda: lib.mlp_layer_bwd_dw(delta64, E, 64, rows64p, 64, partf64))
print(f"dw_ffma n=64 X=act(pre)64  {t:.3f} ms  {2*E*64*64/t/1e9:.1f} TFLOP/s")
w64 = torch.randn(64, 64, device=DEV) * 0.1
dx64 = torch.empty(E, 64, device=DEV)
t = timeit(lambda: lib.mlp_layer_bwd_dx_tc(delta64, E, 64, w64, 64, 0, 64, 64, dx64, False))
print(f"dx_tc  n=64  ko=64         {t:.3f} ms  {2*E*64*64/t/1e9:.1f} TFLOP/s")
t = timeit(lambda: lib.mlp_layer_bwd_dx(delta64, None, 0, None, E, 64, w64, 64, 0, 64, 64, dx64, False, False, None))
print(f"dx_ffma n=64 ko=64         {t:.3f} ms  {2*E*64*64/t/1e9:.1f} TFLOP/s")
