"""Times the level-0 hop of cfg3 (1,013,889 nodes / 3,038,817 edges) with both backends."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mswe_gnn_b200  # noqa
from mswe_gnn_b200 import lib
from mswe_gnn_b200.utils.synthetic import make_single_scale_mesh
DEV = "cuda"
nx = int(os.environ.get("NX", 712))
d = make_single_scale_mesh(nx, nx, seed=1)
n, e = d.x.shape[0], d.edge_index.shape[1]
ei = d.edge_index.to(DEV)
rowptr, src, dst, eid = lib.csr_build(ei[0].contiguous(), ei[1].contiguous(), None, 0, n, 0, n)
o = torch.randn(n, 64, device=DEV); s = torch.randn(e, 64, device=DEV); W = torch.randn(64, 64, device=DEV) / 8
wt = torch.empty(64, 64, device=DEV); lib.pack_linear(W, 64, wt)
img = torch.empty(lib.hop_tc_image_bytes(), dtype=torch.uint8, device=DEV); lib.hop_tc_pack(W, img)
img16 = torch.empty(lib.hop_tc16_image_bytes(), dtype=torch.uint8, device=DEV); lib.hop_tc16_pack(W, float(W.abs().max()), img16)
out = torch.empty_like(o)
bytes_alg = 4 * 64 * (e + 2 * n) + 4 * (e + n + 1)
def run(which):
    if which == "nofilter":
        lib.propagate_hop_fwd(o, o, s, rowptr, src, 0, n, None, 1, 0, None, 0, None, out, 64)
    elif which == "ffma":
        lib.propagate_hop_fwd(o, o, s, rowptr, src, 0, n, wt, 1, 0, None, 0, None, out, 64)
    elif which == "tc16s":
        lib.propagate_hop_tc16s_fwd(o, o, s, rowptr, src, 0, n, img16, 1, 0, None, 0, None, None, out)
    elif which == "tc16":
        lib.propagate_hop_tc16_fwd(o, o, s, rowptr, src, 0, n, img16, 1, 0, None, 0, None, None, out)
    else:
        lib.propagate_hop_tc_fwd(o, o, s, rowptr, src, 0, n, img, 1, 0, None, 0, None, None, out)
for which in os.environ.get("WHICH", "nofilter,ffma,tc,tc16,tc16s").split(","):
    for _ in range(3): run(which)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps): run(which)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{which}: n={n} e={e} {ms:.4f} ms  {bytes_alg / ms / 1e6:.0f} GB/s algorithmic ({bytes_alg / ms / 1e6 / 6541.1:.3f} of measured copy peak)")

if os.environ.get("TRACE"):
    import ctypes as C
    l = lib.load()
    fn = l.swe_propagate_hop_tc_fwd_traced
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 5 + [C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32] + [C.c_void_p] * 5
    trace = torch.zeros(3 * 128, dtype=torch.int64, device=DEV)
    rc = fn(o.data_ptr(), o.data_ptr(), s.data_ptr(), rowptr.data_ptr(), src.data_ptr(), 0, n, img.data_ptr(), 1, 0, None, 0, None,
            None, out.data_ptr(), trace.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    t = trace.cpu().view(3, 16, 8)
    t0 = int(t[0, 2, 0])
    ev = [["start", "slot_free", "round0", "round1", "prev_out"], ["begin", "d_full", "staged"], ["wait", "a_full", "issued"]]
    for tile in range(2, 9):
        for r, name in enumerate(["gather", "epilogue", "mma"]):
            print(tile, name, " ".join(f"{ev[r][e]}={int(t[r, tile, e]) - t0}" for e in range(len(ev[r]))))

if os.environ.get("TRACE16S"):
    import ctypes as C
    l = lib.load()
    l.swe_hop_tc16s_set_trace.restype = None
    l.swe_hop_tc16s_set_trace.argtypes = [C.c_void_p]
    trace = torch.zeros(16 * 16, dtype=torch.int64, device=DEV)
    l.swe_hop_tc16s_set_trace(trace.data_ptr())
    lib.propagate_hop_tc16s_fwd(o, o, s, rowptr, src, 0, n, img16, 1, 0, None, 0, None, None, out)
    torch.cuda.synchronize()
    t = trace.cpu().view(16, 16)
    t0 = int(t[2, 0])
    names = ["start", "ids", "p0_iss", "p0_full", "p0_acc", "a_empty", "prev_out", "p1_iss", "p1_full", "p1_acc", "end"]
    for tile in range(2, 10):
        print(tile, " ".join(f"{names[e]}={int(t[tile, e]) - t0}" for e in range(len(names))))
