"""Times the edge-gate kernels alone on a cfg3-sized finest level (3.04 M edges)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mswe_gnn_b200  # noqa
from mswe_gnn_b200 import lib
from mswe_gnn_b200.engine import PackedGateTC, PackedMLP
from mswe_gnn_b200.models.models import make_mlp
from mswe_gnn_b200.utils.synthetic import make_tri_mesh

DEV = "cuda:0"
nx = int(os.environ.get("NX", "712"))
d = make_tri_mesh(nx, nx, 1)
n = d.x.shape[0]
row, col = d.edge_index.to(DEV)
rowptr, src, dst, eid = lib.csr_build(row, col, None, 0, n, 0, n)
E = int(src.numel())
torch.manual_seed(0)
xs = torch.randn(n, 64, device=DEV); xd = torch.randn(n, 64, device=DEV); a = torch.randn(E, 64, device=DEV)
mlp = make_mlp(320, 64, hidden_size=128, n_layers=3, bias=True, activation="prelu").to(DEV)
tc = PackedGateTC(mlp); codes, slopes = tc.acts_and_slopes(); img = tc.image()
pk = PackedMLP(mlp, [(64, 64)] * 5, {}); st = pk.struct()
s = torch.empty(E, 64, device=DEV)
which = os.environ.get("WHICH", "tc16,tc16ng,tc,ffma").split(",")
img16 = tc.image16(); ws = tc.flag_ws(E)
for name in which:
    fn = {"tc": lambda: lib.edge_gate_tc_fwd(xs, xd, xd, a, src, dst, E, img, 320, codes, slopes, True, s, None),
          "tc16": lambda: lib.edge_gate_tc16_fwd(xs, xd, xd, a, src, dst, E, img16, img, 320, codes, slopes, True, s, None, ws),
          "tc16ng": lambda: lib.edge_gate_tc16_fwd(xs, xd, xd, a, src, dst, E, img16, None, 320, codes, slopes, True, s, None, None),
          "ffma": lambda: lib.edge_gate_fwd(xs, xd, xd, a, src, dst, E, st, True, s, 64)}[name]
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = int(os.environ.get("REPS", "5"))
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * E * (320 * 128 + 128 * 128 + 128 * 64)
    print(f"{name}: E={E} {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s (algorithmic)  {ms * 1e3 / ((E + 127) // 128 / 148):.2f} us/tile/SM")

if os.environ.get("TRACE"):
    import ctypes as C
    l = lib.load()
    fn = l.swe_edge_gate_tc_fwd_traced
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 6 + [C.c_int64, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    trace = torch.zeros(3 * 128, dtype=torch.int64, device=DEV)
    act3 = (C.c_int32 * 3)(*codes); sl3 = (C.c_void_p * 3)(*[None if t is None else t.data_ptr() for t in slopes])
    rc = fn(xs.data_ptr(), xd.data_ptr(), xd.data_ptr(), a.data_ptr(), src.data_ptr(), dst.data_ptr(), E, img.data_ptr(), 320,
            act3, sl3, 1, s.data_ptr(), None, trace.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    t = trace.cpu().view(3, 16, 8)
    t0 = int(t[0, 2, 0])
    names = ["rowA(w0)", "rowB(w4)", "mma"]
    for tile in range(2, 8):
        for r in range(3):
            print(tile, names[r], [(int(v) - t0) if int(v) else None for v in t[r, tile, :6]])

if os.environ.get("TRACE_DEC"):
    import ctypes as C
    l = lib.load()
    l.swe_gate_tc_set_trace.argtypes = [C.c_void_p]
    n = xs.shape[0]
    p_src = torch.empty(n, 128, device=DEV); p_dst = torch.empty(n, 128, device=DEV)
    for name, fn in (("partials", lambda: lib.gate_partials_tc(xs, xd, 0, n, img, 320, 0, p_src)),
                     ("dec", lambda: lib.edge_gate_tc_dec_fwd(p_src, p_dst, a, src, dst, E, img, 320, codes, slopes, True, s))):
        lib.gate_partials_tc(xs, xd, 0, n, img, 320, 1, p_dst)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 10:.3f} ms")
        trace = torch.zeros(3 * 128, dtype=torch.int64, device=DEV)
        l.swe_gate_tc_set_trace(trace.data_ptr())
        fn()
        torch.cuda.synchronize()
        t = trace.cpu().view(3, 16, 8)
        t0 = int(t[0, 2, 0]) if int(t[0, 2, 0]) else int(t[2, 2, 0])
        for tile in range(2, 6):
            for r, nm in enumerate(["rowA(w0)", "rowB(w4)", "mma"]):
                print(tile, nm, [(int(v) - t0) if int(v) else None for v in t[r, tile, :6]])

if os.environ.get("TRACE_STAT"):
    import ctypes as C
    l = lib.load()
    l.swe_gate_tc_set_trace.argtypes = [C.c_void_p]
    tab = torch.empty((E + 127) // 128 * 128, 128, device=DEV)
    for name, fn in (("static partials (once per rollout)", lambda: lib.gate_static_partials_tc(None, a, src, dst, E, img, 320, tab)),
                     ("stat gate (a_e hoisted)", lambda: lib.edge_gate_tc_stat_fwd(tab, xs, xd, xd, src, dst, E, img, 320, codes, slopes, True, s)),
                     ("full gate", lambda: lib.edge_gate_tc_fwd(xs, xd, xd, a, src, dst, E, img, 320, codes, slopes, True, s, None))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / 10:.3f} ms")
        trace = torch.zeros(3 * 128, dtype=torch.int64, device=DEV)
        l.swe_gate_tc_set_trace(trace.data_ptr())
        fn()
        torch.cuda.synchronize()
        t = trace.cpu().view(3, 16, 8)
        t0 = int(t[0, 2, 0]) if int(t[0, 2, 0]) else int(t[2, 2, 0])
        for tile in range(2, 6):
            for r, nm in enumerate(["rowA(w0)", "rowB(w4)", "mma"]):
                print("  ", tile, nm, [(int(v) - t0) if int(v) else None for v in t[r, tile, :6]])

if os.environ.get("TRACE16"):
    import ctypes as C
    l = lib.load()
    l.swe_gate_tc16_set_trace.argtypes = [C.c_void_p]
    trace = torch.zeros(3 * 128 + 64, dtype=torch.int64, device=DEV)
    l.swe_gate_tc16_set_trace(trace.data_ptr())
    lib.edge_gate_tc16_fwd(xs, xd, xd, a, src, dst, E, img16, None, 320, codes, slopes, True, s, None, None)
    torch.cuda.synchronize()
    tc = trace.cpu()[384:]
    t = trace.cpu()[:384].view(3, 16, 8)
    t0 = int(t[0, 0, 0])
    print("tc16 trace (cycles since group 0's first tile): rows = local tile of the role")
    print("  groups: [ids done, gather done, D0 ready, E1 done, D1 ready, E2 done, D2 ready, E3 done]")
    print("  mma   : [P0 start, P0 issued, P1 start, P1 issued, Q1 start, Q1 issued, Q2 start, Q2 issued]  (tile k: group k & 1)")
    for r, nm in enumerate(["group0", "group1", "mma"]):
        for tile in range(8):
            print("  ", nm, tile, [(int(v) - t0) if int(v) else None for v in t[r, tile]])
    print("  gather of group 0, its tile 2, per chunk: [before a_empty wait, slot free, converted + stored, fenced + arrived]")
    for c in range(10):
        print("    chunk", c, [(int(v) - t0) if int(v) else None for v in tc[4 * c:4 * c + 4]])
