"""Times the row-MLP pieces (encoders, W0, decoder head) at cfg3 size with both backends."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mswe_gnn_b200  # noqa
from mswe_gnn_b200 import lib
from mswe_gnn_b200.models.gnn import MSGNN
DEV = "cuda"
N = int(os.environ.get("N", 1346574))
ctor = dict(num_node_features=8, num_edge_features=1, num_scales=4, previous_t=3, hid_features=64, mlp_layers=3, seed=666,
            learned_residuals=True, mlp_activation="prelu", gnn_activation="tanh", with_WL=True, K=4)
m = MSGNN(**ctor).to(DEV)
x = torch.rand(N, 8, device=DEV)
xs = torch.empty(N, 64, device=DEV); xd = torch.empty(N, 64, device=DEV); h = torch.randn(N, 64, device=DEV)
pred = torch.empty(N, 2, device=DEV); xn = torch.empty_like(x)
class P: perm = None; n_nodes = N
plan = P()
def t(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
la = m.gnn_processor[0].launcher()
if os.environ.get("ONLY_W0"):
    print("tc W0: %.3f ms" % t(lambda: la.w0_tc.linear(h, 0, N, xs)))
    import ctypes as C
    l = lib.load(); l.swe_row_mlp_tc_set_trace.argtypes = [C.c_void_p]
    tr = torch.zeros(3 * 128, dtype=torch.int64, device=DEV)
    l.swe_row_mlp_tc_set_trace(tr.data_ptr())
    la.w0_tc.linear(h, 0, N, xs); torch.cuda.synchronize()
    tt = tr.cpu().view(3, 16, 8); t0 = int(tt[0, 3, 0])
    ev = [["start", "loaded", "slot_free", "a_full", "prev_out"], ["begin", "d0_full", "stage_free", "staged"], ["begin", "d0_free", "a_full", "issued"]]
    for tile in range(3, 8):
        for r, nm in enumerate(["row", "epi", "mma"]):
            print(tile, nm, " ".join(f"{ev[r][e]}={int(tt[r, tile, e]) - t0}" for e in range(len(ev[r]))))
    sys.exit(0)
if os.environ.get("ONLY_DEC") or os.environ.get("ONLY_ENC"):
    import ctypes as C
    l = lib.load(); l.swe_row_mlp_tc_set_trace.argtypes = [C.c_void_p]
    fn = (lambda: m._decode(h, "tanh", m.gnn_activation, x, plan, pred, None, 0, xn)) if os.environ.get("ONLY_DEC") else \
         (lambda: m._tc_static.encode(x, 0, 2, True, (1, 6), None, 0, N, xs))
    print("time: %.3f ms" % t(fn))
    tr = torch.zeros(3 * 128, dtype=torch.int64, device=DEV)
    l.swe_row_mlp_tc_set_trace(tr.data_ptr())
    fn(); torch.cuda.synchronize()
    tt = tr.cpu().view(3, 16, 8); t0 = int(tt[0, 3, 0])
    for tile in range(3, 7):
        for r, nm in enumerate(["row", "epi", "mma"]):
            print(tile, nm, " ".join(str(int(tt[r, tile, e]) - t0) for e in range(7)))
    sys.exit(0)
for be in ("tc", "ffma"):
    os.environ["MSWE_ROWMLP"] = be
    print(be, "encode (static N + dynamic N): %.3f ms" % t(lambda: m._encode_nodes(x, plan, N, xs, xd)))
    if be == "tc":
        print(be, "  static only: %.3f ms" % t(lambda: m._tc_static.encode(x, 0, 2, True, (1, 6), None, 0, N, xs)))
        print(be, "  dynamic only: %.3f ms" % t(lambda: m._tc_dynamic.encode(x, 2, 6, False, (0, 0), None, 0, N, xd)))
        os.environ["MSWE_ROWLIN"] = "tc"
        print(be, "W0 (swe_row_mlp_tc, 3xTF32): %.3f ms" % t(lambda: la.w0_tc.linear(h, 0, N, xs)))
        os.environ["MSWE_ROWLIN"] = "tc16"
        print(be, "W0 (swe_row_linear_tc16, streaming): %.3f ms" % t(lambda: la.w0_tc.linear(h, 0, N, xs)))
    else:
        print(be, "W0: %.3f ms" % t(lambda: lib.node_linear_fwd(h, 0, N, la.filters.tensors()[0], xs, 64)))
    print(be, "decode head: %.3f ms" % t(lambda: m._decode(h, "tanh", m.gnn_activation, x, plan, pred, None, 0, xn)))
