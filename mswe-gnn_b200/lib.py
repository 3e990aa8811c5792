"""ctypes binding of ``libswe_gnn_b200.so`` (C ABI in ``include/swe_gnn_b200.h``).

This is the only place the package touches native code.  There is NO fallback: if the shared
object is missing or a symbol declared in the header is not exported, loading raises; every
wrapper converts a non-zero status into ``RuntimeError`` carrying ``swe_last_error()``.
Tensors are passed as raw device pointers (``tensor.data_ptr()``) after dtype / device /
contiguity checks; the CUDA stream is torch's current stream, so calls are captured by
``torch.cuda.graph`` like any other launch.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libswe_gnn_b200.so")

SWE_MAX_LAYERS = 8
ACT_CODES = {None: 0, "prelu": 1, "relu": 2, "tanh": 3, "leakyrelu": 4, "elu": 5, "swish": 6, "sigmoid": 7}


class SweLayer(C.Structure):
    _fields_ = [("wt", C.c_void_p), ("bias", C.c_void_p), ("slope", C.c_void_p),
                ("k_in", C.c_int32), ("n_out", C.c_int32), ("act", C.c_int32), ("_pad", C.c_int32)]


class SweMlp(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("_pad", C.c_int32), ("layer", SweLayer * SWE_MAX_LAYERS)]


SWE_MAX_SEGS = 5


class SweSeg(C.Structure):
    _fields_ = [("base", C.c_void_p), ("idx", C.c_void_p), ("slope", C.c_void_p),
                ("ld", C.c_int32), ("width", C.c_int32), ("act", C.c_int32), ("_pad", C.c_int32)]


class SweRows(C.Structure):
    _fields_ = [("n_seg", C.c_int32), ("_pad", C.c_int32), ("seg", SweSeg * SWE_MAX_SEGS)]


class SweRowMlp(C.Structure):
    _fields_ = [("x_rows", C.c_void_p), ("act_in", C.c_int32), ("_pad0", C.c_int32), ("slope_in", C.c_void_p),
                ("raw", C.c_void_p), ("raw_ld", C.c_int32), ("raw_col0", C.c_int32), ("raw_cols", C.c_int32),
                ("with_wl", C.c_int32), ("wl_col_a", C.c_int32), ("wl_col_b", C.c_int32), ("perm", C.c_void_p),
                ("w_first", C.c_void_p), ("b_first", C.c_void_p), ("act_first", C.c_int32), ("_pad1", C.c_int32),
                ("slope_first", C.c_void_p), ("row_lo", C.c_int32), ("_pad2", C.c_int32), ("n_rows", C.c_int64),
                ("n_tc", C.c_int32), ("_pad3", C.c_int32), ("img", C.c_void_p * 2), ("bias", C.c_void_p * 2),
                ("act", C.c_int32 * 2), ("slope", C.c_void_p * 2), ("out_rows", C.c_void_p),
                ("head", C.c_int32), ("act_head", C.c_int32), ("w_head", C.c_void_p), ("b_head", C.c_void_p),
                ("slope_head", C.c_void_p), ("x0", C.c_void_p), ("n_cols", C.c_int32), ("previous_t", C.c_int32),
                ("head_perm", C.c_void_p), ("res_mode", C.c_int32), ("eps", C.c_float), ("res_w", C.c_void_p),
                ("pred", C.c_void_p), ("step_ptr", C.c_void_p), ("pred_step_stride", C.c_int64), ("x_next", C.c_void_p)]


_p, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
_mlp = C.POINTER(SweMlp)
_rows = C.POINTER(SweRows)
_pi32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes); mirrors include/swe_gnn_b200.h one to one
SIGNATURES = {
    "swe_abi_version": (C.c_int, []),
    "swe_last_error": (C.c_char_p, []),
    "swe_build_arch": (C.c_char_p, []),
    "swe_pack_linear": (C.c_int, [_p, _i32, _i32, _i32, _p, _p]),
    "swe_csr_build_ws_bytes": (_sz, [_i64, _i32]),
    "swe_csr_build": (C.c_int, [_p, _p, _i64, _p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "swe_node_encode_fwd": (C.c_int, [_p, _i32, _p, _i32, _i32, _i32, _i32, _mlp, _mlp, _p, _p, _i32, _p]),
    "swe_edge_encode_fwd": (C.c_int, [_p, _i32, _p, _i64, _mlp, _p, _i32, _p]),
    "swe_edge_gate_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _mlp, _i32, _p, _i32, _p]),
    "swe_gate_tc_image_bytes": (_sz, [_i32]),
    "swe_gate_tc_pack": (C.c_int, [_p, _i32, _p, _p, _p, _p, _p, _p, _p]),
    "swe_edge_gate_tc_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _p, _i32, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_void_p), _i32, _p, _p, _p]),
    "swe_edge_gate_tc_fwd_listed": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _p, _i32, C.POINTER(C.c_int32),
                                              C.POINTER(C.c_void_p), _i32, _p, _p, _p]),
    "swe_gate_tc16_image_bytes": (_sz, [_i32]),
    "swe_gate_tc16_pack": (C.c_int, [_p, _i32, _p, _p, _p, _p, _p, C.POINTER(C.c_float), _p, _p]),
    "swe_edge_gate_tc16_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _p, _p, _i32, C.POINTER(C.c_int32),
                                         C.POINTER(C.c_void_p), _i32, _p, _p, _p, _p]),
    "swe_row_mlp_tc16": (C.c_int, [C.POINTER(SweRowMlp), C.POINTER(C.c_void_p), _p]),
    "swe_row_linear_tc16": (C.c_int, [_p, _i64, _i64, _p, _p, _p]),
    "swe_gate_partials_tc": (C.c_int, [_p, _p, _i32, _i32, _p, _i32, _i32, _p, _p]),
    "swe_edge_gate_tc_dec_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _p, _i32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p),
                                           _i32, _p, _p]),
    "swe_row_mlp_tc": (C.c_int, [C.POINTER(SweRowMlp), _p]),
    "swe_hop_tc_image_bytes": (_sz, []),
    "swe_hop_tc_pack": (C.c_int, [_p, _p, _p]),
    "swe_propagate_hop_tc_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i32, _i32, _p, _i32, _p, _p, _p, _p]),
    "swe_hop_tc16_image_bytes": (_sz, []),
    "swe_hop_tc16_pack": (C.c_int, [_p, _f32, _p, _p]),
    "swe_propagate_hop_tc16_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i32, _i32, _p, _i32, _p, _p, _p, _p]),
    "swe_propagate_hop_tc16s_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i32, _i32, _p, _i32, _p, _p, _p, _p]),
    "swe_node_linear_fwd": (C.c_int, [_p, _i32, _i32, _p, _p, _i32, _p]),
    "swe_propagate_hop_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i32, _i32, _p, _i32, _p, _p, _i32, _p]),
    "swe_pool_mean_fwd": (C.c_int, [_p, _p, _p, _i32, _i32, _p, _i32, _p]),
    "swe_decode_head_fwd": (C.c_int, [_p, _i32, _p, _mlp, _p, _i32, _p, _i32, _i32, _i32, _p, _f32, _p, _p, _i64,
                                      _p, _i32, _p]),
    "swe_apply_bc": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _i32, _p, _i32, _p, _p]),
    "swe_step_advance": (C.c_int, [_p, _p]),
    "swe_pack_rows": (C.c_int, [_p, _p, _i64, _i32, _p, _p]),
    "swe_temporal_window": (C.c_int, [_p, _i32, _p, _p, _i64, _i32, _p, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p]),
    "swe_rollout_metrics_cols": (_i32, []),
    "swe_rollout_metrics_ws_bytes": (_sz, [_i32]),
    "swe_rollout_metrics": (C.c_int, [_p, _p, _i64, _i32, _p, _i32, _p, _p, _p]),
    "swe_train_step_ws_bytes": (_sz, []),
    "swe_loss_fwd_bwd": (C.c_int, [_p, _p, _i64, _p, _i64, _i32, _i32, _f32, _f32, _f32, _i32, _p, _p, _p, _p]),
    "swe_clip_adamw_step": (C.c_int, [_p, _p, _p, _p, _i64, _p, _f32, _f32, _f32, _f32, _f32, _p, _p, _p]),
    "swe_ipc_alloc": (C.c_int, [_sz, C.POINTER(C.c_void_p), C.c_char_p]),
    "swe_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "swe_ipc_close": (C.c_int, [_p]),
    "swe_ipc_free": (C.c_int, [_p]),
    "swe_halo_exchange": (C.c_int, [_p, _i32, _i32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _p, _p, _i32, _i32, _p]),
    # training path
    "swe_mlp_layer_fwd": (C.c_int, [_rows, _i64, _p, _p, _i32, _p, _p]),
    "swe_mlp_layer_bwd_dx": (C.c_int, [_p, _p, _i32, _p, _i64, _i32, _p, _i32, _i32, _i32, _i32, _p, _i32, _i32, _p,
                                       _pi32, _p]),
    "swe_mlp_layer_bwd_dx_grid": (C.c_int, [_i64]),
    "swe_mlp_layer_bwd_dw": (C.c_int, [_p, _i64, _i32, _rows, _i32, _p, _pi32, _p]),
    "swe_mlp_layer_bwd_dw_grid": (C.c_int, [_i64]),
    "swe_gate_static_partials_tc": (C.c_int, [_p, _p, _p, _p, _i64, _p, _i32, _p, _p]),
    "swe_edge_gate_tc_stat_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _p, _i32, C.POINTER(C.c_int32),
                                            C.POINTER(C.c_void_p), _i32, _p, _p]),
    "swe_edge_gate_tc_train_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _p, _i32, C.POINTER(C.c_int32),
                                             C.POINTER(C.c_void_p), _i32, _p, _p, _p, _p, _p, _p, _i32, C.c_float, _p]),
    "swe_gate_fix_preacts": (C.c_int, [_p] * 12 + [_i32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p), _p, _p, _p, _p, _p,
                                       _i32, _p]),
    "swe_mlp_layer_bwd_dx_tc": (C.c_int, [_p, _i64, _i32, _p, _i32, _i32, _i32, _i32, _p, _i32, _p, _i32, _i32, _p]),
    "swe_mlp_layer_bwd_dx_tc_fused": (C.c_int, [_p, _p, _i32, _p, _i64, _i32, _p, _i32, _i32, _i32, _i32, _p, _i32, _p, _i32,
                                                _i32, _p, _pi32, _p]),
    "swe_mlp_layer_bwd_dx_tc_grid": (C.c_int, [_i64]),
    "swe_mlp_layer_bwd_dw_tc": (C.c_int, [_p, _i64, _i32, _rows, _p, _pi32, _p]),
    "swe_mlp_layer_bwd_dw_tc_grid": (C.c_int, [_i64]),
    "swe_reduce_partials": (C.c_int, [_p, _i32, _i64, _i32, _i32, _i32, _i32, _p, _i32, _i32, _p]),
    "swe_gate_norm_fwd": (C.c_int, [_p, _i32, _p, _i32, _i64, _p, _i32, _p]),
    "swe_gate_norm_bwd": (C.c_int, [_p, _p, _i32, _p, _i32, _i64, _i32, _p]),
    "swe_act_fwd": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _i32, _p]),
    "swe_act_bwd": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _pi32, _i32, _p]),
    "swe_propagate_hop_train_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i32, _i32, _p, _i32, _i32, _p, _p, _p, _i32, _p]),
    "swe_row_flags": (C.c_int, [_p, _i32, _i32, _p, _i32, _p]),
    "swe_hop_bwd_dst": (C.c_int, [_p, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _i32, _i32, _i32, _p, _p, _i32, _p]),
    "swe_hop_bwd_src": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _i32, _p]),
    "swe_edge_to_node_sum": (C.c_int, [_p, _p, _p, _i32, _i32, _p, _i32, _i32, _p]),
    "swe_pool_mean_bwd": (C.c_int, [_p, _p, _p, _i32, _i32, _p, _i32, _p, _i32, _i32, _p]),
    "swe_static_inputs_fwd": (C.c_int, [_p, _i32, _p, _i32, _i32, _i32, _p, _i32, _p]),
    "swe_node_inputs_bwd": (C.c_int, [_p, _i32, _p, _i32, _i32, _p, _i32, _i32, _i32, _i32, _p, _p]),
    "swe_head_fwd": (C.c_int, [_p, _i32, _i32, _p, _p, _i32, _p, _i32, _i32, _i32, _p, _f32, _p, _p]),
    "swe_head_bwd": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _i32, _p, _i32, _i32, _i32, _p, _f32, _p, _p, _p, _pi32, _p]),
}

_lib = None


def load(path: Optional[str] = None):
    """Load the shared object and bind every declared symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: the CUDA extension is not built. Run `python -c \"import __graft_entry__ as g; "
            f"g.build()\"` (needs nvcc). There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise RuntimeError(f"{path} does not export `{name}` declared in include/swe_gnn_b200.h") from e
        fn.restype, fn.argtypes = res, args
    if lib.swe_abi_version() != 1:
        raise RuntimeError(f"ABI version mismatch: library {lib.swe_abi_version()}, binding 1")
    _lib = lib
    return lib


launch_count = 0          # C-ABI calls issued so far (each is one kernel launch; csr_build is several)
_launch_log = None        # when a list: (name, meta) of every call is appended (bench instrumentation)


def _check(status: int, what: str):
    global launch_count
    launch_count += 1
    if status != 0:
        msg = load().swe_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {status}): {msg}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream on the current device (also the capturing stream inside
    ``torch.cuda.graph``).  The raw query is ~20x cheaper than building a ``torch.cuda.Stream`` object, which matters
    with ~700 launches per training step."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor], dtype=torch.float32) -> Optional[int]:
    """Device pointer of a contiguous CUDA tensor of the given dtype (None passes NULL)."""
    if t is None or isinstance(t, int):
        return t
    if not t.is_cuda:
        raise RuntimeError("mswe_gnn_b200 kernels need CUDA tensors; got a %s tensor (no CPU fallback)" % t.device)
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    if t.device.index != torch.cuda.current_device():
        # every launch goes to the current device's stream (_stream): a tensor of another device would be an illegal address
        raise RuntimeError(f"tensor on {t.device} but the current CUDA device is {torch.cuda.current_device()}: "
                           "call torch.cuda.set_device / use `with torch.cuda.device(...)` around the call")
    return t.data_ptr()


# ------------------------------------------------------------------------------------------------
# thin wrappers (one per C entry point)
# ------------------------------------------------------------------------------------------------
def pack_linear(w: torch.Tensor, k_pad: int, out: torch.Tensor):
    n_out, k_in = w.shape
    _check(load().swe_pack_linear(ptr(w), n_out, k_in, k_pad, ptr(out), _stream()), "swe_pack_linear")


def csr_build(row, col, node_map, dst_lo, n_dst, src_lo, src_hi, by_row=False):
    """Returns (rowptr, src, dst, eid) int32 tensors; raises ValueError on out-of-range edges."""
    lib = load()
    E = int(row.numel())
    dev = row.device
    rowptr = torch.empty(n_dst + 1, dtype=torch.int32, device=dev)
    src = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    dst = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    eid = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.swe_csr_build_ws_bytes(E, n_dst))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _check(lib.swe_csr_build(ptr(row, torch.int64), ptr(col, torch.int64), E, ptr(node_map, torch.int32),
                             dst_lo, n_dst, src_lo, src_hi, int(by_row), ptr(rowptr, torch.int32),
                             ptr(src, torch.int32), ptr(dst, torch.int32), ptr(eid, torch.int32),
                             ptr(err, torch.int32), ws.data_ptr(), ws_bytes, _stream()), "swe_csr_build")
    n_bad = int(err.item())
    if n_bad:
        raise ValueError(f"{n_bad} edges have an endpoint outside the node range of their scale "
                         f"(dst range [{dst_lo}, {dst_lo + n_dst}), src range [{src_lo}, {src_hi}))")
    return rowptr, src[:E], dst[:E], eid[:E]


def node_encode_fwd(x, perm, n_nodes, n_static_raw, with_wl, n_dyn_rows, mlp_s: SweMlp, mlp_d: SweMlp, xs, xd, F):
    _check(load().swe_node_encode_fwd(ptr(x), x.shape[1], ptr(perm, torch.int32), n_nodes, n_static_raw, int(with_wl),
                                      n_dyn_rows, C.byref(mlp_s), C.byref(mlp_d), ptr(xs), ptr(xd), F, _stream()),
           "swe_node_encode_fwd")


def edge_encode_fwd(edge_attr, eid, n_edges, mlp: SweMlp, a_out, F):
    _check(load().swe_edge_encode_fwd(ptr(edge_attr), edge_attr.shape[1], ptr(eid, torch.int32), n_edges,
                                      C.byref(mlp), ptr(a_out), F, _stream()), "swe_edge_encode_fwd")


def edge_gate_fwd(xs, xd_src, xd_dst, a, src, dst, n_edges, mlp: SweMlp, normalize, s_out, F):
    _check(load().swe_edge_gate_fwd(ptr(xs), ptr(xd_src), ptr(xd_dst), ptr(a), ptr(src, torch.int32),
                                    ptr(dst, torch.int32), n_edges, C.byref(mlp), int(normalize), ptr(s_out), F,
                                    _stream()), "swe_edge_gate_fwd")


def gate_tc_image_bytes(k1: int) -> int:
    return int(load().swe_gate_tc_image_bytes(k1))


def gate_tc_pack(w1, b1, w2, b2, w3, b3, image):
    _check(load().swe_gate_tc_pack(ptr(w1), w1.shape[1], ptr(b1), ptr(w2), ptr(b2), ptr(w3), ptr(b3),
                                   image.data_ptr(), _stream()), "swe_gate_tc_pack")


def edge_gate_tc_fwd(xs, xd_src, xd_dst, a, src, dst, n_edges, image, k1, acts, slopes, normalize, s_out, dbg=None):
    act3 = (C.c_int32 * 3)(*acts)
    slope3 = (C.c_void_p * 3)(*[None if s is None else ptr(s) for s in slopes])
    _check(load().swe_edge_gate_tc_fwd(ptr(xs), ptr(xd_src), ptr(xd_dst), ptr(a), ptr(src, torch.int32),
                                       ptr(dst, torch.int32), n_edges, image.data_ptr(), k1, act3, slope3,
                                       int(normalize), ptr(s_out), ptr(dbg), _stream()), "swe_edge_gate_tc_fwd")


def gate_tc16_image_bytes(k1: int) -> int:
    return int(load().swe_gate_tc16_image_bytes(k1))


def gate_tc16_pack(w1, b1, w2, b2, w3, b3, wmax3, image):
    """wmax3: three host floats, max |w| of the three layers (the power-of-two weight scales are chosen from them)."""
    wm = (C.c_float * 3)(*[float(v) for v in wmax3])
    _check(load().swe_gate_tc16_pack(ptr(w1), w1.shape[1], ptr(b1), ptr(w2), ptr(b2), ptr(w3), ptr(b3), wm,
                                     image.data_ptr(), _stream()), "swe_gate_tc16_pack")


def edge_gate_tc16_fwd(xs, xd_src, xd_dst, a, src, dst, n_edges, image16, image_tf32, k1, acts, slopes, normalize, s_out,
                       dbg=None, flag_ws=None):
    act3 = (C.c_int32 * 3)(*acts)
    slope3 = (C.c_void_p * 3)(*[None if s is None else ptr(s) for s in slopes])
    _check(load().swe_edge_gate_tc16_fwd(ptr(xs), ptr(xd_src), ptr(xd_dst), ptr(a), ptr(src, torch.int32),
                                         ptr(dst, torch.int32), n_edges, image16.data_ptr(),
                                         None if image_tf32 is None else image_tf32.data_ptr(), k1, act3, slope3,
                                         int(normalize), ptr(s_out), ptr(dbg), ptr(flag_ws, torch.int32), _stream()),
           "swe_edge_gate_tc16_fwd")


def row_mlp_tc16(desc, imgs16) -> bool:
    """swe_row_mlp_tc16; False when the shape is not covered by the fp16 streaming kernel (caller falls back)."""
    arr = (C.c_void_p * 2)(*[t.data_ptr() for t in imgs16])
    rc = load().swe_row_mlp_tc16(C.byref(desc), arr, _stream())
    if rc == -3:                                           # SWE_E_UNSUPP
        return False
    _check(rc, "swe_row_mlp_tc16")
    return True


def row_linear_tc16(x, row_lo, n_rows, w_image, out):
    """out[row_lo + r] = x[row_lo + r] · Wᵀ (swe_row_linear_tc16; w_image: hop_tc16_pack of W)."""
    _check(load().swe_row_linear_tc16(ptr(x), int(row_lo), int(n_rows), w_image.data_ptr(), ptr(out), _stream()),
           "swe_row_linear_tc16")


def edge_gate_tc_fwd_listed(xs, xd_src, xd_dst, a, src, dst, n_edges, image, k1, acts, slopes, normalize, s_out, tile_list):
    act3 = (C.c_int32 * 3)(*acts)
    slope3 = (C.c_void_p * 3)(*[None if s is None else ptr(s) for s in slopes])
    _check(load().swe_edge_gate_tc_fwd_listed(ptr(xs), ptr(xd_src), ptr(xd_dst), ptr(a), ptr(src, torch.int32),
                                              ptr(dst, torch.int32), n_edges, image.data_ptr(), k1, act3, slope3,
                                              int(normalize), ptr(s_out), ptr(tile_list, torch.int32), _stream()),
           "swe_edge_gate_tc_fwd_listed")


def gate_static_partials_tc(xs, a, src, dst, n_edges, image, k1, p_out):
    _check(load().swe_gate_static_partials_tc(ptr(xs), ptr(a), ptr(src, torch.int32), ptr(dst, torch.int32), n_edges,
                                              image.data_ptr(), k1, ptr(p_out), _stream()), "swe_gate_static_partials_tc")


def edge_gate_tc_stat_fwd(p_edge, xs, xd_src, xd_dst, src, dst, n_edges, image, k1, acts, slopes, normalize, s_out):
    act3 = (C.c_int32 * 3)(*acts)
    slope3 = (C.c_void_p * 3)(*[None if s is None else ptr(s) for s in slopes])
    _check(load().swe_edge_gate_tc_stat_fwd(ptr(p_edge), ptr(xs), ptr(xd_src), ptr(xd_dst), ptr(src, torch.int32),
                                            ptr(dst, torch.int32), n_edges, image.data_ptr(), k1, act3, slope3,
                                            int(normalize), ptr(s_out), _stream()), "swe_edge_gate_tc_stat_fwd")


def edge_gate_tc_train_fwd(xs, xd_src, xd_dst, a, src, dst, n_edges, image, k1, acts, slopes, normalize, pre1, pre2, pre3,
                           s_out, fix_lists=None, fix_count=None, fix_cap=0, fix_tau=0.0):
    act3 = (C.c_int32 * 3)(*acts)
    slope3 = (C.c_void_p * 3)(*[None if s is None else ptr(s) for s in slopes])
    _check(load().swe_edge_gate_tc_train_fwd(_addr(xs), _addr(xd_src), _addr(xd_dst), _addr(a), ptr(src, torch.int32),
                                             ptr(dst, torch.int32), n_edges, image.data_ptr(), k1, act3, slope3,
                                             int(normalize), ptr(pre1), ptr(pre2), ptr(pre3), ptr(s_out),
                                             None if fix_lists is None else fix_lists.data_ptr(),
                                             None if fix_count is None else fix_count.data_ptr(), int(fix_cap),
                                             float(fix_tau), _stream()),
           "swe_edge_gate_tc_train_fwd")


def gate_fix_preacts(xs, xd_src, xd_dst, a, src, dst, w1, b1, w2, b2, w3, b3, k1, acts, slopes, pre1, pre2, pre3, fix_lists,
                     fix_count, fix_cap):
    act3 = (C.c_int32 * 3)(*acts)
    slope3 = (C.c_void_p * 3)(*[None if s is None else ptr(s) for s in slopes])
    _check(load().swe_gate_fix_preacts(_addr(xs), _addr(xd_src), _addr(xd_dst), _addr(a), ptr(src, torch.int32),
                                       ptr(dst, torch.int32), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(w3), ptr(b3), k1, act3,
                                       slope3, ptr(pre1), ptr(pre2), ptr(pre3), fix_lists.data_ptr(), fix_count.data_ptr(),
                                       int(fix_cap), _stream()), "swe_gate_fix_preacts")


def node_linear_fwd(x, row_lo, n_rows, wt, out, F):
    _check(load().swe_node_linear_fwd(ptr(x), row_lo, n_rows, ptr(wt), ptr(out), F, _stream()), "swe_node_linear_fwd")


def propagate_hop_fwd(o_src, o_dst, s, rowptr, src, dst_lo, n_dst, wt, with_gradient, upwind, addend, act, slope,
                      out, F):
    _check(load().swe_propagate_hop_fwd(ptr(o_src), ptr(o_dst), ptr(s), ptr(rowptr, torch.int32),
                                        ptr(src, torch.int32), dst_lo, n_dst, ptr(wt), int(with_gradient),
                                        int(upwind), ptr(addend), act, ptr(slope), ptr(out), F, _stream()),
           "swe_propagate_hop_fwd")


def gate_partials_tc(xs, xd, row_lo, n_rows, image, k1, role, p_out):
    _check(load().swe_gate_partials_tc(ptr(xs), ptr(xd), row_lo, n_rows, image.data_ptr(), k1, role, ptr(p_out), _stream()),
           "swe_gate_partials_tc")


def edge_gate_tc_dec_fwd(p_src, p_dst, a, src, dst, n_edges, image, k1, acts, slopes, normalize, s_out):
    act3 = (C.c_int32 * 3)(*acts)
    slope3 = (C.c_void_p * 3)(*[None if s is None else ptr(s) for s in slopes])
    _check(load().swe_edge_gate_tc_dec_fwd(ptr(p_src), ptr(p_dst), ptr(a), ptr(src, torch.int32), ptr(dst, torch.int32),
                                           n_edges, image.data_ptr(), k1, act3, slope3, int(normalize), ptr(s_out),
                                           _stream()), "swe_edge_gate_tc_dec_fwd")


def row_mlp_tc(desc: SweRowMlp):
    _check(load().swe_row_mlp_tc(C.byref(desc), _stream()), "swe_row_mlp_tc")


def hop_tc_image_bytes() -> int:
    return int(load().swe_hop_tc_image_bytes())


def hop_tc_pack(w, image):
    _check(load().swe_hop_tc_pack(ptr(w), image.data_ptr(), _stream()), "swe_hop_tc_pack")


def propagate_hop_tc_fwd(o_src, o_dst, s, rowptr, src, dst_lo, n_dst, w_image, with_gradient, upwind, addend, act,
                         slope, agg_out, out):
    _check(load().swe_propagate_hop_tc_fwd(ptr(o_src), ptr(o_dst), ptr(s), ptr(rowptr, torch.int32),
                                           ptr(src, torch.int32), dst_lo, n_dst, w_image.data_ptr(), int(with_gradient),
                                           int(upwind), ptr(addend), act, ptr(slope), ptr(agg_out), ptr(out),
                                           _stream()), "swe_propagate_hop_tc_fwd")


def hop_tc16_image_bytes() -> int:
    return int(load().swe_hop_tc16_image_bytes())


def hop_tc16_pack(w, wmax: float, image):
    _check(load().swe_hop_tc16_pack(ptr(w), float(wmax), image.data_ptr(), _stream()), "swe_hop_tc16_pack")


def propagate_hop_tc16_fwd(o_src, o_dst, s, rowptr, src, dst_lo, n_dst, w_image, with_gradient, upwind, addend, act,
                           slope, agg_out, out):
    _check(load().swe_propagate_hop_tc16_fwd(ptr(o_src), ptr(o_dst), ptr(s), ptr(rowptr, torch.int32), ptr(src, torch.int32),
                                             dst_lo, n_dst, w_image.data_ptr(), int(with_gradient), int(upwind), ptr(addend),
                                             act, ptr(slope), ptr(agg_out), ptr(out), _stream()),
           "swe_propagate_hop_tc16_fwd")


def propagate_hop_tc16s_fwd(o_src, o_dst, s, rowptr, src, dst_lo, n_dst, w_image, with_gradient, upwind, addend, act,
                            slope, agg_out, out):
    """s-ring edition of the fp16 hop (gate rows streamed by cp.async.bulk; same image, bit-identical results)."""
    _check(load().swe_propagate_hop_tc16s_fwd(ptr(o_src), ptr(o_dst), ptr(s), ptr(rowptr, torch.int32), ptr(src, torch.int32),
                                              dst_lo, n_dst, w_image.data_ptr(), int(with_gradient), int(upwind), ptr(addend),
                                              act, ptr(slope), ptr(agg_out), ptr(out), _stream()),
           "swe_propagate_hop_tc16s_fwd")


def pool_mean_fwd(x, rowptr, fine, coarse_lo, n_coarse, out, F):
    _check(load().swe_pool_mean_fwd(ptr(x), ptr(rowptr, torch.int32), ptr(fine, torch.int32), coarse_lo, n_coarse,
                                    ptr(out), F, _stream()), "swe_pool_mean_fwd")


def decode_head_fwd(h, act_in, slope_in, dec: SweMlp, x0, perm, n_nodes, previous_t, res_mode, res_w, eps, pred,
                    step_ptr, pred_step_stride, x_next, F):
    _check(load().swe_decode_head_fwd(ptr(h), act_in, ptr(slope_in), C.byref(dec), ptr(x0), x0.shape[1],
                                      ptr(perm, torch.int32), n_nodes, previous_t, res_mode, ptr(res_w), float(eps),
                                      ptr(pred), ptr(step_ptr, torch.int32), pred_step_stride, ptr(x_next), F,
                                      _stream()), "swe_decode_head_fwd")


def apply_bc(x, n_static_raw, previous_t, type_bc, node_bc, bc, step_ptr):
    if type_bc == 3:
        raise ValueError("Vector boundary conditions are not yet implemented.")
    if type_bc not in (1, 2):
        raise ValueError(f"BC_type={type_bc} is not a valid input. Please select either:\n"
                         "1: Inflow water depth\n2: Inflow discharge")
    _check(load().swe_apply_bc(ptr(x), x.shape[1], n_static_raw, previous_t, type_bc, ptr(node_bc, torch.int64),
                               node_bc.numel(), ptr(bc), bc.shape[-1], ptr(step_ptr, torch.int32), _stream()),
           "swe_apply_bc")


def pack_rows(src, idx, n_rows, dst):
    _check(load().swe_pack_rows(ptr(src), ptr(idx, torch.int32), n_rows, src.shape[1], ptr(dst), _stream()),
           "swe_pack_rows")


def temporal_window(x_static, wd, v, bc, init_time, previous_t, rollout_steps, x, y, bc_out):
    n, t_sim = wd.shape
    n_static = 0 if x_static is None else int(x_static.shape[1])
    n_bc, t_bc = (0, 1) if bc is None else (int(bc.shape[0]), int(bc.shape[1]))
    _check(load().swe_temporal_window(ptr(x_static), n_static, ptr(wd), ptr(v), int(n), int(t_sim), ptr(bc), n_bc, t_bc,
                                      int(init_time), int(previous_t), int(rollout_steps), ptr(x), ptr(y), ptr(bc_out), _stream()),
           "swe_temporal_window")


def rollout_metrics_cols() -> int:
    return int(load().swe_rollout_metrics_cols())


def rollout_metrics(pred, real, thr, out, ws):
    n, _, T = pred.shape
    _check(load().swe_rollout_metrics(ptr(pred), ptr(real), int(n), int(T), ptr(thr), 0 if thr is None else int(thr.numel()),
                                      out.data_ptr(), ws.data_ptr(), _stream()), "swe_rollout_metrics")


def train_step_ws_bytes() -> int:
    return int(load().swe_train_step_ws_bytes())


def loss_fwd_bwd(pred, real, real_stride, rows, n, only_where_water, mae, w0, w1, scale, accumulate, loss, dpred, ws):
    """real: tensor whose data_ptr() is element (0, 0) of the target; rows: optional uint8 / bool [n] tensor."""
    _check(load().swe_loss_fwd_bwd(ptr(pred), real.data_ptr(), int(real_stride), None if rows is None else rows.data_ptr(),
                                   int(n), int(only_where_water), int(mae), float(w0), float(w1), float(scale), int(accumulate),
                                   ptr(loss), ptr(dpred), ws.data_ptr(), _stream()), "swe_loss_fwd_bwd")


def clip_adamw_step(params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, max_norm, state, ws):
    _check(load().swe_clip_adamw_step(ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq), int(params.numel()), ptr(lr),
                                      float(beta1), float(beta2), float(eps), float(weight_decay), float(max_norm), ptr(state),
                                      ws.data_ptr(), _stream()), "swe_clip_adamw_step")


def ipc_alloc(nbytes: int):
    """(device pointer, 64-byte IPC handle) of a zero-filled cudaMalloc'ed arena."""
    out = C.c_void_p()
    handle = C.create_string_buffer(64)
    _check(load().swe_ipc_alloc(int(nbytes), C.byref(out), handle), "swe_ipc_alloc")
    return int(out.value), handle.raw


def ipc_open(handle: bytes) -> int:
    out = C.c_void_p()
    _check(load().swe_ipc_open(handle, C.byref(out)), "swe_ipc_open")
    return int(out.value)


def ipc_close(ptr_: int):
    _check(load().swe_ipc_close(ptr_), "swe_ipc_close")


def ipc_free(ptr_: int):
    _check(load().swe_ipc_free(ptr_), "swe_ipc_free")


def halo_exchange(arr, send_idx_ptrs, n_send, remote_rows, remote_flags, local_flags, seq_ptr, done_ptr, do_push=True,
                  do_wait=True):
    """All *_ptrs / remote_* / *_flags arguments: lists of raw device addresses (ints), one per neighbour."""
    n = len(n_send)
    vp = lambda xs: (C.c_void_p * max(n, 1))(*[x if x else None for x in xs])
    _check(load().swe_halo_exchange(ptr(arr), int(arr.shape[1]), n, vp(send_idx_ptrs), (C.c_int64 * max(n, 1))(*n_send),
                                    vp(remote_rows), vp(remote_flags), vp(local_flags), seq_ptr, done_ptr, int(do_push),
                                    int(do_wait), _stream()), "swe_halo_exchange")


def step_advance(step_ptr):
    _check(load().swe_step_advance(ptr(step_ptr, torch.int32), _stream()), "swe_step_advance")


# ------------------------------------------------------------------------------------------------
# training path
# ------------------------------------------------------------------------------------------------
def vptr(t: torch.Tensor, row_lo: int = 0) -> int:
    """Address of a *virtual* full-size array whose rows [row_lo, row_lo + len(t)) are backed by `t`
    (the kernels only touch the row range they are given)."""
    ptr(t)
    return t.data_ptr() - row_lo * t.shape[1] * t.element_size()


def make_rows(segs) -> SweRows:
    """segs: sequence of (tensor_or_address, idx or None, ld, width, act_code, slope tensor or None)."""
    r = SweRows()
    r.n_seg = len(segs)
    for j, (base, idx, ld, width, act, slope) in enumerate(segs):
        sg = r.seg[j]
        sg.base = base if isinstance(base, int) else ptr(base)
        sg.idx = ptr(idx, torch.int32)
        sg.slope = ptr(slope)
        sg.ld, sg.width, sg.act = int(ld), int(width), int(act)
    return r


def _addr(t):
    return t if (t is None or isinstance(t, int)) else ptr(t)


def mlp_layer_fwd(rows: SweRows, n_rows, wt, bias, n_out, pre):
    _check(load().swe_mlp_layer_fwd(C.byref(rows), n_rows, _addr(wt), _addr(bias), n_out, _addr(pre), _stream()),
           "swe_mlp_layer_fwd")


def mlp_layer_bwd_dx(dh, pre, act, slope, n_rows, n, w, w_ld, k_off, k_valid, ko, dx, accumulate, write_delta, part):
    g = C.c_int32(0)
    _check(load().swe_mlp_layer_bwd_dx(_addr(dh), _addr(pre), act, ptr(slope), n_rows, n, _addr(w), w_ld, k_off, k_valid,
                                       ko, _addr(dx), int(accumulate), int(write_delta), _addr(part), C.byref(g),
                                       _stream()), "swe_mlp_layer_bwd_dx")
    return g.value


def mlp_layer_bwd_dx_grid(n_rows) -> int:
    return int(load().swe_mlp_layer_bwd_dx_grid(n_rows))


def mlp_layer_bwd_dw(delta, n_rows, n, rows: SweRows, ko, part):
    g = C.c_int32(0)
    _check(load().swe_mlp_layer_bwd_dw(_addr(delta), n_rows, n, C.byref(rows), ko, _addr(part), C.byref(g), _stream()),
           "swe_mlp_layer_bwd_dw")
    return g.value


def mlp_layer_bwd_dw_grid(n_rows) -> int:
    return int(load().swe_mlp_layer_bwd_dw_grid(n_rows))


def mlp_layer_bwd_dx_tc(delta, n_rows, n, w, w_ld, k_off, k_valid, ko, dx0, acc0, dx1=None, acc1=False, split=None):
    _check(load().swe_mlp_layer_bwd_dx_tc(_addr(delta), n_rows, n, _addr(w), w_ld, k_off, k_valid, ko, _addr(dx0), int(acc0),
                                          _addr(dx1), int(acc1), ko if split is None else split, _stream()),
           "swe_mlp_layer_bwd_dx_tc")


def mlp_layer_bwd_dx_tc_fused(dh, pre, act, slope, n_rows, n, w, w_ld, k_off, k_valid, ko, dx0, acc0, dx1, acc1, split, part):
    g = C.c_int32(0)
    _check(load().swe_mlp_layer_bwd_dx_tc_fused(_addr(dh), _addr(pre), act, ptr(slope), n_rows, n, _addr(w), w_ld, k_off,
                                                k_valid, ko, _addr(dx0), int(acc0), _addr(dx1), int(acc1),
                                                ko if split is None else split, _addr(part), C.byref(g), _stream()),
           "swe_mlp_layer_bwd_dx_tc_fused")
    return g.value


def mlp_layer_bwd_dx_tc_grid(n_rows) -> int:
    return int(load().swe_mlp_layer_bwd_dx_tc_grid(n_rows))


def mlp_layer_bwd_dw_tc(delta, n_rows, n, rows: SweRows, part):
    g = C.c_int32(0)
    _check(load().swe_mlp_layer_bwd_dw_tc(_addr(delta), n_rows, n, C.byref(rows), _addr(part), C.byref(g), _stream()),
           "swe_mlp_layer_bwd_dw_tc")
    return g.value


def mlp_layer_bwd_dw_tc_grid(n_rows) -> int:
    return int(load().swe_mlp_layer_bwd_dw_tc_grid(n_rows))


def reduce_partials(part, n_parts, part_stride, item_off, n_items, ko, k_valid, out, ld_out, k_off):
    _check(load().swe_reduce_partials(_addr(part), n_parts, part_stride, item_off, n_items, ko, k_valid, _addr(out),
                                      ld_out, k_off, _stream()), "swe_reduce_partials")


def gate_norm_fwd(pre3, act, slope, normalize, n_edges, s_out, F):
    _check(load().swe_gate_norm_fwd(ptr(pre3), act, ptr(slope), int(normalize), n_edges, ptr(s_out), F, _stream()),
           "swe_gate_norm_fwd")


def gate_norm_bwd(ds, pre3, act, slope, normalize, n_edges, F):
    _check(load().swe_gate_norm_bwd(ptr(ds), ptr(pre3), act, ptr(slope), int(normalize), n_edges, F, _stream()),
           "swe_gate_norm_bwd")


def act_fwd(x, row_lo, n_rows, act, slope, y, F):
    _check(load().swe_act_fwd(_addr(x), row_lo, n_rows, act, ptr(slope), _addr(y), F, _stream()), "swe_act_fwd")


def act_bwd(g, x, row_lo, n_rows, act, slope, gx, slope_part, F):
    gr = C.c_int32(0)
    _check(load().swe_act_bwd(_addr(g), _addr(x), row_lo, n_rows, act, ptr(slope), _addr(gx), _addr(slope_part),
                              C.byref(gr), F, _stream()), "swe_act_bwd")
    return gr.value


def propagate_hop_train_fwd(o_src, o_dst, s, rowptr, src, dst_lo, n_dst, wt, with_gradient, upwind, addend, agg_out,
                            out, F):
    _check(load().swe_propagate_hop_train_fwd(_addr(o_src), _addr(o_dst), ptr(s), ptr(rowptr, torch.int32),
                                              ptr(src, torch.int32), dst_lo, n_dst, ptr(wt), int(with_gradient),
                                              int(upwind), _addr(addend), _addr(agg_out), _addr(out), F, _stream()),
           "swe_propagate_hop_train_fwd")


def row_flags(o, row_lo, n_rows, flags, F):
    _check(load().swe_row_flags(_addr(o), row_lo, n_rows, ptr(flags, torch.uint8), F, _stream()), "swe_row_flags")


def hop_bwd_dst(da, o_src, o_dst, s, ds, accumulate_ds, rowptr, src, wet_src, wet_dst, dst_lo, n_dst, with_gradient,
                g_next, g_part, F):
    _check(load().swe_hop_bwd_dst(_addr(da), _addr(o_src), _addr(o_dst), ptr(s), ptr(ds), int(accumulate_ds),
                                  ptr(rowptr, torch.int32), ptr(src, torch.int32), ptr(wet_src, torch.uint8),
                                  ptr(wet_dst, torch.uint8), dst_lo, n_dst, int(with_gradient), _addr(g_next),
                                  _addr(g_part), F, _stream()), "swe_hop_bwd_dst")


def hop_bwd_src(da, s, t_rowptr, t_pos, dst, wet_src, wet_dst, src_lo, n_src, with_gradient, accumulate, g_io, F):
    _check(load().swe_hop_bwd_src(_addr(da), ptr(s), ptr(t_rowptr, torch.int32), ptr(t_pos, torch.int32),
                                  ptr(dst, torch.int32), ptr(wet_src, torch.uint8), ptr(wet_dst, torch.uint8), src_lo,
                                  n_src, int(with_gradient), int(accumulate), _addr(g_io), F, _stream()),
           "swe_hop_bwd_src")


def edge_to_node_sum(e, rowptr, pos, node_lo, n_nodes, out, accumulate, F):
    _check(load().swe_edge_to_node_sum(ptr(e), ptr(rowptr, torch.int32), ptr(pos, torch.int32), node_lo, n_nodes,
                                       _addr(out), int(accumulate), F, _stream()), "swe_edge_to_node_sum")


def pool_mean_bwd(g, f_rowptr, coarse, fine_lo, n_fine, pool_rowptr, coarse_lo, dx, accumulate, F):
    _check(load().swe_pool_mean_bwd(_addr(g), ptr(f_rowptr, torch.int32), ptr(coarse, torch.int32), fine_lo, n_fine,
                                    ptr(pool_rowptr, torch.int32), coarse_lo, _addr(dx), int(accumulate), F, _stream()),
           "swe_pool_mean_bwd")


def static_inputs_fwd(x, perm, n_nodes, n_static_raw, with_wl, xin_s):
    _check(load().swe_static_inputs_fwd(ptr(x), x.shape[1], ptr(perm, torch.int32), n_nodes, n_static_raw, int(with_wl),
                                        ptr(xin_s), xin_s.shape[1], _stream()), "swe_static_inputs_fwd")


def node_inputs_bwd(dxin_s, dxin_d, n_cols, perm, n_nodes, n_dyn_rows, n_static_raw, with_wl, dx):
    _check(load().swe_node_inputs_bwd(ptr(dxin_s), dxin_s.shape[1], ptr(dxin_d), dxin_d.shape[1], n_cols,
                                      ptr(perm, torch.int32), n_nodes, n_dyn_rows, n_static_raw, int(with_wl), ptr(dx),
                                      _stream()), "swe_node_inputs_bwd")


def head_fwd(pre3, act, slope, x0, perm, n_nodes, previous_t, res_mode, res_w, eps, pred):
    _check(load().swe_head_fwd(ptr(pre3), pre3.shape[1], act, ptr(slope), ptr(x0), x0.shape[1], ptr(perm, torch.int32),
                               n_nodes, previous_t, res_mode, ptr(res_w), float(eps), ptr(pred), _stream()),
           "swe_head_fwd")


def head_bwd(dpred, pre3, act, slope, x0, perm, n_nodes, previous_t, res_mode, res_w, eps, dh3, dx0, res_part):
    g = C.c_int32(0)
    _check(load().swe_head_bwd(ptr(dpred), ptr(pre3), pre3.shape[1], act, ptr(slope), ptr(x0), x0.shape[1],
                               ptr(perm, torch.int32), n_nodes, previous_t, res_mode, ptr(res_w), float(eps), ptr(dh3),
                               ptr(dx0), ptr(res_part), C.byref(g), _stream()), "swe_head_bwd")
    return g.value
