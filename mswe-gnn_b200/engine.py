"""Launch sequencing for the models: packed parameters, workspaces and the per-call kernel
sequences.  Everything numeric happens in ``libswe_gnn_b200.so``; this module only decides which
buffer each launch reads and writes.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import lib
from .lib import ACT_CODES, SweLayer, SweMlp
from .models.models import activation_name_of
from .plan import EdgeSet, GraphPlan

SUPPORTED_F = (16, 32, 64)


def padded_width(f: int) -> int:
    for w in SUPPORTED_F:
        if f <= w:
            return w
    raise NotImplementedError(f"hid_features={f} is wider than the largest kernel instantiation ({SUPPORTED_F[-1]})")


def _round4(k: int) -> int:
    return (k + 3) // 4 * 4


class PackedMLP:
    """Kernel-side image of a ``make_mlp`` Sequential: transposed (k-major) weights with the input
    and output widths zero-padded to what the kernels expect, biases copied next to them, PReLU
    slopes referenced in place.  Re-packed automatically when a parameter changed."""

    def __init__(self, seq: nn.Sequential, in_blocks: Sequence[Tuple[int, int]], hidden_pad: Dict[int, int],
                 first_k_pad: Optional[int] = None):
        """in_blocks: (true width, padded width) of every column block of the first layer's
        input; hidden_pad: true output width -> padded output width."""
        self.linears: List[nn.Linear] = []
        self.acts: List[Optional[nn.Module]] = []
        mods = list(seq)
        i = 0
        while i < len(mods):
            assert isinstance(mods[i], nn.Linear), "MLP layout must be (Linear, activation)*"
            self.linears.append(mods[i])
            if i + 1 < len(mods) and not isinstance(mods[i + 1], nn.Linear):
                self.acts.append(mods[i + 1])
                i += 2
            else:
                self.acts.append(None)
                i += 1
        if len(self.linears) > lib.SWE_MAX_LAYERS:
            raise NotImplementedError(f"MLPs deeper than {lib.SWE_MAX_LAYERS} layers are not supported")
        self.in_blocks = list(in_blocks)
        self.hidden_pad = dict(hidden_pad)
        self.first_k_pad = first_k_pad
        self._stamp = None
        self._buf = None
        self._struct = SweMlp()

    def _out_pad(self, n: int) -> int:
        return self.hidden_pad.get(n, n)

    def struct(self) -> SweMlp:
        stamp = tuple((p.data_ptr(), p._version) for lin in self.linears for p in lin.parameters()) + \
            tuple(a.weight.data_ptr() for a in self.acts if isinstance(a, nn.PReLU))
        if stamp == self._stamp:
            return self._struct
        dev = self.linears[0].weight.device
        # layout of the flat buffer
        shapes = []
        for li, lin in enumerate(self.linears):
            n_out, k_in = lin.weight.shape
            if li == 0:
                assert sum(b[0] for b in self.in_blocks) == k_in, "input blocks do not match the first layer"
                k_pad = sum(b[1] for b in self.in_blocks)
                if self.first_k_pad:
                    k_pad = max(k_pad, self.first_k_pad)
                k_pad = _round4(k_pad)
            else:
                k_pad = _round4(self._out_pad(k_in))
            shapes.append((k_pad, self._out_pad(n_out)))
        total = sum(k * n + _round4(n) for k, n in shapes)
        if self._buf is None or self._buf.numel() != total or self._buf.device != dev:
            self._buf = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        st = self._struct
        st.n_layers = len(self.linears)
        with torch.no_grad():
            for li, (lin, act) in enumerate(zip(self.linears, self.acts)):
                k_pad, n_pad = shapes[li]
                n_out, k_in = lin.weight.shape
                w = lin.weight.detach()
                if w.dtype != torch.float32:
                    raise TypeError("kernels are fp32; cast the model with .float()")
                need_pad = (n_pad != n_out) or (li == 0 and any(a != b for a, b in self.in_blocks)) or \
                    (li > 0 and self._out_pad(k_in) != k_in)
                if need_pad:
                    wp = torch.zeros(n_pad, k_pad, dtype=torch.float32, device=dev)
                    if li == 0:
                        c_src = c_dst = 0
                        for true_w, pad_w in self.in_blocks:
                            wp[:n_out, c_dst:c_dst + true_w] = w[:, c_src:c_src + true_w]
                            c_src += true_w
                            c_dst += pad_w
                    else:
                        wp[:n_out, :k_in] = w
                    w_src = wp
                else:
                    w_src = w.contiguous()
                wt = self._buf[off:off + k_pad * n_pad]
                lib.pack_linear(w_src, k_pad, wt.view(k_pad, n_pad))
                L: SweLayer = st.layer[li]
                L.wt = wt.data_ptr()
                off += k_pad * n_pad
                if lin.bias is not None:
                    b = self._buf[off:off + n_pad]
                    b.zero_()
                    b[:n_out] = lin.bias.detach()
                    L.bias = b.data_ptr()
                else:
                    L.bias = None
                off += _round4(n_pad)
                name = activation_name_of(act)
                L.act = ACT_CODES[name]
                L.slope = act.weight.data_ptr() if isinstance(act, nn.PReLU) else None
                if isinstance(act, nn.PReLU) and act.weight.numel() != 1:
                    raise NotImplementedError("per-channel PReLU is not supported")
                L.k_in, L.n_out = k_pad, n_pad
        self._stamp = stamp
        return st


class PackedFilters:
    """Packed (transposed, padded) filter matrices W_0..W_K of one SWEGNN (gnn.py:381-384)."""

    def __init__(self, linears: Sequence[nn.Linear], F: int, FP: int):
        self.linears = list(linears)
        self.F, self.FP = F, FP
        self._stamp = None
        self._buf = None
        self._tc_stamp = None
        self._tc_img = None

    def tc_images(self) -> List[torch.Tensor]:
        """Pre-swizzled hi|lo TF32 images of W_0..W_K for the tcgen05 hop kernel (F = 64 only)."""
        stamp = tuple((l.weight.data_ptr(), l.weight._version) for l in self.linears)
        if stamp != self._tc_stamp:
            dev = self.linears[0].weight.device
            nb = lib.hop_tc_image_bytes()
            if self._tc_img is None or self._tc_img.device != dev:
                self._tc_img = torch.empty(len(self.linears), nb, dtype=torch.uint8, device=dev)
            with torch.no_grad():
                for i, lin in enumerate(self.linears):
                    lib.hop_tc_pack(lin.weight.detach().contiguous(), self._tc_img[i])
            self._tc_stamp = stamp
        return [self._tc_img[i] for i in range(len(self.linears))]

    def tc16_images(self) -> List[torch.Tensor]:
        """fp16 hi|lo images (64-byte swizzle) of W_0..W_K for the kind::f16 hop kernel; the per-matrix power-of-two
        scale comes from max |w|, read on the host when the weights change (never inside a captured step)."""
        stamp = tuple((l.weight.data_ptr(), l.weight._version) for l in self.linears)
        if stamp != getattr(self, "_tc16_stamp", None):
            dev = self.linears[0].weight.device
            nb = lib.hop_tc16_image_bytes()
            if getattr(self, "_tc16_img", None) is None or self._tc16_img.device != dev:
                self._tc16_img = torch.empty(len(self.linears), (nb + 255) // 256 * 256, dtype=torch.uint8, device=dev)
            with torch.no_grad():
                wmax = torch.stack([l.weight.detach().abs().max() for l in self.linears]).tolist()
                for i, lin in enumerate(self.linears):
                    lib.hop_tc16_pack(lin.weight.detach().contiguous(), wmax[i], self._tc16_img[i])
            self._tc16_stamp = stamp
        return [self._tc16_img[i] for i in range(len(self.linears))]

    def tensors(self) -> List[torch.Tensor]:
        stamp = tuple((l.weight.data_ptr(), l.weight._version) for l in self.linears)
        if stamp != self._stamp:
            dev = self.linears[0].weight.device
            if self._buf is None or self._buf.device != dev:
                self._buf = torch.zeros(len(self.linears), self.FP, self.FP, dtype=torch.float32, device=dev)
            with torch.no_grad():
                for i, lin in enumerate(self.linears):
                    w = lin.weight.detach()
                    if self.FP != self.F:
                        wp = torch.zeros(self.FP, self.FP, dtype=torch.float32, device=dev)
                        wp[:self.F, :self.F] = w
                        w = wp
                    lib.pack_linear(w.contiguous(), self.FP, self._buf[i])
            self._stamp = stamp
        return [self._buf[i] for i in range(len(self.linears))]


class PackedGateTC:
    """Weight image of a 3-layer edge MLP for the tcgen05 gate kernel (swe_gate_tc_pack)."""

    def __init__(self, seq: nn.Sequential):
        self.linears = [m for m in seq if isinstance(m, nn.Linear)]
        self.acts = [m for m in seq if not isinstance(m, nn.Linear)]
        self._stamp = None
        self._img = None

    @staticmethod
    def eligible(seq: nn.Sequential, F: int) -> bool:
        lins = [m for m in seq if isinstance(m, nn.Linear)]
        acts = [m for m in seq if not isinstance(m, nn.Linear)]
        if F != 64 or len(lins) != 3 or len(acts) != 3:
            return False
        shapes = [tuple(l.weight.shape) for l in lins]
        return shapes[0][0] == 128 and shapes[0][1] in (256, 320) and shapes[1] == (128, 128) and shapes[2] == (64, 128)

    def image(self):
        stamp = tuple((p.data_ptr(), p._version) for l in self.linears for p in l.parameters())
        if stamp != self._stamp:
            l1, l2, l3 = self.linears
            k1 = l1.weight.shape[1]
            dev = l1.weight.device
            n = lib.gate_tc_image_bytes(k1)
            if self._img is None or self._img.numel() != n or self._img.device != dev:
                self._img = torch.empty(n, dtype=torch.uint8, device=dev)
            with torch.no_grad():
                lib.gate_tc_pack(l1.weight.detach().contiguous(), None if l1.bias is None else l1.bias.detach(),
                                 l2.weight.detach().contiguous(), None if l2.bias is None else l2.bias.detach(),
                                 l3.weight.detach().contiguous(), None if l3.bias is None else l3.bias.detach(), self._img)
            self._stamp = stamp
        return self._img

    def acts_and_slopes(self):
        codes = [ACT_CODES[activation_name_of(a)] for a in self.acts]
        slopes = [a.weight if isinstance(a, nn.PReLU) else None for a in self.acts]
        return codes, slopes

    def image16(self):
        """fp16 hi|lo image for the kind::f16 gate kernel (swe_gate_tc16_pack); the per-matrix power-of-two scales come
        from max |w|, read on the host when the weights change (never inside a captured step)."""
        stamp = tuple((p.data_ptr(), p._version) for l in self.linears for p in l.parameters())
        if stamp != getattr(self, "_stamp16", None):
            l1, l2, l3 = self.linears
            k1 = l1.weight.shape[1]
            dev = l1.weight.device
            n = lib.gate_tc16_image_bytes(k1)
            if getattr(self, "_img16", None) is None or self._img16.numel() != n or self._img16.device != dev:
                self._img16 = torch.empty(n, dtype=torch.uint8, device=dev)
            with torch.no_grad():
                wmax = torch.stack([l.weight.detach().abs().max() for l in self.linears]).tolist()
                lib.gate_tc16_pack(l1.weight.detach().contiguous(), None if l1.bias is None else l1.bias.detach(),
                                   l2.weight.detach().contiguous(), None if l2.bias is None else l2.bias.detach(),
                                   l3.weight.detach().contiguous(), None if l3.bias is None else l3.bias.detach(),
                                   wmax, self._img16)
            self._stamp16 = stamp
        return self._img16

    def flag_ws(self, n_edges: int):
        """Scratch of the fp16 gate's range guard: tile count + one slot per 128-edge tile."""
        need = (n_edges + 127) // 128 + 1
        ws = getattr(self, "_flag_ws", None)
        dev = self.linears[0].weight.device
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = torch.zeros(need, dtype=torch.int32, device=dev)
            self._flag_ws = ws
        return ws


def rowmlp_backend() -> str:
    """'tc' (tcgen05 row MLPs for encoders / W0 / decoder head, default for F = 64) or 'ffma'."""
    import os
    return os.environ.get("MSWE_ROWMLP", "tc")


def rowmlp16_backend() -> str:
    """Which two-layer row MLPs run on the fp16 streaming kernel swe_row_mlp_tc16: '1' all (default), 'enc' the encoders
    only, 'dec' the decoder only, '0' none (swe_row_mlp_tc)."""
    import os
    return os.environ.get("MSWE_ROWMLP16", "1")


def rowlin_backend() -> str:
    """'tc16' (default): o_0 = x_d W_0ᵀ by the streaming kernel swe_row_linear_tc16; 'tc': by swe_row_mlp_tc (3xTF32)."""
    import os
    return os.environ.get("MSWE_ROWLIN", "tc16")


class RowMlpTC:
    """tcgen05 image of a ``make_mlp`` stack for ``swe_row_mlp_tc``.

    kind='encoder': Linear(k<=8 -> 64) on CUDA cores + two 64->64 tensor-core layers;
    kind='decoder': two 64->64 tensor-core layers + the 64->2 head;
    kind='linear' : a single bias-free 64->64 layer (filter_matrix[0])."""

    def __init__(self, linears, acts, kind: str):
        self.linears, self.acts, self.kind = list(linears), list(acts), kind
        self._stamp = None
        self._img = None

    @staticmethod
    def split(seq: nn.Sequential):
        lins = [m for m in seq if isinstance(m, nn.Linear)]
        acts, mods = [], list(seq)
        for i, m in enumerate(mods):
            if isinstance(m, nn.Linear):
                acts.append(mods[i + 1] if i + 1 < len(mods) and not isinstance(mods[i + 1], nn.Linear) else None)
        return lins, acts

    @classmethod
    def for_encoder(cls, seq: nn.Sequential, F: int, max_in: int = 8):
        lins, acts = cls.split(seq)
        ok = F == 64 and len(lins) == 3 and lins[0].in_features <= max_in and \
            [tuple(l.weight.shape) for l in lins[1:]] == [(64, 64), (64, 64)] and lins[0].out_features == 64 and \
            all(not (isinstance(a, nn.PReLU) and a.weight.numel() != 1) for a in acts)
        return cls(lins, acts, "encoder") if ok else None

    @classmethod
    def for_decoder(cls, seq: nn.Sequential, F: int):
        lins, acts = cls.split(seq)
        ok = F == 64 and len(lins) == 3 and [tuple(l.weight.shape) for l in lins] == [(64, 64), (64, 64), (2, 64)] and \
            all(not (isinstance(a, nn.PReLU) and a.weight.numel() != 1) for a in acts)
        return cls(lins, acts, "decoder") if ok else None

    def tc_layers(self):
        return {"encoder": self.linears[1:], "decoder": self.linears[:2], "linear": self.linears}[self.kind]

    def images(self):
        tl = self.tc_layers()
        stamp = tuple((l.weight.data_ptr(), l.weight._version) for l in tl)
        if stamp != self._stamp:
            dev = tl[0].weight.device
            nb = lib.hop_tc_image_bytes()
            if self._img is None or self._img.device != dev:
                self._img = torch.empty(len(tl), nb, dtype=torch.uint8, device=dev)
            with torch.no_grad():
                for i, l in enumerate(tl):
                    lib.hop_tc_pack(l.weight.detach().contiguous(), self._img[i])
            self._stamp = stamp
        return [self._img[i] for i in range(len(tl))]

    def _fill_tc(self, d: "lib.SweRowMlp", layer_ids):
        imgs = self.images()
        d.n_tc = len(layer_ids)
        for j, li in enumerate(layer_ids):
            lin, act = self.linears[li], self.acts[li]
            d.img[j] = imgs[j].data_ptr()
            d.bias[j] = None if lin.bias is None else lin.bias.data_ptr()
            d.act[j] = ACT_CODES[activation_name_of(act)]
            d.slope[j] = act.weight.data_ptr() if isinstance(act, nn.PReLU) else None

    def encode(self, raw, raw_col0, raw_cols, with_wl, wl_cols, perm, row_lo, n_rows, out_rows):
        d = lib.SweRowMlp()
        lin0, act0 = self.linears[0], self.acts[0]
        d.raw, d.raw_ld, d.raw_col0, d.raw_cols = raw.data_ptr(), raw.shape[1], raw_col0, raw_cols
        d.with_wl, d.wl_col_a, d.wl_col_b = int(with_wl), wl_cols[0], wl_cols[1]
        d.perm = None if perm is None else lib.ptr(perm, torch.int32)
        d.w_first = lib.ptr(lin0.weight.detach().contiguous())
        d.b_first = None if lin0.bias is None else lin0.bias.data_ptr()
        d.act_first = ACT_CODES[activation_name_of(act0)]
        d.slope_first = act0.weight.data_ptr() if isinstance(act0, nn.PReLU) else None
        d.row_lo, d.n_rows = row_lo, n_rows
        self._fill_tc(d, [1, 2])
        d.out_rows = lib.ptr(out_rows)
        self._launch(d)

    def images16(self):
        """fp16 hi/lo images (swe_hop_tc16_pack) of the tensor-core layers for the streaming kernel swe_row_mlp_tc16."""
        tl = self.tc_layers()
        stamp = tuple((l.weight.data_ptr(), l.weight._version) for l in tl)
        if stamp != getattr(self, "_stamp16s", None):
            dev = tl[0].weight.device
            if getattr(self, "_img16s", None) is None or self._img16s.device != dev:
                self._img16s = torch.empty(len(tl), lib.hop_tc16_image_bytes(), dtype=torch.uint8, device=dev)
            with torch.no_grad():
                for i, l in enumerate(tl):
                    wd = l.weight.detach().contiguous()
                    lib.hop_tc16_pack(wd, float(wd.abs().max()), self._img16s[i])
            self._stamp16s = stamp
        return [self._img16s[i] for i in range(len(tl))]

    def _launch(self, d):
        """Two-layer stacks go to the fp16 streaming kernel where it covers the shape (MSWE_ROWMLP16=0: never)."""
        mode = rowmlp16_backend()
        if d.n_tc == 2 and (mode == "1" or (mode == "enc" and not d.head) or (mode == "dec" and d.head)) \
                and lib.row_mlp_tc16(d, self.images16()):
            return
        lib.row_mlp_tc(d)

    def image16(self):
        """kind='linear': the fp16 hi/lo image of the single layer for the streaming kernel (swe_row_linear_tc16)."""
        w = self.linears[0].weight
        stamp = (w.data_ptr(), w._version)
        if stamp != getattr(self, "_stamp16", None):
            if getattr(self, "_img16", None) is None or self._img16.device != w.device:
                self._img16 = torch.empty(lib.hop_tc16_image_bytes(), dtype=torch.uint8, device=w.device)
            with torch.no_grad():
                wd = w.detach().contiguous()
                lib.hop_tc16_pack(wd, float(wd.abs().max()), self._img16)
            self._stamp16 = stamp
        return self._img16

    def linear(self, x_rows, row_lo, n_rows, out_rows):
        if (self.kind == "linear" and self.linears[0].bias is None and self.acts[0] is None and rowlin_backend() == "tc16"
                and x_rows.shape[1] == 64 and out_rows.shape[1] == 64):
            lib.row_linear_tc16(x_rows, row_lo, n_rows, self.image16(), out_rows)
            return
        d = lib.SweRowMlp()
        d.x_rows, d.act_in = lib.ptr(x_rows), 0
        d.row_lo, d.n_rows = row_lo, n_rows
        self._fill_tc(d, [0])
        d.out_rows = lib.ptr(out_rows)
        lib.row_mlp_tc(d)

    def decode(self, h, act_in, slope_in, x0, perm, n_nodes, previous_t, res_mode, res_w, eps, pred, step_ptr, pred_stride,
               x_next, row_lo: int = 0):
        """Rows [row_lo, row_lo + n_nodes) (a rank of a partitioned mesh decodes its owned rows only, scale by scale)."""
        d = lib.SweRowMlp()
        d.x_rows, d.act_in = lib.ptr(h), act_in
        d.slope_in = None if slope_in is None else slope_in.data_ptr()
        d.row_lo, d.n_rows = row_lo, n_nodes
        self._fill_tc(d, [0, 1])
        lin, act = self.linears[2], self.acts[2]
        d.head = 1
        d.w_head = lib.ptr(lin.weight.detach().contiguous())
        d.b_head = None if lin.bias is None else lin.bias.data_ptr()
        d.act_head = ACT_CODES[activation_name_of(act)]
        d.slope_head = act.weight.data_ptr() if isinstance(act, nn.PReLU) else None
        d.x0, d.n_cols, d.previous_t = lib.ptr(x0), x0.shape[1], previous_t
        d.head_perm = None if perm is None else lib.ptr(perm, torch.int32)
        d.res_mode, d.eps = res_mode, float(eps)
        d.res_w = None if res_w is None else lib.ptr(res_w)
        d.pred = lib.ptr(pred)
        d.step_ptr = None if step_ptr is None else lib.ptr(step_ptr, torch.int32)
        d.pred_step_stride = pred_stride
        d.x_next = None if x_next is None else lib.ptr(x_next)
        self._launch(d)


def hop_backend() -> str:
    """'tc16s' (default for F = 64: fp16 hi/lo filter on tcgen05, gate rows streamed into per-warp shared-memory buffers by
    cp.async.bulk — 0.301 ms at cfg3 level 0 = 0.67 of the copy peak), 'tc' (3xTF32 filter, per-thread loads: 0.346 ms),
    'tc16' (fp16 hi/lo filter, per-thread loads: 0.411 ms) or 'ffma' (exact-fp32 CUDA cores).  'tc16s' stages at most 12
    edges per 4 nodes and 32 per 8; blocks beyond that (hubs) run one edge at a time, so graphs whose in-degrees exceed
    3 as a rule are better served by 'tc'."""
    import os
    return os.environ.get("MSWE_HOP", "tc16s")


_STATIC_TOKEN = None
_XS_STATIC = False
_TOKEN_COUNTER = [0]
_PSTAT_BYTES = [0]                 # bytes held by all hoisted static-partial tables of this process


def new_static_token() -> int:
    """A fresh identity for 'the static inputs of this rollout' (never reused inside the process)."""
    _TOKEN_COUNTER[0] += 1
    return _TOKEN_COUNTER[0]


class static_inputs:
    """Context manager a rollout loop puts around its steps: inside it the edge features, the static input columns and
    the weights are the same on every call, identified by `token` (a new token = new inputs).  `xs_static` says
    whether the ENCODED static node features x_s are constant too: true for models built with with_WL=False; with
    with_WL=True (config.yaml) x_s is encoded from the current water level and changes every step.
    The tcgen05 gate then evaluates the share of the first edge-MLP layer that belongs to the constant blocks once
    per token (``gate_layer0() == 'static'``) instead of once per step."""

    def __init__(self, token, xs_static: bool = False):
        self.token, self.xs_static = token, bool(xs_static)

    def __enter__(self):
        global _STATIC_TOKEN, _XS_STATIC
        self._old = (_STATIC_TOKEN, _XS_STATIC)
        _STATIC_TOKEN, _XS_STATIC = self.token, self.xs_static

    def __exit__(self, *exc):
        global _STATIC_TOKEN, _XS_STATIC
        _STATIC_TOKEN, _XS_STATIC = self._old
        return False


def gate_layer0() -> str:
    """'static' (default): inside a rollout (``static_inputs``) of a model built with with_WL=False the share of layer 0
    that belongs to the constant input blocks (x_s[r], x_s[c], a_e) is hoisted into a per-edge table computed once;
    the per-step gate multiplies only the x_d blocks (144 instead of 216 MMAs per 128 edges: 1.66 vs 2.10 ms per 3 M
    edges).  With with_WL=True (config.yaml) x_s is encoded from the current water level, nothing worth hoisting is
    constant, and this is 'full' — as it is outside a rollout.
    'full': every edge multiplies its whole 5F-wide input on the tensor core; 'dec': layer 0 is
    decomposed into per-node partial tables (swe_gate_partials_tc) + the edge part.  Measured on cfg3 (r01d): the
    decomposed gate itself is 11 % faster (5.82 vs 6.52 ms/step) but the tables cost 2.70 ms/step, because the
    kernel is bound by its per-tile epilogue chain, not by the layer-0 MMAs — so 'full' stays the default."""
    import os
    return os.environ.get("MSWE_GATE_L0", "static")


def gate_backend() -> str:
    """'tc16' (tcgen05 with fp16 hi/lo splits, kind::f16, default where eligible), 'tc' (tcgen05 3xTF32) or 'ffma'
    (exact-fp32 CUDA cores)."""
    import os
    return os.environ.get("MSWE_GATE", "tc16")


class SweGnnLauncher:
    """Kernel sequence of one ``SWEGNN.forward`` call (reference models/gnn.py:387-445) on a
    destination-CSR edge set: gate once, W0, then K hops ping-ponging between two buffers."""

    def __init__(self, module, F: int):
        self.m = module
        self.F = F
        self.FP = padded_width(F)
        nseg = 5 if module.edge_features > 0 else 4
        if module.edge_features not in (0, F):
            raise NotImplementedError("SWEGNN edge_features must be 0 or equal to the node width "
                                      f"(got {module.edge_features} vs {F})")
        two = {F: self.FP, 2 * F: 2 * self.FP}
        self.mlp = PackedMLP(module.edge_mlp, [(F, self.FP)] * nseg, two)
        self.filters = PackedFilters(list(module.filter_matrix), F, self.FP) if module.with_filter_matrix else None
        self.tc = PackedGateTC(module.edge_mlp) if PackedGateTC.eligible(module.edge_mlp, F) else None
        self._pstat = {}                           # id(edge set) -> (stamp, edge set, per-edge static partial table)
        self.w0_tc = RowMlpTC([module.filter_matrix[0]], [None], "linear") if (module.with_filter_matrix and F == 64) else None

    def _static_partials(self, es, xs, a, img, k1):
        """Per-edge table of the static share of layer 0 for the current ``static_inputs`` token ([E, 128] fp32),
        or None when it does not fit comfortably in device memory (the full gate is used then)."""
        if es.n_edges == 0:
            return None
        stamp = (_STATIC_TOKEN, self.tc._stamp, 0 if xs is None else xs.data_ptr(), 0 if a is None else a.data_ptr())
        dev = (xs if xs is not None else a).device
        ent = self._pstat.get(id(es))
        if ent is not None and ent[0] == stamp and ent[1] is es:
            return ent[2]
        tab = ent[2] if (ent is not None and ent[1] is es) else None
        if tab is None:
            rows = (es.n_edges + 127) // 128 * 128         # whole 128-edge tiles (internal tile-transposed order)
            need = rows * 128 * 4
            free, total = torch.cuda.mem_get_info(dev)
            if need > 0.25 * free or _PSTAT_BYTES[0] + need > 0.35 * total:
                return None
            tab = torch.empty(rows, 128, dtype=torch.float32, device=dev)
            _PSTAT_BYTES[0] += need
        lib.gate_static_partials_tc(xs, a, es.src, es.dst, es.n_edges, img, k1, tab)
        if len(self._pstat) >= 4 and id(es) not in self._pstat:
            old = self._pstat.pop(next(iter(self._pstat)))
            _PSTAT_BYTES[0] -= old[2].numel() * 4
        self._pstat[id(es)] = (stamp, es, tab)
        return tab

    def gate(self, es, xs, xd_src, xd_dst, a, s_buf, dbg=None, ptab=None):
        m = self.m
        backend = gate_backend()
        if self.tc is not None and backend == "tc16" and gate_layer0() != "dec":
            codes, slopes = self.tc.acts_and_slopes()
            k1 = self.tc.linears[0].weight.shape[1]
            lib.edge_gate_tc16_fwd(xs, xd_src, xd_dst, a, es.src, es.dst, es.n_edges, self.tc.image16(), self.tc.image(),
                                   k1, codes, slopes, m.normalize, s_buf, dbg, self.tc.flag_ws(es.n_edges))
            return
        if self.tc is not None and backend in ("tc", "tc16"):
            codes, slopes = self.tc.acts_and_slopes()
            k1 = self.tc.linears[0].weight.shape[1]
            img = self.tc.image()
            mode = gate_layer0()
            # (hoisting a_e alone — all that is constant when x_s carries the water level — was measured on cfg3:
            #  2.13 vs 2.10 ms per 3 M edges, the table read costs what the 24 saved MMAs give; not taken)
            if mode == "static" and dbg is None and _STATIC_TOKEN is not None and _XS_STATIC:
                tab = self._static_partials(es, xs if _XS_STATIC else None, a, img, k1)
                if tab is not None:
                    lib.edge_gate_tc_stat_fwd(tab, None if _XS_STATIC else xs, xd_src, xd_dst, es.src, es.dst, es.n_edges,
                                              img, k1, codes, slopes, m.normalize, s_buf)
                    return
            if ptab is not None and dbg is None and mode == "dec":
                p_src, p_dst = ptab
                lib.gate_partials_tc(xs, xd_src, es.src_lo, es.src_hi - es.src_lo, img, k1, 0, p_src)
                lib.gate_partials_tc(xs, xd_dst, es.dst_lo, es.n_dst, img, k1, 1, p_dst)
                lib.edge_gate_tc_dec_fwd(p_src, p_dst, a, es.src, es.dst, es.n_edges, img, k1, codes, slopes,
                                         m.normalize, s_buf)
            else:
                lib.edge_gate_tc_fwd(xs, xd_src, xd_dst, a, es.src, es.dst, es.n_edges, img, k1, codes, slopes,
                                     m.normalize, s_buf, dbg)
        else:
            lib.edge_gate_fwd(xs, xd_src, xd_dst, a, es.src, es.dst, es.n_edges, self.mlp.struct(), m.normalize,
                              s_buf, self.FP)

    def run(self, es: EdgeSet, xs, xd_src, xd_dst, a, s_buf, o_dst_rows_zero: bool, tmp_a, tmp_b, out,
            addend=None, act_code: int = 0, act_slope=None, halo=None, scale: int = 0, ptab=None):
        """Writes rows [es.dst_lo, es.dst_lo+es.n_dst) of `out`.

        xd_src: array holding x_d[row]; xd_dst: array holding x_d[col] or None when those rows are
        zero (then the hop also treats o[col] as zero).  tmp_a / tmp_b: scratch [N, FP] arrays.
        halo: optional ``parallel.HaloExchanger`` (partitioned meshes): the halo rows of scale `scale`
        are refreshed from their owners after every hop but the last.
        """
        m, FP = self.m, self.FP
        E = es.n_edges
        self.gate(es, xs, xd_src, xd_dst, a, s_buf, ptab=ptab)
        K = m.K
        if m.with_filter_matrix:
            W = self.filters.tensors()
            # o_0 = x_d W0ᵀ on the destination rows and on every source row read by the hops
            def w0(x, lo, n):
                if self.w0_tc is not None and rowmlp_backend() == "tc":
                    self.w0_tc.linear(x, lo, n, tmp_a)
                else:
                    lib.node_linear_fwd(x, lo, n, W[0], tmp_a, FP)
            if xd_dst is not None:
                if es.src_lo == es.dst_lo:
                    # same node set: one launch over the sources (on a rank of a partitioned mesh they extend beyond the
                    # owned destination rows by the halo rows the hops read)
                    w0(xd_dst, es.dst_lo, max(es.n_dst, es.src_hi - es.src_lo))
                else:
                    w0(xd_dst, es.dst_lo, es.n_dst)
                    w0(xd_src, es.src_lo, es.src_hi - es.src_lo)
                o_src, o_dst = tmp_a, tmp_a
            else:
                w0(xd_src, es.src_lo, es.src_hi - es.src_lo)
                o_src, o_dst = tmp_a, None
        else:
            W = [None] * (K + 1)
            o_src, o_dst = xd_src, xd_dst
        if K == 0:
            raise NotImplementedError("SWEGNN with K=0 hops")
        if K > 1 and es.src_lo != es.dst_lo:
            raise NotImplementedError("multi-hop propagation needs source and destination in the same node set")
        bufs = [tmp_b, tmp_a] if o_src is tmp_a else [tmp_a, tmp_b]
        hb = hop_backend()
        if hb == "tc16s" and es.max_block4 > 12:
            hb = "tc"                  # in-degrees beyond the s-ring kernel's staging (not a dual mesh): per-thread loads
        use_tc = m.with_filter_matrix and self.F == 64 and FP == 64 and hb in ("tc", "tc16", "tc16s")
        Wtc = (self.filters.tc16_images() if hb in ("tc16", "tc16s") else self.filters.tc_images()) if use_tc else None
        hop_tc = {"tc16": lib.propagate_hop_tc16_fwd, "tc16s": lib.propagate_hop_tc16s_fwd}.get(hb, lib.propagate_hop_tc_fwd)
        for k in range(K):
            last = k == K - 1
            dst_buf = out if last else bufs[k % 2]
            if use_tc:
                hop_tc(o_src, o_dst, s_buf, es.rowptr, es.src, es.dst_lo, es.n_dst, Wtc[k + 1],
                       m.with_gradient, m.upwind_mode, addend if last else None,
                       act_code if last else 0, act_slope if last else None, None, dst_buf)
            else:
                lib.propagate_hop_fwd(o_src, o_dst, s_buf, es.rowptr, es.src, es.dst_lo, es.n_dst, W[k + 1],
                                      m.with_gradient, m.upwind_mode, addend if last else None,
                                      act_code if last else 0, act_slope if last else None, dst_buf, FP)
            if halo is not None and not last:
                halo.exchange(dst_buf, scale)
            o_src = o_dst = dst_buf
