"""Training path: forward that keeps pre-activations + the hand-derived backward, wired into
``torch.autograd`` so ``loss.backward()`` works on ``GNN`` / ``MSGNN`` / ``SWEGNN`` exactly like on
the reference modules (``training/train.py:125-145``).

torch.autograd is only the tape *between* model calls (loss, window shift, BPTT over rollout
steps).  Inside one model call everything — forward and backward — is a fixed sequence of
kernels from ``libswe_gnn_b200.so`` (``csrc/swe_backward.cu``); the math is SURVEY.md Appendix B:

* hop k:  ``da = g W``, ``dW += gᵀ agg``, ``ds_e += act·da[c]⊙(o[c]−o[r])``,
  ``g ← g + da⊙Σ_in act·s − Σ_out act·s⊙da[dst]`` (the second sum runs over the transposed CSR —
  no atomics); ``act`` is the reference's per-hop wet-edge mask (``gnn.py:408-411``), kept so
  that input gradients equal autograd's;
* gate:  ``du = (ds − s(s·ds))/‖u‖`` then the edge-MLP layers backwards; weight gradients are
  per-CTA partial sums reduced in CTA order (deterministic);
* pooling, un-pooling, encoders, decoder head: the same building blocks.

Saved per SWEGNN call: ``s``, the MLP pre-activations, ``o_k`` and ``agg_k`` of every hop.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import lib
from .lib import ACT_CODES
from .models.models import activation_name_of

TILE_WIDTHS = (16, 32, 64, 128)


def _tile_width(n: int) -> int:
    for w in TILE_WIDTHS:
        if n <= w:
            return w
    raise NotImplementedError(f"layer width {n} exceeds the largest kernel tile (128)")


_TC_ACTS = (ACT_CODES[None], ACT_CODES["prelu"], ACT_CODES["relu"], ACT_CODES["leakyrelu"])


def train_gemm_backend() -> str:
    """'tc' (tcgen05 3xTF32 backward GEMMs for the wide edge-MLP layers, default) or 'ffma'."""
    import os
    return os.environ.get("MSWE_TRAIN_GEMM", "tc")


def _fuse_delta() -> bool:
    """MSWE_TRAIN_FUSE_DELTA=0: form delta = dh ⊙ act'(pre) in a separate element-wise pass instead of inside dx_tc."""
    import os
    return os.environ.get("MSWE_TRAIN_FUSE_DELTA", "1") != "0"


def _min_rows() -> int:
    """Row count below which the training GEMMs stay on CUDA cores (the tensor-core kernels' fixed cost — a weight
    image per CTA, one CTA per 128 rows — is not paid back on the small coarse levels); MSWE_TC_MIN_ROWS overrides."""
    import os
    return int(os.environ.get("MSWE_TC_MIN_ROWS", "16384"))



def _tc_part(name: str) -> bool:
    """MSWE_TRAIN_TC_PARTS selects which training GEMMs of the wide edge MLP use the tensor cores (3xTF32).

    Default ``fwd,dx,dw``: the forward (the inference gate kernel in the variant that stores its pre-activations)
    and both backward GEMMs.  The forward's ~3e-6 error would flip the PReLU/ReLU derivative mask of the few
    pre-activations that close to the kink — an O(1) change of one summand each, ~1e-3 relative L2 in the weight
    gradients — so those entries are listed by the kernel and re-evaluated in exact fp32 (``swe_gate_fix_preacts``,
    ``MSWE_TRAIN_FIX_TAU``); with that the gradients match the fp32 oracle to ~1e-5 like the all-CUDA-core path."""
    import os
    return train_gemm_backend() == "tc" and name in os.environ.get("MSWE_TRAIN_TC_PARTS", "fwd,dx,dw").split(",")


_FIX_BUF = {}


def _fix_tau() -> float:
    """|pre| below tau * max(1, row max) is re-evaluated in exact fp32 after the tensor-core forward (0 = off)."""
    import os
    return float(os.environ.get("MSWE_TRAIN_FIX_TAU", "1e-4"))


def _fix_buffers(n_edges: int, dev):
    """Work lists of swe_edge_gate_tc_train_fwd / swe_gate_fix_preacts: [3][cap] entries + zeroed [3] counters."""
    cap = max(1 << 16, n_edges // 2)
    key = str(dev)
    buf = _FIX_BUF.get(key)
    if buf is None or buf[2] < cap:
        buf = (torch.empty(3 * cap, dtype=torch.int64, device=dev), torch.zeros(4, dtype=torch.int32, device=dev), cap)
        _FIX_BUF[key] = buf
    buf[1].zero_()
    return buf


def _r4(k: int) -> int:
    return (k + 3) // 4 * 4


class Arr:
    """Rows [lo, lo + n) of a virtual ``[*, F]`` fp32 array in plan order, backed by a compact tensor."""

    __slots__ = ("t", "lo", "n", "F")

    def __init__(self, t: torch.Tensor, lo: int):
        assert t.dim() == 2 and t.is_contiguous() and t.dtype == torch.float32
        self.t, self.lo, self.n, self.F = t, lo, t.shape[0], t.shape[1]

    @property
    def addr(self) -> int:                     # address of virtual row 0
        return self.t.data_ptr() - self.lo * self.F * 4

    @property
    def row0(self) -> int:                     # address of the first backed row
        return self.t.data_ptr()

    @staticmethod
    def empty(n, F, lo, dev):
        return Arr(torch.empty(max(n, 1), F, dtype=torch.float32, device=dev)[:n], lo)

    @staticmethod
    def zeros(n, F, lo, dev):
        return Arr(torch.zeros(max(n, 1), F, dtype=torch.float32, device=dev)[:n], lo)


class GradSink:
    """fp32 gradient buffers of the parameters touched by one backward (kernels accumulate)."""

    def __init__(self):
        self.g = {}

    def of(self, p: torch.Tensor) -> torch.Tensor:
        t = self.g.get(id(p))
        if t is None:
            t = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)
            self.g[id(p)] = t
        return t

    def get(self, p):
        return self.g.get(id(p))


class TrainMLP:
    """Per-forward image of a ``make_mlp`` stack for the training kernels: zero-padded row-major
    weights (backward), their k-major transposes (forward), padded biases.  ``col_blocks`` selects
    and orders column blocks of the first layer (edge MLP: one block per gathered segment)."""

    def __init__(self, seq: nn.Sequential, col_blocks=None):
        mods = list(seq)
        self.linears: List[nn.Linear] = []
        self.acts: List[Optional[nn.Module]] = []
        i = 0
        while i < len(mods):
            self.linears.append(mods[i])
            if i + 1 < len(mods) and not isinstance(mods[i + 1], nn.Linear):
                self.acts.append(mods[i + 1]); i += 2
            else:
                self.acts.append(None); i += 1
        for a in self.acts:
            if isinstance(a, nn.PReLU) and a.weight.numel() != 1:
                raise NotImplementedError("per-channel PReLU is not supported")
            if a is not None and activation_name_of(a) not in ACT_CODES:
                raise NotImplementedError(f"activation {a} is not supported by the kernels")
        L = len(self.linears)
        self.n_pad = [_tile_width(l.weight.shape[0]) for l in self.linears]
        k0 = self.linears[0].weight.shape[1]
        self.blocks = col_blocks if col_blocks is not None else [(0, k0)]      # (offset, width) in W1 columns
        self.block_pad = [_r4(w) for _, w in self.blocks]
        dev = self.linears[0].weight.device
        self.w_rm, self.wt, self.bias = [], [], []
        with torch.no_grad():
            for li, lin in enumerate(self.linears):
                w = lin.weight.detach().float()
                n_out, k_in = w.shape
                if li == 0:
                    k_pad = sum(self.block_pad)
                    wp = torch.zeros(self.n_pad[0], k_pad, device=dev)
                    c = 0
                    for (off, wd), bp in zip(self.blocks, self.block_pad):
                        wp[:n_out, c:c + wd] = w[:, off:off + wd]
                        c += bp
                else:
                    k_pad = self.n_pad[li - 1]
                    wp = torch.zeros(self.n_pad[li], k_pad, device=dev)
                    wp[:n_out, :k_in] = w
                self.w_rm.append(wp)
                self.wt.append(wp.t().contiguous())
                if lin.bias is not None:
                    b = torch.zeros(self.n_pad[li], device=dev)
                    b[:n_out] = lin.bias.detach().float()
                    self.bias.append(b)
                else:
                    self.bias.append(None)
        self.L = L

    def act_code(self, li):
        return ACT_CODES[activation_name_of(self.acts[li])]

    def slope(self, li):
        a = self.acts[li]
        return a.weight if isinstance(a, nn.PReLU) else None

    # ---- forward: returns the list of pre-activation tensors [R, n_pad[li]]
    def forward(self, first_segs, n_rows: int, dev):
        """first_segs: list of (base, idx, ld, width) matching ``self.blocks``."""
        pres = []
        for li in range(self.L):
            pre = torch.empty(max(n_rows, 1), self.n_pad[li], dtype=torch.float32, device=dev)
            if li == 0:
                rows = lib.make_rows([(b, i, ld, w, 0, None) for (b, i, ld, w) in first_segs])
            else:
                rows = lib.make_rows([(pres[li - 1], None, self.n_pad[li - 1], self.n_pad[li - 1],
                                       self.act_code(li - 1), self.slope(li - 1))])
            lib.mlp_layer_fwd(rows, n_rows, self.wt[li], self.bias[li], self.n_pad[li], pre)
            pres.append(pre)
        return pres

    # ---- backward: dh = grad w.r.t. the post-activation output of the last layer, [R, n_pad[-1]] (overwritten)
    def backward(self, dh, pres, first_segs, n_rows: int, sink: GradSink, need_dx: List[bool], dx_out=None,
                 dx_accumulate=None):
        """Returns the list of first-layer input gradients (one [R, tile] tensor per block, None
        where ``need_dx`` is False).  ``dx_out[j]`` / ``dx_accumulate[j]`` let a block's gradient be
        accumulated into an existing [R, tile] tensor.

        Wide layers (64/128 outputs, inputs in 32-column multiples: the edge MLP of the default
        model) run their two GEMMs on the tensor cores (``swe_mlp_layer_bwd_dx_tc/_dw_tc``);
        everything else stays on the exact-fp32 CUDA-core kernels.  ``MSWE_TRAIN_GEMM=ffma``
        forces the latter."""
        dev = dh.device
        if n_rows == 0:
            return [None] * len(self.blocks)
        tc_dx_on, tc_dw_on = _tc_part("dx"), _tc_part("dw")
        for li in range(self.L - 1, -1, -1):
            lin = self.linears[li]
            n_out, k_in = lin.weight.shape
            n = self.n_pad[li]
            act, slope = self.act_code(li), self.slope(li)
            grid = lib.mlp_layer_bwd_dx_grid(n_rows)
            part = torch.empty(max(grid, 1) * (n + 1), dtype=torch.float32, device=dev)
            gw = sink.of(lin.weight)
            if li > 0:
                ko = self.n_pad[li - 1]
                dx = torch.empty(max(n_rows, 1), ko, dtype=torch.float32, device=dev)
                prev_act = self.act_code(li - 1)
                tc = tc_dx_on and n in (64, 128) and ko in (64, 128) and n_rows >= _min_rows()
                tc_dw = tc_dw_on and prev_act in _TC_ACTS and n in (64, 128) and ko in (64, 128) and n_rows >= _min_rows()
                if tc and act in _TC_ACTS and _fuse_delta():
                    # delta (in place), bias / slope partials and the GEMM in one pass over the rows
                    grid = lib.mlp_layer_bwd_dx_tc_grid(n_rows)
                    part = torch.empty(max(grid, 1) * (n + 1), dtype=torch.float32, device=dev)
                    lib.mlp_layer_bwd_dx_tc_fused(dh, pres[li], act, slope, n_rows, n, self.w_rm[li], self.w_rm[li].shape[1], 0,
                                                  ko, ko, dx, False, None, False, None, part)
                elif tc:
                    # delta (in place) + bias / slope partials on CUDA cores (element-wise), the GEMM on tensor cores
                    lib.mlp_layer_bwd_dx(dh, pres[li], act, slope, n_rows, n, self.w_rm[li], self.w_rm[li].shape[1], 0, 16, 16,
                                         None, False, True, part)
                    lib.mlp_layer_bwd_dx_tc(dh, n_rows, n, self.w_rm[li], self.w_rm[li].shape[1], 0, ko, ko, dx, False)
                else:
                    lib.mlp_layer_bwd_dx(dh, pres[li], act, slope, n_rows, n, self.w_rm[li], self.w_rm[li].shape[1], 0, ko, ko,
                                         dx, False, True, part)
                rows = lib.make_rows([(pres[li - 1], None, ko, ko, prev_act, self.slope(li - 1))])
                if tc_dw:
                    self._dw_tc(dh, n_rows, n, rows, [(0, ko, 0, k_in)], ko, gw, n_out, k_in, dev)
                else:
                    self._dw(dh, n_rows, n, rows, ko, gw, n_out, k_in, 0, k_in, dev)
                results = None
            else:
                nb = len(self.blocks)
                wide = n in (64, 128) and n_rows >= _min_rows() and all(
                    bp == wd and wd in (64, 128) for (_, wd), bp in zip(self.blocks, self.block_pad))
                tc, tc_dw = tc_dx_on and wide, tc_dw_on and wide
                results = [None] * nb
                outs, accs = [None] * nb, [False] * nb
                for j in range(nb):
                    if need_dx[j]:
                        if dx_out is not None and dx_out[j] is not None:
                            outs[j], accs[j] = dx_out[j], bool(dx_accumulate[j])
                        else:
                            outs[j] = torch.empty(max(n_rows, 1), _tile_width(self.block_pad[j]), dtype=torch.float32,
                                                  device=dev)
                        results[j] = outs[j]
                col0 = [0] * nb
                for j in range(1, nb):
                    col0[j] = col0[j - 1] + self.block_pad[j - 1]
                ld0 = self.w_rm[0].shape[1]
                fuse = tc and act in _TC_ACTS and _fuse_delta() and any(o is not None for o in outs)
                if (tc or tc_dw) and not fuse:
                    lib.mlp_layer_bwd_dx(dh, pres[0], act, slope, n_rows, n, self.w_rm[0], ld0, 0, 16, 16, None, False, True,
                                         part)
                if tc:
                    # input gradients: two 64-wide neighbours per pass where possible; the first pass also forms delta
                    j = 0
                    while j < nb:
                        if outs[j] is None:
                            j += 1
                            continue
                        wj = self.block_pad[j]
                        pair = wj == 64 and j + 1 < nb and outs[j + 1] is not None and self.block_pad[j + 1] == 64
                        args = (n_rows, n, self.w_rm[0], ld0, col0[j], 128 if pair else wj, 128 if pair else wj, outs[j], accs[j],
                                outs[j + 1] if pair else None, accs[j + 1] if pair else False, 64 if pair else None)
                        if fuse:
                            grid = lib.mlp_layer_bwd_dx_tc_grid(n_rows)
                            part = torch.empty(max(grid, 1) * (n + 1), dtype=torch.float32, device=dev)
                            lib.mlp_layer_bwd_dx_tc_fused(dh, pres[0], act, slope, *args, part)
                            fuse = False
                        else:
                            lib.mlp_layer_bwd_dx_tc(dh, *args)
                        j += 2 if pair else 1
                else:
                    first = not tc_dw                       # (delta already finished when the dW side runs on tensor cores)
                    for j in range(nb):
                        bp = self.block_pad[j]
                        if first or outs[j] is not None:
                            lib.mlp_layer_bwd_dx(dh, pres[0] if first else None, act if first else 0, slope if first else None,
                                                 n_rows, n, self.w_rm[0], ld0, col0[j], bp, _tile_width(bp), outs[j], accs[j],
                                                 True, part if first else None)
                            first = False
                if tc_dw:
                    # weight gradients: provider segments in groups of up to 256 columns
                    j = 0
                    while j < nb:
                        grp, wsum = [], 0
                        while j < nb and wsum + self.block_pad[j] <= 256:
                            grp.append(j); wsum += self.block_pad[j]; j += 1
                        while wsum not in (64, 128, 256):            # the kernel takes 64, 128 or 256 columns
                            j -= 1; wsum -= self.block_pad[grp.pop()]
                        rows = lib.make_rows([(first_segs[q][0], first_segs[q][1], first_segs[q][2], first_segs[q][3], 0, None)
                                              for q in grp])
                        items, c = [], 0
                        for q in grp:
                            off, wd = self.blocks[q]
                            items.append((c, wd, off, wd)); c += wd
                        self._dw_tc(dh, n_rows, n, rows, items, wsum, gw, n_out, k_in, dev)
                else:
                    for (off, wd), bp, seg in zip(self.blocks, self.block_pad, first_segs):
                        b, i, ld, w = seg
                        rows = lib.make_rows([(b, i, ld, w, 0, None)])
                        self._dw(dh, n_rows, n, rows, _tile_width(bp), gw, n_out, k_in, off, wd, dev)
            # bias and PReLU slope gradients from the partials of the (first) dx call
            if lin.bias is not None:
                lib.reduce_partials(part, grid, n + 1, 0, n_out, n_out, n_out, sink.of(lin.bias), n_out, 0)
            if slope is not None:
                lib.reduce_partials(part, grid, n + 1, n, 1, 1, 1, sink.of(slope), 1, 0)
            if li > 0:
                dh = dx
        return results

    @staticmethod
    def _dw_square(delta, n_rows, F, rows, gw, dev):
        """dW of a bias-free F x F Linear over node rows (hop filters): tensor cores for F = 64 on large levels."""
        if F == 64 and n_rows >= _min_rows() and _tc_part("dw"):
            TrainMLP._dw_tc(delta, n_rows, F, rows, [(0, F, 0, F)], F, gw, F, F, dev)
        else:
            TrainMLP._dw(delta, n_rows, F, rows, F, gw, F, F, 0, F, dev)

    @staticmethod
    def _dw_tc(delta, n_rows, n, rows, items, width, gw, n_out, k_in, dev):
        """Tensor-core weight gradient of one provider group; ``items``: (column offset inside the group, segment
        width, first column in the Linear weight, valid columns) per segment."""
        grid = lib.mlp_layer_bwd_dw_tc_grid(n_rows)
        part = torch.empty(max(grid, 1) * n * width, dtype=torch.float32, device=dev)
        g = lib.mlp_layer_bwd_dw_tc(delta, n_rows, n, rows, part)
        for c0, w, k_off, k_valid in items:
            lib.reduce_partials(part, g, n * width, c0 * n, n_out * w, w, k_valid, gw, k_in, k_off)

    @staticmethod
    def _dw(delta, n_rows, n, rows, ko, gw, n_out, k_in, k_off, k_valid, dev):
        grid = lib.mlp_layer_bwd_dw_grid(n_rows)
        part = torch.empty(max(grid, 1) * n * ko, dtype=torch.float32, device=dev)
        g = lib.mlp_layer_bwd_dw(delta, n_rows, n, rows, ko, part)
        if n_rows > 0:
            lib.reduce_partials(part, g, n * ko, 0, n_out * ko, ko, k_valid, gw, k_in, k_off)


# ------------------------------------------------------------------------------------------------
# one SWEGNN call
# ------------------------------------------------------------------------------------------------
def _transposed(es):
    """(t_rowptr, t_pos): edges grouped by source node (stable), t_pos = position in dst-CSR order."""
    if es.t_rowptr is None:
        E = es.n_edges
        dev = es.rowptr.device
        n_src = es.src_hi - es.src_lo
        if E == 0:
            es.t_rowptr = torch.zeros(n_src + 1, dtype=torch.int32, device=dev)
            es.t_pos = torch.zeros(1, dtype=torch.int32, device=dev)
        else:
            rp, _, _, order = lib.csr_build(es.src.long(), es.dst.long(), None, es.src_lo, n_src, es.dst_lo,
                                            es.dst_lo + es.n_dst, by_row=True)
            es.t_rowptr, es.t_pos = rp, order.contiguous()       # "original edge id" of this build = CSR position
    return es.t_rowptr, es.t_pos


def _node_linear_bwd(g: Arr, lo: int, n: int, W: torch.Tensor, out: Arr, F: int):
    """out[lo:lo+n] = g[lo:lo+n] · W (input gradient of a bias-free Linear over node rows)."""
    if F == 64 and n >= _min_rows() and _tc_part("dx"):
        lib.mlp_layer_bwd_dx_tc(g.row0, n, F, W.detach().contiguous(), F, 0, F, F, out.addr + lo * F * 4, False)
    else:
        lib.node_linear_fwd(g.addr, lo, n, W.detach().contiguous(), out.addr, F)


class SweCall:
    pass


def swegnn_forward_train(mod, es, xs: Arr, xd_src: Arr, xd_dst: Optional[Arr], a: Optional[torch.Tensor], out: Arr,
                         addend: Optional[Arr]) -> SweCall:
    """Training-mode ``SWEGNN.forward`` (reference ``models/gnn.py:387-445``) on a CSR edge set;
    writes the destination rows of ``out`` and returns what the backward needs."""
    F = mod.edge_output_size
    if F not in (16, 32, 64):
        raise NotImplementedError("training kernels need hid_features in {16, 32, 64}")
    if mod.upwind_mode:
        raise NotImplementedError("upwind_mode=True has no backward kernel")
    dev = xs.t.device
    E = es.n_edges
    c = SweCall()
    c.mod, c.es, c.xs, c.xd_src, c.xd_dst, c.a, c.addend = mod, es, xs, xd_src, xd_dst, a, addend
    blocks, segs = [(0, F), (F, F), (2 * F, F)], [(xs.addr, es.src, F, F), (xs.addr, es.dst, F, F),
                                                  (xd_src.addr, es.src, F, F)]
    c.seg_kind = ["xs_src", "xs_dst", "xd_src"]
    if xd_dst is not None:
        blocks.append((3 * F, F)); segs.append((xd_dst.addr, es.dst, F, F)); c.seg_kind.append("xd_dst")
    if mod.edge_features > 0:
        if a is None:
            raise ValueError("SWEGNN built with edge features needs edge_attr")
        blocks.append((4 * F, F)); segs.append((a, None, F, F)); c.seg_kind.append("a")
    c.mlp = TrainMLP(mod.edge_mlp, blocks)
    c.segs = segs
    c.s = torch.empty(max(E, 1), F, dtype=torch.float32, device=dev)
    gate_tc = mod.launcher().tc if _tc_part("fwd") else None
    if gate_tc is not None and all(c.mlp.act_code(i) in _TC_ACTS for i in range(3)):
        # the default model (F = 64, 5F|4F -> 2F -> 2F -> F): the inference gate kernel on the tensor cores, in the
        # variant that also stores the three pre-activations the backward needs
        codes, slopes = gate_tc.acts_and_slopes()
        c.pres = [torch.empty(max(E, 1), w, dtype=torch.float32, device=dev) for w in (2 * F, 2 * F, F)]
        a_in = a if mod.edge_features > 0 else None
        xd_dst_addr = None if xd_dst is None else xd_dst.addr
        k1 = gate_tc.linears[0].weight.shape[1]
        tau = _fix_tau()
        lists, count, cap = _fix_buffers(E, dev) if tau > 0 else (None, None, 0)
        lib.edge_gate_tc_train_fwd(xs.addr, xd_src.addr, xd_dst_addr, a_in, es.src, es.dst, E, gate_tc.image(), k1, codes,
                                   slopes, mod.normalize, c.pres[0], c.pres[1], c.pres[2], c.s, lists, count, cap, tau)
        if tau > 0 and E > 0:
            # the few pre-activations that lie within the 3xTF32 error of the activation's kink are re-evaluated in
            # exact fp32, so that the derivative masks of the backward are those of the fp32 reference
            l1, l2, l3 = gate_tc.linears
            wb = [t.detach().contiguous() if t is not None else None
                  for t in (l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias)]
            lib.gate_fix_preacts(xs.addr, xd_src.addr, xd_dst_addr, a_in, es.src, es.dst, *wb, k1, codes, slopes, c.pres[0],
                                 c.pres[1], c.pres[2], lists, count, cap)
    else:
        c.pres = c.mlp.forward(segs, E, dev)
        last = c.mlp.L - 1
        lib.gate_norm_fwd(c.pres[last], c.mlp.act_code(last), c.mlp.slope(last), mod.normalize, E, c.s, F)
    K = mod.K
    lo, n = es.dst_lo, es.n_dst
    c.o, c.agg = [], []
    if mod.with_filter_matrix:
        if xd_dst is None or es.src_lo != es.dst_lo:
            raise NotImplementedError("filter matrices need source and destination rows in the same node set")
        W = mod.launcher().filters.tensors()
        o0 = Arr.empty(n, F, lo, dev)
        lib.node_linear_fwd(xd_dst.addr, lo, n, W[0], o0.addr, F)
        o_src = o_dst = o0
    else:
        W = [None] * (K + 1)
        o_src, o_dst = xd_src, xd_dst
        if K > 1:
            raise NotImplementedError("K > 1 without filter matrices is not supported in training")
    for k in range(K):
        last_hop = k == K - 1
        dst = out if last_hop else Arr.empty(n, F, lo, dev)
        agg = Arr.empty(n, F, lo, dev) if W[k + 1] is not None else None
        lib.propagate_hop_train_fwd(o_src.addr, None if o_dst is None else o_dst.addr, c.s, es.rowptr, es.src, lo, n,
                                    W[k + 1], mod.with_gradient, 0, addend.addr if (addend is not None and last_hop) else None,
                                    None if agg is None else agg.addr, dst.addr, F)
        c.o.append((o_src, o_dst))
        c.agg.append(agg)
        o_src = o_dst = dst
    return c


def swegnn_backward(c: SweCall, g_out: Arr, d_xs: Arr, d_xd_src: Arr, d_xd_src_accumulate: bool,
                    d_a: Optional[torch.Tensor], d_a_accumulate: bool, sink: GradSink) -> Optional[Arr]:
    """Backward of one call.  g_out: gradient w.r.t. the output rows.  Accumulates into ``d_xs``
    (always) and ``d_xd_src`` (gradient w.r.t. the x_d array the SOURCE rows were read from; written
    or accumulated).  Returns the gradient w.r.t. the destination rows of x_d when that is a
    different array (``None`` when destination == source array, in which case it is in d_xd_src)."""
    mod, es = c.mod, c.es
    F = mod.edge_output_size
    dev = c.s.device
    E, K = es.n_edges, mod.K
    lo, n = es.dst_lo, es.n_dst
    src_lo, n_src = es.src_lo, es.src_hi - es.src_lo
    t_rowptr, t_pos = _transposed(es)
    ds = torch.empty(max(E, 1), F, dtype=torch.float32, device=dev)
    same = c.xd_dst is not None and es.src_lo == es.dst_lo
    g = g_out
    if mod.with_gradient:
        if not same:
            raise NotImplementedError("with_gradient needs source and destination rows in the same node set")
        for k in range(K - 1, -1, -1):
            o_src, o_dst = c.o[k]
            if mod.with_filter_matrix:
                Wk = mod.filter_matrix[k + 1].weight
                rows = lib.make_rows([(c.agg[k].row0, None, F, F, 0, None)])
                TrainMLP._dw_square(g.row0, n, F, rows, sink.of(Wk), dev)
                da = Arr.empty(n, F, lo, dev)
                _node_linear_bwd(g, lo, n, Wk, da, F)
            else:
                da = g
            flags = torch.empty(lo + n, dtype=torch.uint8, device=dev)
            lib.row_flags(o_src.addr, lo, n, flags, F)
            gp = Arr.empty(n, F, lo, dev)
            lib.hop_bwd_dst(da.addr, o_src.addr, o_dst.addr, c.s, ds, k != K - 1, es.rowptr, es.src, flags, flags, lo, n, 1,
                            g.addr, gp.addr, F)
            lib.hop_bwd_src(da.addr, c.s, t_rowptr, t_pos, es.dst, flags, flags, src_lo, n_src, 1, 0, gp.addr, F)
            g = gp
    else:
        if K != 1 or mod.with_filter_matrix:
            raise NotImplementedError("with_gradient=False is supported for K=1 without filter matrices (un-pooling)")
        o_src, o_dst = c.o[0]
        hi = max(src_lo + n_src, lo + n)
        wet_src = torch.zeros(hi, dtype=torch.uint8, device=dev)
        lib.row_flags(o_src.addr, src_lo, n_src, wet_src, F)
        wet_dst = None
        if o_dst is not None:
            wet_dst = torch.zeros(hi, dtype=torch.uint8, device=dev)
            lib.row_flags(o_dst.addr, lo, n, wet_dst, F)
        lib.hop_bwd_dst(g.addr, o_src.addr, None if o_dst is None else o_dst.addr, c.s, ds, 0, es.rowptr, es.src, wet_src,
                        wet_dst, lo, n, 0, None, None, F)
        # gradient w.r.t. the source rows of x_d through  s ⊙ o[src]
        lib.hop_bwd_src(g.addr, c.s, t_rowptr, t_pos, es.dst, wet_src, wet_dst, src_lo, n_src, 0, d_xd_src_accumulate,
                        d_xd_src.addr, F)
        d_xd_src_accumulate = True
    # ---- o_0 = x_d W0ᵀ  (or x_d itself)
    d_dst_out = None
    if mod.with_gradient:
        if mod.with_filter_matrix:
            W0 = mod.filter_matrix[0].weight
            rows = lib.make_rows([(c.xd_dst.addr + lo * F * 4, None, F, F, 0, None)])
            TrainMLP._dw_square(g.row0, n, F, rows, sink.of(W0), dev)
            if d_xd_src_accumulate:
                tmp = Arr.empty(n, F, lo, dev)
                _node_linear_bwd(g, lo, n, W0, tmp, F)
                _add_rows(d_xd_src, tmp, lo, n, F)
            else:
                _node_linear_bwd(g, lo, n, W0, d_xd_src, F)
        else:
            if d_xd_src_accumulate:
                _add_rows(d_xd_src, g, lo, n, F)
            else:
                d_xd_src.t.copy_(g.t)
        d_xd_src_accumulate = True
    else:
        d_dst_out = g                      # out = o_dst + agg (+ addend): identity w.r.t. the destination rows
    # ---- gate
    last = c.mlp.L - 1
    lib.gate_norm_bwd(ds, c.pres[last], c.mlp.act_code(last), c.mlp.slope(last), mod.normalize, E, F)
    need = [True] * len(c.segs)
    dx_out = [None] * len(c.segs)
    dx_acc = [False] * len(c.segs)
    for j, kind in enumerate(c.seg_kind):
        if kind == "a":
            if d_a is None:
                need[j] = False
            else:
                dx_out[j], dx_acc[j] = d_a, d_a_accumulate
    dz = c.mlp.backward(ds, c.pres, c.segs, E, sink, need, dx_out, dx_acc)
    for j, kind in enumerate(c.seg_kind):
        if kind == "xs_src":
            lib.edge_to_node_sum(dz[j], t_rowptr, t_pos, src_lo, n_src, d_xs.addr, 1, F)
        elif kind == "xs_dst":
            lib.edge_to_node_sum(dz[j], es.rowptr, None, lo, n, d_xs.addr, 1, F)
        elif kind == "xd_src":
            lib.edge_to_node_sum(dz[j], t_rowptr, t_pos, src_lo, n_src, d_xd_src.addr, int(d_xd_src_accumulate), F)
            d_xd_src_accumulate = True
        elif kind == "xd_dst":
            lib.edge_to_node_sum(dz[j], es.rowptr, None, lo, n, d_xd_src.addr, 1, F)
    return d_dst_out


def _add_rows(dst: Arr, src: Arr, lo: int, n: int, F: int):
    """dst[lo:lo+n] += src[lo:lo+n] through the identity-CSR of edge_to_node_sum (one 'edge' per row)."""
    dev = dst.t.device
    rp = _iota(n + 1, dev)
    lib.edge_to_node_sum(src.addr + lo * F * 4, rp, None, lo, n, dst.addr, 1, F)


_IOTA = {}


def _iota(n: int, dev) -> torch.Tensor:
    key = (dev, )
    t = _IOTA.get(key)
    if t is None or t.numel() < n:
        t = torch.arange(max(n, 1024), dtype=torch.int32, device=dev)
        _IOTA[key] = t
    return t


# ------------------------------------------------------------------------------------------------
# whole-model tapes
# ------------------------------------------------------------------------------------------------
class _Tape:
    pass


def _encode_train(model, plan, graph, x, n_dyn_rows, dev):
    """Encoders in training mode.  Returns tape pieces and the encoded arrays."""
    F = model.hid_features
    N = plan.n_nodes
    t = _Tape()
    n_static_raw = model.static_node_features - int(bool(model.with_WL))
    t.n_static_raw = n_static_raw
    ks = _r4(model.static_node_features)
    t.xin_s = torch.empty(max(N, 1), ks, dtype=torch.float32, device=dev)
    lib.static_inputs_fwd(x, plan.perm, N, n_static_raw, bool(model.with_WL), t.xin_s)
    t.mlp_s = TrainMLP(model.static_node_encoder)
    t.seg_s = [(t.xin_s, None, ks, model.static_node_features)]
    t.pres_s = t.mlp_s.forward(t.seg_s, N, dev)
    xs = Arr.empty(N, F, 0, dev)
    Ls = t.mlp_s.L - 1
    lib.act_fwd(t.pres_s[Ls], 0, N, t.mlp_s.act_code(Ls), t.mlp_s.slope(Ls), xs.addr, F)
    n_dyn = model.dynamic_node_features
    t.mlp_d = TrainMLP(model.dynamic_node_encoder)
    t.seg_d = [(x.data_ptr() + 4 * n_static_raw, plan.perm, x.shape[1], n_dyn)]
    t.pres_d = t.mlp_d.forward(t.seg_d, n_dyn_rows, dev)
    xd = Arr.empty(n_dyn_rows, F, 0, dev)
    Ld = t.mlp_d.L - 1
    lib.act_fwd(t.pres_d[Ld], 0, n_dyn_rows, t.mlp_d.act_code(Ld), t.mlp_d.slope(Ld), xd.addr, F)
    # edge encoder (rows in CSR order of every scale)
    t.a = None
    E = plan.n_edges_total
    if getattr(model, "_pk_edge", None) is not None:
        if getattr(plan, "eid_global", None) is None:
            plan.eid_global = torch.cat([es.eid + lo for (lo, hi), es in zip(plan.edge_slices, plan.edges)]).to(torch.int32).contiguous() \
                if E else torch.zeros(1, dtype=torch.int32, device=dev)
        ea = graph.edge_attr
        if ea.dtype != torch.float32 or not ea.is_contiguous():
            raise TypeError("edge_attr must be a contiguous float32 tensor")
        t.mlp_e = TrainMLP(model.edge_encoder)
        t.seg_e = [(ea, plan.eid_global, ea.shape[1], ea.shape[1])]
        t.pres_e = t.mlp_e.forward(t.seg_e, E, dev)
        t.a = torch.empty(max(E, 1), F, dtype=torch.float32, device=dev)
        Le = t.mlp_e.L - 1
        lib.act_fwd(t.pres_e[Le], 0, E, t.mlp_e.act_code(Le), t.mlp_e.slope(Le), t.a, F)
    else:
        ea = graph.edge_attr
        if ea is not None and ea.shape[1] == F:
            t.a = torch.cat([ea[lo:hi][es.eid.long()] for (lo, hi), es in zip(plan.edge_slices, plan.edges)]).contiguous()
    return t, xs, xd


def _encode_backward(model, plan, t, d_xs: Arr, d_xd: Arr, d_a, n_dyn_rows, dx0, sink, dev):
    N = plan.n_nodes
    want_dx = dx0 is not None
    dzs = t.mlp_s.backward(d_xs.t, t.pres_s, t.seg_s, N, sink, [want_dx])
    dzd = t.mlp_d.backward(d_xd.t, t.pres_d, t.seg_d, n_dyn_rows, sink, [want_dx])
    if want_dx:
        lib.node_inputs_bwd(dzs[0], dzd[0], dx0.shape[1], plan.perm, N, n_dyn_rows, t.n_static_raw, bool(model.with_WL), dx0)
    if d_a is not None and getattr(t, "mlp_e", None) is not None:
        t.mlp_e.backward(d_a, t.pres_e, t.seg_e, plan.n_edges_total, sink, [False])


def _head_train(model, plan, h: Arr, act_name, act_module, x, dev):
    """Decoder + residual + ReLU + dry mask (reference gnn.py:332-348 / 141-150)."""
    F = model.hid_features
    N = plan.n_nodes
    t = _Tape()
    t.h, t.act_name, t.act_module = h, act_name, act_module
    t.mlp = TrainMLP(model.node_decoder)
    slope = act_module.weight if isinstance(act_module, nn.PReLU) else None
    rows_first = [(h.addr, None, F, F)]
    # the gnn activation in front of the decoder is applied by the first layer's row provider
    pres = []
    for li in range(t.mlp.L):
        pre = torch.empty(max(N, 1), t.mlp.n_pad[li], dtype=torch.float32, device=dev)
        if li == 0:
            rows = lib.make_rows([(h.addr, None, F, F, ACT_CODES[act_name], slope)])
        else:
            rows = lib.make_rows([(pres[li - 1], None, t.mlp.n_pad[li - 1], t.mlp.n_pad[li - 1], t.mlp.act_code(li - 1),
                                   t.mlp.slope(li - 1))])
        lib.mlp_layer_fwd(rows, N, t.mlp.wt[li], t.mlp.bias[li], t.mlp.n_pad[li], pre)
        pres.append(pre)
    t.pres = pres
    t.res_mode = model._residual_mode()
    t.res_w = model.residual_weights.detach().contiguous() if t.res_mode in (1, 2) else None
    pred = torch.empty(N, model.out_dim, dtype=torch.float32, device=dev)
    L = t.mlp.L - 1
    lib.head_fwd(pres[L], t.mlp.act_code(L), t.mlp.slope(L), x, plan.perm, N, model.previous_t, t.res_mode, t.res_w,
                 1e-4, pred)
    t.x = x
    return t, pred


def _head_backward(model, plan, t, dpred, dx0, sink, dev) -> Arr:
    """Returns the gradient w.r.t. the processor output h (before the gnn activation)."""
    F = model.hid_features
    N = plan.n_nodes
    L = t.mlp.L - 1
    dh = torch.empty(max(N, 1), t.mlp.n_pad[L], dtype=torch.float32, device=dev)
    n_part = 2 * model.previous_t
    res_part = torch.zeros(4096 * 16, dtype=torch.float32, device=dev)
    grid = lib.head_bwd(dpred, t.pres[L], t.mlp.act_code(L), t.mlp.slope(L), t.x, plan.perm, N, model.previous_t,
                        t.res_mode, t.res_w, 1e-4, dh, dx0, res_part)
    if t.res_mode == 2:
        lib.reduce_partials(res_part, grid, 16, 0, n_part, n_part, n_part, sink.of(model.residual_weights), n_part, 0)
    elif t.res_mode == 1:
        gw = sink.of(model.residual_weights)
        lib.reduce_partials(res_part, grid, 16, 0, n_part, 2, 1, gw, 1, 0)
        lib.reduce_partials(res_part, grid, 16, 1, n_part, 2, 1, gw, 1, 0)
    slope = t.act_module.weight if isinstance(t.act_module, nn.PReLU) else None
    # decoder layers; the first layer's input is act(h): its gradient still needs act'
    mlp = t.mlp
    first_seg = [(None, None, F, F)]
    # custom chain because the first-layer provider applies the gnn activation
    dcur = dh
    for li in range(mlp.L - 1, -1, -1):
        lin = mlp.linears[li]
        n_out, k_in = lin.weight.shape
        n = mlp.n_pad[li]
        grid_dx = lib.mlp_layer_bwd_dx_grid(N)
        part = torch.empty(max(grid_dx, 1) * (n + 1), dtype=torch.float32, device=dev)
        ko = mlp.n_pad[li - 1] if li > 0 else _tile_width(F)
        dx = torch.empty(max(N, 1), ko, dtype=torch.float32, device=dev)
        lib.mlp_layer_bwd_dx(dcur, t.pres[li], mlp.act_code(li), mlp.slope(li), N, n, mlp.w_rm[li], mlp.w_rm[li].shape[1],
                             0, ko if li > 0 else F, ko, dx, False, True, part)
        if li > 0:
            rows = lib.make_rows([(t.pres[li - 1], None, ko, ko, mlp.act_code(li - 1), mlp.slope(li - 1))])
        else:
            rows = lib.make_rows([(t.h.addr, None, F, F, ACT_CODES[t.act_name], slope)])
        TrainMLP._dw(dcur, N, n, rows, ko, sink.of(lin.weight), n_out, k_in, 0, k_in, dev)
        if lin.bias is not None:
            lib.reduce_partials(part, grid_dx, n + 1, 0, n_out, n_out, n_out, sink.of(lin.bias), n_out, 0)
        if mlp.slope(li) is not None:
            lib.reduce_partials(part, grid_dx, n + 1, n, 1, 1, 1, sink.of(mlp.slope(li)), 1, 0)
        dcur = dx
    d_h = Arr.empty(N, F, 0, dev)
    if t.act_name is None:
        return Arr(dcur[:N] if dcur.shape[0] != N else dcur, 0)
    gpart = torch.zeros(4096, dtype=torch.float32, device=dev) if slope is not None else None
    g = lib.act_bwd(dcur, t.h.addr, 0, N, ACT_CODES[t.act_name], slope, d_h.addr, gpart, F)
    if slope is not None:
        lib.reduce_partials(gpart, g, 1, 0, 1, 1, 1, sink.of(slope), 1, 0)
    return d_h


def _a_slice(a, plan, s):
    if a is None:
        return None
    lo, hi = plan.edge_slices[s]
    return a[lo:hi] if hi > lo else None


# ---- MSGNN ------------------------------------------------------------------------------------
def msgnn_forward_train(model, plan, graph, x):
    dev = x.device
    F, S = model.hid_features, model.num_scales
    lo, nn_ = plan.scale_lo, plan.scale_n
    T = _Tape()
    T.enc, xs, xd0 = _encode_train(model, plan, graph, x, nn_[0], dev)
    T.xs = xs
    a = T.enc.a
    N = plan.n_nodes
    up = torch.empty(max(N, 1), F, dtype=torch.float32, device=dev)
    down = torch.empty(max(N, 1), F, dtype=torch.float32, device=dev)
    up_s = [Arr(up[lo[s]:lo[s] + nn_[s]], lo[s]) for s in range(S)]
    down_s = [Arr(down[lo[s]:lo[s] + nn_[s]], lo[s]) for s in range(S)]
    T.calls_down, T.calls_up, T.calls_unpool = [], [], []
    cur = Arr(xd0.t, lo[0])
    for i in range(S - 1):
        c = swegnn_forward_train(model.gnn_processor[i], plan.edges[i], xs, cur, cur, _a_slice(a, plan, i), down_s[i], None)
        T.calls_down.append(c)
        pe = plan.pool[i]
        nxt = Arr.empty(nn_[i + 1], F, lo[i + 1], dev)
        lib.pool_mean_fwd(down_s[i].addr, pe.rowptr, pe.src, pe.dst_lo, pe.n_dst, nxt.addr, F)
        cur = nxt
    for i in range(S):
        s = S - 1 - i
        c = swegnn_forward_train(model.gnn_processor[S - 1 + i], plan.edges[s], xs, cur, cur, _a_slice(a, plan, s), up_s[s], None)
        T.calls_up.append(c)
        if i < S - 1:
            ue = plan.unpool[s - 1]
            nxt = Arr.empty(nn_[s - 1], F, lo[s - 1], dev)
            cu = swegnn_forward_train(model.intra_scale_gnn[i], ue, xs, up_s[s], None, None, nxt,
                                      down_s[s - 1] if model.skip_connections else None)
            T.calls_unpool.append(cu)
            cur = nxt
    T.up, T.up_s, T.down_s = up, up_s, down_s
    T.head, pred = _head_train(model, plan, Arr(up, 0), model._gnn_activation_name, model.gnn_activation, x, dev)
    return T, pred


def msgnn_backward(model, plan, T, dpred, want_dx: bool, sink: GradSink):
    dev = dpred.device
    F, S = model.hid_features, model.num_scales
    lo, nn_ = plan.scale_lo, plan.scale_n
    N = plan.n_nodes
    dx0 = torch.zeros(T.head.x.shape, dtype=torch.float32, device=dev) if want_dx else None
    d_up_full = _head_backward(model, plan, T.head, dpred, dx0, sink, dev)
    d_up = [Arr(d_up_full.t[lo[s]:lo[s] + nn_[s]], lo[s]) for s in range(S)]
    d_xs = Arr.zeros(N, F, 0, dev)
    a = T.enc.a
    d_a = torch.zeros_like(a) if (a is not None and getattr(T.enc, "mlp_e", None) is not None) else None
    a_used = [False] * S
    d_cur = [None] * S                       # gradient w.r.t. the x_d rows of each scale at the current stage
    d_down = [None] * S

    def da_of(s):
        if d_a is None:
            return None, False
        lo_e, hi_e = plan.edge_slices[s]
        if hi_e <= lo_e:
            return None, False
        acc = a_used[s]
        a_used[s] = True
        return d_a[lo_e:hi_e], acc

    for i in range(S - 1, -1, -1):
        s = S - 1 - i
        if i < S - 1:
            # cur[s-1] = unpool(up[s]) + down[s-1]
            g = d_cur[s - 1]
            if model.skip_connections:
                d_down[s - 1] = g
            swegnn_backward(T.calls_unpool[i], g, d_xs, d_up[s], True, None, False, sink)
        dslice, dacc = da_of(s)
        d_in = Arr.empty(nn_[s], F, lo[s], dev)
        swegnn_backward(T.calls_up[i], d_up[s], d_xs, d_in, False, dslice, dacc, sink)
        d_cur[s] = d_in
    for i in range(S - 2, -1, -1):
        pe, ue = plan.pool[i], plan.unpool[i]
        if d_down[i] is None:
            d_down[i] = Arr.empty(nn_[i], F, lo[i], dev)
            acc = 0
        else:
            # the skip gradient aliases d_cur[i]; copy before accumulating so d_cur stays intact for nobody (consumed)
            acc = 1
        lib.pool_mean_bwd(d_cur[i + 1].addr, ue.rowptr, ue.src, lo[i], nn_[i], pe.rowptr, lo[i + 1], d_down[i].addr, acc, F)
        dslice, dacc = da_of(i)
        d_in = Arr.empty(nn_[i], F, lo[i], dev)
        swegnn_backward(T.calls_down[i], d_down[i], d_xs, d_in, False, dslice, dacc, sink)
        d_cur[i] = d_in
    _encode_backward(model, plan, T.enc, d_xs, Arr(d_cur[0].t, 0), d_a, nn_[0], dx0, sink, dev)
    return dx0


# ---- GNN --------------------------------------------------------------------------------------
def gnn_forward_train(model, plan, graph, x):
    dev = x.device
    F = model.hid_features
    N = plan.n_nodes
    T = _Tape()
    T.enc, xs, xd = _encode_train(model, plan, graph, x, N, dev)
    a = T.enc.a
    es = plan.edges[0]
    T.calls, T.raw = [], []
    cur = xd
    act = ACT_CODES[model._gnn_activation_name]
    slope = model.gnn_activation.weight if isinstance(model.gnn_activation, nn.PReLU) else None
    for conv in model.gnn_processor:
        raw = Arr.empty(N, F, 0, dev)
        T.calls.append(swegnn_forward_train(conv, es, xs, cur, cur, a, raw, None))
        T.raw.append(raw)
        if act:
            nxt = Arr.empty(N, F, 0, dev)
            lib.act_fwd(raw.addr, 0, N, act, slope, nxt.addr, F)
        else:
            nxt = raw
        cur = nxt
    T.head, pred = _head_train(model, plan, cur, None, None, x, dev)
    return T, pred


def gnn_backward(model, plan, T, dpred, want_dx: bool, sink: GradSink):
    dev = dpred.device
    F = model.hid_features
    N = plan.n_nodes
    dx0 = torch.zeros(T.head.x.shape, dtype=torch.float32, device=dev) if want_dx else None
    g = _head_backward(model, plan, T.head, dpred, dx0, sink, dev)
    d_xs = Arr.zeros(N, F, 0, dev)
    a = T.enc.a
    d_a = torch.zeros_like(a) if (a is not None and getattr(T.enc, "mlp_e", None) is not None) else None
    act = ACT_CODES[model._gnn_activation_name]
    slope = model.gnn_activation.weight if isinstance(model.gnn_activation, nn.PReLU) else None
    first = True
    for c, raw in zip(reversed(T.calls), reversed(T.raw)):
        if act:
            g_raw = Arr.empty(N, F, 0, dev)
            gpart = torch.zeros(4096, dtype=torch.float32, device=dev) if slope is not None else None
            gr = lib.act_bwd(g.addr, raw.addr, 0, N, act, slope, g_raw.addr, gpart, F)
            if slope is not None:
                lib.reduce_partials(gpart, gr, 1, 0, 1, 1, 1, sink.of(slope), 1, 0)
        else:
            g_raw = g
        d_in = Arr.empty(N, F, 0, dev)
        swegnn_backward(c, g_raw, d_xs, d_in, False, d_a, not first, sink)
        first = False
        g = d_in
    _encode_backward(model, plan, T.enc, d_xs, g, d_a, N, dx0, sink, dev)
    return dx0


# ------------------------------------------------------------------------------------------------
# torch.autograd glue
# ------------------------------------------------------------------------------------------------
class _ModelFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, graph, plan, x, *params):
        x = x.detach().contiguous()
        if model.type_model == "MSGNN":
            T, pred = msgnn_forward_train(model, plan, graph, x)
        else:
            T, pred = gnn_forward_train(model, plan, graph, x)
        ctx.model, ctx.plan, ctx.T = model, plan, T
        ctx.params = params
        return pred

    @staticmethod
    def backward(ctx, dpred):
        model, plan, T = ctx.model, ctx.plan, ctx.T
        dpred = dpred.detach().contiguous().float()
        sink = GradSink()
        want_dx = ctx.needs_input_grad[3]
        if model.type_model == "MSGNN":
            dx = msgnn_backward(model, plan, T, dpred, want_dx, sink)
        else:
            dx = gnn_backward(model, plan, T, dpred, want_dx, sink)
        grads = []
        for p, need in zip(ctx.params, ctx.needs_input_grad[4:]):
            g = sink.get(p)
            grads.append(g if (need and g is not None) else None)
        ctx.T = None
        return (None, None, None, dx, *grads)


def model_autograd(model, graph):
    multiscale = model.type_model == "MSGNN"
    plan = model._plans.get(graph, getattr(model, "num_scales", 1), multiscale)
    params = [p for p in model.parameters()]
    return _ModelFn.apply(model, graph, plan, graph.x, *params)


class _SweFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, es, x_s, x_d, a_csr, *params):
        dev = x_d.device
        F = mod.edge_output_size
        N = x_d.shape[0]
        xs = Arr(x_s.detach().float().contiguous(), 0)
        xd = Arr(x_d.detach().float().contiguous(), 0)
        out = Arr.empty(N, F, 0, dev)
        a = None if a_csr is None else a_csr.detach().float().contiguous()
        ctx.call = swegnn_forward_train(mod, es, xs, xd, xd, a, out, None)
        ctx.N, ctx.has_a = N, a is not None
        ctx.params = params
        return out.t

    @staticmethod
    def backward(ctx, g_out):
        c = ctx.call
        dev = g_out.device
        F = c.mod.edge_output_size
        N = ctx.N
        sink = GradSink()
        d_xs = Arr.zeros(N, F, 0, dev)
        d_xd = Arr.empty(N, F, 0, dev)
        d_a = torch.zeros_like(c.a) if ctx.has_a else None
        r = swegnn_backward(c, Arr(g_out.detach().float().contiguous(), 0), d_xs, d_xd, False, d_a, False, sink)
        if r is not None and c.xd_dst is not None:      # out = x_d[dst] + agg: identity path of the destination rows
            _add_rows(d_xd, r, 0, N, F)
        grads = [sink.get(p) if need else None for p, need in zip(ctx.params, ctx.needs_input_grad[5:])]
        ctx.call = None
        return (None, None, d_xs.t, d_xd.t, d_a, *grads)


def swegnn_autograd(mod, x_s, x_d, edge_index, edge_attr):
    """Stand-alone differentiable ``SWEGNN.forward`` (all rows are sources and destinations)."""
    F = mod.edge_output_size

    class _G:
        pass
    g = _G()
    g.edge_index, g.x = edge_index, x_d
    plan = mod._plans.get(g, 1, False)
    es = plan.edges[0]
    a_csr = None
    if mod.edge_features > 0:
        a_csr = edge_attr[es.eid.long()]          # differentiable gather back to the caller's edge order
    params = [p for p in mod.parameters()]
    return _SweFn.apply(mod, es, x_s, x_d, a_csr, *params)
