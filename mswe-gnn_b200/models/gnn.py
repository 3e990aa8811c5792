"""Host-side mirror of the reference's ``models/gnn.py``: ``SWEGNN``, ``GNN``, ``MSGNN``.

Constructors, parameter names / shapes, ``forward`` signatures and results are the reference's
(``/root/reference/models/gnn.py:39-42,181-185,363-365``); the arithmetic is not: one forward is
a fixed sequence of fused sm_100a kernels over destination-CSR edge sets (see ``engine.py`` and
``include/swe_gnn_b200.h``).  Differences in *how* the same numbers are produced:

* the edge weights ``s_ij`` do not depend on the hop index (their inputs are loop invariants,
  reference ``gnn.py:414-422``), so they are evaluated once per ``SWEGNN`` call instead of K
  times, and the per-hop wet-edge mask (``gnn.py:408-411``) is dropped because a masked edge
  contributes ``s·(0−0)``;
* each ``SWEGNN`` call only touches the rows of the scale it works on (the reference computes
  all N rows and multiplies by ``(mask == i)``, ``gnn.py:307,322,331``);
* aggregation is a sequential sum over a stable destination-CSR segment — deterministic and in
  the same order as CPU ``scatter_add_``.

There is no CPU path: inputs must be CUDA tensors.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from .. import lib
from ..engine import PackedMLP, RowMlpTC, SweGnnLauncher, gate_layer0, padded_width, rowmlp_backend
from ..lib import ACT_CODES
from ..plan import PlanCache, build_plan
from .models import BaseFloodModel, activation_functions, make_mlp


class SWEGNN(nn.Module):
    r"""Shallow-Water-Equation inspired graph operator (reference ``gnn.py:352-451``)

    .. math::
        \mathbf{o}^{k+1}_i = \mathbf{o}^k_i + W_{k+1}\sum_{j \to i}
            \mathbf{s}_{ji} \odot (\mathbf{o}^k_i - \mathbf{o}^k_j), \qquad
        \mathbf{s}_{ji} = \frac{\mathrm{MLP}(\mathbf{x}_{sj}, \mathbf{x}_{si}, \mathbf{x}_{dj},
            \mathbf{x}_{di}, \mathbf{e}_{ji})}{\lVert\cdot\rVert_2}
    """

    def __init__(self, static_node_features: int, dynamic_node_features: int, edge_features: int,
                 K: int = 2, normalize=True, with_filter_matrix=True, with_gradient=True,
                 upwind_mode=False, device='cpu', **mlp_kwargs):
        super().__init__()
        if static_node_features != dynamic_node_features:
            raise NotImplementedError("SWEGNN kernels need static and dynamic node widths to be equal")
        self.edge_features = edge_features
        self.edge_input_size = edge_features + static_node_features * 2 + dynamic_node_features * 2
        self.edge_output_size = dynamic_node_features
        self.normalize = normalize
        self.K = K
        self.with_filter_matrix = with_filter_matrix
        self.device = device
        self.with_gradient = with_gradient
        self.upwind_mode = upwind_mode
        self.edge_mlp = make_mlp(self.edge_input_size, self.edge_output_size,
                                 hidden_size=self.edge_output_size * 2, device=device, **mlp_kwargs)
        if with_filter_matrix:
            self.filter_matrix = nn.ModuleList(
                nn.Linear(dynamic_node_features, dynamic_node_features, bias=False, device=device)
                for _ in range(K + 1))
        self._launcher: Optional[SweGnnLauncher] = None
        self._plans = PlanCache()

    def launcher(self) -> SweGnnLauncher:
        if self._launcher is None:
            self._launcher = SweGnnLauncher(self, self.edge_output_size)
        return self._launcher

    def forward(self, x_s: Tensor, x_d: Tensor, edge_index: Tensor, edge_attr: Optional[Tensor] = None) -> Tensor:
        """Stand-alone operator call on arbitrary node features (edges may connect any nodes).
        Inside ``GNN`` / ``MSGNN`` the launcher is driven directly on pre-built plans."""
        if torch.is_grad_enabled() and (x_d.requires_grad or any(p.requires_grad for p in self.parameters())):
            from ..autograd import swegnn_autograd
            return swegnn_autograd(self, x_s, x_d, edge_index, edge_attr)
        F = self.edge_output_size
        la = self.launcher()
        FP = la.FP
        N = x_d.shape[0]

        class _G:
            pass
        g = _G()
        g.edge_index, g.x = edge_index, x_d
        plan = self._plans.get(g, 1, False)
        es = plan.edges[0]

        def widen(t):
            t = t.detach().to(torch.float32)
            return t.contiguous() if FP == F else torch.nn.functional.pad(t, (0, FP - F)).contiguous()
        xs, xd = widen(x_s), widen(x_d)
        a = None
        if self.edge_features > 0:
            a = widen(edge_attr)[es.eid.long()].contiguous()
        s_buf = torch.empty(max(es.n_edges, 1), FP, device=xd.device)
        ta, tb, out = torch.empty_like(xd), torch.empty_like(xd), torch.empty_like(xd)
        ptab = (torch.empty(N, 2 * FP, device=xd.device), torch.empty(N, 2 * FP, device=xd.device)) \
            if (FP == 64 and gate_layer0() == "dec") else None
        la.run(es, xs, xd, xd, a, s_buf, False, ta, tb, out, ptab=ptab)
        return out if FP == F else out[:, :F].contiguous()

    def __repr__(self):
        return '{}(node_features={}, edge_features={}, K={}, with_filter_matrix={}, with_gradient={})'.format(
            self.__class__.__name__, self.edge_output_size, self.edge_features, self.K,
            self.with_filter_matrix, self.with_gradient)


class _EncodeDecodeMixin:
    """Pieces shared by GNN and MSGNN: packed encoders / decoder and the workspace."""

    def _setup_packing(self):
        F, FP = self.hid_features, padded_width(self.hid_features)
        self._FP = FP
        pad = {F: FP}
        n_dyn = self.dynamic_node_features
        self._pk_static = PackedMLP(self.static_node_encoder, [(self.static_node_features,) * 2], pad)
        self._pk_dynamic = PackedMLP(self.dynamic_node_encoder, [(n_dyn, n_dyn)], pad)
        self._pk_decoder = PackedMLP(self.node_decoder, [(F, FP)], pad)
        self._pk_edge = None
        if getattr(self, "edge_mlp", False) and hasattr(self, "edge_encoder"):
            ne = self.edge_encoder[0].in_features
            self._pk_edge = PackedMLP(self.edge_encoder, [(ne, ne)], pad)
        if self.static_node_features + 0 > 32 or n_dyn > 32:
            raise NotImplementedError("encoders take at most 32 input columns")
        self._plans = PlanCache()
        self._ws = {}
        self._edge_cache = None
        self._edge_cache_ref = None
        # tcgen05 row MLPs (F = 64, 3-layer encoders / decoder): None where the stack does not fit
        self._tc_static = RowMlpTC.for_encoder(self.static_node_encoder, F)
        self._tc_dynamic = RowMlpTC.for_encoder(self.dynamic_node_encoder, F)
        self._tc_decoder = RowMlpTC.for_decoder(self.node_decoder, F)
        self._tc_edge = RowMlpTC.for_encoder(self.edge_encoder, F) if self._pk_edge is not None else None

    def _encode_nodes(self, x, plan, n_dyn_rows, xs, xd):
        """static_node_encoder / dynamic_node_encoder (reference gnn.py:284-294)."""
        n_static_raw = self.static_node_features - int(bool(self.with_WL))
        if self._tc_static is not None and self._tc_dynamic is not None and rowmlp_backend() == "tc":
            n_cols = x.shape[1]
            # inside a rollout (engine.static_inputs) the static columns and the encoder weights do not change:
            # without the water-level input, x_s is encoded on the first step only
            from .. import engine
            tok = engine._STATIC_TOKEN if engine._XS_STATIC else None
            stamp = None if tok is None else (tok, xs.data_ptr(), id(plan), tuple(
                (p.data_ptr(), p._version) for p in self.static_node_encoder.parameters()))
            if stamp is None or stamp != getattr(self, "_xs_stamp", None):
                self._tc_static.encode(x, 0, n_static_raw, self.with_WL, (n_static_raw - 1, n_cols - 2), plan.perm, 0,
                                       plan.n_nodes, xs)
                self._xs_stamp = stamp
            self._tc_dynamic.encode(x, n_static_raw, n_cols - n_static_raw, False, (0, 0), plan.perm, 0, n_dyn_rows, xd)
        else:
            lib.node_encode_fwd(x, plan.perm, plan.n_nodes, n_static_raw, self.with_WL, n_dyn_rows,
                                self._pk_static.struct(), self._pk_dynamic.struct(), xs, xd, self._FP)

    def _check_input(self, graph):
        x = graph.x
        if not x.is_cuda:
            raise RuntimeError("mswe_gnn_b200 models run on CUDA tensors only (no CPU fallback); "
                               f"graph.x is on {x.device}")
        if x.device.index is not None and x.device.index != torch.cuda.current_device():
            # every launch goes to the raw stream of the CURRENT device (lib._stream); pointers of another device
            # there end in an illegal address
            raise RuntimeError(f"graph.x is on {x.device} but the current CUDA device is {torch.cuda.current_device()}; "
                               f"wrap the call in `with torch.cuda.device({x.device.index}):`")
        if x.dtype != torch.float32:
            raise TypeError("graph.x must be float32")
        if x.shape[1] != self.num_node_features:
            raise ValueError(f"graph.x has {x.shape[1]} columns, model expects {self.num_node_features}")
        for p in self.parameters():
            if p.device != x.device:
                raise RuntimeError(f"model parameters are on {p.device}, graph on {x.device}; call model.to(device)")
            break

    def _workspace(self, plan, names):
        key = (plan.key, plan.n_nodes)
        ws = self._ws.get(key)
        if ws is not None and ws.get("_plan") is not plan:      # same key, different plan object: a recycled address
            ws = None
        if ws is None:
            dev = plan.edges[0].rowptr.device
            FP, N = self._FP, plan.n_nodes
            # node arrays of a partitioned mesh live in the rank's IPC arena (parallel.PeerHalo): neighbours store their
            # boundary rows straight into the halo rows
            alloc = getattr(self, "_ws_alloc", None)
            ws = {n: (alloc(n, N, FP) if alloc is not None else torch.empty(N, FP, dtype=torch.float32, device=dev))
                  for n in names}
            ws["s"] = torch.empty(plan.max_edges, FP, dtype=torch.float32, device=dev)
            # per-node partial tables of the decomposed edge-MLP layer 0 (tcgen05 gate, F = 64)
            ws["ptab"] = (torch.empty(N, 2 * FP, dtype=torch.float32, device=dev),
                          torch.empty(N, 2 * FP, dtype=torch.float32, device=dev)) \
                if (FP == 64 and gate_layer0() == "dec") else None
            ws["a"] = None
            ws["_plan"] = plan
            if len(self._ws) >= 2 and key not in self._ws:
                self._ws.pop(next(iter(self._ws)))
            self._ws[key] = ws
        return ws

    def _encoded_edges(self, plan, graph, ws):
        """Encoded edge features in CSR order, [E_total, FP]; the edge attributes and (outside
        training) the encoder weights are constant over a rollout, so the result is cached."""
        if self._pk_edge is None:
            ea = graph.edge_attr
            if ea is None or ea.shape[1] != self.hid_features or self._FP != self.hid_features:
                raise NotImplementedError("edge_mlp=False needs raw edge features of width hid_features")
            parts = [ea[lo:hi][es.eid.long()] for (lo, hi), es in zip(plan.edge_slices, plan.edges)]
            return torch.cat(parts).contiguous()
        st = self._pk_edge.struct()
        ea = graph.edge_attr
        stamp = (ea.data_ptr(), ea._version, tuple(ea.shape), self._pk_edge._stamp, plan.key, id(plan))
        if ws["a"] is not None and self._edge_cache == stamp and self._edge_cache_ref is ea:
            return ws["a"]
        if ea.dtype != torch.float32 or not ea.is_contiguous():
            raise TypeError("edge_attr must be a contiguous float32 tensor")
        if ws["a"] is None:
            ws["a"] = torch.empty(max(plan.n_edges_total, 1), self._FP, dtype=torch.float32, device=ea.device)
        for (lo, hi), es in zip(plan.edge_slices, plan.edges):
            if es.n_edges:
                if self._tc_edge is not None and rowmlp_backend() == "tc":
                    self._tc_edge.encode(ea[lo:hi], 0, ea.shape[1], False, (0, 0), es.eid, 0, es.n_edges, ws["a"][lo:hi])
                else:
                    lib.edge_encode_fwd(ea[lo:hi], es.eid, es.n_edges, st, ws["a"][lo:hi], self._FP)
        self._edge_cache = stamp
        self._edge_cache_ref = ea            # identity, not just address: a freed tensor's address can be reused
        return ws["a"]

    def _decode(self, h, act_name, act_module, x, plan, pred, step_ptr=None, pred_stride=0, x_next=None, owned_only=False):
        slope = act_module.weight if isinstance(act_module, nn.PReLU) else None
        res_w = self.residual_weights.detach().contiguous() if self._residual_mode() in (1, 2) else None
        if self._tc_decoder is not None and rowmlp_backend() == "tc" and x.shape[1] <= 16:
            if owned_only and plan.scale_owned is not None:
                # the halo rows of x / pred belong to their owners (who store them into our copy): decode owned rows only
                for lo, n in zip(plan.scale_lo, plan.scale_owned):
                    self._tc_decoder.decode(h, ACT_CODES[act_name], slope, x, plan.perm, n, self.previous_t,
                                            self._residual_mode(), res_w, 1e-4, pred, step_ptr, pred_stride, x_next, row_lo=lo)
                return
            self._tc_decoder.decode(h, ACT_CODES[act_name], slope, x, plan.perm, plan.n_nodes, self.previous_t,
                                    self._residual_mode(), res_w, 1e-4, pred, step_ptr, pred_stride, x_next)
            return
        if owned_only:
            raise NotImplementedError("the peer-memory halo transport needs the tcgen05 decoder (F = 64, 3-layer decoder)")
        lib.decode_head_fwd(h, ACT_CODES[act_name], slope, self._pk_decoder.struct(), x, plan.perm, plan.n_nodes,
                            self.previous_t, self._residual_mode(), res_w, 1e-4, pred, step_ptr, pred_stride,
                            x_next, self._FP)

    def _needs_grad(self, graph) -> bool:
        return torch.is_grad_enabled() and (graph.x.requires_grad or any(p.requires_grad for p in self.parameters()))


class GNN(BaseFloodModel, _EncodeDecodeMixin):
    '''
    SWE-GNN encoder-processor-decoder (reference ``gnn.py:13-152``)
    ------
    num_node_features: int, number of features per node
    num_edge_features: int, number of features per edge
    hid_features: int, number of features per node (and edge) in the GNN layers
    K: int, K-hop neighbourhood
    n_GNN_layers: int, number of GNN layers
    type_GNN: only "SWEGNN" (the learned graph shift operator) is part of the hot path; the PyG
        baselines "GNN_A", "GNN_L", "GAT" are not built
    edge_mlp: bool, adds MLP as edge encoder
    mlp_layers / mlp_activation / gnn_activation / with_WL / normalize / with_filter_matrix /
    with_gradient / base_model_kwargs: as in the reference
    '''

    _WS_NAMES = ["xs", "h0", "h1", "ta", "tb"]          # node arrays of one forward (engine / parallel.PeerHalo)

    def __init__(self, num_node_features, num_edge_features, hid_features=32, K=2, n_GNN_layers=2, type_GNN="SWEGNN",
                 mlp_layers=1, mlp_activation='prelu', gnn_activation='prelu', dropout=0,
                 with_WL=True, normalize=True, with_filter_matrix=True, edge_mlp=True,
                 with_gradient=True, **base_model_kwargs):
        super(GNN, self).__init__(**base_model_kwargs)
        if type_GNN != "SWEGNN":
            if type_GNN in ("GNN_A", "GNN_L", "GAT"):
                raise NotImplementedError(f"type_GNN='{type_GNN}' (a PyG library convolution) is outside the "
                                          "B200 hot path; only 'SWEGNN' is built")
            raise ValueError("Only 'GNN_A', 'GNN_L', 'GAT', and 'SWEGNN' are valid for now")
        self.type_model = "GNN"
        self.hid_features = hid_features
        self.num_node_features = num_node_features
        self.num_edge_features = num_edge_features
        self.type_GNN = type_GNN
        self.edge_mlp = edge_mlp
        self.with_WL = with_WL
        self._gnn_activation_name = gnn_activation
        self.dynamic_node_features = self.previous_t * self.out_dim
        self.static_node_features = num_node_features - self.dynamic_node_features + self.with_WL
        dev = self.device

        if edge_mlp:
            self.num_edge_features = hid_features
            self.edge_encoder = make_mlp(num_edge_features, hid_features, hid_features, n_layers=mlp_layers, bias=True,
                                         activation=mlp_activation, device=dev)
        self.dynamic_node_encoder = make_mlp(self.dynamic_node_features, hid_features, hid_features,
                                             n_layers=mlp_layers, activation=mlp_activation, device=dev)
        self.static_node_encoder = make_mlp(self.static_node_features, hid_features, hid_features, n_layers=2,
                                            bias=True, activation=mlp_activation, device=dev)
        self.gnn_processor = nn.ModuleList(
            SWEGNN(hid_features, hid_features, self.num_edge_features, K=K, device=dev, n_layers=mlp_layers,
                   activation=mlp_activation, bias=True, normalize=normalize,
                   with_filter_matrix=with_filter_matrix, with_gradient=with_gradient)
            for _ in range(n_GNN_layers))
        self.gnn_activation = activation_functions(gnn_activation, device=dev)
        self.node_decoder = make_mlp(hid_features, self.out_dim, hid_features, n_layers=mlp_layers, dropout=dropout,
                                     activation=mlp_activation, device=dev)
        self._setup_packing()

    def forward(self, graph):
        """graph: PyG-like ``Data``/``Batch`` with ``x [N, C]``, ``edge_index [2, E]``,
        ``edge_attr [E, n_e]`` → ``[N, 2]`` (water depth, |discharge|)."""
        self._check_input(graph)
        if self._needs_grad(graph):
            from ..autograd import model_autograd
            return model_autograd(self, graph)
        plan = self._plans.get(graph, 1, False)
        pred = torch.empty(plan.n_nodes, self.out_dim, dtype=torch.float32, device=graph.x.device)
        self._launch(plan, graph, graph.x.contiguous(), pred)
        return pred

    def _launch(self, plan, graph, x, pred, step_ptr=None, pred_stride=0, x_next=None, halo=None):
        FP = self._FP
        n_layers = len(self.gnn_processor)
        ws = self._workspace(plan, self._WS_NAMES)
        a = self._encoded_edges(plan, graph, ws)
        N = plan.n_nodes
        self._encode_nodes(x, plan, N, ws["xs"], ws["h0"])
        cur, nxt = ws["h0"], ws["h1"]
        act_mod = self.gnn_activation
        slope = act_mod.weight if isinstance(act_mod, nn.PReLU) else None
        es = plan.edges[0]
        for li, conv in enumerate(self.gnn_processor):
            if halo is not None and li > 0:
                halo.exchange(cur, 0)
            conv.launcher().run(es, ws["xs"], cur, cur, a, ws["s"], False, ws["ta"], ws["tb"], nxt,
                                act_code=ACT_CODES[self._gnn_activation_name], act_slope=slope, halo=halo, scale=0,
                                ptab=ws["ptab"])
            cur, nxt = nxt, cur
        self._decode(cur, None, None, x, plan, pred, step_ptr, pred_stride, x_next,
                     owned_only=bool(getattr(halo, "owned_only", False)))


class MSGNN(BaseFloodModel, _EncodeDecodeMixin):
    '''
    Multi-scale mSWE-GNN encoder-processor-decoder (reference ``gnn.py:154-350``): one SWEGNN per
    scale on the way down (fine → coarse, mean-pooled between scales) and on the way up (coarse →
    fine, un-pooled by a learned SWEGNN on the inter-scale edges plus skip connections).
    ------
    num_node_features, num_edge_features, num_scales, hid_features, K (int or list of length
    num_scales or 2*num_scales-1), mlp_layers, mlp_activation, gnn_activation, learned_pooling,
    skip_connections, with_WL, normalize, with_filter_matrix, edge_mlp, with_gradient,
    base_model_kwargs: as in the reference
    '''

    _WS_NAMES = ["xs", "cur", "down", "up", "ta", "tb"]  # node arrays of one forward (engine / parallel.PeerHalo)

    def __init__(self, num_node_features, num_edge_features, num_scales, hid_features=32, K=2,
                 mlp_layers=2, mlp_activation='prelu', gnn_activation='tanh',
                 learned_pooling=False, skip_connections=True,
                 with_WL=False, normalize=True, with_filter_matrix=True, edge_mlp=True,
                 with_gradient=True, **base_model_kwargs):
        super(MSGNN, self).__init__(**base_model_kwargs)
        self.type_model = "MSGNN"
        self.hid_features = hid_features
        self.num_node_features = num_node_features
        self.edge_mlp = edge_mlp
        self.with_WL = with_WL
        self.num_scales = num_scales
        self._gnn_activation_name = gnn_activation
        self.dynamic_node_features = self.previous_t * self.NUM_WATER_VARS
        self.static_node_features = num_node_features - self.dynamic_node_features + self.with_WL
        self.learned_pooling = learned_pooling
        self.skip_connections = skip_connections
        self.K = [K] * num_scales if isinstance(K, int) else list(K)
        self.K = self.K + self.K[::-1][1:]          # mirrored for the coarse-to-fine pass
        assert len(self.K) == num_scales * 2 - 1, "K must be an int or a list of length num_scales or num_scales*2-1"
        if learned_pooling:
            raise NotImplementedError("learned_pooling=True is not implemented by the B200 kernels "
                                      "(config.yaml default is False)")
        dev = self.device

        if edge_mlp:
            self.edge_encoder = make_mlp(num_edge_features, hid_features, hid_features, n_layers=mlp_layers, bias=True,
                                         activation=mlp_activation, device=dev)
            num_edge_features = hid_features
        self.dynamic_node_encoder = make_mlp(self.dynamic_node_features, hid_features, hid_features,
                                             n_layers=mlp_layers, activation=mlp_activation, device=dev)
        self.static_node_encoder = make_mlp(self.static_node_features, hid_features, hid_features,
                                            n_layers=mlp_layers, bias=True, activation=mlp_activation, device=dev)
        self.intra_scale_gnn = nn.ModuleList(
            SWEGNN(hid_features, hid_features, 0, K=1, n_layers=mlp_layers, activation=mlp_activation, bias=True,
                   normalize=True, with_filter_matrix=False, with_gradient=False, device=dev)
            for _ in range(num_scales - 1))
        # (the reference builds these without device=, i.e. on the CPU, and relies on model.to())
        self.gnn_processor = nn.ModuleList(
            SWEGNN(hid_features, hid_features, num_edge_features, K=k, n_layers=mlp_layers,
                   activation=mlp_activation, bias=True, normalize=normalize,
                   with_filter_matrix=with_filter_matrix, with_gradient=with_gradient)
            for k in self.K)
        self.gnn_activation = activation_functions(gnn_activation, device=dev)
        self.node_decoder = make_mlp(hid_features, self.out_dim, hid_features, n_layers=mlp_layers, dropout=0,
                                     activation=mlp_activation, device=dev)
        self._setup_packing()

    def _create_scale_mask(self, data):
        """int32 [N] scale id per node (reference ``utils/dataset.py:615-638``); the kernels do
        not need it, it is kept for callers (``training/train.py:164``)."""
        n = data.x.size(0)
        ptr = data.node_ptr.reshape(-1, data.node_ptr.shape[-1]).tolist()
        mask = torch.zeros(n, dtype=torch.int, device=data.x.device)
        for s in range(self.num_scales):
            for row in ptr:
                mask[row[s]:row[s + 1]] = s
        return mask

    def forward(self, graph):
        """graph: PyG-like ``Data`` (or a batch adapted by ``adapt_batch_training``) with the
        fields of SURVEY.md Appendix C → ``[N, 2]`` predictions at every scale."""
        self._check_input(graph)
        if self._needs_grad(graph):
            from ..autograd import model_autograd
            return model_autograd(self, graph)
        plan = self._plans.get(graph, self.num_scales, True)
        pred = torch.empty(plan.n_nodes, self.out_dim, dtype=torch.float32, device=graph.x.device)
        self._launch(plan, graph, graph.x.contiguous(), pred)
        return pred

    def _launch(self, plan, graph, x, pred, step_ptr=None, pred_stride=0, x_next=None, halo=None):
        """halo: ``parallel.HaloExchanger`` when `graph` is one rank's part of a partitioned mesh."""
        FP, S = self._FP, self.num_scales
        ws = self._workspace(plan, self._WS_NAMES)
        a = self._encoded_edges(plan, graph, ws)
        xs, cur, down, up, ta, tb, s_buf = ws["xs"], ws["cur"], ws["down"], ws["up"], ws["ta"], ws["tb"], ws["s"]
        self._encode_nodes(x, plan, plan.scale_n[0], xs, cur)

        def a_of(s):
            lo, hi = plan.edge_slices[s]
            return a[lo:hi] if hi > lo else None

        # fine -> coarse
        cross = halo.part.inter_cross if halo is not None else None
        for i in range(S - 1):
            es = plan.edges[i]
            if halo is not None and i > 0:
                halo.exchange(cur, i)                     # pooled rows of the halo nodes come from their owners
            self.gnn_processor[i].launcher().run(es, xs, cur, cur, a_of(i), s_buf, False, ta, tb, down, halo=halo, scale=i,
                                                 ptab=ws["ptab"])
            if halo is not None and cross[i]:
                halo.exchange(down, i)                    # a child owned by another rank (real meshes, App. D-5)
            pe = plan.pool[i]
            lib.pool_mean_fwd(down, pe.rowptr, pe.src, pe.dst_lo, pe.n_dst, cur, FP)
        # coarse -> fine
        for i in range(S):
            s = S - 1 - i
            es = plan.edges[s]
            if halo is not None and S > 1:
                halo.exchange(cur, s)
            self.gnn_processor[S - 1 + i].launcher().run(es, xs, cur, cur, a_of(s), s_buf, False, ta, tb, up, halo=halo, scale=s,
                                                         ptab=ws["ptab"])
            if i < S - 1:
                if halo is not None and cross[s - 1]:
                    halo.exchange(up, s)
                ue = plan.unpool[s - 1]
                # x_d[fine] is identically zero here (nothing wrote the finer rows since the last
                # pooling zeroed them), so the gate skips that block and the hop starts from 0
                self.intra_scale_gnn[i].launcher().run(ue, xs, up, None, None, s_buf, True, ta, tb, cur,
                                                       addend=down if self.skip_connections else None, ptab=ws["ptab"])
        self._decode(up, self._gnn_activation_name, self.gnn_activation, x, plan, pred, step_ptr, pred_stride, x_next,
                     owned_only=bool(getattr(halo, "owned_only", False)))
