/*
 * swe_gnn_b200.h — C ABI of libswe_gnn_b200.so (sm_100a kernels for the mSWE-GNN hot path).
 *
 * The reference (sdat2/mSWE-GNN) is pure Python; it has no FFI.  The "interface each entry point
 * replaces" is therefore a span of ATen calls inside the reference's Python methods; each
 * declaration below cites that span (paths relative to the reference root).  The reference-side
 * binding a maintainer would add is a ctypes stub — see INTEGRATION.md.
 *
 * Conventions:
 *   - plain pointers and sizes only; every buffer is caller-allocated DEVICE memory, contiguous
 *     row-major, 16-byte aligned; feature width F ∈ {16, 32, 64} (host pads other widths);
 *   - node / edge ids are int32 in "plan order" (see swe_csr_build); features are fp32;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), performs no host
 *     synchronisation and no data-dependent host control flow, so call sequences are CUDA-graph
 *     capturable; the library keeps no mutable global state except a thread-local error string;
 *   - return value: 0 = ok, <0 = invalid argument (SWE_E_*), >0 = cudaError_t of the launch;
 *     no C++ exception crosses the boundary; swe_last_error() describes the last failure on the
 *     calling thread.
 */
#ifndef SWE_GNN_B200_H
#define SWE_GNN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWE_ABI_VERSION 1
#define SWE_MAX_LAYERS 8

#define SWE_E_INVAL   (-1)   /* bad size / null pointer / unsupported width                */
#define SWE_E_ALIGN   (-2)   /* pointer not 16-byte aligned                                 */
#define SWE_E_UNSUPP  (-3)   /* configuration outside what the kernels implement            */

/* activation codes — reference models/models.py:149-169 (`activation_functions`) */
enum swe_act {
    SWE_ACT_NONE = 0, SWE_ACT_PRELU = 1, SWE_ACT_RELU = 2, SWE_ACT_TANH = 3,
    SWE_ACT_LEAKYRELU = 4 /* slope 0.1 */, SWE_ACT_ELU = 5, SWE_ACT_SWISH = 6, SWE_ACT_SIGMOID = 7
};

/* One Linear(+bias)+activation layer of a `make_mlp` stack (models/models.py:121-146).
 * `wt` is the TRANSPOSED weight, k-major: wt[k * n_out + n] = Linear.weight[n][k], with k_in
 * rounded up to a multiple of 4 and zero rows appended (swe_pack_linear produces it).
 * `slope` points at the PReLU parameter in device memory (read on the device; may be NULL for
 * non-PReLU activations); `bias` may be NULL. */
typedef struct swe_layer {
    const float* wt;
    const float* bias;
    const float* slope;
    int32_t k_in;      /* padded to a multiple of 4 */
    int32_t n_out;
    int32_t act;       /* enum swe_act */
    int32_t _pad;
} swe_layer_t;

typedef struct swe_mlp {
    int32_t n_layers;
    int32_t _pad;
    swe_layer_t layer[SWE_MAX_LAYERS];
} swe_mlp_t;

int         swe_abi_version(void);
const char* swe_last_error(void);
/* Compile-time facts a caller may assert on: target arch string ("sm_100a"). */
const char* swe_build_arch(void);

/* ---------------------------------------------------------------------------------------------
 * Plan construction (integer work; results are bit-exact against oracle/plan_oracle.py)
 * ------------------------------------------------------------------------------------------- */

/* wt[k][n] = w[n][k] for k < k_in, 0 for k_in <= k < k_pad.  Packs a torch Linear weight
 * [n_out, k_in] for the kernels. */
int swe_pack_linear(const float* w, int32_t n_out, int32_t k_in, int32_t k_pad, float* wt, void* stream);

/* Destination-CSR of one edge set, STABLE in the original edge order, so that a sequential
 * in-segment sum reproduces CPU `Tensor.scatter_add_` ordering bit for bit.
 * Replaces: the implicit edge traversal order of `scatter(shift_sum, col[...], reduce='sum')`
 * (models/gnn.py:437-438) and of `MSGNN._pooling` (models/gnn.py:256).
 *   row, col      : int64 [E] original global node ids (edge = row -> col, aggregated at col)
 *   node_map      : int32 [n_nodes_total] original id -> plan id, or NULL for identity
 *   dst_lo,n_dst  : plan-id range the destinations must fall in
 *   src_lo,src_hi : plan-id range the sources must fall in
 *   rowptr        : int32 [n_dst+1] out;  src,dst,eid : int32 [E] out (plan ids, original edge id)
 *   err_flag      : int32 [1] device, incremented for every edge outside the ranges
 *   ws            : scratch of at least swe_csr_build_ws_bytes(E, n_dst) bytes
 * by_row != 0 builds the CSR keyed on `row` instead (transposed CSR used by the backward pass);
 * then `dst` receives the row ids and `src` the col ids. */
size_t swe_csr_build_ws_bytes(int64_t n_edges, int32_t n_dst);
int swe_csr_build(const int64_t* row, const int64_t* col, int64_t n_edges, const int32_t* node_map,
                  int32_t dst_lo, int32_t n_dst, int32_t src_lo, int32_t src_hi, int32_t by_row,
                  int32_t* rowptr, int32_t* src, int32_t* dst, int32_t* eid, int32_t* err_flag,
                  void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Forward kernels
 * ------------------------------------------------------------------------------------------- */

/* Node encoders.  Replaces models/gnn.py:284-294 (MSGNN) / :113-125 (GNN): split x into static
 * and dynamic columns, append WL = x[:, n_static_raw-1] + x[:, n_cols-2] when with_wl, run
 * static_node_encoder and dynamic_node_encoder.
 *   x        : [*, n_cols] input rows; row of plan node i is perm[i] (perm NULL = identity)
 *   xs_out   : [n_nodes, F] ; xd_out : [n_dyn_rows, F] (dynamic encoding is only needed for the
 *              first n_dyn_rows plan nodes: the finest scale of an MSGNN, all nodes of a GNN) */
int swe_node_encode_fwd(const float* x, int32_t n_cols, const int32_t* perm, int32_t n_nodes,
                        int32_t n_static_raw, int32_t with_wl, int32_t n_dyn_rows,
                        const swe_mlp_t* static_mlp, const swe_mlp_t* dynamic_mlp,
                        float* xs_out, float* xd_out, int32_t F, void* stream);

/* Edge encoder.  Replaces models/gnn.py:281-282 / :109-110: edge_encoder(edge_attr), written in
 * CSR order: a_out[p] = MLP(edge_attr[eid[p]]). */
int swe_edge_encode_fwd(const float* edge_attr, int32_t n_edge_feat, const int32_t* eid, int64_t n_edges,
                        const swe_mlp_t* mlp, float* a_out, int32_t F, void* stream);

/* Edge gate s_ij.  Replaces models/gnn.py:414-426 (gather x_s[row], x_s[col], x_d[row],
 * x_d[col], edge_attr -> cat -> edge_mlp -> L2 normalise -> NaN->0), evaluated ONCE per SWEGNN
 * call (its inputs do not depend on the hop index).
 *   xs : [n_nodes, F];  xd_src / xd_dst : [n_nodes, F] arrays x_d[row] / x_d[col] are read from
 *   (normally the same array; xd_dst == NULL means "x_d[col] is known to be zero", which is the
 *   case for every un-pool call of MSGNN.forward (gnn.py:327), and skips that input block)
 *   a : [E, F] encoded edge features in CSR order or NULL (edge_features=0, gnn.py:419)
 *   src, dst : int32 [E] plan ids in CSR order (src = `row`, dst = `col`);  s_out : [E, F]
 *   mlp layer 0 always has k_in = 5F (4F when a == NULL): its weight rows follow the reference
 *   column order [x_s[row] | x_s[col] | x_d[row] | x_d[col] | e]. */
int swe_edge_gate_fwd(const float* xs, const float* xd_src, const float* xd_dst, const float* a,
                      const int32_t* src, const int32_t* dst, int64_t n_edges, const swe_mlp_t* mlp,
                      int32_t normalize, float* s_out, int32_t F, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Edge gate on tcgen05 tensor cores (F = 64, 3-layer edge MLP 5F|4F -> 2F -> 2F -> F: the default
 * config.yaml model).  Same contract and reference span as swe_edge_gate_fwd (models/gnn.py:414-426);
 * products are 3xTF32 (error-free hi/lo splits, fp32 accumulation in TMEM), see DESIGN.md.
 * ------------------------------------------------------------------------------------------- */

/* Bytes of the packed weight image for a first layer with k1 input columns (256 or 320). */
size_t swe_gate_tc_image_bytes(int32_t k1);

/* Packs edge_mlp.{0,2,4}.{weight,bias} (torch Linear layout [n_out, k_in]) into the image the
 * kernel streams: per 32-column K-chunk a hi tile and a lo tile in the UMMA K-major SWIZZLE_128B
 * shared-memory layout, followed by the three bias vectors. */
int swe_gate_tc_pack(const float* w1, int32_t k1, const float* b1, const float* w2, const float* b2,
                     const float* w3, const float* b3, void* image, void* stream);

/* act3: HOST array of 3 activation codes; slope3: HOST array of 3 DEVICE pointers to the PReLU
 * parameters (NULL for other activations).  dbg: NULL, or a device buffer of 128*128*2 + 128*64
 * floats receiving the raw accumulators of the first tile (tests only). */
int swe_edge_gate_tc_fwd(const float* xs, const float* xd_src, const float* xd_dst, const float* a,
                         const int32_t* src, const int32_t* dst, int64_t n_edges, const void* image,
                         int32_t k1, const int32_t* act3, const float* const* slope3, int32_t normalize,
                         float* s_out, float* dbg, void* stream);

/* swe_edge_gate_tc_fwd restricted to the 128-edge tiles tile_list[1 .. tile_list[0]] (device memory; tile t = edges
 * [128 t, 128 t + 128)).  A fixed launch whatever the list holds: used as the range-guard fallback of
 * swe_edge_gate_tc16_fwd inside captured graphs. */
int swe_edge_gate_tc_fwd_listed(const float* xs, const float* xd_src, const float* xd_dst, const float* a,
                                const int32_t* src, const int32_t* dst, int64_t n_edges, const void* image,
                                int32_t k1, const int32_t* act3, const float* const* slope3, int32_t normalize,
                                float* s_out, const int32_t* tile_list, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Edge gate on tcgen05 with fp16 hi/lo splits (kind::f16, K = 16 per instruction: half the tensor-core
 * instructions of the 3xTF32 kernel, same three products, same fp32 accumulation).  Same contract and
 * reference span as swe_edge_gate_fwd (models/gnn.py:414-426).  Operands are scaled by powers of two into fp16's
 * exponent window: per matrix for the weights (wmax3 = HOST array of max |w| of the three layers, read when the
 * image is packed), per row for the hidden activations, none for the gathered layer-0 inputs — whose
 * rows are range-checked while they are converted: tiles with a row maximum outside [2^-5, 2^15] are listed in
 * flag_ws ((number of tiles + 1) int32 of scratch) and re-evaluated by swe_edge_gate_tc_fwd_listed from image_tf32
 * (the swe_gate_tc_pack image of the same weights).  flag_ws == NULL skips the guard (tests).
 * ------------------------------------------------------------------------------------------- */
size_t swe_gate_tc16_image_bytes(int32_t k1);
int swe_gate_tc16_pack(const float* w1, int32_t k1, const float* b1, const float* w2, const float* b2,
                       const float* w3, const float* b3, const float* wmax3, void* image, void* stream);
int swe_edge_gate_tc16_fwd(const float* xs, const float* xd_src, const float* xd_dst, const float* a,
                           const int32_t* src, const int32_t* dst, int64_t n_edges, const void* image16,
                           const void* image_tf32, int32_t k1, const int32_t* act3, const float* const* slope3,
                           int32_t normalize, float* s_out, float* dbg, int32_t* flag_ws, void* stream);

/* Row linear out[row_lo + r, :] = x[row_lo + r, :] · Wᵀ for r < n_rows (x, out: [*, 64] fp32; no bias, no activation): the
 * o_0 = x_d W_0ᵀ at the head of every SWEGNN call (models/gnn.py:401-402) as a STREAMING kernel — whole 128-row tiles
 * enter a 3-deep shared-memory ring by cp.async.bulk, fp16 hi/lo operands with a per-row power-of-two scale, 12
 * tcgen05.mma per tile (same arithmetic contract as swe_propagate_hop_tc16_fwd's filter: rel 1e-5 of fp32).
 * w_image: swe_hop_tc16_pack of W [64, 64]. */
int swe_row_linear_tc16(const float* x, int64_t row_lo, int64_t n_rows, const void* w_image, float* out, void* stream);

/* Decomposed first layer of the edge MLP: W1·[x_s[r]|x_s[c]|x_d[r]|x_d[c]|a] = P_src[r] + P_dst[c] + E·a with
 *   P_src[n] = A·x_s[n] + C·x_d[n]   (role 0),   P_dst[n] = B·x_s[n] + D·x_d[n]   (role 1; xd NULL drops D·x_d),
 * evaluated once per NODE (2·128² MAC) instead of once per edge.  swe_gate_partials_tc writes one table
 * p_out[row_lo + i, 0:128) for i < n_rows; swe_edge_gate_tc_dec_fwd is swe_edge_gate_tc_fwd reading the tables
 * (same reference span models/gnn.py:414-426; the sum is re-associated, ~1e-7 relative). */
int swe_gate_partials_tc(const float* xs, const float* xd, int32_t row_lo, int32_t n_rows, const void* image,
                         int32_t k1, int32_t role, float* p_out, void* stream);
int swe_edge_gate_tc_dec_fwd(const float* p_src, const float* p_dst, const float* a, const int32_t* src,
                             const int32_t* dst, int64_t n_edges, const void* image, int32_t k1,
                             const int32_t* act3, const float* const* slope3, int32_t normalize, float* s_out,
                             void* stream);

/* ---------------------------------------------------------------------------------------------
 * Hop with the F×F filter on tcgen05 tensor cores (F = 64; 3xTF32, fp32 accumulation in TMEM).
 * Same contract and reference span as swe_propagate_hop_fwd with wt != NULL (models/gnn.py:428-443);
 * the aggregation is bit-identical to it, the filter product differs by ~1e-7 relative.
 * ------------------------------------------------------------------------------------------- */
size_t swe_hop_tc_image_bytes(void);
/* w: filter_matrix[k].weight, torch layout [64, 64] -> pre-swizzled hi|lo TF32 image */
int swe_hop_tc_pack(const float* w, void* image, void* stream);
int swe_propagate_hop_tc_fwd(const float* o_src, const float* o_dst, const float* s, const int32_t* rowptr,
                             const int32_t* src, int32_t dst_lo, int32_t n_dst, const void* w_image,
                             int32_t with_gradient, int32_t upwind, const float* addend, int32_t act,
                             const float* slope, float* agg_out, float* out, void* stream);

/* The same hop with the filter as fp16 hi/lo splits on kind::f16 (K = 16 per instruction, 64-byte swizzle):
 * agg rows are scaled per row, the filter per matrix (wmax = max |w|, a HOST value read when the image is packed) by powers
 * of two; same contract, reference span (models/gnn.py:428-443) and accuracy as swe_propagate_hop_tc_fwd. */
size_t swe_hop_tc16_image_bytes(void);
int swe_hop_tc16_pack(const float* w, float wmax, void* image, void* stream);
int swe_propagate_hop_tc16_fwd(const float* o_src, const float* o_dst, const float* s, const int32_t* rowptr,
                               const int32_t* src, int32_t dst_lo, int32_t n_dst, const void* w_image,
                               int32_t with_gradient, int32_t upwind, const float* addend, int32_t act,
                               const float* slope, float* agg_out, float* out, void* stream);
/* s-ring edition of the fp16 hop (same image, same contract, bit-identical results): every gather warp owns 8 contiguous
 * nodes of a tile and streams their run of gate rows s_p (CSR order: one contiguous run, 768 of the node's 1280 bytes)
 * into its own shared-memory buffers with cp.async.bulk one tile ahead; rowptr two tiles and src ids one tile ahead. */
int swe_propagate_hop_tc16s_fwd(const float* o_src, const float* o_dst, const float* s, const int32_t* rowptr,
                                const int32_t* src, int32_t dst_lo, int32_t n_dst, const void* w_image,
                                int32_t with_gradient, int32_t upwind, const float* addend, int32_t act,
                                const float* slope, float* agg_out, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Row MLPs on tcgen05 (F = 64): encoders (models/gnn.py:281-294), filter_matrix[0] (gnn.py:401-402) and the decoder
 * head (gnn.py:339-348 + models/models.py:50-91 + utils/dataset.py:508-529) as ONE row-streaming kernel:
 *   X0 = act_in(x_rows[row_lo + r])   |   act_first(w_first · raw(r) + b_first)   (first Linear of an encoder, <= 8 raw
 *        inputs incl. the optional WL = raw[wl_col_a] + raw[wl_col_b], evaluated on CUDA cores)
 *   X1 = act[0](W_0 X0 + bias[0]),  X2 = act[1](W_1 X1 + bias[1]) (n_tc = 2)      64 -> 64 layers, 3xTF32 on tcgen05
 *   out_rows[row_lo + r] = X_last     |   head != 0: pred / x_next exactly like swe_decode_head_fwd
 * img[l]: swe_hop_tc_pack image of the layer's [64, 64] weight.
 * ------------------------------------------------------------------------------------------- */
typedef struct swe_rowmlp {
    const float* x_rows; int32_t act_in; int32_t _pad0; const float* slope_in;
    const float* raw; int32_t raw_ld; int32_t raw_col0; int32_t raw_cols; int32_t with_wl; int32_t wl_col_a; int32_t wl_col_b;
    const int32_t* perm;
    const float* w_first; const float* b_first; int32_t act_first; int32_t _pad1; const float* slope_first;
    int32_t row_lo; int32_t _pad2; int64_t n_rows;
    int32_t n_tc; int32_t _pad3; const void* img[2]; const float* bias[2]; int32_t act[2]; const float* slope[2];
    float* out_rows;
    int32_t head; int32_t act_head; const float* w_head; const float* b_head; const float* slope_head;
    const float* x0; int32_t n_cols; int32_t previous_t; const int32_t* head_perm; int32_t res_mode; float eps;
    const float* res_w; float* pred; const int32_t* step_ptr; int64_t pred_step_stride; float* x_next;
} swe_rowmlp_t;
int swe_row_mlp_tc(const swe_rowmlp_t* desc, void* stream);

/* The two-layer shapes of swe_row_mlp_tc (node encoders: raw rows of 8 floats -> 64 -> 64 -> 64; decoder: rows -> 64 -> 64 ->
 * head) as a STREAMING kernel with fp16 hi/lo operands (per-row power-of-two scale, 12 MMAs per layer, layer-1 operand
 * in TMEM, inputs through a shared-memory ring filled by asynchronous copies several tiles ahead).  Same descriptor;
 * img16[0], img16[1] = swe_hop_tc16_pack images of the two [64, 64] matrices (desc->img is not read).  Returns
 * SWE_E_UNSUPP for shapes it does not cover (n_tc != 2, activations outside none / relu / leakyrelu / prelu in the
 * layers or the head — tanh is accepted as the decoder's input activation —, raw rows that are not 8 aligned floats);
 * the caller then uses swe_row_mlp_tc. */
int swe_row_mlp_tc16(const swe_rowmlp_t* desc, const void* const* img16, void* stream);

/* out[dst_lo + i] = x[dst_lo + i] · Wᵀ for i < n_rows.  Replaces models/gnn.py:401-402
 * (filter_matrix[0]).  wt is the packed (k-major) F×F weight. */
int swe_node_linear_fwd(const float* x, int32_t row_lo, int32_t n_rows, const float* wt, float* out,
                        int32_t F, void* stream);

/* One hop.  Replaces models/gnn.py:428-443:
 *   agg[c]  = Σ_{p in CSR segment of c, original edge order} s[p] ⊙ (o[c] − o[src[p]])   (with_gradient)
 *           | Σ s[p] ⊙ o[src[p]]                                                        (otherwise)
 *   out[c]  = act( o[c] + agg[c]·Wᵀ (or agg[c] when wt == NULL) + addend[c] )
 * for c in [dst_lo, dst_lo + n_dst).  upwind != 0 clamps (o[c] − o[src]) at >= 0 (gnn.py:431-432).
 * `o_dst` may differ from `o_src` (un-pool reads coarse rows and fine rows of different arrays);
 * pass o_dst == NULL when the destination rows are known to be zero.  `addend` (skip connection,
 * gnn.py:330-331) and the output activation (GNN: gnn.py:135-136) may be NULL / SWE_ACT_NONE. */
int swe_propagate_hop_fwd(const float* o_src, const float* o_dst, const float* s, const int32_t* rowptr,
                          const int32_t* src, int32_t dst_lo, int32_t n_dst, const float* wt,
                          int32_t with_gradient, int32_t upwind, const float* addend,
                          int32_t act, const float* slope, float* out, int32_t F, void* stream);

/* Mean pooling onto the coarser scale.  Replaces MSGNN._pooling, models/gnn.py:256
 * (scatter(x[row_fine], col_coarse, reduce='mean')): out[c] = Σ x[fine[p]] / max(1, count). */
int swe_pool_mean_fwd(const float* x, const int32_t* rowptr, const int32_t* fine, int32_t coarse_lo,
                      int32_t n_coarse, float* out, int32_t F, void* stream);

/* Decoder head.  Replaces models/gnn.py:332-348 (MSGNN) / :141-150 (GNN) and, when x_next is
 * given, utils/dataset.py:508-529 (`use_prediction`):
 *   y = relu( node_decoder(act_in(h)) + residual(x0) );  h' = y_h·[|y_h| > eps];  q' = y_q·[y_h != 0]
 *   pred[perm[i]] = (h', q');  x_next[perm[i]] = [static cols, window shifted left by 2, h', q']
 *   h        : [n_nodes, F] processor output (x_up) in plan order
 *   x0       : [*, n_cols] the step's input rows (original order)
 *   res_mode : 0 none, 1 learned (res_w [previous_t] shared by both vars), 2 'all' (res_w
 *              [previous_t,2]), 3 unweighted last step (models/models.py:50-77)
 *   pred     : [*, 2] original order;  x_next : [*, n_cols] or NULL (may alias x0)
 *   step_ptr : NULL, or int32 [1] DEVICE rollout step counter: the prediction is then written at
 *              pred + (*step_ptr) * pred_step_stride (floats), which lets one captured CUDA graph
 *              serve every step of training/train.py:87-93 */
int swe_decode_head_fwd(const float* h, int32_t act_in, const float* slope_in, const swe_mlp_t* decoder,
                        const float* x0, int32_t n_cols, const int32_t* perm, int32_t n_nodes,
                        int32_t previous_t, int32_t res_mode, const float* res_w, float eps,
                        float* pred, const int32_t* step_ptr, int64_t pred_step_stride,
                        float* x_next, int32_t F, void* stream);

/* Boundary-condition injection of the rollout loop.  Replaces utils/dataset.py:486-497:
 * x[node_bc[b], n_static_raw + 2*t + (type_bc-1)] = bc[b][t][step] for t < previous_t.
 *   bc : [n_bc, previous_t, n_steps_total] ; step_ptr : int32 [1] DEVICE step counter (NULL = 0) so
 *   that a captured CUDA graph can be replayed for every time step. */
int swe_apply_bc(float* x, int32_t n_cols, int32_t n_static_raw, int32_t previous_t, int32_t type_bc,
                 const int64_t* node_bc, int32_t n_bc, const float* bc, int32_t n_steps_total,
                 const int32_t* step_ptr, void* stream);

/* dst[i, :] = src[idx[i], :] for rows of `width` floats: packs the boundary rows a rank sends to a peer in the
 * partitioned large-mesh rollout (no reference counterpart: the reference has no graph partitioning). */
int swe_pack_rows(const float* src, const int32_t* idx, int64_t n_rows, int32_t width, float* dst, void* stream);

/* *step_ptr += 1 (end of one rollout step, training/train.py:87). */
int swe_step_advance(int32_t* step_ptr, void* stream);


/* ---------------------------------------------------------------------------------------------
 * Training path (forward that keeps what the backward needs + backward kernels).
 * The reference has no hand-written backward: it is whatever torch.autograd derives for
 * models/gnn.py:387-445 etc. inside LightningTrainer.training_step (training/train.py:125-145).
 * The kernels below are that derivative, written out; parity is checked
 * against torch.autograd run on the oracle.  Every reduction is a fixed-order sum (per-CTA
 * partials reduced in CTA order, CSR segments in edge order): no atomics, bit-reproducible.
 * ------------------------------------------------------------------------------------------- */

/* Row provider: X[r, :] = concat_j act_j( base_j[(idx_j ? idx_j[r] : r) * ld_j + (0..width_j)) ),
 * every block zero-padded to a multiple of 4 columns (<= 128). */
typedef struct swe_seg {
    const float*   base;
    const int32_t* idx;     /* NULL = identity */
    const float*   slope;   /* PReLU parameter (device) or NULL */
    int32_t ld;
    int32_t width;
    int32_t act;            /* enum swe_act applied on load */
    int32_t _pad;
} swe_seg_t;

#define SWE_MAX_SEGS 5
typedef struct swe_rows {
    int32_t n_seg;
    int32_t _pad;
    swe_seg_t seg[SWE_MAX_SEGS];
} swe_rows_t;

/* pre[r, 0:n_out) = X[r, :] · Wᵀ + bias  (one Linear of a make_mlp stack, models/models.py:133-145,
 * WITHOUT its activation: the consumer applies it through its row provider, so the saved tensor
 * is the pre-activation the backward needs).  wt: packed k-major weight (swe_pack_linear) whose
 * row blocks follow the padded segment widths.  n_out in {16, 32, 64, 128}. */
int swe_mlp_layer_fwd(const swe_rows_t* X, int64_t n_rows, const float* wt, const float* bias, int32_t n_out,
                      float* pre, void* stream);

/* Backward of one Linear+activation w.r.t. its input:
 *   delta = dh ⊙ act'(pre)            (skipped when pre == NULL: dh already holds delta)
 *   dx[r, 0:KO) (+)= delta[r, :] · W[:, k_off : k_off + KO)      (columns >= k_valid give 0)
 * dh: [n_rows, n] in/out (overwritten with delta when write_delta); W: torch Linear.weight layout
 * [n, w_ld] row-major; KO in {16, 32, 64, 128}; dx: [n_rows, KO] or NULL (then only delta and
 * the partial sums are produced).  part: NULL or [grid_ctas][n + 1] per-CTA partial sums of
 * delta (bias gradient) and of dh·min(pre,0)-style PReLU slope gradient (last slot); reduce them
 * with swe_reduce_partials.  Returns the number of CTAs through *grid_out. */
int swe_mlp_layer_bwd_dx(float* dh, const float* pre, int32_t act, const float* slope, int64_t n_rows, int32_t n,
                         const float* w, int32_t w_ld, int32_t k_off, int32_t k_valid, int32_t ko,
                         float* dx, int32_t accumulate, int32_t write_delta, float* part, int32_t* grid_out,
                         void* stream);
int swe_mlp_layer_bwd_dx_grid(int64_t n_rows);

/* Weight gradient of one Linear: part[cta][n_i * KO + k] = Σ_{rows of this CTA} delta[r, n_i] · X[r, k]
 * (X: provider with ONE segment, zero-padded to KO).  Reduce with swe_reduce_partials. */
int swe_mlp_layer_bwd_dw(const float* delta, int64_t n_rows, int32_t n, const swe_rows_t* X, int32_t ko,
                         float* part, int32_t* grid_out, void* stream);
int swe_mlp_layer_bwd_dw_grid(int64_t n_rows);

/* Static share of the edge MLP's first layer, hoisted out of the rollout: the encoded edge features a_e (and, for
 * with_WL=False models, the encoded static node features x_s) do not change over the steps of a rollout
 * (training/train.py:67-95 re-evaluates them every step), so P[e] = W1[:, a_e block] · a[e] (+ W1[:, x_s blocks] ·
 * (x_s[src[e]], x_s[dst[e]]) when xs != NULL) is computed once (no bias) into a table in an internal tile-transposed
 * order (128 floats per edge, padded to whole 128-edge tiles: p_out holds ceil(n_edges / 128) * 128 * 128 floats) ... */
int swe_gate_static_partials_tc(const float* xs, const float* a, const int32_t* src, const int32_t* dst,
                                int64_t n_edges, const void* image, int32_t k1, float* p_out, void* stream);
/* ... and every step evaluates s_ij from P[e] plus the remaining blocks: x_d (xd_dst NULL: un-pool call) and, when it
 * is not part of the table (xs != NULL here: models with with_WL=True, whose x_s contains the current water level), x_s. */
int swe_edge_gate_tc_stat_fwd(const float* p_edge, const float* xs, const float* xd_src, const float* xd_dst,
                              const int32_t* src, const int32_t* dst, int64_t n_edges, const void* image, int32_t k1,
                              const int32_t* act3, const float* const* slope3, int32_t normalize, float* s_out,
                              void* stream);

/* swe_edge_gate_tc_fwd that also stores the pre-activations of the three edge-MLP layers (pre1, pre2: [E, 128],
 * pre3: [E, 64]; bias included, activation not applied) — the forward of the training step for the default model.
 * fix_count != NULL: [3] counters (zeroed by the caller) and fix_lists [3][fix_cap] receive, per layer, the entries
 * (edge << 8 | column) with |pre| < fix_tau * max(1, max |pre| of the row piece): candidates for a wrong sign. */
int swe_edge_gate_tc_train_fwd(const float* xs, const float* xd_src, const float* xd_dst, const float* a,
                               const int32_t* src, const int32_t* dst, int64_t n_edges, const void* image, int32_t k1,
                               const int32_t* act3, const float* const* slope3, int32_t normalize, float* pre1,
                               float* pre2, float* pre3, float* s_out, unsigned long long* fix_lists,
                               int32_t* fix_count, int32_t fix_cap, float fix_tau, void* stream);
/* Recomputes the listed pre-activations in exact fp32 (w1 [128, k1], w2 [128, 128], w3 [64, 128]: torch Linear
 * layout), layer by layer, so that the backward's derivative masks (v > 0) are those of an fp32 forward. */
int swe_gate_fix_preacts(const float* xs, const float* xd_src, const float* xd_dst, const float* a, const int32_t* src,
                         const int32_t* dst, const float* w1, const float* b1, const float* w2, const float* b2,
                         const float* w3, const float* b3, int32_t k1, const int32_t* act3, const float* const* slope3,
                         float* pre1, float* pre2, float* pre3, const unsigned long long* fix_lists,
                         const int32_t* fix_count, int32_t fix_cap, void* stream);

/* Tensor-core (tcgen05, 3xTF32 = fp32-accurate products, fp32 accumulation in TMEM) forms of the two GEMMs above
 * for the wide edge-MLP layers; same mathematics, relative error ~1e-6 instead of exact-fp32 summation order.
 *   dx_tc: dx[r, 0:ko) (+)= delta[r, 0:n) · W[0:n, k_off : k_off + ko)  (columns >= k_valid give 0); delta is the
 *          finished delta (run swe_mlp_layer_bwd_dx with dx == NULL first for delta and the bias partials);
 *          n, ko in {64, 128}; columns [0, split) go to dx0 [n_rows, split], the rest to dx1 [n_rows, ko - split]
 *          (split == ko: dx1 unused; otherwise split == 64, ko == 128: two 64-wide blocks in one pass).
 *   dw_tc: part[cta][seg_col0 * n + n_i * w_seg + k] = Σ_{rows of this CTA} delta[r, n_i] · X_seg[r, k] for every
 *          segment of the provider X (widths multiples of 32, none/relu/leakyrelu/prelu on load); n in {64, 128},
 *          provider 64, 128 or 256 columns wide.  Reduce each segment with
 *          swe_reduce_partials(part, grid, n * width(X), seg_col0 * n, n_out * w_seg, w_seg, ...). */
int swe_mlp_layer_bwd_dx_tc(const float* delta, int64_t n_rows, int32_t n, const float* w, int32_t w_ld,
                            int32_t k_off, int32_t k_valid, int32_t ko, float* dx0, int32_t accumulate0, float* dx1,
                            int32_t accumulate1, int32_t split, void* stream);
/* dx_tc that forms delta = dh ⊙ act'(pre) itself while staging the rows (act in none/relu/leakyrelu/prelu): dh is
 * overwritten with delta (the input of dw_tc), part receives [grid][n + 1] per-CTA partial sums of delta (bias
 * gradient) and, last slot, of dh·pre over pre <= 0 (PReLU slope gradient) — what swe_mlp_layer_bwd_dx returns. */
int swe_mlp_layer_bwd_dx_tc_fused(float* dh, const float* pre, int32_t act, const float* slope, int64_t n_rows,
                                  int32_t n, const float* w, int32_t w_ld, int32_t k_off, int32_t k_valid, int32_t ko,
                                  float* dx0, int32_t accumulate0, float* dx1, int32_t accumulate1, int32_t split,
                                  float* part, int32_t* grid_out, void* stream);
int swe_mlp_layer_bwd_dx_tc_grid(int64_t n_rows);
int swe_mlp_layer_bwd_dw_tc(const float* delta, int64_t n_rows, int32_t n, const swe_rows_t* X, float* part,
                            int32_t* grid_out, void* stream);
int swe_mlp_layer_bwd_dw_tc_grid(int64_t n_rows);

/* out[(j / ko) * ld_out + k_off + j % ko] += Σ_{c < n_parts, in order} part[c * part_stride + item_off + j]
 * for j < n_items with (j % ko) < k_valid. */
int swe_reduce_partials(const float* part, int32_t n_parts, int64_t part_stride, int32_t item_off, int32_t n_items,
                        int32_t ko, int32_t k_valid, float* out, int32_t ld_out, int32_t k_off, void* stream);

/* s = act(pre3) / ||act(pre3)||₂, NaN -> 0 (models/gnn.py:425-426); F-wide rows. */
int swe_gate_norm_fwd(const float* pre3, int32_t act, const float* slope, int32_t normalize, int64_t n_edges,
                      float* s_out, int32_t F, void* stream);
/* ds (in place) -> du = (ds − s (s·ds)) / ||u||  (0 where ||u|| = 0); normalize == 0: unchanged. */
int swe_gate_norm_bwd(float* ds, const float* pre3, int32_t act, const float* slope, int32_t normalize,
                      int64_t n_edges, int32_t F, void* stream);

/* y = act(x) and its backward g ⊙ act'(x) over rows [row_lo, row_lo + n_rows) of [*, F] arrays
 * (gnn_activation, models/gnn.py:135-136,332-336).  slope_part: NULL or [grid] per-CTA PReLU
 * slope-gradient partials. */
int swe_act_fwd(const float* x, int32_t row_lo, int32_t n_rows, int32_t act, const float* slope, float* y,
                int32_t F, void* stream);
int swe_act_bwd(const float* g, const float* x, int32_t row_lo, int32_t n_rows, int32_t act, const float* slope,
                float* gx, float* slope_part, int32_t* grid_out, int32_t F, void* stream);

/* swe_propagate_hop_fwd that also stores agg (the filter's input) for the backward. */
int swe_propagate_hop_train_fwd(const float* o_src, const float* o_dst, const float* s, const int32_t* rowptr,
                                const int32_t* src, int32_t dst_lo, int32_t n_dst, const float* wt,
                                int32_t with_gradient, int32_t upwind, const float* addend, float* agg_out,
                                float* out, int32_t F, void* stream);

/* flags[i] = (Σ_f o[i, f] != 0) for rows [row_lo, row_lo + n_rows): the reference's per-hop wet-node
 * mask (models/gnn.py:408), which the backward needs to reproduce autograd's input gradients. */
int swe_row_flags(const float* o, int32_t row_lo, int32_t n_rows, uint8_t* flags, int32_t F, void* stream);

/* Backward of one hop, destination-centric part (models/gnn.py:428-443 differentiated):
 *   act_p = wet_dst[c] | wet_src[src[p]]
 *   with_gradient: ds[p] (+)= act_p · da[c] ⊙ (o_dst[c] − o_src[src[p]]);  g_part[c] = g_next[c] + da[c] ⊙ Σ_p act_p s[p]
 *   otherwise    : ds[p] (+)= act_p · da[c] ⊙ o_src[src[p]]                (g_part untouched)
 * wet_dst / o_dst may be NULL (destination rows identically zero). */
int swe_hop_bwd_dst(const float* da, const float* o_src, const float* o_dst, const float* s, float* ds,
                    int32_t accumulate_ds, const int32_t* rowptr, const int32_t* src, const uint8_t* wet_src,
                    const uint8_t* wet_dst, int32_t dst_lo, int32_t n_dst, int32_t with_gradient,
                    const float* g_next, float* g_part, int32_t F, void* stream);
/* Source-centric part over the transposed CSR (t_rowptr over source nodes, t_pos = position of the
 * edge in destination-CSR order):  acc = Σ_q act · s[p] ⊙ da[dst[p]], p = t_pos[q];
 *   with_gradient: g_io[n] −= acc;   otherwise: g_io[n] (+)= acc (accumulate flag). */
int swe_hop_bwd_src(const float* da, const float* s, const int32_t* t_rowptr, const int32_t* t_pos,
                    const int32_t* dst, const uint8_t* wet_src, const uint8_t* wet_dst, int32_t src_lo,
                    int32_t n_src, int32_t with_gradient, int32_t accumulate, float* g_io, int32_t F, void* stream);

/* out[node_lo + i] (+)= Σ_{q in segment i} e[pos ? pos[q] : q]   (edge -> node sums of the gate's input
 * gradients over a CSR; fixed order). */
int swe_edge_to_node_sum(const float* e, const int32_t* rowptr, const int32_t* pos, int32_t node_lo, int32_t n_nodes,
                         float* out, int32_t accumulate, int32_t F, void* stream);

/* Backward of swe_pool_mean_fwd: dx[fine_lo + i] (+)= Σ_{q in fine segment i} g[coarse[q]] / max(1, count(coarse[q])),
 * count from the pooling CSR rowptr (keyed by coarse node, first key coarse_lo). */
int swe_pool_mean_bwd(const float* g, const int32_t* f_rowptr, const int32_t* coarse, int32_t fine_lo, int32_t n_fine,
                      const int32_t* pool_rowptr, int32_t coarse_lo, float* dx, int32_t accumulate, int32_t F,
                      void* stream);

/* Encoder inputs of the training path: xin_s[i] = [x[perm[i], 0:n_static_raw], WL, 0..] (ks columns),
 * and the scatter of their gradients back into dx (original row order, accumulated):
 *   dx[perm[i], c] += dxin_s[i, c] (c < n_static_raw); WL: dx[.., n_static_raw-1] and dx[.., n_cols-2];
 *   dx[perm[i], n_static_raw + c] += dxin_d[i, c] for i < n_dyn_rows. */
int swe_static_inputs_fwd(const float* x, int32_t n_cols, const int32_t* perm, int32_t n_nodes, int32_t n_static_raw,
                          int32_t with_wl, float* xin_s, int32_t ks, void* stream);
int swe_node_inputs_bwd(const float* dxin_s, int32_t ks, const float* dxin_d, int32_t kd, int32_t n_cols,
                        const int32_t* perm, int32_t n_nodes, int32_t n_dyn_rows, int32_t n_static_raw,
                        int32_t with_wl, float* dx, void* stream);

/* Head of the training path: pred from the decoder's last pre-activation (models/gnn.py:339-348,
 * models/models.py:50-91) and its backward.  pre3: [n_nodes, ldp] (first two columns used).
 *   dh3[i, 0:2] = dpred[perm[i]] ⊙ masks ⊙ relu'  (other columns 0), dx0[perm[i], window] += residual path,
 *   res_part[cta][2*previous_t]: per-CTA partials of d residual_weights laid out [t][var]. */
int swe_head_fwd(const float* pre3, int32_t ldp, int32_t act, const float* slope, const float* x0, int32_t n_cols,
                 const int32_t* perm, int32_t n_nodes, int32_t previous_t, int32_t res_mode, const float* res_w,
                 float eps, float* pred, void* stream);
int swe_head_bwd(const float* dpred, const float* pre3, int32_t ldp, int32_t act, const float* slope, const float* x0,
                 int32_t n_cols, const int32_t* perm, int32_t n_nodes, int32_t previous_t, int32_t res_mode,
                 const float* res_w, float eps, float* dh3, float* dx0, float* res_part, int32_t* grid_out,
                 void* stream);

/* ---------------------------------------------------------------------------------------------
 * Rest of the training step on the device: loss with its gradient, gradient clipping and AdamW, so that
 * forward + backward + update is a fixed kernel sequence (capturable in a CUDA graph).
 * swe_loss_fwd_bwd: /root/reference/training/loss.py:76-118 with conservation = 0: err_v = RMSE (mae = 0) or MAE (mae = 1) of
 *   pred - real over the rows with rows[i] != 0 (NULL: all) and, with only_where_water, a non-zero difference
 *   (mask_on_water, loss.py:30-36); *loss (+)= scale (w0 err_0 + w1 err_1) / (w0 + w1); dpred [n, 2] = d loss / d pred.
 *   real element (i, v) = real[i * real_stride + v * real_stride / 2] (a [n, 2, T] target at one time step: stride 2 T).
 *   ws: swe_train_step_ws_bytes() bytes of scratch.
 * swe_clip_adamw_step: torch.nn.utils.clip_grad_norm_(max_norm) (Lightning gradient_clip_val, /root/reference/main.py:109;
 *   max_norm <= 0: none) followed by one torch.optim.AdamW step (/root/reference/training/train.py:147-155) on flat fp32
 *   buffers of n elements; lr is a DEVICE scalar; state = {step count, last gradient norm, last clip coefficient}.
 * ------------------------------------------------------------------------------------------- */
size_t swe_train_step_ws_bytes(void);
int swe_loss_fwd_bwd(const float* pred, const float* real, int64_t real_stride, const unsigned char* rows, int64_t n,
                     int32_t only_where_water, int32_t mae, float w0, float w1, float scale, int32_t accumulate,
                     float* loss, float* dpred, void* ws, void* stream);
int swe_clip_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr,
                        float beta1, float beta2, float eps, float weight_decay, float max_norm, float* state,
                        void* ws, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Dataset-side helpers on the device.
 * swe_temporal_window: the temporal sample starting at init_time of a resident simulation — x [n, n_static + 2 previous_t],
 *   y [n, 2, rollout_steps], bc_out [n_bc, previous_t, rollout_steps + 1] — as /root/reference/utils/dataset.py:410-471
 *   (`to_temporal`) builds it: dry-bed prefix of previous_t - 1 zero steps, BC's last step repeated once.
 *   wd, v: [n, t_sim]; bc: [n_bc, t_bc].
 * swe_rollout_metrics: per time step, for up to 4 depth thresholds, TP / TN / FP / FN of pred vs real
 *   (/root/reference/utils/miscellaneous.py:123-151), and the error sums of get_rollout_loss (miscellaneous.py:177-199):
 *   out [T, swe_rollout_metrics_cols()] doubles = {TP, TN, FP, FN} x 4 thresholds | sum d0^2, sum d1^2, sum |d0|, sum |d1| |
 *   the same four over rows with a non-zero difference | number of those rows.  pred, real: [n, 2, T].
 * ------------------------------------------------------------------------------------------- */
int swe_temporal_window(const float* x_static, int32_t n_static, const float* wd, const float* v, int64_t n, int32_t t_sim,
                        const float* bc, int32_t n_bc, int32_t t_bc, int32_t init_time, int32_t previous_t,
                        int32_t rollout_steps, float* x, float* y, float* bc_out, void* stream);
int32_t swe_rollout_metrics_cols(void);
size_t swe_rollout_metrics_ws_bytes(int32_t T);
int swe_rollout_metrics(const float* pred, const float* real, int64_t n, int32_t T, const float* thr, int32_t n_thr,
                        double* out, void* ws, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: halo exchange of a partitioned mesh over peer memory (one process per GPU; no reference counterpart,
 * contract = SURVEY.md §8(e): owned rows bit-identical to the single-GPU result).
 * swe_ipc_*: a cudaMalloc'ed, zero-filled arena with its 64-byte CUDA IPC handle; peers map it with swe_ipc_open.
 * swe_halo_exchange: rows send_idx[q][0..n_send[q]) of `arr` ([*, width] fp32, width % 4 == 0) are stored to
 * remote_rows[q] (peer-mapped first halo row of this rank in neighbour q's copy of the array), then *seq + 1 is stored
 * (release, system scope) to remote_flags[q]; do_wait: the launch then waits until local_flags[q] >= *seq + 1 for every
 * neighbour and sets *seq += 1.  All pointer arrays are HOST arrays of n_peers (<= 16) DEVICE pointers; `seq` and `done`
 * are device words owned by the caller (zero-initialised).  One kernel, capturable in a CUDA graph.  do_push / do_wait
 * select the two halves (tests drive them separately with a host barrier in between).
 * ------------------------------------------------------------------------------------------- */
int swe_ipc_alloc(size_t bytes, void** ptr_out, unsigned char* handle64);
int swe_ipc_open(const unsigned char* handle64, void** ptr_out);
int swe_ipc_close(void* ptr);
int swe_ipc_free(void* ptr);
int swe_halo_exchange(const float* arr, int32_t width, int32_t n_peers, const int32_t* const* send_idx,
                      const int64_t* n_send, float* const* remote_rows, uint32_t* const* remote_flags,
                      const uint32_t* const* local_flags, uint32_t* seq, uint32_t* done, int32_t do_push,
                      int32_t do_wait, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SWE_GNN_B200_H */
