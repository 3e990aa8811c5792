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

/* *step_ptr += 1 (end of one rollout step, training/train.py:87). */
int swe_step_advance(int32_t* step_ptr, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SWE_GNN_B200_H */
