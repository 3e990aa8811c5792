// Row MLPs on the 5th-generation tensor cores (F = 64): node / edge encoders, filter_matrix[0] and the
// decoder head.  On CUDA cores these streams are FP32-FMA-bound (2·64² FLOP per layer per row against
// ~0.5 KB of traffic per row); here the 64→64 layers are tcgen05.mma (3xTF32, fp32 accumulation in TMEM)
// and the kernel is a row stream again.
//
//   X0   = act_in(x_rows[r])                                    (INPUT_ROWS)      [models/gnn.py:401-402, 339]
//        | act_f(W_f · raw(r) + b_f),  raw = ≤ 8 raw input columns (+ WL)  (first Linear of an encoder on CUDA
//                                                                cores: 64·k MACs)  [models/gnn.py:281-294]
//   X1   = act_0(W_0 · X0 + b_0)                                 tcgen05, SS (A from shared memory)
//   X2   = act_1(W_1 · X1 + b_1)                                 tcgen05, TS (A stays in TMEM)         (n_tc = 2)
//   out  = X_last rows                                           (ROW output)
//        | decoder head: relu(act_h(w_h · X_last) + residual(x0)), dry mask, prediction + window shift
//                                                                [models/gnn.py:339-348, models/models.py:50-91,
//                                                                 utils/dataset.py:508-529]
//
// Persistent, one CTA per SM, 21 warps:
//   warps 0-15  row warps  : build X0 (16 lanes × 16 B per row, coalesced), split into TF32 hi/lo, write the A
//                            operand; one tile later read the result from the shared-memory stage and store the
//                            rows / run the head — fully coalesced global traffic on both sides
//   warps 16-19 epilogue   : thread = TMEM lane = row: D → bias, activation → (TS operand for the next layer | stage)
//   warp  20    MMA issuer
#include "swe_tc.cuh"

namespace swe {
namespace tc {

constexpr int RF = 64;
constexpr int R_TILE = 128;
constexpr int R_KC = 32;
constexpr int R_A_TILE = R_TILE * 128;          // [128 x 32] tf32 tile, bytes
constexpr int R_A_SLOT = 4 * R_A_TILE;          // 2 chunks x (hi | lo)
constexpr int R_W_TILE = RF * 128;              // [64 x 32] tf32 tile
constexpr size_t R_W_IMAGE = 4 * (size_t)R_W_TILE;       // 32 KB per layer (same layout as swe_hop_tc_pack)
#ifndef SWE_R_EPI_WARPS
#define SWE_R_EPI_WARPS 4
#endif
constexpr int R_ROW_WARPS = 16, R_ROW_THREADS = 512, R_EPI_WARPS = SWE_R_EPI_WARPS;   // 4 or 8 (2 per TMEM lane quarter, 32 columns each)
constexpr int R_THREADS = R_ROW_THREADS + R_EPI_WARPS * 32 + 32;    // 800
constexpr int R_STAGE_LD = RF + 4;
constexpr size_t R_STAGE_BYTES = (size_t)R_TILE * R_STAGE_LD * 4;
// TMEM columns (512 allocated): everything between the two tensor-core layers is double-buffered by tile parity, so
// that the epilogue warps turn D0 of tile i + 1 into its layer-1 operand while layer 1 of tile i is still running
constexpr uint32_t RC_D0 = 0, RC_D1 = 128, RC_AHI = 256, RC_ALO = 384;      // + 64 * (tile & 1)

// Activations inside the per-element loops: the "leaky family" (none / relu / leakyrelu / prelu) is two FP32
// instructions; everything else goes through ONE out-of-line function.  Inlining the 8-way act_apply switch at the
// ~200 call sites of this kernel produced 24 k SASS instructions (392 KB) and the epilogue ran out of the
// instruction cache (12 k cycles per tile instead of ~1 k).
__device__ __noinline__ float act_generic(int act, float v, float slope) { return act_apply(act, v, slope); }
// tanh for the decoder's input activation (64 per row, 1.35 M rows per step): inline, ~20 instructions instead of a call into
// libdevice's tanhf behind an out-of-line switch (~45).  |x| < 0.25: odd polynomial up to x^9 (next term < 2e-9 relative);
// otherwise (1 - t) / (1 + t) with t = exp(-2|x|) (ex2.approx, ~2 ulp; no cancellation since t <= 0.61): ~3e-7 relative.
__device__ __forceinline__ float tanh_inline(float x) {
    const float ax = fabsf(x), x2 = x * x;
    const float poly = x * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 0.021869488f, -0.053968254f), 0.13333333f), -0.33333333f), 1.f);
    const float t = __expf(-2.f * ax);
    const float big = copysignf(__fdividef(1.f - t, 1.f + t), x);        // rcp.approx + mul: 2 ulp, no IEEE-division subroutine
    return ax < 0.25f ? poly : big;
}
struct ActSel { int act; float slope; bool leaky; };
__device__ __forceinline__ ActSel act_select(int act, const float* slope_p) {
    ActSel a;
    a.act = act;
    a.leaky = (act == SWE_ACT_NONE || act == SWE_ACT_PRELU || act == SWE_ACT_RELU || act == SWE_ACT_LEAKYRELU);
    a.slope = act == SWE_ACT_NONE ? 1.f : act == SWE_ACT_RELU ? 0.f : act == SWE_ACT_LEAKYRELU ? 0.1f
              : (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
    return a;
}
__device__ __forceinline__ float act_do(const ActSel& a, float v) {
    return a.leaky ? fmaxf(v, 0.f) + a.slope * fminf(v, 0.f) : act_generic(a.act, v, a.slope);
}
// the row warps' input activation: tanh (config.yaml's gnn_activation in front of the decoder) gets the inline version
__device__ __forceinline__ float act_in_do(const ActSel& a, float v) {
    return a.act == SWE_ACT_TANH ? tanh_inline(v) : act_do(a, v);
}

struct __align__(8) RowBarriers {
    uint64_t a_full, a_empty;          // A operand written (512) / consumed by layer-0 MMAs (commit)
    uint64_t d0_full[2], d1_full[2];   // layer-0 / layer-1 accumulators complete (commit), per buffer
    uint64_t x1_ready[2];              // layer-1 operand written to TMEM (128), per buffer
    uint64_t d0_free[2], d1_free[2];   // accumulators read by the epilogue (128): the buffer may be overwritten
    uint64_t st_full[2], st_empty[2];  // stage buffer (tile parity) holds a tile (128) / consumed (512)
};

constexpr size_t ROWMLP_SMEM = 1024 + (size_t)R_A_SLOT + 2 * R_W_IMAGE + 2 * R_STAGE_BYTES + 64 * 8 * 4 + 64 * 4 +
                               2 * 64 * 4 + 2 * 64 * 4 + sizeof(RowBarriers) + 16;

struct RowMlpParams {
    // ---- input
    const float* x_rows; int act_in; const float* slope_in;          // INPUT_ROWS: [*, 64] rows row_lo + r
    const float* raw; int raw_ld; int raw_col0; int raw_cols; int with_wl; int wl_col_a, wl_col_b;
    const int32_t* perm;                                              // raw row of row r = perm[row_lo + r] (NULL: identity)
    const float* w_first; const float* b_first; int act_first; const float* slope_first;   // [64][raw_k] torch layout
    int raw_k;                                                        // raw_cols + with_wl  (<= 8)
    int row_lo; long long n_rows;
    // ---- tensor-core layers
    int n_tc; const unsigned char* img[2]; const float* bias[2]; int act[2]; const float* slope[2];
    // ---- output
    float* out_rows;                                                  // [*, 64] or NULL
    // decoder head
    int head; const float* w_head; const float* b_head; int act_head; const float* slope_head;   // [2][64]
    const float* x0; int n_cols; const int32_t* head_perm; int previous_t; int res_mode; const float* res_w; float eps;
    float* pred; const int32_t* step_ptr; long long pred_step_stride; float* x_next;
    long long* trace;          // profiling aid: [3 roles][16 tiles][8 events] clock64 stamps of CTA 0
};

__global__ void __launch_bounds__(R_THREADS, 1) row_mlp_tc_kernel(const __grid_constant__ RowMlpParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_slot = smem;
    unsigned char* w_tile = smem + R_A_SLOT;                           // layer l at + l * 32 KB
    float* stage = reinterpret_cast<float*>(w_tile + 2 * R_W_IMAGE);
    // two stage buffers (tile parity): the epilogue warps fill one while the row warps write the other one out — with a
    // single buffer the two took turns and their hand-over latencies added up to the whole tile time
    float* s_wf = stage + 2 * R_TILE * R_STAGE_LD;                      // [8][64] first-layer weights (k-major, padded with 0)
    float* s_bf = s_wf + 64 * 8;                                        // [64]
    float* s_bias = s_bf + 64;                                          // [2][64]
    float* s_wh = s_bias + 128;                                         // [2][64] head weights
    RowBarriers* bar = reinterpret_cast<RowBarriers*>(s_wh + 128);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar->a_full, R_ROW_THREADS); mbar_init(&bar->a_empty, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&bar->d0_full[b], 1); mbar_init(&bar->d1_full[b], 1);
            mbar_init(&bar->x1_ready[b], R_EPI_WARPS * 32);
            mbar_init(&bar->d0_free[b], R_EPI_WARPS * 32); mbar_init(&bar->d1_free[b], R_EPI_WARPS * 32);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(&bar->st_full[b], R_EPI_WARPS * 32); mbar_init(&bar->st_empty[b], R_ROW_THREADS); }
        fence_barrier_init();
    }
    for (int l = 0; l < p.n_tc; ++l)
        for (int i = threadIdx.x * 16; i < (int)R_W_IMAGE; i += R_THREADS * 16)
            *reinterpret_cast<float4*>(w_tile + l * R_W_IMAGE + i) = *reinterpret_cast<const float4*>(p.img[l] + i);
    // First encoder layer, stored [k][64] (a lane reads its 4 columns as one float4).  When a node's raw row is exactly
    // 8 floats (the default models) the weights are re-indexed by ABSOLUTE raw column, and the water-level input
    // WL = x[a] + x[b] is folded in by adding its weight to columns a and b: the row workers then load the row as two
    // float4 and run a plain 8 -> 64 layer (before: 10 predicated scalar loads per row, 7 k cycles per 128 rows).
    const bool row8 = p.raw != nullptr && p.raw_ld == 8 && p.raw_col0 + p.raw_cols <= 8 &&
                      (!p.with_wl || (p.wl_col_a < 8 && p.wl_col_b < 8)) && ((reinterpret_cast<uintptr_t>(p.raw) & 15u) == 0);
    if (row8) {
        if (threadIdx.x < 64) {
            const int n = threadIdx.x;
            float w8[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) w8[c] = 0.f;
            for (int j = 0; j < p.raw_cols; ++j) {
                const float w = p.w_first ? p.w_first[n * p.raw_k + j] : 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) if (c == p.raw_col0 + j) w8[c] += w;
            }
            if (p.with_wl) {
                const float w = p.w_first ? p.w_first[n * p.raw_k + p.raw_cols] : 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) if (c == p.wl_col_a || c == p.wl_col_b) w8[c] += w;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) s_wf[c * 64 + n] = w8[c];
        }
    } else {
        for (int i = threadIdx.x; i < 64 * 8; i += R_THREADS) {
            const int n = i >> 3, k = i & 7;
            s_wf[k * 64 + n] = (p.w_first && k < p.raw_k) ? p.w_first[n * p.raw_k + k] : 0.f;
        }
    }
    for (int i = threadIdx.x; i < 64; i += R_THREADS) {
        s_bf[i] = p.b_first ? p.b_first[i] : 0.f;
        s_bias[i] = p.bias[0] ? p.bias[0][i] : 0.f;
        s_bias[64 + i] = (p.n_tc > 1 && p.bias[1]) ? p.bias[1][i] : 0.f;
        s_wh[i] = p.head ? p.w_head[i] : 0.f;
        s_wh[64 + i] = p.head ? p.w_head[64 + i] : 0.f;
    }
    fence_proxy_async_smem();
    if (warp == R_ROW_WARPS + R_EPI_WARPS) tmem_alloc(tmem_holder, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_holder;
    const long long n_tiles = (p.n_rows + R_TILE - 1) / R_TILE;
    const long long tpc = (n_tiles + gridDim.x - 1) / gridDim.x;
    const long long tile0 = (long long)blockIdx.x * tpc;
    const int n_my = (int)max(0ll, min(tpc, n_tiles - tile0));

#define R_STAMP(role_, i_, ev_) do { if (p.trace && blockIdx.x == 0 && lane == 0 && (i_) < 16) p.trace[(role_) * 128 + (i_) * 8 + (ev_)] = clock64(); } while (0)
    if (warp < R_ROW_WARPS) {
        // =====================================================================================
        // row warps
        // =====================================================================================
        const int g = threadIdx.x >> 4, q = threadIdx.x & 15, q4 = 4 * q;
        const int chunk = q >> 3, piece = q & 7;
        uint32_t a_off[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) a_off[k] = sw128_offset(g + 32 * k, piece * 4);
        const ActSel a_in = act_select(p.act_in, p.slope_in), a_f = act_select(p.act_first, p.slope_first),
                     a_h = act_select(p.act_head, p.slope_head);
        // this lane's 4 output columns of the CUDA-core first layer
        float bf[4];
        if (p.raw) {
#pragma unroll
            for (int c = 0; c < 4; ++c) bf[c] = s_bf[q4 + c];
        }
        const int n_static_raw = p.n_cols - 2 * p.previous_t;
        // head constants of this lane (column q of a node's inputs, columns [4 q, 4 q + 4) of the last hidden layer), read once
        float4 wh0 = make_float4(0.f, 0.f, 0.f, 0.f), wh1 = wh0;
        float res_w0 = 0.f, res_w1 = 0.f, bh0 = 0.f, bh1 = 0.f;
        if (p.head) {
            wh0 = *reinterpret_cast<const float4*>(s_wh + q4); wh1 = *reinterpret_cast<const float4*>(s_wh + 64 + q4);
            const int rel = q - n_static_raw;
            if (rel >= 0 && q < p.n_cols && p.res_mode != 0) {
                const int t = rel >> 1, jv = rel & 1;
                const float w = p.res_mode == 1 ? __ldg(p.res_w + t) : p.res_mode == 2 ? __ldg(p.res_w + 2 * t + jv)
                                : (t == p.previous_t - 1 ? 1.f : 0.f);
                if (jv) res_w1 = w; else res_w0 = w;
            }
            if (p.b_head) { bh0 = __ldg(p.b_head); bh1 = __ldg(p.b_head + 1); }
        }
        auto finish_tile = [&](int j) {
            const long long r0 = (tile0 + j) * R_TILE;
            // head: lane c (< n_cols <= 16) of a row's 16 lanes owns input column c of that row: one coalesced
            // load of the node's inputs before the wait, one coalesced store of its shifted window after it
            float xr[4];
            long long orow[4];
            if (p.head) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const long long row = r0 + g + 32 * k;
                    xr[k] = 0.f; orow[k] = 0;
                    if (row < p.n_rows) {
                        const long long node = (long long)p.row_lo + row;
                        orow[k] = p.head_perm ? p.head_perm[node] : node;
                        if (q < p.n_cols) xr[k] = p.x0[orow[k] * p.n_cols + q];
                    }
                }
            }
            const float* stg = stage + (j & 1) * (R_TILE * R_STAGE_LD);
            float* pred_step = (p.head && q < 2) ? p.pred + (p.step_ptr ? (long long)(*p.step_ptr) * p.pred_step_stride : 0) : nullptr;
            mbar_wait(&bar->st_full[j & 1], ((uint32_t)j >> 1) & 1);
            if (warp == 0) R_STAMP(0, j + 2, 6);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = g + 32 * k;
                const long long row = r0 + r;
                const float4 d = *reinterpret_cast<const float4*>(stg + r * R_STAGE_LD + q4);
                if (!p.head) {
                    if (row < p.n_rows) stg4(p.out_rows + ((long long)p.row_lo + row) * RF + q4, d);
                } else {
                    // last decoder layer 64 -> 2: every lane holds 4 of the 64 inputs
                    float y0 = d.x * wh0.x + d.y * wh0.y + d.z * wh0.z + d.w * wh0.w;
                    float y1 = d.x * wh1.x + d.y * wh1.y + d.z * wh1.z + d.w * wh1.w;
                    // residual (models/models.py:50-77): this lane's column contributes to variable (c - n_static) & 1
                    float c0 = xr[k] * res_w0, c1 = xr[k] * res_w1;
#pragma unroll
                    for (int off = 8; off >= 1; off >>= 1) {
                        y0 += __shfl_xor_sync(0xffffffffu, y0, off);
                        y1 += __shfl_xor_sync(0xffffffffu, y1, off);
                        c0 += __shfl_xor_sync(0xffffffffu, c0, off);
                        c1 += __shfl_xor_sync(0xffffffffu, c1, off);
                    }
                    const float shifted = __shfl_down_sync(0xffffffffu, xr[k], 2, 16);       // column c + 2 of the same row
                    y0 += bh0; y1 += bh1;
                    y0 = fmaxf(act_do(a_h, y0) + c0, 0.f);
                    y1 = fmaxf(act_do(a_h, y1) + c1, 0.f);
                    const float oh = (fabsf(y0) > p.eps) ? y0 : 0.f;                        // h · [|h| > eps]
                    const float oq = (y0 != 0.f) ? y1 : 0.f;                                 // q · [h != 0] (un-thresholded h)
                    if (row < p.n_rows) {
                        if (q < 2) pred_step[orow[k] * 2 + q] = q ? oq : oh;
                        if (p.x_next && q < p.n_cols) {
                            const float v = q < n_static_raw ? xr[k] : (q < p.n_cols - 2 ? shifted : (q == p.n_cols - 2 ? oh : oq));
                            p.x_next[orow[k] * p.n_cols + q] = v;
                        }
                    }
                }
            }
            mbar_arrive(&bar->st_empty[j & 1]);
        };
        // the rows of tile i + 1 are requested as soon as tile i has been handed to the tensor core, so that their
        // latency (and the first encoder layer) overlaps with writing out tile i - 1
        float4 x[4];
        auto load_tile = [&](int i) {
            const long long r0 = (tile0 + i) * R_TILE;
            if (p.x_rows) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const long long row = r0 + g + 32 * k;
                    x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    // (the input activation is applied one phase later, activate_tile: the loads fly while tile i - 2 is
                    //  written out; rows past the end stay zero or become act(0) — they are computed but never stored)
                    if (row < p.n_rows) x[k] = ldg4(p.x_rows + ((long long)p.row_lo + row) * RF + q4);
                }
            } else {
                // first encoder layer on CUDA cores: the 8 raw inputs of this thread's 4 rows, then one pass over k with
                // the weights of its 4 columns read once per k (a float4 from shared memory) and used for all 4 rows
                // (two rows per pass keeps the live set — inputs, accumulators, the weights of one k — inside the kernel's
                //  80-register budget)
#pragma unroll
                for (int kp = 0; kp < 4; kp += 2) {
                    // the 8 inputs of a row live in two float4 NAMED per row (an indexed float[2][8] was placed in
                    // local memory: its STL/LDL were a quarter of all stall samples of the kernel)
                    float4 lo0 = make_float4(0.f, 0.f, 0.f, 0.f), hi0 = lo0, lo1 = lo0, hi1 = lo0;
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const long long row = r0 + g + 32 * (kp + u);
                        if (row < p.n_rows) {
                            const long long node = (long long)p.row_lo + row;
                            const float* xr = p.raw + (long long)(p.perm ? p.perm[node] : node) * p.raw_ld;
                            float4 l4, h4;
                            if (row8) {
                                l4 = ldg4(xr); h4 = ldg4(xr + 4);
                            } else {
                                float t[8];
#pragma unroll
                                for (int c = 0; c < 8; ++c) t[c] = c < p.raw_cols ? __ldg(xr + p.raw_col0 + c) : 0.f;
                                if (p.with_wl) {
                                    const float wl = __ldg(xr + p.wl_col_a) + __ldg(xr + p.wl_col_b);
#pragma unroll
                                    for (int c = 0; c < 8; ++c) t[c] = (c == p.raw_cols) ? wl : t[c];
                                }
                                l4 = make_float4(t[0], t[1], t[2], t[3]); h4 = make_float4(t[4], t[5], t[6], t[7]);
                            }
                            if (u == 0) { lo0 = l4; hi0 = h4; } else { lo1 = l4; hi1 = h4; }
                        }
                    }
                    float acc[2][4];
#pragma unroll
                    for (int u = 0; u < 2; ++u) { acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 0.f; }
#define SWE_IN8(l_, h_, kk_) ((kk_) == 0 ? l_.x : (kk_) == 1 ? l_.y : (kk_) == 2 ? l_.z : (kk_) == 3 ? l_.w : \
                              (kk_) == 4 ? h_.x : (kk_) == 5 ? h_.y : (kk_) == 6 ? h_.z : h_.w)
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const float4 w = *reinterpret_cast<const float4*>(s_wf + kk * 64 + q4);
                        const float i0 = SWE_IN8(lo0, hi0, kk), i1 = SWE_IN8(lo1, hi1, kk);
                        acc[0][0] = fmaf(i0, w.x, acc[0][0]); acc[0][1] = fmaf(i0, w.y, acc[0][1]);
                        acc[0][2] = fmaf(i0, w.z, acc[0][2]); acc[0][3] = fmaf(i0, w.w, acc[0][3]);
                        acc[1][0] = fmaf(i1, w.x, acc[1][0]); acc[1][1] = fmaf(i1, w.y, acc[1][1]);
                        acc[1][2] = fmaf(i1, w.z, acc[1][2]); acc[1][3] = fmaf(i1, w.w, acc[1][3]);
                    }
#undef SWE_IN8
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const long long row = r0 + g + 32 * (kp + u);
                        const float4 y = row < p.n_rows ? make_float4(act_do(a_f, acc[u][0] + bf[0]), act_do(a_f, acc[u][1] + bf[1]),
                                                                      act_do(a_f, acc[u][2] + bf[2]), act_do(a_f, acc[u][3] + bf[3]))
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                        if (kp == 0) { if (u == 0) x[0] = y; else x[1] = y; } else { if (u == 0) x[2] = y; else x[3] = y; }
                    }
                }
            }
        };
        auto activate_tile = [&]() {
            if (p.x_rows && p.act_in != SWE_ACT_NONE) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    x[k].x = act_in_do(a_in, x[k].x); x[k].y = act_in_do(a_in, x[k].y);
                    x[k].z = act_in_do(a_in, x[k].z); x[k].w = act_in_do(a_in, x[k].w);
                }
            }
        };
        // one loop with ONE call site per phase (the phases are big inlined lambdas; with a prologue and two epilogue
        // copies the kernel was 10.9 k SASS instructions): iteration i stores tile i, requests tile i + 1 and writes
        // out tile i - 2.  The output lags two tiles behind the input: waiting for tile i - 1 (whose layer 1 is only
        // issued after layer 0 of tile i) would hold back the store of tile i + 1 and serialise the whole chain.
#pragma unroll 1
        for (int i = -1; i < n_my + 2; ++i) {
            if (i >= 0 && i < n_my) {
                if (warp == 0) R_STAMP(0, i, 0);
                activate_tile();
                mbar_wait(&bar->a_empty, ((uint32_t)i & 1) ^ 1);           // layer-0 MMAs of the previous tile are done
                if (warp == 0) R_STAMP(0, i, 2);
                unsigned char* base = a_slot + (size_t)chunk * 2 * R_A_TILE;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float4 hh, ll;
                    split_tf32(x[k].x, hh.x, ll.x); split_tf32(x[k].y, hh.y, ll.y); split_tf32(x[k].z, hh.z, ll.z); split_tf32(x[k].w, hh.w, ll.w);
                    *reinterpret_cast<float4*>(base + a_off[k]) = hh;
                    *reinterpret_cast<float4*>(base + R_A_TILE + a_off[k]) = ll;
                }
                fence_proxy_async_smem();
                mbar_arrive(&bar->a_full);
                if (warp == 0) R_STAMP(0, i, 3);
            }
            if (i + 1 < n_my) load_tile(i + 1);
            if (warp == 0 && i >= 0) R_STAMP(0, i, 5);
            if (i >= 2 && i - 2 < n_my) finish_tile(i - 2);
            if (warp == 0 && i >= 0) R_STAMP(0, i, 4);
        }
    } else if (warp < R_ROW_WARPS + R_EPI_WARPS) {
        // =====================================================================================
        // epilogue warps.  Two-layer stacks:  E0(0) ; for i: { E0(i + 1) ; E1(i) }  — the operand of layer 1 of the next
        // tile is produced before this tile's layer-1 result is awaited, so the tensor pipe always has work queued.
        // =====================================================================================
        const int lq = warp & 3, hf0 = (warp - R_ROW_WARPS) >> 2;       // lane quarter, first column half
        const uint32_t lane_addr = tmem_base + ((uint32_t)(lq * 32) << 16);
        // everything indexed by the layer is selected into registers HERE: a runtime index into the parameter
        // struct (or a local array) becomes a local-memory load + a dependent branch per element in the loops below
        const ActSel a0 = act_select(p.act[0], p.slope[0]), a1 = act_select(p.act[1], p.slope[1]);
        const bool two = p.n_tc > 1;
        const ActSel a_l = two ? a1 : a0;
        const float* bias_l = s_bias + (two ? 64 : 0);
        float* my_row0 = stage + (lq * 32 + lane) * R_STAGE_LD;
        // X1 = act(D0 + b0) -> TF32 hi/lo -> TMEM operand of layer 1 (buffer = tile parity)
        auto e0 = [&](int i) {
            const uint32_t b = (uint32_t)i & 1, bph = ((uint32_t)i >> 1) & 1, bo = b * 64;
            mbar_wait(&bar->d0_full[b], bph);
            tc_fence_after_sync();
#pragma unroll 1
            for (int hf = hf0; hf < 2; hf += R_EPI_WARPS / 4) {
                uint32_t v[32], lo[32];
                tmem_ld32(lane_addr + RC_D0 + bo + hf * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float y = act_do(a0, __uint_as_float(v[j]) + s_bias[hf * 32 + j]);
                    const float hh = round_tf32(y);
                    v[j] = __float_as_uint(hh);
                    lo[j] = __float_as_uint(y - hh);
                }
                tmem_st32(lane_addr + RC_AHI + bo + hf * 32, v);
                tmem_st32(lane_addr + RC_ALO + bo + hf * 32, lo);
            }
            tmem_wait_st();
            tc_fence_before_sync();
            mbar_arrive(&bar->d0_free[b]);
            mbar_arrive(&bar->x1_ready[b]);
        };
#pragma unroll 1
        for (int i = -1; i < n_my; ++i) {
            if (two && i + 1 < n_my) e0(i + 1);
            if (i < 0) continue;
            const uint32_t b = (uint32_t)i & 1, bph = ((uint32_t)i >> 1) & 1, bo = b * 64;
            if (warp == R_ROW_WARPS) R_STAMP(1, i, 1);
            if (two) mbar_wait(&bar->d1_full[b], bph); else mbar_wait(&bar->d0_full[b], bph);
            tc_fence_after_sync();
            mbar_wait(&bar->st_empty[b], bph ^ 1);                        // this stage buffer consumed (tile i-2)
            float* my_row = my_row0 + b * (R_TILE * R_STAGE_LD);
            if (warp == R_ROW_WARPS) R_STAMP(1, i, 2);
            const uint32_t dcol = (two ? RC_D1 : RC_D0) + bo;
#pragma unroll 1
            for (int hf = hf0; hf < 2; hf += R_EPI_WARPS / 4) {
                uint32_t v[32];
                tmem_ld32(lane_addr + dcol + hf * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 r;
                    r.x = act_do(a_l, __uint_as_float(v[j]) + bias_l[hf * 32 + j]);
                    r.y = act_do(a_l, __uint_as_float(v[j + 1]) + bias_l[hf * 32 + j + 1]);
                    r.z = act_do(a_l, __uint_as_float(v[j + 2]) + bias_l[hf * 32 + j + 2]);
                    r.w = act_do(a_l, __uint_as_float(v[j + 3]) + bias_l[hf * 32 + j + 3]);
                    *reinterpret_cast<float4*>(my_row + hf * 32 + j) = r;
                }
            }
            tc_fence_before_sync();
            if (two) mbar_arrive(&bar->d1_free[b]); else mbar_arrive(&bar->d0_free[b]);
            mbar_arrive(&bar->st_full[b]);
            if (warp == R_ROW_WARPS) R_STAMP(1, i, 3);
        }
    } else {
        // =====================================================================================
        // MMA issuer.  L0(0) ; for i: { L0(i + 1) ; L1(i) }
        // =====================================================================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(R_TILE, RF);
            const uint32_t a_u32 = smem_u32(a_slot), w_u32 = smem_u32(w_tile);
            auto l0 = [&](int i) {
                const uint32_t ph = (uint32_t)i & 1;
                const uint32_t b = (uint32_t)i & 1, bph = ((uint32_t)i >> 1) & 1;
                R_STAMP(2, i, 0);
                mbar_wait(&bar->d0_free[b], bph ^ 1);                     // D0[b] of tile i-2 has been read
                R_STAMP(2, i, 1);
                mbar_wait(&bar->a_full, ph);
                tc_fence_after_sync();
                R_STAMP(2, i, 2);
                const uint32_t d0 = tmem_base + RC_D0 + b * 64;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t a_hi = a_u32 + c * 2 * R_A_TILE, a_lo = a_hi + R_A_TILE;
                    const uint32_t w_hi = w_u32 + c * 2 * R_W_TILE, w_lo = w_hi + R_W_TILE;
#pragma unroll
                    for (int ks = 0; ks < R_KC / 8; ++ks) {
                        const uint64_t dah = make_desc_sw128(a_hi + ks * 32), dal = make_desc_sw128(a_lo + ks * 32);
                        const uint64_t dwh = make_desc_sw128(w_hi + ks * 32), dwl = make_desc_sw128(w_lo + ks * 32);
                        mma_tf32_ss(d0, dal, dwh, idesc, (c | ks) ? 1u : 0u);
                        mma_tf32_ss(d0, dah, dwl, idesc, 1u);
                        mma_tf32_ss(d0, dah, dwh, idesc, 1u);
                    }
                }
                mma_commit(&bar->a_empty);
                mma_commit(&bar->d0_full[b]);
                R_STAMP(2, i, 3);
            };
            for (int i = -1; i < n_my; ++i) {
                if (i + 1 < n_my) l0(i + 1);
                if (i < 0) continue;
                const uint32_t b = (uint32_t)i & 1, bph = ((uint32_t)i >> 1) & 1;
                if (p.n_tc > 1) {
                    mbar_wait(&bar->x1_ready[b], bph);
                    mbar_wait(&bar->d1_free[b], bph ^ 1);                 // D1[b] of tile i-2 has been staged
                    tc_fence_after_sync();
                    const uint32_t d1 = tmem_base + RC_D1 + b * 64, xh = tmem_base + RC_AHI + b * 64, xl = tmem_base + RC_ALO + b * 64;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint32_t w_hi = w_u32 + (uint32_t)R_W_IMAGE + c * 2 * R_W_TILE, w_lo = w_hi + R_W_TILE;
#pragma unroll
                        for (int ks = 0; ks < R_KC / 8; ++ks) {
                            const uint32_t kcol = c * R_KC + ks * 8;
                            const uint64_t dwh = make_desc_sw128(w_hi + ks * 32), dwl = make_desc_sw128(w_lo + ks * 32);
                            mma_tf32_ts(d1, xl + kcol, dwh, idesc, (c | ks) ? 1u : 0u);
                            mma_tf32_ts(d1, xh + kcol, dwl, idesc, 1u);
                            mma_tf32_ts(d1, xh + kcol, dwh, idesc, 1u);
                        }
                    }
                    mma_commit(&bar->d1_full[b]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == R_ROW_WARPS + R_EPI_WARPS) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace swe

using namespace swe;

static long long* g_rowmlp_trace = nullptr;
extern "C" void swe_row_mlp_tc_set_trace(long long* t) { g_rowmlp_trace = t; }

extern "C" int swe_row_mlp_tc(const swe_rowmlp_t* d, void* stream) {
    SWE_REQUIRE(d, SWE_E_INVAL, "row_mlp_tc: null descriptor");
    SWE_REQUIRE(d->n_rows >= 0 && d->row_lo >= 0 && (d->n_tc == 1 || d->n_tc == 2), SWE_E_INVAL, "row_mlp_tc: bad sizes");
    SWE_REQUIRE((d->x_rows != nullptr) != (d->raw != nullptr), SWE_E_INVAL, "row_mlp_tc: exactly one of x_rows / raw");
    SWE_REQUIRE(!d->raw || (d->w_first && d->raw_cols >= 1 && d->raw_cols + (d->with_wl ? 1 : 0) <= 8), SWE_E_UNSUPP,
                "row_mlp_tc: the CUDA-core first layer takes at most 8 inputs");
    SWE_REQUIRE((d->out_rows != nullptr) != (d->head != 0), SWE_E_INVAL, "row_mlp_tc: exactly one of out_rows / head");
    SWE_REQUIRE(!d->head || d->n_cols <= 16, SWE_E_UNSUPP, "row_mlp_tc: the head handles at most 16 input columns");
    SWE_REQUIRE(!d->head || (d->w_head && d->x0 && d->pred && d->previous_t >= 1 && d->n_cols > 2 * d->previous_t &&
                             d->res_mode >= 0 && d->res_mode <= 3 && (d->res_mode == 0 || d->res_mode == 3 || d->res_w)),
                SWE_E_INVAL, "row_mlp_tc: bad head arguments");
    for (int l = 0; l < d->n_tc; ++l)
        SWE_REQUIRE(d->img[l] && aligned16(d->img[l]), SWE_E_ALIGN, "row_mlp_tc: layer %d image null/unaligned", l);
    SWE_REQUIRE((!d->x_rows || aligned16(d->x_rows)) && (!d->out_rows || aligned16(d->out_rows)), SWE_E_ALIGN,
                "row_mlp_tc: unaligned rows");
    if (d->n_rows == 0) return 0;
    tc::RowMlpParams p;
    memset(&p, 0, sizeof(p));
    p.x_rows = d->x_rows; p.act_in = d->act_in; p.slope_in = d->slope_in;
    p.raw = d->raw; p.raw_ld = d->raw_ld; p.raw_col0 = d->raw_col0; p.raw_cols = d->raw_cols; p.with_wl = d->with_wl;
    p.wl_col_a = d->wl_col_a; p.wl_col_b = d->wl_col_b; p.perm = d->perm;
    p.w_first = d->w_first; p.b_first = d->b_first; p.act_first = d->act_first; p.slope_first = d->slope_first;
    p.raw_k = d->raw_cols + (d->with_wl ? 1 : 0);
    p.row_lo = d->row_lo; p.n_rows = d->n_rows; p.n_tc = d->n_tc;
    for (int l = 0; l < 2; ++l) {
        p.img[l] = (const unsigned char*)d->img[l]; p.bias[l] = d->bias[l]; p.act[l] = d->act[l]; p.slope[l] = d->slope[l];
    }
    p.out_rows = d->out_rows; p.head = d->head; p.w_head = d->w_head; p.b_head = d->b_head; p.act_head = d->act_head;
    p.slope_head = d->slope_head; p.x0 = d->x0; p.n_cols = d->n_cols; p.head_perm = d->head_perm; p.previous_t = d->previous_t;
    p.res_mode = d->res_mode; p.res_w = d->res_w; p.eps = d->eps; p.pred = d->pred; p.step_ptr = d->step_ptr;
    p.pred_step_stride = d->pred_step_stride; p.x_next = d->x_next;
    p.trace = g_rowmlp_trace; g_rowmlp_trace = nullptr;
    cudaError_t e = cudaFuncSetAttribute(tc::row_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::ROWMLP_SMEM);
    if (e != cudaSuccess) { set_error("row_mlp_tc smem opt-in (%zu B): %s", tc::ROWMLP_SMEM, cudaGetErrorString(e)); return (int)e; }
    const long long n_tiles = (d->n_rows + tc::R_TILE - 1) / tc::R_TILE;
    tc::row_mlp_tc_kernel<<<grid_for(n_tiles, 1), tc::R_THREADS, tc::ROWMLP_SMEM, (cudaStream_t)stream>>>(p);
    return check_launch("row_mlp_tc");
}
