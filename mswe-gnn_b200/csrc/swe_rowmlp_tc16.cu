// Two-layer row MLPs (F = 64) as a STREAMING kernel with fp16 hi/lo operands: the node encoders
//     X0 = act_f(W_f · raw(r) + b_f)   (≤ 8 raw inputs, CUDA cores)      [models/gnn.py:281-294, models/models.py:121-146]
// and the decoder
//     X0 = act_in(x_rows[r])                                             [models/gnn.py:339]
// followed by  X1 = act_0(W_0 X0 + b_0),  X2 = act_1(W_1 X1 + b_1)  on tcgen05 (kind::f16, 3 products per layer, fp32
// accumulation), then ROW output or the decoder head (swe_rowmlp_tc.cu's head, same arithmetic)
//                                                                        [models/gnn.py:339-348, utils/dataset.py:508-529].
// Same contract as swe_row_mlp_tc for these shapes (leaky-family activations in the tensor-core layers; everything else
// stays on swe_row_mlp_tc).  Structure of swe_rowlin_tc16.cu, i.e. what made o_0 = x_d W_0ᵀ 1.8x faster:
//   * inputs arrive by asynchronous copies into a shared-memory ring several tiles ahead (decoder: whole 128-row tiles
//     by cp.async.bulk, two in flight; encoders: the rows' 32 raw bytes by cp.async through the node permutation, four
//     tiles in flight) — no thread waits for a global load on the critical path;
//   * fp16 hi/lo operands with a per-row power-of-two scale: 12 MMAs per layer instead of 24, a 32 KB operand slot
//     instead of 64 KB, so operand slots, accumulators and the layer-1 operand in TMEM are all double-buffered;
//   * layer 1 reads its operand from TMEM (the epilogue that produces it owns the row: maximum, scale and split without
//     any exchange).
// Weight images: swe_hop_tc16_pack of the two [64, 64] matrices.
#include <stdlib.h>
#include "swe_tc.cuh"

namespace swe {
namespace rm16 {
using namespace swe::tc;

constexpr int F = 64, TILE = 128, KC = 32;
constexpr int A_TILE = TILE * 64;               // [128 x 32] fp16 = 8 KB
constexpr int A_SLOT = 4 * A_TILE;              // 2 chunks x (hi | lo) = 32 KB
constexpr int W_TILE = F * 64;                  // [64 x 32] fp16 = 4 KB
constexpr size_t W_IMAGE = 4 * (size_t)W_TILE;  // 16 KB (+ 16 B: descale): swe_hop_tc16_pack
constexpr int RAW_STAGES = 4, ROW_STAGES = 2;
constexpr size_t RAW_TILE = (size_t)TILE * 32, ROW_TILE = (size_t)TILE * F * 4;
constexpr size_t IN_BYTES = ROW_STAGES * ROW_TILE;                     // 64 KB (>= RAW_STAGES * RAW_TILE)
constexpr int MAX_STAGES = 4;
constexpr int STAGE_LD = F + 4;
constexpr size_t STAGE_BYTES = (size_t)TILE * STAGE_LD * 4;
constexpr int CONV_WARPS = 8, CONV_THREADS = 256, EPI_WARPS = 4;
constexpr int THREADS = CONV_THREADS + EPI_WARPS * 32 + 64;            // + MMA issuer warp + loader warp = 448
constexpr uint32_t C_D0 = 0, C_D1 = 128, C_AHI = 256, C_ALO = 320;     // + 64 b (D) / + 32 b (A) by tile parity

struct __align__(8) Bar {
    uint64_t in_full[MAX_STAGES], in_empty[MAX_STAGES];
    uint64_t a_full[2], a_empty[2];
    uint64_t d0_full[2], d0_free[2], x1_ready[2], d1_full[2], d1_free[2];
    uint64_t st_full, st_empty;
};

constexpr size_t SMEM = 1024 + 2 * (size_t)A_SLOT + 2 * W_IMAGE + IN_BYTES + STAGE_BYTES + 6 * TILE * 4 +
                        (8 * 64 + 64 + 128 + 128 + 8) * 4 + sizeof(Bar) + 16;

struct Params {
    const float* x_rows; int act_in;
    uint32_t zero_mask;                      // 0 at run time, opaque to the compiler (see the ring-slot release)
    const float* raw; int raw_col0, raw_cols, with_wl, wl_col_a, wl_col_b, raw_k;
    const int32_t* perm;
    const float* w_first; const float* b_first; int act_first; const float* slope_first;
    long long row_lo, n_rows;
    const unsigned char* img[2]; const float* bias[2]; int act[2]; const float* slope[2];
    float* out_rows;
    int head; const float* w_head; const float* b_head; int act_head; const float* slope_head;
    const float* x0; int n_cols; const int32_t* head_perm; int previous_t; int res_mode; const float* res_w; float eps;
    float* pred; const int32_t* step_ptr; long long pred_step_stride; float* x_next;
};

struct Leaky { float slope; };
__device__ __forceinline__ Leaky leaky_of(int act, const float* slope_p) {
    Leaky a;
    a.slope = act == SWE_ACT_NONE ? 1.f : act == SWE_ACT_RELU ? 0.f : act == SWE_ACT_LEAKYRELU ? 0.1f
              : (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
    return a;
}
// v > 0 ? v : slope v  (slope <= 1: the larger of the two, else the smaller)
__device__ __forceinline__ float leaky_do(const Leaky& a, float v) {
    const float t = a.slope * v;
    return a.slope <= 1.f ? fmaxf(v, t) : fminf(v, t);
}
// tanh of the decoder's input activation (see swe_rowmlp_tc.cu): ~3e-7 relative
__device__ __forceinline__ float tanh16(float x) {
    const float ax = fabsf(x), x2 = x * x;
    const float poly = x * fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 0.021869488f, -0.053968254f), 0.13333333f), -0.33333333f), 1.f);
    const float t = __expf(-2.f * ax);
    const float big = copysignf(__fdividef(1.f - t, 1.f + t), x);
    return ax < 0.25f ? poly : big;
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <bool RAW_IN, bool HEAD>
__global__ void __launch_bounds__(THREADS, 1) row_mlp_tc16_kernel(const __grid_constant__ Params p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_slots = smem;
    unsigned char* w_tile = a_slots + 2 * (size_t)A_SLOT;                 // layer l at + l * 16 KB
    unsigned char* in_ring = w_tile + 2 * W_IMAGE;
    float* stage = reinterpret_cast<float*>(in_ring + IN_BYTES);          // [128][68]
    // [4][128] descale of D0 rows, indexed by tile & 3: without the stage hand-over (HEAD) nothing keeps the converters from
    // running two tiles ahead of the epilogue that reads it; slot i & 3 is rewritten for tile i + 4, whose operand store
    // waits for the MMAs of tile i + 2, which were issued after the epilogue of tile i released its accumulator
    float* s_inv0 = stage + TILE * STAGE_LD;
    float* s_inv1 = s_inv0 + 4 * TILE;                                    // [2][128] descale of D1 rows (epilogue-private)
    float* s_wf = s_inv1 + 2 * TILE;                                      // [8][64] first-layer weights by absolute raw column
    float* s_bf = s_wf + 8 * 64;                                          // [64]
    float* s_bias = s_bf + 64;                                            // [2][64]
    float* s_wh = s_bias + 128;                                           // [2][64] head weights
    float* s_misc = s_wh + 128;                                           // max |bias_0|
    Bar* bar = reinterpret_cast<Bar*>(s_misc + 8);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar + 1);

    constexpr int STAGES = RAW_IN ? RAW_STAGES : ROW_STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&bar->in_full[i], RAW_IN ? 32 : 1); mbar_init(&bar->in_empty[i], CONV_THREADS); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar->a_full[i], CONV_THREADS); mbar_init(&bar->a_empty[i], 1);
            mbar_init(&bar->d0_full[i], 1); mbar_init(&bar->d1_full[i], 1); mbar_init(&bar->x1_ready[i], EPI_WARPS * 32);
            mbar_init(&bar->d0_free[i], EPI_WARPS * 32); mbar_init(&bar->d1_free[i], EPI_WARPS * 32);
        }
        mbar_init(&bar->st_full, EPI_WARPS * 32); mbar_init(&bar->st_empty, CONV_THREADS);
        fence_barrier_init();
    }
    for (int l = 0; l < 2; ++l)
        for (int i = threadIdx.x * 16; i < (int)W_IMAGE; i += THREADS * 16)
            *reinterpret_cast<float4*>(w_tile + l * W_IMAGE + i) = *reinterpret_cast<const float4*>(p.img[l] + i);
    if (RAW_IN && threadIdx.x < 64) {
        // first layer re-indexed by ABSOLUTE raw column (the row is 8 floats), the water-level input WL = x[a] + x[b] folded
        // in by adding its weight to columns a and b (swe_rowmlp_tc.cu)
        const int n = threadIdx.x;
        float w8[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) w8[c] = 0.f;
        for (int j = 0; j < p.raw_cols; ++j) {
            const float w = p.w_first[n * p.raw_k + j];
#pragma unroll
            for (int c = 0; c < 8; ++c) if (c == p.raw_col0 + j) w8[c] += w;
        }
        if (p.with_wl) {
            const float w = p.w_first[n * p.raw_k + p.raw_cols];
#pragma unroll
            for (int c = 0; c < 8; ++c) if (c == p.wl_col_a || c == p.wl_col_b) w8[c] += w;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) s_wf[c * 64 + n] = w8[c];
    }
    for (int i = threadIdx.x; i < 64; i += THREADS) {
        s_bf[i] = (RAW_IN && p.b_first) ? p.b_first[i] : 0.f;
        s_bias[i] = p.bias[0] ? p.bias[0][i] : 0.f;
        s_bias[64 + i] = p.bias[1] ? p.bias[1][i] : 0.f;
        s_wh[i] = HEAD ? p.w_head[i] : 0.f;
        s_wh[64 + i] = HEAD ? p.w_head[64 + i] : 0.f;
    }
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int i = 0; i < 64; ++i) m = fmaxf(m, p.bias[0] ? fabsf(p.bias[0][i]) : 0.f);
        s_misc[0] = m;
    }
    fence_proxy_async_smem();
    if (warp == CONV_WARPS) tmem_alloc(tmem_holder, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_holder;
    const long long n_tiles = (p.n_rows + TILE - 1) / TILE;
    const int n_my = (int)((n_tiles - (long long)blockIdx.x + gridDim.x - 1) / gridDim.x);
    auto tile_row0 = [&](int i) { return ((long long)blockIdx.x + (long long)i * gridDim.x) * TILE; };

    if (warp < CONV_WARPS) {
        // =====================================================================================
        // converter warps: X0 rows -> per-row scale -> fp16 hi/lo operand; two tiles later: stage -> rows / head
        // =====================================================================================
        const int g = threadIdx.x >> 4, q = threadIdx.x & 15, q4 = 4 * q;          // 16 groups x 16 lanes; rows g + 16 k
        const int chunk = q >> 3, piece = q & 7;
        const uint32_t a_sub = ((uint32_t)(q & 1)) * 8u;
        const float w_descale0 = *reinterpret_cast<const float*>(p.img[0] + W_IMAGE);
        const Leaky a_f = leaky_of(p.act_first, p.slope_first), a_in = leaky_of(p.act_in, nullptr);
        float4 wf[8];                                                   // this lane's 4 output columns of the first layer
        float bf[4] = {0.f, 0.f, 0.f, 0.f};
        if (RAW_IN) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) wf[kk] = *reinterpret_cast<const float4*>(s_wf + kk * 64 + q4);
#pragma unroll
            for (int c = 0; c < 4; ++c) bf[c] = s_bf[q4 + c];
        }
        auto write_out = [&](int j) {                                   // ROW output only (the head runs in the epilogue warps)
            const long long r0 = tile_row0(j);
            {
                mbar_wait(&bar->st_full, (uint32_t)j & 1);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = g + 16 * k;
                    const float4 d = *reinterpret_cast<const float4*>(stage + r * STAGE_LD + q4);
                    if (r0 + r < p.n_rows) stg4(p.out_rows + (p.row_lo + r0 + r) * F + q4, d);
                }
            }
            mbar_arrive(&bar->st_empty);
        };
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const int s = i % STAGES, slot = i & 1;
            const long long r0 = tile_row0(i);
            const int rows = (int)min((long long)TILE, p.n_rows - r0);
            mbar_wait(&bar->in_full[s], (uint32_t)(i / STAGES) & 1);
            float4 x[8];
            if (RAW_IN) {
                const float* rawt = reinterpret_cast<const float*>(in_ring + (size_t)s * RAW_TILE);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = g + 16 * k;
                    const float4 lo = *reinterpret_cast<const float4*>(rawt + r * 8), hi = *reinterpret_cast<const float4*>(rawt + r * 8 + 4);
                    float4 acc = make_float4(bf[0], bf[1], bf[2], bf[3]);
#define RM16_FMA(in_, w_) acc.x = fmaf(in_, w_.x, acc.x); acc.y = fmaf(in_, w_.y, acc.y); acc.z = fmaf(in_, w_.z, acc.z); acc.w = fmaf(in_, w_.w, acc.w)
                    RM16_FMA(lo.x, wf[0]); RM16_FMA(lo.y, wf[1]); RM16_FMA(lo.z, wf[2]); RM16_FMA(lo.w, wf[3]);
                    RM16_FMA(hi.x, wf[4]); RM16_FMA(hi.y, wf[5]); RM16_FMA(hi.z, wf[6]); RM16_FMA(hi.w, wf[7]);
#undef RM16_FMA
                    x[k] = r < rows ? make_float4(leaky_do(a_f, acc.x), leaky_do(a_f, acc.y), leaky_do(a_f, acc.z), leaky_do(a_f, acc.w))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else {
                const float* src = reinterpret_cast<const float*>(in_ring + (size_t)s * ROW_TILE);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int r = g + 16 * k;
                    x[k] = r < rows ? *reinterpret_cast<const float4*>(src + r * F + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            // The ring slot goes back to the loader only when this thread's shared-memory reads have RETURNED: the arrive's
            // address is made to depend on the loaded values (an arrive issued right behind the LDS could overtake them in
            // the memory pipeline, and the loader's next copy then lands on rows that are still being read — seen as a
            // run-to-run difference of the decoder on a 21-tiles-per-CTA mesh).
            {
                uint32_t dep = 0u, zero;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    dep |= __float_as_uint(x[k].x) | __float_as_uint(x[k].y) | __float_as_uint(x[k].z) | __float_as_uint(x[k].w);
                asm volatile("and.b32 %0, %1, %2;" : "=r"(zero) : "r"(dep), "r"(p.zero_mask));
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar->in_empty[s]) + zero) : "memory");
            }
            if (!RAW_IN && p.act_in != SWE_ACT_NONE) {
                if (p.act_in == SWE_ACT_TANH) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) { x[k].x = tanh16(x[k].x); x[k].y = tanh16(x[k].y); x[k].z = tanh16(x[k].z); x[k].w = tanh16(x[k].w); }
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        x[k].x = leaky_do(a_in, x[k].x); x[k].y = leaky_do(a_in, x[k].y);
                        x[k].z = leaky_do(a_in, x[k].z); x[k].w = leaky_do(a_in, x[k].w);
                    }
                }
            }
            float sc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float m = fmaxf(fmaxf(fabsf(x[k].x), fabsf(x[k].y)), fmaxf(fabsf(x[k].z), fabsf(x[k].w)));
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
                uint32_t sb = 267u - (__float_as_uint(m) >> 23);
                sb = sb > 253u ? 253u : sb;
                sc[k] = __uint_as_float(sb << 23);
            }
            mbar_wait(&bar->a_empty[slot], (((uint32_t)i >> 1) & 1) ^ 1);
            if (q == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k)                              // 2^-e = 1 / sc exactly: exponent 254 - biased exponent of sc
                    s_inv0[(i & 3) * TILE + g + 16 * k] = __uint_as_float((254u << 23) - __float_as_uint(sc[k])) * w_descale0;
            }
            unsigned char* base = a_slots + (size_t)slot * A_SLOT + (size_t)chunk * 2 * A_TILE;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t off = sw64_piece_offset(g + 16 * k, piece >> 1) + a_sub;
                uint2 hh, ll;
                split_f16x2(x[k].x * sc[k], x[k].y * sc[k], hh.x, ll.x);
                split_f16x2(x[k].z * sc[k], x[k].w * sc[k], hh.y, ll.y);
                *reinterpret_cast<uint2*>(base + off) = hh;
                *reinterpret_cast<uint2*>(base + A_TILE + off) = ll;
            }
            fence_proxy_async_smem();
            mbar_arrive(&bar->a_full[slot]);
            if (!HEAD && i >= 2) write_out(i - 2);
        }
        if (!HEAD && n_my >= 2) write_out(n_my - 2);
        if (!HEAD && n_my >= 1) write_out(n_my - 1);
    } else if (warp < CONV_WARPS + EPI_WARPS) {
        // =====================================================================================
        // epilogue warps (thread = TMEM lane = row):  E0(0) ; for i: { E0(i + 1) ; E1(i) }
        // =====================================================================================
        const int lq = warp & 3, row = lq * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(lq * 32) << 16);
        const Leaky a0 = leaky_of(p.act[0], p.slope[0]), a1 = leaky_of(p.act[1], p.slope[1]);
        const float w_descale1 = *reinterpret_cast<const float*>(p.img[1] + W_IMAGE);
        const float bias0_max = s_misc[0];
        float* my_row = stage + row * STAGE_LD;
        const Leaky a_h = leaky_of(p.act_head, p.slope_head);
        const int n_static_raw = p.n_cols - 2 * p.previous_t;
        float bh0 = 0.f, bh1 = 0.f;
        if (HEAD && p.b_head) { bh0 = __ldg(p.b_head); bh1 = __ldg(p.b_head + 1); }
        // X1 = act_0(D0 descaled + b_0) -> row scale -> fp16 hi/lo -> TMEM operand of layer 1
        auto e0 = [&](int i) {
            const uint32_t b = (uint32_t)i & 1, bph = ((uint32_t)i >> 1) & 1;
            mbar_wait(&bar->d0_full[b], bph);
            tc_fence_after_sync();
            const float dsc = s_inv0[(i & 3) * TILE + row];
            // pass 1: an upper bound of the row's largest |activation| (|act(v)| <= max(1, |slope|) (|D| dsc + max |b|))
            float m = 0.f;
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t v[32];
                tmem_ld32(lane_addr + C_D0 + b * 64 + hf * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) m = fmaxf(m, fabsf(__uint_as_float(v[j])));
            }
            m = fmaf(m, fabsf(dsc), bias0_max) * fmaxf(1.f, fabsf(a0.slope));
            uint32_t sb = 267u - (__float_as_uint(m) >> 23);
            sb = sb > 253u ? 253u : sb;
            const float scale = __uint_as_float(sb << 23);
            s_inv1[b * TILE + row] = __uint_as_float((254u - sb) << 23) * w_descale1;
            // pass 2: 16 columns at a time -> 8 packed columns of hi and of lo
#pragma unroll 1
            for (int cb = 0; cb < 4; ++cb) {
                uint32_t v[16], hi[8], lo[8];
                tmem_ld16(lane_addr + C_D0 + b * 64 + cb * 16, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float t0 = leaky_do(a0, fmaf(__uint_as_float(v[2 * j]), dsc, s_bias[cb * 16 + 2 * j]));
                    const float t1 = leaky_do(a0, fmaf(__uint_as_float(v[2 * j + 1]), dsc, s_bias[cb * 16 + 2 * j + 1]));
                    split_f16x2(t0 * scale, t1 * scale, hi[j], lo[j]);
                }
                tmem_st8(lane_addr + C_AHI + b * 32 + cb * 8, hi);
                tmem_st8(lane_addr + C_ALO + b * 32 + cb * 8, lo);
            }
            tmem_wait_st();
            tc_fence_before_sync();
            mbar_arrive(&bar->d0_free[b]);
            mbar_arrive(&bar->x1_ready[b]);
        };
#pragma unroll 1
        for (int i = -1; i < n_my; ++i) {
            if (i + 1 < n_my) e0(i + 1);
            if (i < 0) continue;
            const uint32_t b = (uint32_t)i & 1, bph = ((uint32_t)i >> 1) & 1;
            // head: this row's inputs (original row number through the permutation, then its n_cols <= 16 input columns) are
            // requested before the wait for the accumulator
            float hx[16];
            long long horow = 0;
            if (HEAD) {
#pragma unroll
                for (int c = 0; c < 16; ++c) hx[c] = 0.f;
                const long long grow = tile_row0(i) + row;
                if (grow < p.n_rows) {
                    const long long node = p.row_lo + grow;
                    horow = p.head_perm ? p.head_perm[node] : node;
                    const float* xr = p.x0 + horow * p.n_cols;
#pragma unroll
                    for (int c = 0; c < 16; ++c) if (c < p.n_cols) hx[c] = __ldg(xr + c);
                }
            }
            mbar_wait(&bar->d1_full[b], bph);
            tc_fence_after_sync();
            const float dsc = s_inv1[b * TILE + row];
            if (!HEAD) {
                mbar_wait(&bar->st_empty, ((uint32_t)i & 1) ^ 1);       // stage written out (tile i - 1)
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[32];
                    tmem_ld32(lane_addr + C_D1 + b * 64 + hf * 32, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 r;
                        r.x = leaky_do(a1, fmaf(__uint_as_float(v[j]), dsc, s_bias[64 + hf * 32 + j]));
                        r.y = leaky_do(a1, fmaf(__uint_as_float(v[j + 1]), dsc, s_bias[64 + hf * 32 + j + 1]));
                        r.z = leaky_do(a1, fmaf(__uint_as_float(v[j + 2]), dsc, s_bias[64 + hf * 32 + j + 2]));
                        r.w = leaky_do(a1, fmaf(__uint_as_float(v[j + 3]), dsc, s_bias[64 + hf * 32 + j + 3]));
                        *reinterpret_cast<float4*>(my_row + hf * 32 + j) = r;
                    }
                }
            } else {
                // decoder head in the thread that owns the row: 64 -> 2 dot products straight from the accumulator, residual,
                // ReLU, dry mask, prediction and shifted window (models/gnn.py:339-348, models/models.py:50-91,
                // utils/dataset.py:508-529) — no stage, no hand-over
                float y0 = 0.f, y1 = 0.f;
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t v[32];
                    tmem_ld32(lane_addr + C_D1 + b * 64 + hf * 32, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float t = leaky_do(a1, fmaf(__uint_as_float(v[j]), dsc, s_bias[64 + hf * 32 + j]));
                        y0 = fmaf(t, s_wh[hf * 32 + j], y0);
                        y1 = fmaf(t, s_wh[64 + hf * 32 + j], y1);
                    }
                }
                const long long grow = tile_row0(i) + row;
                if (grow < p.n_rows) {
                    float c0 = 0.f, c1 = 0.f;
                    if (p.res_mode != 0) {
                        // column c of the inputs contributes to variable (c - n_static) & 1 of time step (c - n_static) >> 1
                        // (static register indices: a runtime index would put hx in local memory)
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const int rel = c - n_static_raw;
                            if (rel >= 0 && c < p.n_cols) {
                                const int t = rel >> 1, jv = rel & 1;
                                const float w = p.res_mode == 1 ? __ldg(p.res_w + t) : p.res_mode == 2 ? __ldg(p.res_w + 2 * t + jv)
                                                : (t == p.previous_t - 1 ? 1.f : 0.f);
                                if (jv) c1 = fmaf(hx[c], w, c1); else c0 = fmaf(hx[c], w, c0);
                            }
                        }
                    }
                    y0 = fmaxf(leaky_do(a_h, y0 + bh0) + c0, 0.f);
                    y1 = fmaxf(leaky_do(a_h, y1 + bh1) + c1, 0.f);
                    const float oh = (fabsf(y0) > p.eps) ? y0 : 0.f;     // h · [|h| > eps]
                    const float oq = (y0 != 0.f) ? y1 : 0.f;              // q · [h != 0] (un-thresholded h)
                    float* pr = p.pred + (p.step_ptr ? (long long)(*p.step_ptr) * p.pred_step_stride : 0) + horow * 2;
                    *reinterpret_cast<float2*>(pr) = make_float2(oh, oq);
                    if (p.x_next) {
                        float* xn = p.x_next + horow * p.n_cols;
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            if (c < p.n_cols)
                                xn[c] = c < n_static_raw ? hx[c] : (c < p.n_cols - 2 ? hx[c + 2 < 16 ? c + 2 : 15] : (c == p.n_cols - 2 ? oh : oq));
                    }
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&bar->d1_free[b]);
            if (!HEAD) mbar_arrive(&bar->st_full);
        }
    } else if (warp == CONV_WARPS + EPI_WARPS) {
        // =====================================================================================
        // MMA issuer (converged warp):  L0(0) ; for i: { L0(i + 1) ; L1(i) }
        // =====================================================================================
        const uint32_t idesc = make_idesc_f16(TILE, F);
        const uint64_t desc_hi = make_desc_sw64(0) & 0xFFFFFFFF00000000ull;
        const uint32_t desc_lo0 = (uint32_t)make_desc_sw64(0);
        auto dsc = [&](uint32_t lo) { return desc_hi | (uint64_t)lo; };
        const uint32_t a_d = desc_lo0 + (smem_u32(a_slots) >> 4), w_d = desc_lo0 + (smem_u32(w_tile) >> 4);
        auto l0 = [&](int i) {
            const uint32_t b = (uint32_t)i & 1, bph = ((uint32_t)i >> 1) & 1;
            mbar_wait(&bar->d0_free[b], bph ^ 1);                       // D0[b] of tile i - 2 has been read
            mbar_wait(&bar->a_full[b], bph);
            tc_fence_after_sync();
            const uint32_t d0 = tmem_base + C_D0 + b * 64;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t a_hi = a_d + (b * A_SLOT + (uint32_t)c * 2 * A_TILE) / 16, a_lo = a_hi + A_TILE / 16;
                const uint32_t w_hi = w_d + ((uint32_t)c * 2 * W_TILE) / 16, w_lo = w_hi + W_TILE / 16;
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                    mma_f16_ss_warp(d0, dsc(a_lo + 2 * ks), dsc(w_hi + 2 * ks), idesc, (c | ks) ? 1u : 0u);
                    mma_f16_ss_warp(d0, dsc(a_hi + 2 * ks), dsc(w_lo + 2 * ks), idesc, 1u);
                    mma_f16_ss_warp(d0, dsc(a_hi + 2 * ks), dsc(w_hi + 2 * ks), idesc, 1u);
                }
            }
            mma_commit_warp(&bar->a_empty[b]);
            mma_commit_warp(&bar->d0_full[b]);
        };
#pragma unroll 1
        for (int i = -1; i < n_my; ++i) {
            if (i + 1 < n_my) l0(i + 1);
            if (i < 0) continue;
            const uint32_t b = (uint32_t)i & 1, bph = ((uint32_t)i >> 1) & 1;
            mbar_wait(&bar->x1_ready[b], bph);
            mbar_wait(&bar->d1_free[b], bph ^ 1);                       // D1[b] of tile i - 2 has been staged
            tc_fence_after_sync();
            const uint32_t d1 = tmem_base + C_D1 + b * 64, xh = tmem_base + C_AHI + b * 32, xl = tmem_base + C_ALO + b * 32;
            const uint32_t w1 = w_d + (uint32_t)(W_IMAGE / 16);
#pragma unroll
            for (int ks = 0; ks < F / 16; ++ks) {
                const uint32_t w_hi = w1 + ((uint32_t)(ks >> 1) * 2 * W_TILE) / 16 + 2 * (ks & 1), w_lo = w_hi + W_TILE / 16;
                mma_f16_ts_warp(d1, xl + ks * 8, dsc(w_hi), idesc, ks ? 1u : 0u);
                mma_f16_ts_warp(d1, xh + ks * 8, dsc(w_lo), idesc, 1u);
                mma_f16_ts_warp(d1, xh + ks * 8, dsc(w_hi), idesc, 1u);
            }
            mma_commit_warp(&bar->d1_full[b]);
        }
    } else {
        // =====================================================================================
        // loader warp
        // =====================================================================================
        if (RAW_IN) {
            // a row's 32 raw bytes through the node permutation, two 16-byte cp.async per row, four rows per lane
#pragma unroll 1
            for (int i = 0; i < n_my; ++i) {
                const int s = i % STAGES;
                const long long r0 = tile_row0(i);
                const int rows = (int)min((long long)TILE, p.n_rows - r0);
                mbar_wait(&bar->in_empty[s], (((uint32_t)(i / STAGES)) & 1) ^ 1);
                unsigned char* dstt = in_ring + (size_t)s * RAW_TILE;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = lane + 32 * u;
                    if (r < rows) {
                        const long long node = p.row_lo + r0 + r;
                        const float* xr = p.raw + (long long)(p.perm ? __ldg(p.perm + node) : node) * 8;
                        cp_async16(dstt + r * 32, xr);
                        cp_async16(dstt + r * 32 + 16, xr + 4);
                    }
                }
                cp_async_arrive_noinc(&bar->in_full[s]);
            }
        } else if (lane == 0) {
#pragma unroll 1
            for (int i = 0; i < n_my; ++i) {
                const int s = i % STAGES;
                const long long r0 = tile_row0(i);
                const uint32_t bytes = (uint32_t)(min((long long)TILE, p.n_rows - r0) * F * 4);
                mbar_wait(&bar->in_empty[s], (((uint32_t)(i / STAGES)) & 1) ^ 1);
                mbar_arrive_expect_tx(&bar->in_full[s], bytes);
                bulk_g2s(in_ring + (size_t)s * ROW_TILE, p.x_rows + (p.row_lo + r0) * F, bytes, &bar->in_full[s]);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == CONV_WARPS) tmem_dealloc(tmem_base, 512);
}

__host__ inline bool leaky_family(int act) {
    return act == SWE_ACT_NONE || act == SWE_ACT_PRELU || act == SWE_ACT_RELU || act == SWE_ACT_LEAKYRELU;
}

}  // namespace rm16
}  // namespace swe

using namespace swe;

// swe_row_mlp_tc's descriptor with the two tensor-core layers as swe_hop_tc16_pack images (img16[0], img16[1]).
// Returns SWE_E_UNSUPP for the shapes this kernel does not cover (the caller keeps swe_row_mlp_tc for those):
// n_tc != 2, activations outside none / relu / leakyrelu / prelu in the layers or the head (tanh allowed as the
// decoder's input activation), raw rows that are not 8 aligned floats.
extern "C" int swe_row_mlp_tc16(const swe_rowmlp_t* d, const void* const* img16, void* stream) {
    SWE_REQUIRE(d && img16, SWE_E_INVAL, "row_mlp_tc16: null descriptor");
    SWE_REQUIRE(d->n_rows >= 0 && d->row_lo >= 0, SWE_E_INVAL, "row_mlp_tc16: bad sizes");
    SWE_REQUIRE((d->x_rows != nullptr) != (d->raw != nullptr), SWE_E_INVAL, "row_mlp_tc16: exactly one of x_rows / raw");
    SWE_REQUIRE((d->out_rows != nullptr) != (d->head != 0), SWE_E_INVAL, "row_mlp_tc16: exactly one of out_rows / head");
    const bool raw_ok = !d->raw || (d->w_first && d->raw_ld == 8 && d->raw_cols >= 1 && d->raw_col0 >= 0 &&
                                    d->raw_col0 + d->raw_cols <= 8 && (!d->with_wl || (d->wl_col_a < 8 && d->wl_col_b < 8 &&
                                    d->wl_col_a >= 0 && d->wl_col_b >= 0)) && d->raw_cols + (d->with_wl ? 1 : 0) <= 8 &&
                                    aligned16(d->raw) && rm16::leaky_family(d->act_first));
    const bool in_ok = !d->x_rows || (aligned16(d->x_rows) && (d->act_in == SWE_ACT_TANH || (rm16::leaky_family(d->act_in) && d->act_in != SWE_ACT_PRELU)));
    const bool head_ok = !d->head || (d->w_head && d->x0 && d->pred && d->previous_t >= 1 && d->n_cols > 2 * d->previous_t &&
                                      d->n_cols <= 16 && d->res_mode >= 0 && d->res_mode <= 3 &&
                                      (d->res_mode == 0 || d->res_mode == 3 || d->res_w) && rm16::leaky_family(d->act_head));
    if (!(d->n_tc == 2 && raw_ok && in_ok && head_ok && rm16::leaky_family(d->act[0]) && rm16::leaky_family(d->act[1]) &&
          img16[0] && img16[1] && aligned16(img16[0]) && aligned16(img16[1]) && (!d->out_rows || aligned16(d->out_rows)))) {
        set_error("row_mlp_tc16: shape not covered by the fp16 streaming kernel");
        return SWE_E_UNSUPP;
    }
    if (d->n_rows == 0) return 0;
    rm16::Params p;
    memset(&p, 0, sizeof(p));
    p.x_rows = d->x_rows; p.act_in = d->act_in;
    p.zero_mask = 0u;
    p.raw = d->raw; p.raw_col0 = d->raw_col0; p.raw_cols = d->raw_cols; p.with_wl = d->with_wl;
    p.wl_col_a = d->wl_col_a; p.wl_col_b = d->wl_col_b; p.raw_k = d->raw_cols + (d->with_wl ? 1 : 0); p.perm = d->perm;
    p.w_first = d->w_first; p.b_first = d->b_first; p.act_first = d->act_first; p.slope_first = d->slope_first;
    p.row_lo = d->row_lo; p.n_rows = d->n_rows;
    for (int l = 0; l < 2; ++l) {
        p.img[l] = (const unsigned char*)img16[l]; p.bias[l] = d->bias[l]; p.act[l] = d->act[l]; p.slope[l] = d->slope[l];
    }
    p.out_rows = d->out_rows; p.head = d->head; p.w_head = d->w_head; p.b_head = d->b_head; p.act_head = d->act_head;
    p.slope_head = d->slope_head; p.x0 = d->x0; p.n_cols = d->n_cols; p.head_perm = d->head_perm; p.previous_t = d->previous_t;
    p.res_mode = d->res_mode; p.res_w = d->res_w; p.eps = d->eps; p.pred = d->pred; p.step_ptr = d->step_ptr;
    p.pred_step_stride = d->pred_step_stride; p.x_next = d->x_next;
    void (*kern)(const rm16::Params) = d->raw ? (d->head ? rm16::row_mlp_tc16_kernel<true, true> : rm16::row_mlp_tc16_kernel<true, false>)
                                              : (d->head ? rm16::row_mlp_tc16_kernel<false, true> : rm16::row_mlp_tc16_kernel<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rm16::SMEM);
    if (e != cudaSuccess) { set_error("row_mlp_tc16 smem opt-in (%zu B): %s", rm16::SMEM, cudaGetErrorString(e)); return (int)e; }
    const long long n_tiles = (d->n_rows + rm16::TILE - 1) / rm16::TILE;
    kern<<<grid_for(n_tiles, 1), rm16::THREADS, rm16::SMEM, (cudaStream_t)stream>>>(p);
    return check_launch("row_mlp_tc16");
}
