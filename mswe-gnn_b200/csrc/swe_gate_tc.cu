// Edge gate on the 5th-generation tensor cores (tcgen05 + TMEM), F = 64, 3-layer edge MLP
// (5F|4F -> 2F -> 2F -> F) — the default config.yaml model.  Replaces models/gnn.py:414-426.
//
// Precision: fp32 parity (north_star: rel 1e-5 per layer) rules out a single TF32 pass (2^-11).
// Every product is evaluated as 3 TF32 MMAs on error-free splits x = hi + lo
// (hi = x with the low 13 mantissa bits cleared, lo = x - hi, both exact in fp32):
//     A·W ≈ A_hi·W_hi + A_lo·W_hi + A_hi·W_lo          (dropped term |A_lo·W_lo| <= 2^-22 |A||W|)
// accumulated in fp32 in TMEM — per-layer relative error ≈ 1e-6.
//
// One CTA works on 128-edge tiles (persistent, two tiles in flight), 320 threads:
//   warps 0-7  row workers : gather the layer-0 input K-chunks (x_s[r], x_s[c], x_d[r], x_d[c], a_e),
//                            split hi/lo, store them to shared memory in the UMMA K-major SWIZZLE_128B
//                            layout; later the epilogues (one thread = one TMEM lane = one edge):
//                            tcgen05.ld accumulators -> bias + PReLU -> split -> tcgen05.st as the next
//                            layer's A operand (A stays in TMEM, never touches shared memory) and
//                            finally L2-normalise + store s_ij.
//   warp 8     weight loader: streams pre-packed weight chunk images (already swizzled, hi|lo) from
//                            global/L2 into a shared-memory ring with cp.async.bulk + mbarrier tx.
//   warp 9     MMA issuer   : one thread issues tcgen05.mma (SS for layer 0, TS for layers 1-2) and
//                            tcgen05.commit to release ring slots / publish accumulators.
// TMEM (512 columns): [0,128) A_hi, [128,256) A_lo, [256,384) D_a (layer 0), [384,512) D_b (layers 1, 2).
#include "swe_tc.cuh"

namespace swe {
namespace tc {

constexpr int GF = 64;                    // feature width this kernel is built for
constexpr int GH = 128;                   // hidden width 2F
constexpr int KC = 32;                    // K elements per chunk = one 128-byte swizzled row
constexpr int TILE_ROWS = 128;
constexpr int TILE16K = TILE_ROWS * 128;  // bytes of a [128 x 32] tf32 tile
constexpr int SLOT_BYTES = 2 * TILE16K;   // hi tile + lo tile
constexpr int A_STAGES = 2;
constexpr int W_STAGES = 4;
constexpr int N_ROW_THREADS = 256;        // warps 0-7
constexpr int N_THREADS = 320;            // + loader warp 8 + MMA warp 9
constexpr uint32_t COL_A_HI = 0, COL_A_LO = 128, COL_D_A = 256, COL_D_B = 384;

struct __align__(8) Barriers {
    uint64_t a_full[A_STAGES], a_empty[A_STAGES];
    uint64_t w_full[W_STAGES], w_empty[W_STAGES];
    uint64_t d_full[3];        // accumulators of layer 0/1/2 complete (tcgen05.commit)
    uint64_t a_ready[2];       // next layer's A operand written to TMEM (256 arrivals)
};

constexpr size_t GATE_TC_SMEM = 1024 /*align slack*/ + (size_t)(A_STAGES + W_STAGES) * SLOT_BYTES +
                                sizeof(float) * 320 + sizeof(int32_t) * 4 * TILE_ROWS + sizeof(Barriers) + 16;

// image layout of the packed weights (bytes)
__host__ __device__ constexpr size_t img_l1_off(int chunk) { return (size_t)chunk * SLOT_BYTES; }
__host__ __device__ constexpr size_t img_l2_off(int n_l1, int chunk) { return (size_t)(n_l1 + chunk) * SLOT_BYTES; }
__host__ __device__ constexpr size_t img_l3_off(int n_l1, int chunk) { return (size_t)(n_l1 + 4) * SLOT_BYTES + (size_t)chunk * (SLOT_BYTES / 2); }
__host__ __device__ constexpr size_t img_bias_off(int n_l1) { return (size_t)(n_l1 + 4) * SLOT_BYTES + 4 * (size_t)(SLOT_BYTES / 2); }
__host__ __device__ constexpr size_t img_bytes(int n_l1) { return img_bias_off(n_l1) + 320 * sizeof(float); }

// ---------------------------------------------------------------------------------------------
// weight packing: Linear weights [n_out, k_in] (row-major = K-major) -> swizzled hi|lo chunk images
// ---------------------------------------------------------------------------------------------
__global__ void gate_tc_pack_kernel(const float* __restrict__ w1, int k1, const float* __restrict__ b1,
                                    const float* __restrict__ w2, const float* __restrict__ b2,
                                    const float* __restrict__ w3, const float* __restrict__ b3,
                                    unsigned char* __restrict__ img) {
    const int n_l1 = k1 / KC;
    const int total = GH * k1 + GH * GH + GF * GH;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const float* w; int n, k, kin, rows; size_t base;
        if (idx < GH * k1) { w = w1; kin = k1; n = idx / k1; k = idx % k1; rows = GH; base = img_l1_off(k / KC); }
        else if (idx < GH * k1 + GH * GH) { int j = idx - GH * k1; w = w2; kin = GH; n = j / GH; k = j % GH; rows = GH; base = img_l2_off(n_l1, k / KC); }
        else { int j = idx - GH * k1 - GH * GH; w = w3; kin = GH; n = j / GH; k = j % GH; rows = GF; base = img_l3_off(n_l1, k / KC); }
        float hi, lo;
        split_tf32(w[(size_t)n * kin + k], hi, lo);
        const uint32_t off = sw128_offset(n, k % KC);
        *reinterpret_cast<float*>(img + base + off) = hi;
        *reinterpret_cast<float*>(img + base + (size_t)rows * 128 + off) = lo;
    }
    float* bias = reinterpret_cast<float*>(img + img_bias_off(n_l1));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 320; i += gridDim.x * blockDim.x)
        bias[i] = i < 128 ? (b1 ? b1[i] : 0.f) : i < 256 ? (b2 ? b2[i - 128] : 0.f) : (b3 ? b3[i - 256] : 0.f);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
struct GateTcParams {
    const float* xs; const float* xd_src; const float* xd_dst; const float* a;
    const int32_t* src; const int32_t* dst;
    long long n_edges;
    const unsigned char* img;   // packed weights
    int n_l1_img;               // chunks of layer 0 present in the image (k1 / 32)
    int act[3]; const float* slope[3];
    int normalize;
    float* s_out;
    float* dbg;                 // optional [128*128 + 128*128 + 128*64] raw accumulators of tile 0
    long long* trace;           // optional [3 roles][16 tiles][8 events] clock64 stamps of CTA 0 (profiling aid)
    // layer-0 input segments gathered and multiplied on the tensor core, in order (0 x_s[r], 1 x_s[c], 2 x_d[r],
    // 3 x_d[c], 4 a_e); MODE 0: all present ones; MODE 1 (decomposed): only a_e, the node part comes from the
    // per-node partial tables; MODE 2 (partials): the kernel IS the producer of such a table for a node range
    int n_seg; int segs[5];
    const float* p_src; const float* p_dst;      // MODE 1: [*, 128] partial tables indexed by plan id; MODE 3: p_src = per-EDGE table
    float* p_out; int row_lo;                    // MODE 2: output table rows [row_lo, row_lo + n_edges); src/dst given: per-edge table
    float* pre_out[3];                           // TRAIN: pre-activations of the three layers ([E,128], [E,128], [E,64])
    // TRAIN: work lists of the pre-activations that lie within the 3xTF32 error of the activation's kink (|pre| < tau):
    // entry = edge << 8 | column; swe_gate_fix_preacts re-evaluates exactly those in exact fp32 (see there)
    unsigned long long* fix_list[3]; int* fix_count; int fix_cap; float fix_tau;
    // list mode (MODE 0): only the tiles tile_list[1 .. tile_list[0]] are evaluated — the tiles the fp16 kernel
    // (swe_gate_tc16.cu) found outside its input window
    const int32_t* tile_list;
};

__device__ __forceinline__ long long gate_tile_of(const GateTcParams& p, int i) {
    const long long t = (long long)blockIdx.x + (long long)i * gridDim.x;
    return p.tile_list ? (long long)p.tile_list[1 + t] : t;
}

__device__ __forceinline__ int l1_chunk_segment(const GateTcParams& p, int i) {
    // i-th ACTIVE chunk of layer 0 -> input segment (two 32-column chunks per 64-wide segment)
    return p.segs[i >> 1];
}

// "leaky family" activations (none / relu / leakyrelu / prelu) are v > 0 ? v : slope * v
template <bool GENERIC>
__device__ __forceinline__ float gate_act(int act, float v, float slope) {
    if (GENERIC) return act_apply(act, v, slope);
    return fmaxf(v, 0.f) + slope * fminf(v, 0.f);
}
__device__ __forceinline__ float leaky_slope(int act, const float* slope_p) {
    switch (act) {
        case SWE_ACT_PRELU:     return slope_p ? __ldg(slope_p) : 0.25f;
        case SWE_ACT_RELU:      return 0.f;
        case SWE_ACT_LEAKYRELU: return 0.1f;
        case SWE_ACT_NONE:      return 1.f;
        default:                return (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
    }
}

// Tile schedule shared by the three roles (tiles t_0 .. t_{n-1} of this CTA, h = n_l1 / 2):
//   L1(t_0) ;  for i: { L2(t_i) ; L1(t_{i+1})[0,h) ; L3(t_i) ; L1(t_{i+1})[h,n_l1) }
// row workers:  G(t_0) ;  for i: { E1(t_i) ; G(t_{i+1})[0,h) ; E2(t_i) ; G(t_{i+1})[h,n_l1) ; E3(t_i) }
// so the tensor pipe works on the next tile's layer 0 while the row workers run this tile's
// epilogues, and on this tile's layers 1-2 while they gather the next tile's inputs.
//
// MODE 1 (decomposed layer 0): W1·z = (A·x_s[r] + C·x_d[r]) + (B·x_s[c] + D·x_d[c]) + E·a_e.  The two brackets only
// depend on ONE node, so they are evaluated once per node (MODE 2 launches of this kernel, 2·128² MAC per node)
// instead of once per edge (4·64·128 MAC per edge, ~3 edges per node); the per-edge layer-0 work shrinks from 10
// K-chunks to the 2 of a_e (none for un-pool calls), and epilogue 1 adds P_src[r] + P_dst[c] in fp32.
// TRAIN: the forward of the training step — additionally stores every layer's pre-activation (what the backward
// kernels differentiate through), otherwise identical.
//
// MODE 3 (static part hoisted out of the rollout): the encoded edge features a_e — and, for models built with
// with_WL=False, the encoded static node features x_s too — do not change over the steps of a rollout (mesh geometry,
// static inputs, frozen weights), so their share of layer 0, P[e] = E·a_e (+ A·x_s[r] + B·x_s[c]), is computed ONCE per
// rollout and call site into a per-edge table (a MODE 2 launch gathering through src/dst) and every step multiplies only
// the other blocks: 8 (or 4) K-chunks instead of 10 per tile, streaming P[e] (512 B per edge, coalesced) in epilogue 1.
// With with_WL=True (config.yaml) x_s contains the water level of the current step and is NOT static.
template <bool GENERIC, int MODE, bool TRAIN = false>
__global__ void __launch_bounds__(N_THREADS, 1) edge_gate_tc_kernel(const __grid_constant__ GateTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    // 1 KB alignment by OFFSETTING the shared array: integer arithmetic on the pointer value loses the address
    // space and turns every later shared access into a generic LD/ST
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_ring = smem;
    unsigned char* w_ring = smem + (size_t)A_STAGES * SLOT_BYTES;
    float* s_bias = reinterpret_cast<float*>(w_ring + (size_t)W_STAGES * SLOT_BYTES);
    int32_t* s_ids = reinterpret_cast<int32_t*>(s_bias + 320);            // [2 tiles][src 128 | dst 128]
    Barriers* bar = reinterpret_cast<Barriers*>(s_ids + 4 * TILE_ROWS);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_l1 = 2 * p.n_seg;                                            // active layer-0 chunks
    const int h_l1 = n_l1 / 2;
    long long n_tiles_all = (p.n_edges + TILE_ROWS - 1) / TILE_ROWS;
    if (p.tile_list) {
        n_tiles_all = min((long long)p.tile_list[0], n_tiles_all);
        if (n_tiles_all <= (long long)blockIdx.x) return;                             // (uniform over the CTA; usually: no tile at all)
    }
    const int n_my = (int)((n_tiles_all - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles of this CTA

    if (threadIdx.x == 0) {
        for (int i = 0; i < A_STAGES; ++i) { mbar_init(&bar->a_full[i], N_ROW_THREADS); mbar_init(&bar->a_empty[i], 1); }
        for (int i = 0; i < W_STAGES; ++i) { mbar_init(&bar->w_full[i], 1); mbar_init(&bar->w_empty[i], 1); }
        for (int i = 0; i < 3; ++i) mbar_init(&bar->d_full[i], 1);
        for (int i = 0; i < 2; ++i) mbar_init(&bar->a_ready[i], N_ROW_THREADS);
        fence_barrier_init();
    }
    {
        const float* gb = reinterpret_cast<const float*>(p.img + img_bias_off(p.n_l1_img));
        for (int i = threadIdx.x; i < 320; i += N_THREADS) s_bias[i] = gb[i];
    }
    if (warp == 8) tmem_alloc(tmem_holder, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_holder;

    if (warp < 8) {
        // =====================================================================================
        // row workers
        // =====================================================================================
        const int q = warp & 3, hf = warp >> 2;
        const int row = q * 32 + lane;                                // TMEM lane / edge within the tile
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        float sl[3];
#pragma unroll
        for (int l = 0; l < 3; ++l) sl[l] = GENERIC ? ((p.act[l] == SWE_ACT_PRELU && p.slope[l]) ? __ldg(p.slope[l]) : 0.f)
                                                    : leaky_slope(p.act[l], p.slope[l]);
        uint32_t a_cnt = 0;
        const int piece = threadIdx.x & 7, r0 = threadIdx.x >> 3;     // gather: 4 rows r0 + 32 i, one 16-B piece
        const uint32_t g_off0 = sw128_offset(r0, piece * 4);          // + i * 4096 for row r0 + 32 i

        auto load_ids = [&](int i) {                                  // endpoints of tile t_i -> s_ids[i & 1]
            const long long e0 = gate_tile_of(p, i) * TILE_ROWS;
            int32_t* ids = s_ids + (i & 1) * 2 * TILE_ROWS;
            if (threadIdx.x < TILE_ROWS) {
                long long e = e0 + threadIdx.x;
                if (e >= p.n_edges) e = p.n_edges - 1;
                ids[threadIdx.x] = (MODE == 2 && !p.src) ? (int32_t)(p.row_lo + e) : __ldg(p.src + e);
            } else {
                long long e = e0 + threadIdx.x - TILE_ROWS;
                if (e >= p.n_edges) e = p.n_edges - 1;
                ids[threadIdx.x] = (MODE == 2 && !p.src) ? (int32_t)(p.row_lo + e) : __ldg(p.dst + e);     // ids[128 + r]
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        };
        auto gather = [&](int i, int c_lo, int c_hi) {                // layer-0 input chunks [c_lo, c_hi) of tile t_i
            const long long e0 = gate_tile_of(p, i) * TILE_ROWS;
            const int32_t* ids = s_ids + (i & 1) * 2 * TILE_ROWS;
            auto issue = [&](int c, float4 (&v)[4]) {                 // 4 independent 16-B loads per thread
                const int sg = l1_chunk_segment(p, c);
                const int koff = (c & 1) * KC + piece * 4;
                if (sg == 4) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        long long e = e0 + r0 + 32 * j;
                        if (e >= p.n_edges) e = p.n_edges - 1;
                        v[j] = ldg4_stream(p.a + e * GF + koff);
                    }
                } else {
                    const float* base = sg < 2 ? p.xs : (sg == 2 ? p.xd_src : p.xd_dst);
                    const int32_t* idp = ids + ((sg & 1) ? TILE_ROWS : 0) + r0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = ldg4(base + (long long)idp[32 * j] * GF + koff);
                }
            };
            if (c_lo >= c_hi) return;                                  // decomposed layer 0 of an un-pool call: nothing to gather
            float4 cur[4], nxt[4];
            issue(c_lo, cur);
#pragma unroll 1
            for (int c = c_lo; c < c_hi; ++c, ++a_cnt) {
                if (c + 1 < c_hi) issue(c + 1, nxt);                  // next chunk's loads fly while this one is packed
                const uint32_t slot = a_cnt % A_STAGES;
                mbar_wait(&bar->a_empty[slot], ((a_cnt / A_STAGES) & 1) ^ 1);
                unsigned char* hi_t = a_ring + (size_t)slot * SLOT_BYTES + g_off0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float4 hh, ll;
                    split_tf32(cur[j].x, hh.x, ll.x); split_tf32(cur[j].y, hh.y, ll.y);
                    split_tf32(cur[j].z, hh.z, ll.z); split_tf32(cur[j].w, hh.w, ll.w);
                    *reinterpret_cast<float4*>(hi_t + j * 4096) = hh;
                    *reinterpret_cast<float4*>(hi_t + TILE16K + j * 4096) = ll;
                }
                fence_proxy_async_smem();
                mbar_arrive(&bar->a_full[slot]);
#pragma unroll
                for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
            }
        };
        // D -> bias, activation -> hi/lo -> TMEM A operand; this thread owns 64 columns of its lane
        auto epilogue_mid = [&](int layer, uint32_t ph, bool dump, int i) {
            // MODE 1, layer 0: the node part of W1·z comes from the per-node tables; fetch it before waiting for D
            float4 padd[16];
            const bool with_p = (MODE == 1 || MODE == 3) && layer == 0;
            if (MODE == 3 && layer == 0) {
                // per-edge table in tile-transposed order [tile][column half][16-B chunk][row]: the 32 rows of a warp
                // read 512 contiguous bytes per chunk (the table is padded to whole tiles)
                const long long tile = gate_tile_of(p, i);
                const float* pe = p.p_src + tile * (TILE_ROWS * GH) + hf * (TILE_ROWS * 64) + row * 4;
#pragma unroll
                for (int j = 0; j < 16; ++j) padd[j] = ldg4_stream(pe + j * (TILE_ROWS * 4));
            } else if (with_p) {
                const int32_t* ids = s_ids + (i & 1) * 2 * TILE_ROWS;
                const float* ps = p.p_src + (long long)ids[row] * GH + hf * 64;
                const float* pd = p.p_dst + (long long)ids[TILE_ROWS + row] * GH + hf * 64;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 x = ldg4(ps + 4 * j), y = ldg4(pd + 4 * j);
                    padd[j] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
                }
            }
            const bool have_d = !(with_p && n_l1 == 0);               // un-pool calls have no a_e: no layer-0 MMA at all
            if (have_d) {
                mbar_wait(&bar->d_full[layer], ph);
                tc_fence_after_sync();
            } else if (i > 0) {
                // nothing of this tile to wait for, but the activations written below are still being read by the
                // layer-2 MMA of the previous tile until that one completes
                mbar_wait(&bar->d_full[2], (uint32_t)(i - 1) & 1);
                tc_fence_after_sync();
            }
            const uint32_t dcol = (layer == 0 ? COL_D_A : COL_D_B) + hf * 64;
            const float* bias = s_bias + layer * 128 + hf * 64;
            const int act = p.act[layer];
            const float slope = sl[layer];
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
                uint32_t v[32];
                if (have_d) {
                    tmem_ld32(lane_addr + dcol + cb * 32, v);
                    tmem_wait_ld();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
                if (with_p) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 t = padd[cb * 8 + j];
                        v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + t.x);
                        v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + t.y);
                        v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + t.z);
                        v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + t.w);
                    }
                }
                if (dump) {
                    float* d = p.dbg + (size_t)layer * 128 * 128 + (size_t)row * 128 + hf * 64 + cb * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) d[j] = __uint_as_float(v[j]);
                }
                if (TRAIN) {
                    const long long e = gate_tile_of(p, i) * TILE_ROWS + row;
                    if (e < p.n_edges) {
                        float* d = p.pre_out[layer] + e * GH + hf * 64 + cb * 32;
                        float mn = 3.4e38f, mx = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = *reinterpret_cast<const float4*>(bias + cb * 32 + j);
                            const float4 q = make_float4(__uint_as_float(v[j]) + b4.x, __uint_as_float(v[j + 1]) + b4.y,
                                                         __uint_as_float(v[j + 2]) + b4.z, __uint_as_float(v[j + 3]) + b4.w);
                            stg4(d + j, q);
                            mn = fminf(fminf(mn, fminf(fabsf(q.x), fabsf(q.y))), fminf(fabsf(q.z), fabsf(q.w)));
                            mx = fmaxf(fmaxf(mx, fmaxf(fabsf(q.x), fabsf(q.y))), fmaxf(fabsf(q.z), fabsf(q.w)));
                        }
                        const float tau = p.fix_tau * fmaxf(1.f, mx);
                        if (p.fix_count && mn < tau) {                  // rare: some pre-activation sits on the kink
                            uint32_t mask = 0;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                mask |= (fabsf(__uint_as_float(v[j]) + bias[cb * 32 + j]) < tau ? 1u : 0u) << j;
                            while (mask) {
                                const int j = __ffs(mask) - 1;
                                mask &= mask - 1;
                                const int k = atomicAdd(p.fix_count + layer, 1);
                                if (k < p.fix_cap)
                                    p.fix_list[layer][k] = ((unsigned long long)e << 8) | (unsigned)(hf * 64 + cb * 32 + j);
                            }
                        }
                    }
                }
                uint32_t lo[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias + cb * 32 + j);
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float y = gate_act<GENERIC>(act, __uint_as_float(v[j + u]) + bb[u], slope);
                        const float hh = round_tf32(y);
                        v[j + u] = __float_as_uint(hh);
                        lo[j + u] = __float_as_uint(y - hh);
                    }
                }
                tmem_st32(lane_addr + COL_A_HI + hf * 64 + cb * 32, v);
                tmem_st32(lane_addr + COL_A_LO + hf * 64 + cb * 32, lo);
            }
            tmem_wait_st();
            tc_fence_before_sync();
            mbar_arrive(&bar->a_ready[layer]);
        };
        auto epilogue_final = [&](int i, uint32_t ph, bool dump) {    // warps 0-3: 64 columns each
            if (hf != 0) return;
            mbar_wait(&bar->d_full[2], ph);
            tc_fence_after_sync();
            uint32_t v0[32], v1[32];
            tmem_ld32(lane_addr + COL_D_B, v0);
            tmem_ld32(lane_addr + COL_D_B + 32, v1);
            tmem_wait_ld();
            tc_fence_before_sync();
            if (dump) {
                float* d = p.dbg + (size_t)2 * 128 * 128 + (size_t)row * 64;
#pragma unroll
                for (int j = 0; j < 32; ++j) { d[j] = __uint_as_float(v0[j]); d[32 + j] = __uint_as_float(v1[j]); }
            }
            const float* bias = s_bias + 256;
            if (TRAIN) {
                const long long e = gate_tile_of(p, i) * TILE_ROWS + row;
                if (e < p.n_edges) {
                    float* d = p.pre_out[2] + e * GF;
                    float mn = 3.4e38f, mx = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b0 = *reinterpret_cast<const float4*>(bias + j), b1 = *reinterpret_cast<const float4*>(bias + 32 + j);
                        const float4 q0 = make_float4(__uint_as_float(v0[j]) + b0.x, __uint_as_float(v0[j + 1]) + b0.y,
                                                      __uint_as_float(v0[j + 2]) + b0.z, __uint_as_float(v0[j + 3]) + b0.w);
                        const float4 q1 = make_float4(__uint_as_float(v1[j]) + b1.x, __uint_as_float(v1[j + 1]) + b1.y,
                                                      __uint_as_float(v1[j + 2]) + b1.z, __uint_as_float(v1[j + 3]) + b1.w);
                        stg4(d + j, q0);
                        stg4(d + 32 + j, q1);
                        mn = fminf(fminf(mn, fminf(fabsf(q0.x), fabsf(q0.y))), fminf(fabsf(q0.z), fabsf(q0.w)));
                        mn = fminf(fminf(mn, fminf(fabsf(q1.x), fabsf(q1.y))), fminf(fabsf(q1.z), fabsf(q1.w)));
                        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(q0.x), fabsf(q0.y))), fmaxf(fabsf(q0.z), fabsf(q0.w)));
                        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(q1.x), fabsf(q1.y))), fmaxf(fabsf(q1.z), fabsf(q1.w)));
                    }
                    const float tau = p.fix_tau * fmaxf(1.f, mx);
                    if (p.fix_count && mn < tau) {
                        unsigned long long mask = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            mask |= (unsigned long long)(fabsf(__uint_as_float(v0[j]) + bias[j]) < tau ? 1u : 0u) << j;
                            mask |= (unsigned long long)(fabsf(__uint_as_float(v1[j]) + bias[32 + j]) < tau ? 1u : 0u) << (32 + j);
                        }
                        while (mask) {
                            const int j = __ffsll((long long)mask) - 1;
                            mask &= mask - 1;
                            const int k = atomicAdd(p.fix_count + 2, 1);
                            if (k < p.fix_cap) p.fix_list[2][k] = ((unsigned long long)e << 8) | (unsigned)j;
                        }
                    }
                }
            }
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b0 = *reinterpret_cast<const float4*>(bias + j), b1 = *reinterpret_cast<const float4*>(bias + 32 + j);
                const float bb0[4] = {b0.x, b0.y, b0.z, b0.w}, bb1[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float y0 = gate_act<GENERIC>(p.act[2], __uint_as_float(v0[j + u]) + bb0[u], sl[2]);
                    const float y1 = gate_act<GENERIC>(p.act[2], __uint_as_float(v1[j + u]) + bb1[u], sl[2]);
                    ss = fmaf(y0, y0, ss); ss = fmaf(y1, y1, ss);
                    v0[j + u] = __float_as_uint(y0); v1[j + u] = __float_as_uint(y1);
                }
            }
            float inv = 1.f;
            if (p.normalize) inv = 1.f / sqrtf(ss);                   // ss == 0 -> inf -> 0*inf = NaN -> 0 below
            const long long e = gate_tile_of(p, i) * TILE_ROWS + row;
            if (e < p.n_edges) {
                float* o = p.s_out + e * GF;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 r;
                    r.x = __uint_as_float(v0[j]) * inv; r.y = __uint_as_float(v0[j + 1]) * inv;
                    r.z = __uint_as_float(v0[j + 2]) * inv; r.w = __uint_as_float(v0[j + 3]) * inv;
                    r.x = (r.x != r.x) ? 0.f : r.x; r.y = (r.y != r.y) ? 0.f : r.y;      // NaN -> 0 (gnn.py:426)
                    r.z = (r.z != r.z) ? 0.f : r.z; r.w = (r.w != r.w) ? 0.f : r.w;
                    stg4(o + j, r);
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 r;
                    r.x = __uint_as_float(v1[j]) * inv; r.y = __uint_as_float(v1[j + 1]) * inv;
                    r.z = __uint_as_float(v1[j + 2]) * inv; r.w = __uint_as_float(v1[j + 3]) * inv;
                    r.x = (r.x != r.x) ? 0.f : r.x; r.y = (r.y != r.y) ? 0.f : r.y;
                    r.z = (r.z != r.z) ? 0.f : r.z; r.w = (r.w != r.w) ? 0.f : r.w;
                    stg4(o + 32 + j, r);
                }
            }
        };

        const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4);
        long long* trp = p.trace + (warp == 0 ? 0 : 1) * 128;
#define SWE_STAMP(i_, ev_) do { if (tr && (i_) < 16) trp[(i_) * 8 + (ev_)] = clock64(); } while (0)
        if (MODE == 2) {
            // partial-table producer: D alternates between the two accumulators, so the tensor pipe works on tile
            // i+1 while tile i is written out.   G(t_0) ; for i: { G(t_{i+1}) ; OUT(t_i) }
            if (n_my > 0) { load_ids(0); gather(0, 0, n_l1); }
#pragma unroll 1
            for (int i = 0; i < n_my; ++i) {
                if (i + 1 < n_my) { load_ids(i + 1); gather(i + 1, 0, n_l1); }
                const int buf = i & 1;
                mbar_wait(&bar->d_full[buf], (uint32_t)(i >> 1) & 1);
                tc_fence_after_sync();
                const long long tile = gate_tile_of(p, i);
                const long long e = tile * TILE_ROWS + row;
                // per-node table: row-major [node][128]; per-edge table (src/dst given): the tile-transposed order
                // MODE 3 reads (see there), all rows of the (padded) last tile included
                const bool per_edge = p.src != nullptr;
                float* o = per_edge ? p.p_out + tile * (TILE_ROWS * GH) + hf * (TILE_ROWS * 64) + row * 4
                                    : p.p_out + ((long long)p.row_lo + e) * GH + hf * 64;
                const int jstride = per_edge ? TILE_ROWS : 1;             // in 16-B chunks
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
                    uint32_t v[32];
                    tmem_ld32(lane_addr + (buf ? COL_D_B : COL_D_A) + hf * 64 + cb * 32, v);
                    tmem_wait_ld();
                    if (per_edge || e < p.n_edges) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            stg4(o + (cb * 8 + j / 4) * jstride * 4, make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                                __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
                    }
                }
                tc_fence_before_sync();
                mbar_arrive(&bar->a_ready[buf]);                      // accumulator `buf` may be overwritten
            }
        } else {
        if (n_my > 0) { load_ids(0); gather(0, 0, n_l1); }
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const uint32_t ph = i & 1;
            const bool dump = p.dbg != nullptr && i == 0 && blockIdx.x == 0;
            const bool more = i + 1 < n_my;
            SWE_STAMP(i, 0);
            epilogue_mid(0, ph, dump, i);
            SWE_STAMP(i, 1);
            if (more) { load_ids(i + 1); gather(i + 1, 0, h_l1); }
            SWE_STAMP(i, 2);
            epilogue_mid(1, ph, dump, i);
            SWE_STAMP(i, 3);
            if (more) gather(i + 1, h_l1, n_l1);
            SWE_STAMP(i, 4);
            epilogue_final(i, ph, dump);
            SWE_STAMP(i, 5);
        }
        }
#undef SWE_STAMP
    } else if (warp == 8) {
        // =====================================================================================
        // weight loader (same chunk order as the MMA issuer)
        // =====================================================================================
        if (lane == 0) {
            uint32_t w_cnt = 0;
            bool n_my_done = false;
            auto load = [&](const unsigned char* srcp, uint32_t bytes) {
                const uint32_t slot = w_cnt % W_STAGES;
                mbar_wait(&bar->w_empty[slot], ((w_cnt / W_STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(&bar->w_full[slot], bytes);
                bulk_g2s(w_ring + (size_t)slot * SLOT_BYTES, srcp, bytes, &bar->w_full[slot]);
                ++w_cnt;
            };
            auto load_l1 = [&](int c_lo, int c_hi) {
                for (int c = c_lo; c < c_hi; ++c)
                    load(p.img + img_l1_off(2 * l1_chunk_segment(p, c) + (c & 1)), SLOT_BYTES);
            };
            if (n_my > 0) load_l1(0, n_l1);
            if (MODE == 2) {
                for (int i = 1; i < n_my; ++i) load_l1(0, n_l1);
                n_my_done = true;
            }
            for (int i = 0; i < n_my && !n_my_done; ++i) {
                const bool more = i + 1 < n_my;
                for (int c = 0; c < 4; ++c) load(p.img + img_l2_off(p.n_l1_img, c), SLOT_BYTES);
                if (more) load_l1(0, h_l1);
                for (int c = 0; c < 4; ++c) load(p.img + img_l3_off(p.n_l1_img, c), SLOT_BYTES / 2);
                if (more) load_l1(h_l1, n_l1);
            }
        }
    } else {
        // =====================================================================================
        // MMA issuer
        // =====================================================================================
        if (lane == 0) {
            const uint32_t idesc128 = make_idesc_tf32(128, 128), idesc64 = make_idesc_tf32(128, 64);
            const uint32_t a_ring_u32 = smem_u32(a_ring), w_ring_u32 = smem_u32(w_ring);
            uint32_t a_cnt = 0, w_cnt = 0;
            // layer 0 (SS): D_a (+)= A_chunk · W_chunkᵀ for chunks [c_lo, c_hi)
            uint32_t d_l1 = COL_D_A;                                   // accumulator of layer 0 (MODE 2 alternates)
            int dfull_l1 = 0;
            auto mma_l1 = [&](int c_lo, int c_hi, bool last) {
                for (int c = c_lo; c < c_hi; ++c, ++a_cnt, ++w_cnt) {
                    const uint32_t sa = a_cnt % A_STAGES, sw = w_cnt % W_STAGES;
                    mbar_wait(&bar->a_full[sa], (a_cnt / A_STAGES) & 1);
                    mbar_wait(&bar->w_full[sw], (w_cnt / W_STAGES) & 1);
                    tc_fence_after_sync();
                    const uint32_t a_hi = a_ring_u32 + sa * SLOT_BYTES, a_lo = a_hi + TILE16K;
                    const uint32_t w_hi = w_ring_u32 + sw * SLOT_BYTES, w_lo = w_hi + TILE16K;
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        const uint64_t dah = make_desc_sw128(a_hi + ks * 32), dal = make_desc_sw128(a_lo + ks * 32);
                        const uint64_t dwh = make_desc_sw128(w_hi + ks * 32), dwl = make_desc_sw128(w_lo + ks * 32);
                        mma_tf32_ss(tmem_base + d_l1, dal, dwh, idesc128, (c | ks) ? 1u : 0u);
                        mma_tf32_ss(tmem_base + d_l1, dah, dwl, idesc128, 1u);
                        mma_tf32_ss(tmem_base + d_l1, dah, dwh, idesc128, 1u);
                    }
                    mma_commit(&bar->a_empty[sa]);
                    mma_commit(&bar->w_empty[sw]);
                }
                if (last && n_l1 > 0) mma_commit(&bar->d_full[dfull_l1]);
            };
            // layers 1 / 2 (TS): A from TMEM, D_b (layer 1: 128 columns, layer 2: 64 columns)
            auto mma_ts = [&](int layer, uint32_t ph) {
                mbar_wait(&bar->a_ready[layer - 1], ph);
                tc_fence_after_sync();
                const uint32_t idesc = layer == 1 ? idesc128 : idesc64;
                const uint32_t lo_off = layer == 1 ? TILE16K : TILE16K / 2;           // lo tile follows hi tile
                for (int c = 0; c < 4; ++c, ++w_cnt) {
                    const uint32_t sw = w_cnt % W_STAGES;
                    mbar_wait(&bar->w_full[sw], (w_cnt / W_STAGES) & 1);
                    tc_fence_after_sync();
                    const uint32_t w_hi = w_ring_u32 + sw * SLOT_BYTES, w_lo = w_hi + lo_off;
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        const uint32_t kcol = c * KC + ks * 8;
                        const uint64_t dwh = make_desc_sw128(w_hi + ks * 32), dwl = make_desc_sw128(w_lo + ks * 32);
                        mma_tf32_ts(tmem_base + COL_D_B, tmem_base + COL_A_LO + kcol, dwh, idesc, (c | ks) ? 1u : 0u);
                        mma_tf32_ts(tmem_base + COL_D_B, tmem_base + COL_A_HI + kcol, dwl, idesc, 1u);
                        mma_tf32_ts(tmem_base + COL_D_B, tmem_base + COL_A_HI + kcol, dwh, idesc, 1u);
                    }
                    mma_commit(&bar->w_empty[sw]);
                }
                mma_commit(&bar->d_full[layer]);
            };
            const bool tr = p.trace != nullptr && blockIdx.x == 0;
            long long* trp = p.trace + 2 * 128;
#define SWE_STAMP(i_, ev_) do { if (tr && (i_) < 16) trp[(i_) * 8 + (ev_)] = clock64(); } while (0)
            if (MODE == 2) {
                for (int i = 0; i < n_my; ++i) {
                    const int buf = i & 1;
                    mbar_wait(&bar->a_ready[buf], (((uint32_t)(i >> 1)) & 1) ^ 1);     // tile i-2 has been written out
                    tc_fence_after_sync();
                    d_l1 = buf ? COL_D_B : COL_D_A;
                    dfull_l1 = buf;
                    mma_l1(0, n_l1, true);
                }
            } else {
            if (n_my > 0) mma_l1(0, n_l1, true);
            for (int i = 0; i < n_my; ++i) {
                const uint32_t ph = i & 1;
                const bool more = i + 1 < n_my;
                SWE_STAMP(i, 0);
                mma_ts(1, ph);
                SWE_STAMP(i, 1);
                if (more) mma_l1(0, h_l1, false);
                SWE_STAMP(i, 2);
                mma_ts(2, ph);
                SWE_STAMP(i, 3);
                if (more) mma_l1(h_l1, n_l1, true);
                SWE_STAMP(i, 4);
            }
            }
#undef SWE_STAMP
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace swe

using namespace swe;

extern "C" int swe_edge_gate_tc_fwd_traced(const float*, const float*, const float*, const float*, const int32_t*,
                                           const int32_t*, int64_t, const void*, int32_t, const int32_t*,
                                           const float* const*, int32_t, float*, float*, long long*, void*);

static long long* g_next_trace = nullptr;      // profiling aid: consumed by the next launch (tools/bench_gate.py)
extern "C" void swe_gate_tc_set_trace(long long* t) { g_next_trace = t; }

static int gate_tc_launch(const tc::GateTcParams& p_in, int mode, void* stream) {
    tc::GateTcParams p = p_in;
    if (g_next_trace && !p.trace) { p.trace = g_next_trace; g_next_trace = nullptr; }
    bool generic = false;
    if (mode != 2)
        for (int i = 0; i < 3; ++i)
            generic |= !(p.act[i] == SWE_ACT_NONE || p.act[i] == SWE_ACT_PRELU || p.act[i] == SWE_ACT_RELU || p.act[i] == SWE_ACT_LEAKYRELU);
    void (*kern)(const tc::GateTcParams) = nullptr;
    if (mode == 0 && p.pre_out[0]) kern = generic ? tc::edge_gate_tc_kernel<true, 0, true> : tc::edge_gate_tc_kernel<false, 0, true>;
    else if (mode == 0) kern = generic ? tc::edge_gate_tc_kernel<true, 0> : tc::edge_gate_tc_kernel<false, 0>;
    else if (mode == 1) kern = generic ? tc::edge_gate_tc_kernel<true, 1> : tc::edge_gate_tc_kernel<false, 1>;
    else if (mode == 3) kern = generic ? tc::edge_gate_tc_kernel<true, 3> : tc::edge_gate_tc_kernel<false, 3>;
    else kern = tc::edge_gate_tc_kernel<false, 2>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::GATE_TC_SMEM);
    if (e != cudaSuccess) { set_error("edge_gate_tc smem opt-in (%zu B): %s", tc::GATE_TC_SMEM, cudaGetErrorString(e)); return (int)e; }
    const long long n_tiles = (p.n_edges + tc::TILE_ROWS - 1) / tc::TILE_ROWS;
    kern<<<grid_for(n_tiles, 1), tc::N_THREADS, tc::GATE_TC_SMEM, (cudaStream_t)stream>>>(p);
    return 0;
}

extern "C" size_t swe_gate_tc_image_bytes(int32_t k1) { return tc::img_bytes(k1 / tc::KC); }

extern "C" int swe_gate_tc_pack(const float* w1, int32_t k1, const float* b1, const float* w2, const float* b2,
                                const float* w3, const float* b3, void* image, void* stream) {
    SWE_REQUIRE(w1 && w2 && w3 && image, SWE_E_INVAL, "gate_tc_pack: null pointer");
    SWE_REQUIRE(k1 == 4 * tc::GF || k1 == 5 * tc::GF, SWE_E_UNSUPP, "gate_tc_pack: k1=%d (expected 256 or 320)", k1);
    SWE_REQUIRE(aligned16(image), SWE_E_ALIGN, "gate_tc_pack: image unaligned");
    tc::gate_tc_pack_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(w1, k1, b1, w2, b2, w3, b3, (unsigned char*)image);
    return check_launch("gate_tc_pack");
}

extern "C" int swe_edge_gate_tc_fwd(const float* xs, const float* xd_src, const float* xd_dst, const float* a,
                                    const int32_t* src, const int32_t* dst, int64_t n_edges, const void* image,
                                    int32_t k1, const int32_t* act3, const float* const* slope3, int32_t normalize,
                                    float* s_out, float* dbg, void* stream) {
    return swe_edge_gate_tc_fwd_traced(xs, xd_src, xd_dst, a, src, dst, n_edges, image, k1, act3, slope3, normalize,
                                       s_out, dbg, nullptr, stream);
}

// same as swe_edge_gate_tc_fwd plus an optional device buffer of 3*16*8 int64 receiving clock64 stamps of
// the phase boundaries of CTA 0 (row workers warp 0 / warp 4, MMA issuer) — a profiling aid, not part of the ABI
extern "C" int swe_edge_gate_tc_fwd_traced(const float* xs, const float* xd_src, const float* xd_dst, const float* a,
                                           const int32_t* src, const int32_t* dst, int64_t n_edges, const void* image,
                                           int32_t k1, const int32_t* act3, const float* const* slope3,
                                           int32_t normalize, float* s_out, float* dbg, long long* trace, void* stream) {
    SWE_REQUIRE(xs && xd_src && src && dst && s_out && image && act3 && slope3 && n_edges >= 0, SWE_E_INVAL,
                "edge_gate_tc: bad arguments");
    SWE_REQUIRE(aligned16(xs) && aligned16(xd_src) && aligned16(s_out) && aligned16(image) && (!a || aligned16(a)) &&
                (!xd_dst || aligned16(xd_dst)), SWE_E_ALIGN, "edge_gate_tc: unaligned buffer");
    SWE_REQUIRE(k1 == (a ? 5 : 4) * tc::GF, SWE_E_UNSUPP, "edge_gate_tc: k1=%d does not match the inputs", k1);
    if (n_edges == 0) return 0;
    tc::GateTcParams p;
    memset(&p, 0, sizeof(p));
    p.xs = xs; p.xd_src = xd_src; p.xd_dst = xd_dst; p.a = a; p.src = src; p.dst = dst; p.n_edges = n_edges;
    p.img = (const unsigned char*)image; p.n_l1_img = k1 / tc::KC;
    for (int i = 0; i < 3; ++i) { p.act[i] = act3[i]; p.slope[i] = slope3[i]; }
    p.normalize = normalize; p.s_out = s_out; p.dbg = dbg; p.trace = trace;
    p.n_seg = 0;
    for (int sg = 0; sg < 5; ++sg)
        if (sg < 3 || (sg == 3 && xd_dst) || (sg == 4 && a)) p.segs[p.n_seg++] = sg;
    if (int r = gate_tc_launch(p, 0, stream)) return r;
    return check_launch("edge_gate_tc_fwd");
}

// list mode: the tiles tile_list[1 .. tile_list[0]] only (fallback of swe_edge_gate_tc16_fwd's range guard; the list lives
// in device memory, so the launch is a fixed node of a captured graph and costs a few microseconds when it is empty)
extern "C" int swe_edge_gate_tc_fwd_listed(const float* xs, const float* xd_src, const float* xd_dst, const float* a,
                                           const int32_t* src, const int32_t* dst, int64_t n_edges, const void* image,
                                           int32_t k1, const int32_t* act3, const float* const* slope3, int32_t normalize,
                                           float* s_out, const int32_t* tile_list, void* stream) {
    SWE_REQUIRE(xs && xd_src && src && dst && s_out && image && act3 && slope3 && tile_list && n_edges >= 0, SWE_E_INVAL,
                "edge_gate_tc_listed: bad arguments");
    SWE_REQUIRE(aligned16(xs) && aligned16(xd_src) && aligned16(s_out) && aligned16(image) && (!a || aligned16(a)) &&
                (!xd_dst || aligned16(xd_dst)), SWE_E_ALIGN, "edge_gate_tc_listed: unaligned buffer");
    SWE_REQUIRE(k1 == (a ? 5 : 4) * tc::GF, SWE_E_UNSUPP, "edge_gate_tc_listed: k1=%d does not match the inputs", k1);
    if (n_edges == 0) return 0;
    tc::GateTcParams p;
    memset(&p, 0, sizeof(p));
    p.xs = xs; p.xd_src = xd_src; p.xd_dst = xd_dst; p.a = a; p.src = src; p.dst = dst; p.n_edges = n_edges;
    p.img = (const unsigned char*)image; p.n_l1_img = k1 / tc::KC;
    for (int i = 0; i < 3; ++i) { p.act[i] = act3[i]; p.slope[i] = slope3[i]; }
    p.normalize = normalize; p.s_out = s_out; p.tile_list = tile_list;
    p.n_seg = 0;
    for (int sg = 0; sg < 5; ++sg)
        if (sg < 3 || (sg == 3 && xd_dst) || (sg == 4 && a)) p.segs[p.n_seg++] = sg;
    if (int r = gate_tc_launch(p, 0, stream)) return r;
    return check_launch("edge_gate_tc_fwd_listed");
}

// ---------------------------------------------------------------------------------------------
// static part of layer 0 hoisted out of the rollout (MODE 3, see the kernel comment)
// ---------------------------------------------------------------------------------------------
extern "C" int swe_gate_static_partials_tc(const float* xs, const float* a, const int32_t* src, const int32_t* dst,
                                           int64_t n_edges, const void* image, int32_t k1, float* p_out, void* stream) {
    SWE_REQUIRE((xs || a) && src && dst && image && p_out && n_edges >= 0, SWE_E_INVAL, "gate_static_partials_tc: bad arguments");
    SWE_REQUIRE((!xs || aligned16(xs)) && aligned16(image) && aligned16(p_out) && (!a || aligned16(a)), SWE_E_ALIGN,
                "gate_static_partials_tc: unaligned buffer");
    SWE_REQUIRE(k1 == 4 * tc::GF || k1 == 5 * tc::GF, SWE_E_UNSUPP, "gate_static_partials_tc: k1=%d", k1);
    SWE_REQUIRE(!a || k1 == 5 * tc::GF, SWE_E_UNSUPP, "gate_static_partials_tc: edge features need k1 = 320");
    if (n_edges == 0) return 0;
    tc::GateTcParams p;
    memset(&p, 0, sizeof(p));
    p.xs = xs; p.a = a; p.src = src; p.dst = dst; p.n_edges = n_edges; p.row_lo = 0; p.p_out = p_out;
    p.img = (const unsigned char*)image; p.n_l1_img = k1 / tc::KC;
    p.n_seg = 0;
    if (xs) { p.segs[p.n_seg++] = 0; p.segs[p.n_seg++] = 1; }
    if (a) p.segs[p.n_seg++] = 4;
    if (int r = gate_tc_launch(p, 2, stream)) return r;
    return check_launch("gate_static_partials_tc");
}

extern "C" int swe_edge_gate_tc_stat_fwd(const float* p_edge, const float* xs, const float* xd_src, const float* xd_dst,
                                         const int32_t* src, const int32_t* dst, int64_t n_edges, const void* image,
                                         int32_t k1, const int32_t* act3, const float* const* slope3, int32_t normalize,
                                         float* s_out, void* stream) {
    SWE_REQUIRE(p_edge && xd_src && src && dst && s_out && image && act3 && slope3 && n_edges >= 0, SWE_E_INVAL,
                "edge_gate_tc_stat: bad arguments");
    SWE_REQUIRE(aligned16(p_edge) && aligned16(xd_src) && aligned16(s_out) && aligned16(image) && (!xd_dst || aligned16(xd_dst)) &&
                (!xs || aligned16(xs)), SWE_E_ALIGN, "edge_gate_tc_stat: unaligned buffer");
    SWE_REQUIRE(k1 == 4 * tc::GF || k1 == 5 * tc::GF, SWE_E_UNSUPP, "edge_gate_tc_stat: k1=%d", k1);
    if (n_edges == 0) return 0;
    tc::GateTcParams p;
    memset(&p, 0, sizeof(p));
    p.xs = xs; p.xd_src = xd_src; p.xd_dst = xd_dst; p.src = src; p.dst = dst; p.n_edges = n_edges; p.p_src = p_edge;
    p.img = (const unsigned char*)image; p.n_l1_img = k1 / tc::KC;
    for (int i = 0; i < 3; ++i) { p.act[i] = act3[i]; p.slope[i] = slope3[i]; }
    p.normalize = normalize; p.s_out = s_out;
    p.n_seg = 0;
    if (xs) { p.segs[p.n_seg++] = 0; p.segs[p.n_seg++] = 1; }      // x_s is not static (with_WL): multiplied every step
    p.segs[p.n_seg++] = 2;
    if (xd_dst) p.segs[p.n_seg++] = 3;
    if (int r = gate_tc_launch(p, 3, stream)) return r;
    return check_launch("edge_gate_tc_stat_fwd");
}

// ---------------------------------------------------------------------------------------------
// exact-fp32 re-evaluation of the pre-activations flagged by the TRAIN forward (one warp per entry)
// ---------------------------------------------------------------------------------------------
namespace swe { namespace tc {
struct FixParams {
    const float* xs; const float* xd_src; const float* xd_dst; const float* a;
    const int32_t* src; const int32_t* dst;
    const float* w[3]; const float* b[3]; int k1;
    int act[3]; const float* slope[3];
    float* pre[3];
    const unsigned long long* list; const int* count; int cap;
};
template <int LAYER>
__global__ void __launch_bounds__(256) gate_fix_kernel(const __grid_constant__ FixParams p) {
    const int lane = threadIdx.x & 31;
    const int n = min(*p.count, p.cap);
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < n; it += warps) {
        const unsigned long long ent = p.list[it];
        const long long e = (long long)(ent >> 8);
        const int col = (int)(ent & 255u);
        float acc = 0.f;
        if (LAYER == 0) {
            const float* wr = p.w[0] + (long long)col * p.k1;
            const long long r = __ldg(p.src + e), c = __ldg(p.dst + e);
#pragma unroll
            for (int sg = 0; sg < 5; ++sg) {
                const float* z = sg == 0 ? p.xs + r * GF : sg == 1 ? p.xs + c * GF : sg == 2 ? p.xd_src + r * GF
                               : sg == 3 ? (p.xd_dst ? p.xd_dst + c * GF : nullptr) : (p.a ? p.a + e * GF : nullptr);
                if (z) {
                    const float2 zv = *reinterpret_cast<const float2*>(z + 2 * lane);
                    const float2 wv = *reinterpret_cast<const float2*>(wr + sg * GF + 2 * lane);
                    acc = fmaf(zv.x, wv.x, acc); acc = fmaf(zv.y, wv.y, acc);
                }
            }
        } else {
            const float sl = (p.act[LAYER - 1] == SWE_ACT_PRELU && p.slope[LAYER - 1]) ? __ldg(p.slope[LAYER - 1]) : 0.f;
            const float4 h = *reinterpret_cast<const float4*>(p.pre[LAYER - 1] + e * GH + 4 * lane);
            const float4 wv = *reinterpret_cast<const float4*>(p.w[LAYER] + (long long)col * GH + 4 * lane);
            acc = fmaf(act_apply(p.act[LAYER - 1], h.x, sl), wv.x, acc); acc = fmaf(act_apply(p.act[LAYER - 1], h.y, sl), wv.y, acc);
            acc = fmaf(act_apply(p.act[LAYER - 1], h.z, sl), wv.z, acc); acc = fmaf(act_apply(p.act[LAYER - 1], h.w, sl), wv.w, acc);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) p.pre[LAYER][e * (LAYER == 2 ? GF : GH) + col] = acc + (p.b[LAYER] ? __ldg(p.b[LAYER] + col) : 0.f);
    }
}
}}  // namespace swe::tc

// The tensor-core forward evaluates a pre-activation to ~3e-6 (3xTF32); an entry that close to 0 may come out with the
// other sign, and the backward's derivative mask (PReLU / ReLU: v > 0) would then differ from the fp32 reference's in
// an O(1) way for that summand.  swe_edge_gate_tc_train_fwd therefore lists every |pre| < tau * max(1, row max) and
// this call recomputes exactly those (typically 1e-4 of all entries) with fp32 FMAs, layer by layer (a layer-l entry is
// recomputed from the already repaired pre-activations of layer l-1, so the result does not depend on the order in
// which the lists were filled: bit-reproducible).
extern "C" int swe_gate_fix_preacts(const float* xs, const float* xd_src, const float* xd_dst, const float* a,
                                    const int32_t* src, const int32_t* dst, const float* w1, const float* b1,
                                    const float* w2, const float* b2, const float* w3, const float* b3, int32_t k1,
                                    const int32_t* act3, const float* const* slope3, float* pre1, float* pre2, float* pre3,
                                    const unsigned long long* fix_lists, const int32_t* fix_count, int32_t fix_cap,
                                    void* stream) {
    SWE_REQUIRE(xs && xd_src && src && dst && w1 && w2 && w3 && act3 && slope3 && pre1 && pre2 && pre3 && fix_lists &&
                fix_count && fix_cap > 0, SWE_E_INVAL, "gate_fix_preacts: bad arguments");
    SWE_REQUIRE(k1 == (a ? 5 : 4) * tc::GF, SWE_E_UNSUPP, "gate_fix_preacts: k1=%d does not match the inputs", k1);
    tc::FixParams p;
    memset(&p, 0, sizeof(p));
    p.xs = xs; p.xd_src = xd_src; p.xd_dst = xd_dst; p.a = a; p.src = src; p.dst = dst;
    p.w[0] = w1; p.w[1] = w2; p.w[2] = w3; p.b[0] = b1; p.b[1] = b2; p.b[2] = b3; p.k1 = k1;
    for (int i = 0; i < 3; ++i) { p.act[i] = act3[i]; p.slope[i] = slope3[i]; }
    p.pre[0] = pre1; p.pre[1] = pre2; p.pre[2] = pre3; p.cap = fix_cap;
    const int grid = 2 * NUM_SMS;
    p.list = fix_lists; p.count = fix_count;
    tc::gate_fix_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    p.list = fix_lists + fix_cap; p.count = fix_count + 1;
    tc::gate_fix_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    p.list = fix_lists + 2 * (size_t)fix_cap; p.count = fix_count + 2;
    tc::gate_fix_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("gate_fix_preacts");
}

// forward of the training step: s_ij plus the three pre-activations the backward differentiates through
extern "C" int swe_edge_gate_tc_train_fwd(const float* xs, const float* xd_src, const float* xd_dst, const float* a,
                                          const int32_t* src, const int32_t* dst, int64_t n_edges, const void* image,
                                          int32_t k1, const int32_t* act3, const float* const* slope3, int32_t normalize,
                                          float* pre1, float* pre2, float* pre3, float* s_out,
                                          unsigned long long* fix_lists, int32_t* fix_count, int32_t fix_cap, float fix_tau,
                                          void* stream) {
    SWE_REQUIRE(xs && xd_src && src && dst && s_out && image && act3 && slope3 && pre1 && pre2 && pre3 && n_edges >= 0,
                SWE_E_INVAL, "edge_gate_tc_train: bad arguments");
    SWE_REQUIRE(!fix_count || (fix_lists && fix_cap > 0 && fix_tau >= 0.f), SWE_E_INVAL, "edge_gate_tc_train: work list");
    SWE_REQUIRE(aligned16(xs) && aligned16(xd_src) && aligned16(s_out) && aligned16(image) && (!a || aligned16(a)) &&
                (!xd_dst || aligned16(xd_dst)) && aligned16(pre1) && aligned16(pre2) && aligned16(pre3), SWE_E_ALIGN,
                "edge_gate_tc_train: unaligned buffer");
    SWE_REQUIRE(k1 == (a ? 5 : 4) * tc::GF, SWE_E_UNSUPP, "edge_gate_tc_train: k1=%d does not match the inputs", k1);
    if (n_edges == 0) return 0;
    tc::GateTcParams p;
    memset(&p, 0, sizeof(p));
    p.xs = xs; p.xd_src = xd_src; p.xd_dst = xd_dst; p.a = a; p.src = src; p.dst = dst; p.n_edges = n_edges;
    p.img = (const unsigned char*)image; p.n_l1_img = k1 / tc::KC;
    for (int i = 0; i < 3; ++i) { p.act[i] = act3[i]; p.slope[i] = slope3[i]; }
    p.normalize = normalize; p.s_out = s_out;
    p.pre_out[0] = pre1; p.pre_out[1] = pre2; p.pre_out[2] = pre3;
    if (fix_count) {
        for (int i = 0; i < 3; ++i) p.fix_list[i] = fix_lists + (size_t)i * fix_cap;
        p.fix_count = fix_count; p.fix_cap = fix_cap; p.fix_tau = fix_tau;
    }
    p.n_seg = 0;
    for (int sg = 0; sg < 5; ++sg)
        if (sg < 3 || (sg == 3 && xd_dst) || (sg == 4 && a)) p.segs[p.n_seg++] = sg;
    if (int r = gate_tc_launch(p, 0, stream)) return r;
    return check_launch("edge_gate_tc_train_fwd");
}

// ---------------------------------------------------------------------------------------------
// decomposed layer 0 (see the kernel comment): per-node partial tables + gate reading them
// ---------------------------------------------------------------------------------------------
extern "C" int swe_gate_partials_tc(const float* xs, const float* xd, int32_t row_lo, int32_t n_rows, const void* image,
                                    int32_t k1, int32_t role, float* p_out, void* stream) {
    SWE_REQUIRE(xs && image && p_out && row_lo >= 0 && n_rows >= 0 && (role == 0 || role == 1), SWE_E_INVAL,
                "gate_partials_tc: bad arguments");
    SWE_REQUIRE(role == 1 || xd, SWE_E_INVAL, "gate_partials_tc: the source role needs x_d");
    SWE_REQUIRE(aligned16(xs) && aligned16(image) && aligned16(p_out) && (!xd || aligned16(xd)), SWE_E_ALIGN,
                "gate_partials_tc: unaligned buffer");
    SWE_REQUIRE(k1 == 4 * tc::GF || k1 == 5 * tc::GF, SWE_E_UNSUPP, "gate_partials_tc: k1=%d", k1);
    if (n_rows == 0) return 0;
    tc::GateTcParams p;
    memset(&p, 0, sizeof(p));
    p.xs = xs; p.xd_src = xd; p.xd_dst = xd; p.n_edges = n_rows; p.row_lo = row_lo; p.p_out = p_out;
    p.img = (const unsigned char*)image; p.n_l1_img = k1 / tc::KC;
    p.n_seg = 0;
    p.segs[p.n_seg++] = role;                         // x_s block of this role (0: x_s[r], 1: x_s[c])
    if (xd) p.segs[p.n_seg++] = 2 + role;             // x_d block (2: x_d[r], 3: x_d[c])
    if (int r = gate_tc_launch(p, 2, stream)) return r;
    return check_launch("gate_partials_tc");
}

extern "C" int swe_edge_gate_tc_dec_fwd(const float* p_src, const float* p_dst, const float* a, const int32_t* src,
                                        const int32_t* dst, int64_t n_edges, const void* image, int32_t k1,
                                        const int32_t* act3, const float* const* slope3, int32_t normalize, float* s_out,
                                        void* stream) {
    SWE_REQUIRE(p_src && p_dst && src && dst && s_out && image && act3 && slope3 && n_edges >= 0, SWE_E_INVAL,
                "edge_gate_tc_dec: bad arguments");
    SWE_REQUIRE(aligned16(p_src) && aligned16(p_dst) && aligned16(s_out) && aligned16(image) && (!a || aligned16(a)),
                SWE_E_ALIGN, "edge_gate_tc_dec: unaligned buffer");
    SWE_REQUIRE(k1 == (a ? 5 : 4) * tc::GF, SWE_E_UNSUPP, "edge_gate_tc_dec: k1=%d does not match the inputs", k1);
    if (n_edges == 0) return 0;
    tc::GateTcParams p;
    memset(&p, 0, sizeof(p));
    p.a = a; p.src = src; p.dst = dst; p.n_edges = n_edges; p.p_src = p_src; p.p_dst = p_dst;
    p.img = (const unsigned char*)image; p.n_l1_img = k1 / tc::KC;
    for (int i = 0; i < 3; ++i) { p.act[i] = act3[i]; p.slope[i] = slope3[i]; }
    p.normalize = normalize; p.s_out = s_out;
    p.n_seg = 0;
    if (a) p.segs[p.n_seg++] = 4;
    if (int r = gate_tc_launch(p, 1, stream)) return r;
    return check_launch("edge_gate_tc_dec_fwd");
}
