// Edge gate on tcgen05, second edition: fp16 hi/lo splits on kind::f16 (K = 16 per instruction) instead of TF32
// hi/lo splits on kind::tf32 (K = 8).  Same arithmetic contract as swe_gate_tc.cu (models/gnn.py:414-426, F = 64,
// 3-layer edge MLP 5F|4F -> 2F -> 2F -> F, fp32 accumulation in TMEM, result within rel 1e-5 per layer of fp32),
// half the tensor-core instructions: 108 instead of 216 per 128-edge tile — the instruction count is what bounds the
// TF32 kernel (tools/microbench/mma_rate3.cu: one tcgen05.mma with M = 128, N = 128 retires per 64 cycles whatever its
// kind, operand source or swizzle, so K = 16 per instruction does twice the work of K = 8).
//
// Precision.  x' = 2^e x is split as hi = rn16(x'), lo = rn16(x' - hi); |x' - hi - lo| <= max(2^-24 |x'|, 2^-25).
// A·W ≈ A_lo·W_hi + A_hi·W_lo + A_hi·W_hi (dropped A_lo·W_lo <= 2^-22 |A||W|): the same three products as 3xTF32.
// fp16 has 5 exponent bits, so every operand is scaled by a power of two (exact) into the window where both halves are
// normal numbers:
//   * weights: one exponent per matrix (max |w'| in [2^13, 2^14)), chosen when the image is packed;
//   * hidden activations (layers 1, 2): one exponent per ROW (max |y'| in [2^13, 2^14)), formed in the epilogue that
//     produces the row (the thread owns it) and undone in the next epilogue — no range restriction at all;
//   * layer-0 inputs (five gathered segments, converted by different threads before the row's maximum is known):
//     unscaled, i.e. an absolute resolution of 2^-25 (<= 2^-20 of the row's maximum inside the guarded window) and
//     a ceiling of 65504.  The kernel tracks max |x| per
//     row while it converts; a tile with a row outside [2^-5, 2^15] (never seen with O(1) encoder outputs, but
//     nothing forbids it) is appended to a work list and redone by the TF32 kernel (swe_gate_tc.cu in list mode),
//     so the result never depends on the window.
//
// One CTA, 20 warps, persistent over 128-edge tiles, TWO tiles in flight (one per group, independent pipelines):
//   warps 0-7 / 8-15 : row-worker group 0 / 1 (tiles k = g, g + 2, ...): gather + convert the layer-0 input into the
//                      group's shared-memory ring (UMMA K-major, 64-byte swizzle, 32 fp16 per row), then the three
//                      epilogues of the tile (thread = TMEM lane = edge, half of the columns): tcgen05.ld ->
//                      descale, bias, activation -> row scale -> fp16 hi/lo -> tcgen05.st (next layer's A operand
//                      stays in TMEM) / L2-normalise + store s_ij
//   warps 16, 17     : MMA issuer of group 0 / 1 (one thread each): layer 0 chunk by chunk as the gather publishes them
//                      (SS), layers 1 and 2 (TS) as the epilogues publish their operand.  Two issuers because the
//                      tensor pipe's instruction queue is shallow: a single thread's barrier waits and commits
//                      (~400 cycles per 6-instruction chunk, on a scheduler it shares with four busy worker warps)
//                      left the pipe idle half of the time; with one issuer per group the other group's instructions
//                      fill those gaps, and every epilogue of one group runs under MMAs of the other
//   warps 18, 19     : weight loader of group 0 / 1: W2 resident in shared memory (64 KB, loaded once); per tile the
//                      W1 chunks and W3 stream through the group's 2-slot ring (cp.async.bulk + mbarrier tx)
// TMEM (512 columns): group g owns [256 g, 256 g + 256): D (128 fp32 columns, reused by the three layers),
// A_hi (64 columns = 128 fp16), A_lo (64 columns).
#include <stdlib.h>
#include "swe_tc.cuh"

namespace swe {
namespace tc16 {
using namespace swe::tc;

constexpr int GF = 64, GH = 128;
constexpr int KC = 32;                          // K elements per chunk = one 64-byte swizzled row
constexpr int TILE = 128;
constexpr int A_TILE = TILE * 64;               // [128 x 32] fp16 = 8 KB
constexpr int A_SLOT = 2 * A_TILE;              // hi | lo
constexpr int W1_CHUNK = 2 * GH * 64;           // [128 x 32] fp16 hi | lo = 16 KB
constexpr int W3_CHUNK = 2 * GF * 64;           // [64 x 32]  fp16 hi | lo = 8 KB
constexpr int W_RES_BYTES = 4 * W1_CHUNK;       // W2 resident = 64 KB (W3 rides the ring: 2 slots of 2 K-chunks per tile)
constexpr int A_STAGES = 2, W_STAGES = 2;           // per group
constexpr int GROUP_THREADS = 256;
constexpr int N_THREADS = 2 * GROUP_THREADS + 128;            // + MMA issuers (warps 16, 17) + weight loaders (18, 19)
constexpr float L0_SCALE = 1.f;                 // layer-0 inputs go in unscaled (a multiply per element less in the gather)
constexpr int N_IMG_FLOATS = 328;               // bias[320] | descale[3] | pad | max |bias| of layers 0, 1 | pad

struct __align__(8) Bar {
    uint64_t a_full[2][A_STAGES], a_empty[2][A_STAGES];
    uint64_t w_full[2][W_STAGES], w_empty[2][W_STAGES];
    uint64_t w_res;                             // resident W2 has landed
    uint64_t d_full[2];                         // group g: accumulator complete (commit; 3 per tile)
    uint64_t a_ready[2];                        // group g: next layer's A operand is in TMEM (256 arrivals; 2 per tile)
    uint64_t d_free[2];                         // group g: the final epilogue has read D (256 arrivals; 1 per tile)
};

constexpr size_t SMEM_BYTES = 1024 + (size_t)W_RES_BYTES + 2 * (size_t)W_STAGES * W1_CHUNK + 2 * (size_t)A_STAGES * A_SLOT +
                              sizeof(float) * N_IMG_FLOATS + sizeof(int32_t) * 512 + sizeof(float) * 512 +
                              sizeof(uint32_t) * 256 + sizeof(int) * 4 + sizeof(Bar) + 16;
static_assert((size_t)A_STAGES * A_SLOT >= (size_t)TILE * GF * 4, "the output stage lives in the group's A ring");

// image layout (bytes): W1 chunks | W2 chunks | W3 chunks | floats
__host__ __device__ constexpr size_t img_w1_off(int chunk) { return (size_t)chunk * W1_CHUNK; }
__host__ __device__ constexpr size_t img_w2_off(int n_l1) { return (size_t)n_l1 * W1_CHUNK; }
__host__ __device__ constexpr size_t img_w3_off(int n_l1) { return (size_t)(n_l1 + 4) * W1_CHUNK; }
__host__ __device__ constexpr size_t img_float_off(int n_l1) { return (size_t)(n_l1 + 4) * W1_CHUNK + 4 * (size_t)W3_CHUNK; }
__host__ __device__ constexpr size_t img_bytes(int n_l1) { return img_float_off(n_l1) + N_IMG_FLOATS * sizeof(float); }

// ---------------------------------------------------------------------------------------------
// weight packing: Linear weights [n_out, k_in] (row-major = K-major) -> swizzled fp16 hi|lo chunk images;
// exp3[l] = power-of-two exponent applied to layer l's weights (chosen by the host from max |w|)
// ---------------------------------------------------------------------------------------------
__global__ void gate_tc16_pack_kernel(const float* __restrict__ w1, int k1, const float* __restrict__ b1,
                                      const float* __restrict__ w2, const float* __restrict__ b2,
                                      const float* __restrict__ w3, const float* __restrict__ b3,
                                      int e1, int e2, int e3, unsigned char* __restrict__ img) {
    const int n_l1 = k1 / KC;
    const int total = GH * k1 + GH * GH + GF * GH;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const float* w; int n, k, kin, rows, ex; size_t base, chunk_bytes;
        if (idx < GH * k1) { w = w1; kin = k1; n = idx / k1; k = idx % k1; rows = GH; ex = e1; base = 0; chunk_bytes = W1_CHUNK; }
        else if (idx < GH * k1 + GH * GH) { int j = idx - GH * k1; w = w2; kin = GH; n = j / GH; k = j % GH; rows = GH; ex = e2; base = img_w2_off(n_l1); chunk_bytes = W1_CHUNK; }
        else { int j = idx - GH * k1 - GH * GH; w = w3; kin = GH; n = j / GH; k = j % GH; rows = GF; ex = e3; base = img_w3_off(n_l1); chunk_bytes = W3_CHUNK; }
        const float v = ldexpf(w[(size_t)n * kin + k], ex);
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        const int kk = k % KC;
        const size_t off = base + (size_t)(k / KC) * chunk_bytes + sw64_piece_offset(n, kk >> 3) + (kk & 7) * 2;
        *reinterpret_cast<__half*>(img + off) = h;
        *reinterpret_cast<__half*>(img + off + (size_t)rows * 64) = l;
    }
    float* f = reinterpret_cast<float*>(img + img_float_off(n_l1));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N_IMG_FLOATS; i += gridDim.x * blockDim.x) {
        float v = 0.f;
        if (i < 128) v = b1 ? b1[i] : 0.f;
        else if (i < 256) v = b2 ? b2[i - 128] : 0.f;
        else if (i < 320) v = b3 ? b3[i - 256] : 0.f;
        else if (i == 320) v = ldexpf(1.f / L0_SCALE, -e1);          // D0 -> pre-activation of layer 0
        else if (i == 321) v = ldexpf(1.f, -e2);                     // x the row's own 2^-e
        else if (i == 322) v = ldexpf(1.f, -e3);
        else if (i == 324 || i == 325) {                             // max |bias| of layers 0 / 1 (the epilogues' row-scale bound)
            const float* b = i == 324 ? b1 : b2;
            for (int j = 0; b && j < 128; ++j) v = fmaxf(v, fabsf(b[j]));
        }
        f[i] = v;
    }
}

struct Gate16Params {
    const float* xs; const float* xd_src; const float* xd_dst; const float* a;
    const int32_t* src; const int32_t* dst;
    long long n_edges;
    const unsigned char* img; int n_l1_img;
    int act[3]; const float* slope[3];
    int normalize;
    float* s_out;
    float* dbg;                                   // optional [128*128 + 128*128 + 128*64] pre-activations (no bias) of tile 0
    int n_seg; int segs[5];                       // layer-0 input segments in order (0 x_s[r], 1 x_s[c], 2 x_d[r], 3 x_d[c], 4 a_e)
    int* flag_ws;                                 // [0] = number of flagged tiles, [1 + i] = tile ids (capacity: all tiles)
    long long* trace;                             // optional [3 roles][16 tiles][8 events] clock64 stamps of CTA 0 (profiling aid)
    unsigned stagger_ns;                          // start offset of group 1 (see the kernel)
};

__device__ __noinline__ float act_generic16(int act, float v, float slope) { return act_apply(act, v, slope); }
struct ActSel { int act; float slope; bool leaky; };
__device__ __forceinline__ ActSel act_select(int act, const float* slope_p) {
    ActSel a;
    a.act = act;
    a.leaky = (act == SWE_ACT_NONE || act == SWE_ACT_PRELU || act == SWE_ACT_RELU || act == SWE_ACT_LEAKYRELU);
    a.slope = act == SWE_ACT_NONE ? 1.f : act == SWE_ACT_RELU ? 0.f : act == SWE_ACT_LEAKYRELU ? 0.1f
              : (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
    return a;
}
template <bool GENERIC>
__device__ __forceinline__ float act_do(const ActSel& a, float v) {
    if (GENERIC) { if (!a.leaky) return act_generic16(a.act, v, a.slope); }
    // v > 0 ? v : slope v  as one multiply and one min/max (slope <= 1: the larger of v and slope v, else the smaller);
    // the selection is loop-invariant and becomes FMNMX's predicate operand
    const float t = a.slope * v;
    return a.slope <= 1.f ? fmaxf(v, t) : fminf(v, t);
}

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); }

template <bool GENERIC>
__global__ void __launch_bounds__(N_THREADS, 1) edge_gate_tc16_kernel(const __grid_constant__ Gate16Params p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* w_res = smem;                                           // W2 (4 x 16 KB)
    unsigned char* w_ring = w_res + W_RES_BYTES;
    unsigned char* a_ring = w_ring + 2 * (size_t)W_STAGES * W1_CHUNK;      // [group][slot]   (w_ring: [group][slot] too)
    float* s_f = reinterpret_cast<float*>(a_ring + 2 * (size_t)A_STAGES * A_SLOT);   // bias[320] | descale[3]
    int32_t* s_ids = reinterpret_cast<int32_t*>(s_f + N_IMG_FLOATS);       // [group][src 128 | dst 128]
    float* s_xch = reinterpret_cast<float*>(s_ids + 512);                  // [group][half][128]
    uint32_t* s_rowmax = reinterpret_cast<uint32_t*>(s_xch + 512);         // [group][128] max |x| of the layer-0 input row (bits)
    int* s_flag = reinterpret_cast<int*>(s_rowmax + 256);                  // [group]
    Bar* bar = reinterpret_cast<Bar*>(s_flag + 4);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_l1 = 2 * p.n_seg;
    const long long n_tiles_all = (p.n_edges + TILE - 1) / TILE;
    const int n_my = (int)((n_tiles_all - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int g = 0; g < 2; ++g) {
            for (int i = 0; i < A_STAGES; ++i) { mbar_init(&bar->a_full[g][i], GROUP_THREADS); mbar_init(&bar->a_empty[g][i], 1); }
            mbar_init(&bar->d_full[g], 1); mbar_init(&bar->a_ready[g], GROUP_THREADS); mbar_init(&bar->d_free[g], GROUP_THREADS);
        }
        for (int g = 0; g < 2; ++g)
            for (int i = 0; i < W_STAGES; ++i) { mbar_init(&bar->w_full[g][i], 1); mbar_init(&bar->w_empty[g][i], 1); }
        mbar_init(&bar->w_res, 1);
        fence_barrier_init();
    }
    {
        const float* gf = reinterpret_cast<const float*>(p.img + img_float_off(p.n_l1_img));
        for (int i = threadIdx.x; i < N_IMG_FLOATS; i += N_THREADS) s_f[i] = gf[i];
        for (int i = threadIdx.x; i < 256; i += N_THREADS) s_rowmax[i] = 0u;
        if (threadIdx.x < 4) s_flag[threadIdx.x] = 0;
    }
    if (warp == 16) tmem_alloc(tmem_holder, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_holder;

    if (warp < 16) {
        // =====================================================================================
        // row-worker groups
        // =====================================================================================
        const int g = warp >> 3, gw = warp & 7, q = gw & 3, hf = gw >> 2;
        const int tg = threadIdx.x & 255;
        const int row = q * 32 + lane;                                   // TMEM lane / edge within the tile
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 256u;
        const uint32_t COL_D = 0, COL_AHI = 128, COL_ALO = 192;
        const ActSel a0 = act_select(p.act[0], p.slope[0]), a1 = act_select(p.act[1], p.slope[1]), a2 = act_select(p.act[2], p.slope[2]);
        int32_t* ids = s_ids + g * 256;
        float* xch = s_xch + g * 256;
        uint32_t* rowmax = s_rowmax + g * 128;
        unsigned char* my_ring = a_ring + (size_t)g * A_STAGES * A_SLOT;
        float* stage = reinterpret_cast<float*>(my_ring);               // [128][64] output rows, 16-byte pieces XOR-swizzled by row
        // gather: 8 lanes x 16 B = the 128 contiguous bytes of a row's 32-float chunk (4 rows = 4 full lines per
        // instruction); this thread: floats [4 l8, 4 l8 + 4) of rows rr + 32 u
        const int l8 = tg & 7, rr = tg >> 3;
        const uint32_t g_off = sw64_piece_offset(rr, l8 >> 1) + (l8 & 1) * 8;     // row rr + 32 u: + 2048 u
        uint32_t a_cnt = 0;
        const int n_g = (n_my - g + 1) / 2;
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && tg == 0;
#define G_STAMP(j_, ev_) do { if (tr && (j_) < 16) p.trace[g * 128 + (j_) * 8 + (ev_)] = clock64(); } while (0)

        // the two groups start half a tile period apart, so that one gathers (LSU, conversion ALU, layer-0 MMAs) while the
        // other runs its epilogues (TMEM, ALU) instead of both doing the same thing at the same time
        if (g == 1 && n_g > 0 && p.stagger_ns > 0) __nanosleep(p.stagger_ns);
        // endpoint (src for threads 0-127, dst for 128-255) of this thread's edge of tile j_: fetched one tile ahead, while the
        // epilogue chain of the tile before runs (it used to cost ~1.5 k cycles at the head of every tile)
        auto endpoint_of = [&](int j_) -> int32_t {
            if (j_ >= n_g) return 0;
            long long e = ((long long)blockIdx.x + (long long)(g + 2 * j_) * gridDim.x) * TILE + (tg & 127);
            if (e >= p.n_edges) e = p.n_edges - 1;
            return tg < 128 ? __ldg(p.src + e) : __ldg(p.dst + e);
        };
        int32_t next_id = endpoint_of(0);
#pragma unroll 1
        for (int j = 0; j < n_g; ++j) {
            const long long tile = (long long)blockIdx.x + (long long)(g + 2 * j) * gridDim.x;
            const long long e0 = tile * TILE;
            // ---------------------------------------------------------------- endpoints of the tile
            {
                ids[tg] = next_id;
                group_sync(g);
            }
            G_STAMP(j, 0);
            // ---------------------------------------------------------------- gather + convert layer-0 input chunks
            {
                float m[4] = {0.f, 0.f, 0.f, 0.f};
                // this thread's four rows: endpoint ids and edge numbers once per tile
                int32_t id_s[4], id_d[4], ea[4];                        // ea: edge number relative to the tile (clamped)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    id_s[u] = ids[rr + 32 * u]; id_d[u] = ids[128 + rr + 32 * u];
                    const long long e = e0 + rr + 32 * u;
                    ea[u] = (int32_t)((e < p.n_edges ? e : p.n_edges - 1) - e0);
                }
                auto issue = [&](int c, float4 (&v)[4]) {
                    const int sg = p.segs[c >> 1];
                    const int koff = (c & 1) * KC + l8 * 4;
                    if (sg == 4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = ldg4_stream(p.a + (e0 + ea[u]) * GF + koff);
                    } else {
                        const float* base = (sg < 2 ? p.xs : (sg == 2 ? p.xd_src : p.xd_dst)) + koff;
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = ldg4(base + (long long)((sg & 1) ? id_d[u] : id_s[u]) * GF);
                    }
                };
                auto convert = [&](const float4 (&v)[4]) {              // chunk a_cnt: split, store, publish
                    const uint32_t slot = a_cnt % A_STAGES;
                    mbar_wait(&bar->a_empty[g][slot], ((a_cnt / A_STAGES) & 1) ^ 1);
                    unsigned char* hi_t = my_ring + (size_t)slot * A_SLOT + g_off;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 x = v[u];
                        m[u] = fmaxf(m[u], fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
                        uint2 hh, ll;
                        split_f16x2(x.x, x.y, hh.x, ll.x);
                        split_f16x2(x.z, x.w, hh.y, ll.y);
                        *reinterpret_cast<uint2*>(hi_t + u * 2048) = hh;
                        *reinterpret_cast<uint2*>(hi_t + A_TILE + u * 2048) = ll;
                    }
                    fence_proxy_async_smem();
                    mbar_arrive(&bar->a_full[g][slot]);
                    ++a_cnt;
                };
                // Loads run two chunks ahead of the conversion (three register sets rotating BY NAME: a register copy
                // at the end of an iteration would wait for the loads it copies), the ring three chunks ahead of the
                // tensor core.
                float4 b0[4], b1[4], b2[4];
                issue(0, b0);
                if (n_l1 > 1) issue(1, b1);
#pragma unroll 1
                for (int c = 0; c < n_l1; c += 3) {
                    if (c + 2 < n_l1) issue(c + 2, b2);
                    convert(b0);
                    if (c + 1 < n_l1) {
                        if (c + 3 < n_l1) issue(c + 3, b0);
                        convert(b1);
                    }
                    if (c + 2 < n_l1) {
                        if (c + 4 < n_l1) issue(c + 4, b1);
                        convert(b2);
                    }
                }
                // row maxima of the layer-0 input (range guard); NaN / inf inputs compare as huge and flag the tile
#pragma unroll
                for (int u = 0; u < 4; ++u) atomicMax(rowmax + rr + 32 * u, __float_as_uint(m[u]) & 0x7fffffffu);
            }
            next_id = endpoint_of(j + 1);                                // consumed at the head of the next tile
            G_STAMP(j, 1);
            const bool dump = p.dbg != nullptr && blockIdx.x == 0 && g == 0 && j == 0;
            float inv_scale = 1.f;                                       // 2^-e of the row's current A operand
            // ---------------------------------------------------------------- epilogues of layers 0 and 1
#pragma unroll 1
            for (int layer = 0; layer < 2; ++layer) {
                const ActSel& al = layer == 0 ? a0 : a1;
                const float* bias = s_f + layer * 128 + hf * 64;
                const float dsc = layer == 0 ? s_f[320] : s_f[321] * inv_scale;
                mbar_wait(&bar->d_full[g], (uint32_t)(3 * j + layer) & 1);
                tc_fence_after_sync();
                G_STAMP(j, 2 + 2 * layer);
                // pass 1: an UPPER BOUND of the row's largest |activation| (this thread's 64 columns), nothing kept in
                // registers.  Leaky family: |act(v)| <= |v| <= dsc max|D| + max|bias| — one FMNMX per element; the bound only
                // has to keep the scaled row inside fp16 (it is at most the row maximum plus the largest bias, i.e. it costs a
                // bit or two of the 14 bits of headroom above fp16's normal range, never accuracy of the leading part).
                float m = 0.f;
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
                    uint32_t v[32];
                    tmem_ld32(lane_addr + COL_D + hf * 64 + cb * 32, v);
                    tmem_wait_ld();
                    if (dump) {
                        float* d = p.dbg + (size_t)layer * 128 * 128 + (size_t)row * 128 + hf * 64 + cb * 32;
#pragma unroll
                        for (int i = 0; i < 32; ++i) d[i] = __uint_as_float(v[i]) * dsc;
                    }
                    if (GENERIC && !al.leaky) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            m = fmaxf(m, fabsf(act_generic16(al.act, fmaf(__uint_as_float(v[i]), dsc, bias[cb * 32 + i]), al.slope)));
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) m = fmaxf(m, fabsf(__uint_as_float(v[i])));
                    }
                }
                if (!(GENERIC && !al.leaky)) m = fmaf(m, fabsf(dsc), s_f[324 + layer]) * fmaxf(1.f, fabsf(al.slope));   // + max |bias|; a learned slope may exceed 1
                xch[hf * 128 + row] = m;
                group_sync(g);
                m = fmaxf(m, xch[(hf ^ 1) * 128 + row]);
                if (layer == 0 && hf == 0) {
                    // range guard of the layer-0 conversion: every row's max |x| inside [2^-5, 2^15] (or exactly 0)
                    const uint32_t mb = rowmax[row];
                    rowmax[row] = 0u;
                    if (mb != 0u && (mb < 0x3D000000u || mb > 0x47000000u)) atomicOr(s_flag + g, 1);
                }
                // row scale: max |y'| in [2^13, 2^14)  (biased exponent E of m -> 2^(140 - E), clamped)
                const uint32_t E = __float_as_uint(m) >> 23;
                uint32_t sb = 267u - E;
                sb = sb > 253u ? 253u : sb;
                const float scale = __uint_as_float(sb << 23);
                inv_scale = __uint_as_float((254u - sb) << 23);
                // (no second barrier: a thread rewrites its xch slot only after the next d_full, which needs the a_ready
                //  arrival every thread of the group makes after this read)
                // pass 2: activations again (TMEM reads are cheap, 64 live values are not), scaled, split, stored as the
                // next layer's A operand: 8 packed columns = 16 K-elements per store
#pragma unroll 1
                for (int cb = 0; cb < 4; ++cb) {
                    uint32_t v[16], hi[8], lo[8];
                    tmem_ld16(lane_addr + COL_D + hf * 64 + cb * 16, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float t0 = act_do<GENERIC>(al, fmaf(__uint_as_float(v[2 * i]), dsc, bias[cb * 16 + 2 * i]));
                        const float t1 = act_do<GENERIC>(al, fmaf(__uint_as_float(v[2 * i + 1]), dsc, bias[cb * 16 + 2 * i + 1]));
                        split_f16x2(t0 * scale, t1 * scale, hi[i], lo[i]);
                    }
                    tmem_st8(lane_addr + COL_AHI + hf * 32 + cb * 8, hi);
                    tmem_st8(lane_addr + COL_ALO + hf * 32 + cb * 8, lo);
                }
                tmem_wait_st();
                tc_fence_before_sync();
                mbar_arrive(&bar->a_ready[g]);
                G_STAMP(j, 3 + 2 * layer);
            }
            // ---------------------------------------------------------------- final epilogue: 32 columns per thread
            {
                mbar_wait(&bar->d_full[g], (uint32_t)(3 * j + 2) & 1);
                tc_fence_after_sync();
                G_STAMP(j, 6);
                const float dsc = s_f[322] * inv_scale;
                const float* bias = s_f + 256 + hf * 32;
                uint32_t v[32];
                tmem_ld32(lane_addr + COL_D + hf * 32, v);
                tmem_wait_ld();
                tc_fence_before_sync();
                mbar_arrive(&bar->d_free[g]);                            // D may be overwritten by the group's next tile
                if (dump) {
                    float* d = p.dbg + (size_t)2 * 128 * 128 + (size_t)row * 64 + hf * 32;
#pragma unroll
                    for (int i = 0; i < 32; ++i) d[i] = __uint_as_float(v[i]) * dsc;
                }
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float t = act_do<GENERIC>(a2, fmaf(__uint_as_float(v[i]), dsc, bias[i]));
                    ss = fmaf(t, t, ss);
                    v[i] = __float_as_uint(t);
                }
                xch[hf * 128 + row] = ss;
                group_sync(g);
                ss += xch[(hf ^ 1) * 128 + row];
                float inv = 1.f;
                if (p.normalize) inv = ss > 0.f ? 1.f / sqrtf(ss) : 0.f; // all-zero row: 0 / 0 = NaN -> 0 in the reference (gnn.py:425-426)
                // rows go through a shared-memory stage (the group's idle A ring) so that the global stores are whole
                // 256-byte rows, two per warp instruction, instead of 32 scattered 16-byte pieces (16-byte pieces of a row
                // XOR-ed with the row number: conflict-free on both sides without padding — the ring is exactly 32 KB)
                float* st = stage + row * GF;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    float4 r;
                    r.x = __uint_as_float(v[i]) * inv; r.y = __uint_as_float(v[i + 1]) * inv;
                    r.z = __uint_as_float(v[i + 2]) * inv; r.w = __uint_as_float(v[i + 3]) * inv;
                    *reinterpret_cast<float4*>(st + ((((hf * 32 + i) >> 2) ^ (row & 7)) << 2)) = r;
                }
                if (tg == 0 && s_flag[g]) {                              // (set during epilogue 0, group barriers ago)
                    s_flag[g] = 0;
                    if (p.flag_ws) { const int k = atomicAdd(p.flag_ws, 1); p.flag_ws[1 + k] = (int)tile; }
                }
                group_sync(g);
                {
                    const int c4 = (tg & 15) * 4, r16 = tg >> 4;
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int r = it * 16 + r16;
                        const long long e = e0 + r;
                        const float4 val = *reinterpret_cast<const float4*>(stage + r * GF + (((c4 >> 2) ^ (r & 7)) << 2));
                        if (e < p.n_edges) stg4(p.s_out + e * GF + c4, val);
                    }
                }
                // (the group barrier after the next tile's endpoint load orders these reads before the next gather's
                //  writes into the ring, and this tile's xch reads before its next write)
                G_STAMP(j, 7);
            }
        }
#undef G_STAMP
    } else if (warp >= 18) {
        // =====================================================================================
        // weight loader of group g: W2 once (loader 0); per tile of the group the W1 chunks, then W3 in two halves
        // =====================================================================================
        const int g = warp - 18;
        const int n_g = (n_my - g + 1) / 2;
        if (lane == 0 && n_my > 0) {
            if (g == 0) {
                mbar_arrive_expect_tx(&bar->w_res, (uint32_t)W_RES_BYTES);
                const unsigned char* src = p.img + img_w2_off(p.n_l1_img);
                for (int i = 0; i < W_RES_BYTES / W1_CHUNK; ++i)
                    bulk_g2s(w_res + (size_t)i * W1_CHUNK, src + (size_t)i * W1_CHUNK, W1_CHUNK, &bar->w_res);
            }
            unsigned char* ring = w_ring + (size_t)g * W_STAGES * W1_CHUNK;
            uint32_t slot = 0, ph = 0;
            auto load = [&](const unsigned char* from) {
                mbar_wait_spin(&bar->w_empty[g][slot], ph ^ 1);
                mbar_arrive_expect_tx(&bar->w_full[g][slot], (uint32_t)W1_CHUNK);
                bulk_g2s(ring + (size_t)slot * W1_CHUNK, from, W1_CHUNK, &bar->w_full[g][slot]);
                if (++slot == W_STAGES) { slot = 0; ph ^= 1; }
            };
            for (int j = 0; j < n_g; ++j) {
                for (int c = 0; c < n_l1; ++c) load(p.img + img_w1_off(2 * p.segs[c >> 1] + (c & 1)));
                load(p.img + img_w3_off(p.n_l1_img));
                load(p.img + img_w3_off(p.n_l1_img) + W1_CHUNK);
            }
        }
    } else {
        // =====================================================================================
        // MMA issuer of group g
        // =====================================================================================
        const int g = warp - 16;
        const int n_g = (n_my - g + 1) / 2;
        if (n_g > 0) {                                                 // the whole warp, converged (see mma_f16_ss_warp)
            const uint32_t idesc128 = make_idesc_f16(128, 128), idesc64 = make_idesc_f16(128, 64);
            // shared-memory descriptors = constant high word | (address >> 4) | LBO bit: one 32-bit add per operand
            const uint64_t desc_hi = make_desc_sw64(0) & 0xFFFFFFFF00000000ull;
            const uint32_t desc_lo0 = (uint32_t)make_desc_sw64(0);
            auto dsc = [&](uint32_t lo) { return desc_hi | (uint64_t)lo; };
            const uint32_t a_ring_d = desc_lo0 + ((smem_u32(a_ring) + (uint32_t)g * A_STAGES * A_SLOT) >> 4),
                           w_ring_d = desc_lo0 + ((smem_u32(w_ring) + (uint32_t)g * W_STAGES * W1_CHUNK) >> 4),
                           w_res_d = desc_lo0 + (smem_u32(w_res) >> 4);
            const uint32_t d = tmem_base + (uint32_t)g * 256u, ta_hi = d + 128u, ta_lo = d + 192u;
            uint32_t a_slot = 0, a_ph = 0, w_slot = 0, w_ph = 0;       // ring positions stepped by compare-and-wrap
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && lane == 0;
#define M_STAMP(j_, ev_) do { if (tr && (j_) < 8) p.trace[2 * 128 + (2 * (j_) + g) * 8 + (ev_)] = clock64(); } while (0)
            mbar_wait_spin(&bar->w_res, 0);
            tc_fence_after_sync();
#pragma unroll 1
            for (int j = 0; j < n_g; ++j) {
                // ---- layer 0 (SS): D (+)= A_chunk · W_chunkᵀ, chunk by chunk as the gather publishes them
                if (j >= 1) {
                    mbar_wait_spin(&bar->d_free[g], (uint32_t)(j - 1) & 1);  // the final epilogue of the previous tile has read D
                    tc_fence_after_sync();
                }
                M_STAMP(j, 0);
#pragma unroll 1
                for (int c = 0; c < n_l1; ++c) {
                    mbar_wait_spin(&bar->a_full[g][a_slot], a_ph);
                    mbar_wait_spin(&bar->w_full[g][w_slot], w_ph);
                    tc_fence_after_sync();
                    const uint32_t a_hi = a_ring_d + a_slot * (A_SLOT >> 4), a_lo = a_hi + (A_TILE >> 4);
                    const uint32_t w_hi = w_ring_d + w_slot * (W1_CHUNK >> 4), w_lo = w_hi + (W1_CHUNK >> 5);
                    mma_f16_ss_warp(d, dsc(a_lo), dsc(w_hi), idesc128, c ? 1u : 0u);        // + 32 B per K = 16 step
                    mma_f16_ss_warp(d, dsc(a_hi), dsc(w_lo), idesc128, 1u);
                    mma_f16_ss_warp(d, dsc(a_hi), dsc(w_hi), idesc128, 1u);
                    mma_f16_ss_warp(d, dsc(a_lo + 2), dsc(w_hi + 2), idesc128, 1u);
                    mma_f16_ss_warp(d, dsc(a_hi + 2), dsc(w_lo + 2), idesc128, 1u);
                    mma_f16_ss_warp(d, dsc(a_hi + 2), dsc(w_hi + 2), idesc128, 1u);
                    mma_commit_warp(&bar->a_empty[g][a_slot]);
                    mma_commit_warp(&bar->w_empty[g][w_slot]);
                    if (++a_slot == A_STAGES) { a_slot = 0; a_ph ^= 1; }
                    if (++w_slot == W_STAGES) { w_slot = 0; w_ph ^= 1; }
                }
                mma_commit_warp(&bar->d_full[g]);
                M_STAMP(j, 1);
                // ---- layer 1 (TS): A from TMEM, W2 resident
                mbar_wait_spin(&bar->a_ready[g], 0u);                        // (2 j) & 1
                tc_fence_after_sync();
                M_STAMP(j, 2);
#pragma unroll
                for (int ks = 0; ks < GH / 16; ++ks) {
                    const uint32_t w_hi = w_res_d + (uint32_t)(ks >> 1) * (W1_CHUNK >> 4) + (uint32_t)(ks & 1) * 2u, w_lo = w_hi + (W1_CHUNK >> 5);
                    mma_f16_ts_warp(d, ta_lo + ks * 8, dsc(w_hi), idesc128, ks ? 1u : 0u);
                    mma_f16_ts_warp(d, ta_hi + ks * 8, dsc(w_lo), idesc128, 1u);
                    mma_f16_ts_warp(d, ta_hi + ks * 8, dsc(w_hi), idesc128, 1u);
                }
                mma_commit_warp(&bar->d_full[g]);
                M_STAMP(j, 3);
                // ---- layer 2 (TS, N = 64): W3 in the group's two ring slots (two K-chunks each)
                mbar_wait_spin(&bar->a_ready[g], 1u);                        // (2 j + 1) & 1
                tc_fence_after_sync();
                M_STAMP(j, 4);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    mbar_wait_spin(&bar->w_full[g][w_slot], w_ph);
                    tc_fence_after_sync();
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int ks = half * 4 + kk;
                        const uint32_t w_hi = w_ring_d + w_slot * (W1_CHUNK >> 4) + (uint32_t)(kk >> 1) * (W3_CHUNK >> 4) + (uint32_t)(kk & 1) * 2u,
                                       w_lo = w_hi + (W3_CHUNK >> 5);
                        mma_f16_ts_warp(d, ta_lo + ks * 8, dsc(w_hi), idesc64, ks ? 1u : 0u);
                        mma_f16_ts_warp(d, ta_hi + ks * 8, dsc(w_lo), idesc64, 1u);
                        mma_f16_ts_warp(d, ta_hi + ks * 8, dsc(w_hi), idesc64, 1u);
                    }
                    mma_commit_warp(&bar->w_empty[g][w_slot]);
                    if (++w_slot == W_STAGES) { w_slot = 0; w_ph ^= 1; }
                }
                mma_commit_warp(&bar->d_full[g]);
                M_STAMP(j, 5);
            }
#undef M_STAMP
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc16
}  // namespace swe

using namespace swe;

static long long* g_tc16_trace = nullptr;      // profiling aid: consumed by the next launch (tools/bench_gate.py)
extern "C" void swe_gate_tc16_set_trace(long long* t) { g_tc16_trace = t; }

extern "C" size_t swe_gate_tc16_image_bytes(int32_t k1) { return tc16::img_bytes(k1 / tc16::KC); }

extern "C" int swe_gate_tc16_pack(const float* w1, int32_t k1, const float* b1, const float* w2, const float* b2,
                                  const float* w3, const float* b3, const float* wmax3, void* image, void* stream) {
    SWE_REQUIRE(w1 && w2 && w3 && image && wmax3, SWE_E_INVAL, "gate_tc16_pack: null pointer");
    SWE_REQUIRE(k1 == 4 * tc16::GF || k1 == 5 * tc16::GF, SWE_E_UNSUPP, "gate_tc16_pack: k1=%d (expected 256 or 320)", k1);
    SWE_REQUIRE(aligned16(image), SWE_E_ALIGN, "gate_tc16_pack: image unaligned");
    int ex[3];
    for (int l = 0; l < 3; ++l) {
        // max |w'| in [2^13, 2^14): 2^ex with ex = 13 - floor(log2 max|w|); an all-zero (or non-finite) matrix keeps ex = 0
        const float m = wmax3[l];
        int e = 0;
        if (m > 0.f && m < 3.0e38f) { (void)frexpf(m, &e); e = 14 - e; }       // m = f * 2^e', f in [0.5, 1): floor(log2 m) = e' - 1
        ex[l] = e < -100 ? -100 : (e > 100 ? 100 : e);
    }
    tc16::gate_tc16_pack_kernel<<<148, 256, 0, (cudaStream_t)stream>>>(w1, k1, b1, w2, b2, w3, b3, ex[0], ex[1], ex[2],
                                                                       (unsigned char*)image);
    return check_launch("gate_tc16_pack");
}

extern "C" int swe_edge_gate_tc_fwd_listed(const float*, const float*, const float*, const float*, const int32_t*, const int32_t*,
                                           int64_t, const void*, int32_t, const int32_t*, const float* const*, int32_t, float*,
                                           const int32_t*, void*);

// image16: swe_gate_tc16_pack; image_tf32: swe_gate_tc_pack of the same weights (the fallback of tiles whose layer-0
// inputs leave the fp16 window); flag_ws: (number of 128-edge tiles + 1) int32 of scratch
extern "C" int swe_edge_gate_tc16_fwd(const float* xs, const float* xd_src, const float* xd_dst, const float* a,
                                      const int32_t* src, const int32_t* dst, int64_t n_edges, const void* image16,
                                      const void* image_tf32, int32_t k1, const int32_t* act3, const float* const* slope3,
                                      int32_t normalize, float* s_out, float* dbg, int32_t* flag_ws, void* stream) {
    SWE_REQUIRE(xs && xd_src && src && dst && s_out && image16 && act3 && slope3 && n_edges >= 0, SWE_E_INVAL,
                "edge_gate_tc16: bad arguments");
    SWE_REQUIRE(aligned16(xs) && aligned16(xd_src) && aligned16(s_out) && aligned16(image16) && (!a || aligned16(a)) &&
                (!xd_dst || aligned16(xd_dst)), SWE_E_ALIGN, "edge_gate_tc16: unaligned buffer");
    SWE_REQUIRE(k1 == (a ? 5 : 4) * tc16::GF, SWE_E_UNSUPP, "edge_gate_tc16: k1=%d does not match the inputs", k1);
    SWE_REQUIRE(!flag_ws || image_tf32, SWE_E_INVAL, "edge_gate_tc16: the range-guard work list needs the TF32 image");
    if (n_edges == 0) return 0;
    tc16::Gate16Params p;
    memset(&p, 0, sizeof(p));
    p.xs = xs; p.xd_src = xd_src; p.xd_dst = xd_dst; p.a = a; p.src = src; p.dst = dst; p.n_edges = n_edges;
    p.img = (const unsigned char*)image16; p.n_l1_img = k1 / tc16::KC;
    bool generic = false;
    for (int i = 0; i < 3; ++i) {
        p.act[i] = act3[i]; p.slope[i] = slope3[i];
        generic |= !(act3[i] == SWE_ACT_NONE || act3[i] == SWE_ACT_PRELU || act3[i] == SWE_ACT_RELU || act3[i] == SWE_ACT_LEAKYRELU);
    }
    p.normalize = normalize; p.s_out = s_out; p.dbg = dbg; p.flag_ws = flag_ws;
    p.trace = g_tc16_trace; g_tc16_trace = nullptr;
    { static int st = -1; if (st < 0) { const char* e = getenv("MSWE_TC16_STAGGER_NS"); st = e ? atoi(e) : 8000; } p.stagger_ns = (unsigned)st; }
    p.n_seg = 0;
    for (int sg = 0; sg < 5; ++sg)
        if (sg < 3 || (sg == 3 && xd_dst) || (sg == 4 && a)) p.segs[p.n_seg++] = sg;
    void (*kern)(const tc16::Gate16Params) = generic ? tc16::edge_gate_tc16_kernel<true> : tc16::edge_gate_tc16_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc16::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("edge_gate_tc16 smem opt-in (%zu B): %s", tc16::SMEM_BYTES, cudaGetErrorString(e)); return (int)e; }
    if (flag_ws) {
        e = cudaMemsetAsync(flag_ws, 0, sizeof(int32_t), (cudaStream_t)stream);
        if (e != cudaSuccess) { set_error("edge_gate_tc16 work-list reset: %s", cudaGetErrorString(e)); return (int)e; }
    }
    const long long n_tiles = (n_edges + tc16::TILE - 1) / tc16::TILE;
    kern<<<grid_for(n_tiles, 1), tc16::N_THREADS, tc16::SMEM_BYTES, (cudaStream_t)stream>>>(p);
    if (int r = check_launch("edge_gate_tc16_fwd")) return r;
    if (flag_ws)
        return swe_edge_gate_tc_fwd_listed(xs, xd_src, xd_dst, a, src, dst, n_edges, image_tf32, k1, act3, slope3, normalize,
                                           s_out, flag_ws, stream);
    return 0;
}
