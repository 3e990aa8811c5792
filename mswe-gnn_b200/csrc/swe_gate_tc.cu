// Edge gate on the 5th-generation tensor cores (tcgen05 + TMEM), F = 64, 3-layer edge MLP
// (5F|4F -> 2F -> 2F -> F) — the default config.yaml model.  Replaces models/gnn.py:414-426.
//
// Precision: fp32 parity (north_star: rel 1e-5 per layer) rules out a single TF32 pass (2^-11).
// Every product is evaluated as 3 TF32 MMAs on error-free splits x = hi + lo
// (hi = x with the low 13 mantissa bits cleared, lo = x - hi, both exact in fp32):
//     A·W ≈ A_hi·W_hi + A_lo·W_hi + A_hi·W_lo          (dropped term |A_lo·W_lo| <= 2^-22 |A||W|)
// accumulated in fp32 in TMEM — per-layer relative error ≈ 1e-6.
//
// One CTA = one 128-edge tile at a time (persistent over tiles), 192 threads:
//   warps 0-3  row workers : gather the layer-0 input K-chunks (x_s[r], x_s[c], x_d[r], x_d[c], a_e),
//                            split hi/lo, store them to shared memory in the UMMA K-major SWIZZLE_128B
//                            layout; later the epilogues (one thread = one TMEM lane = one edge):
//                            tcgen05.ld accumulators -> bias + PReLU -> split -> tcgen05.st as the next
//                            layer's A operand (A stays in TMEM, never touches shared memory) and
//                            finally L2-normalise + store s_ij.
//   warp 4     weight loader: streams pre-packed weight chunk images (already swizzled, hi|lo) from
//                            global/L2 into a shared-memory ring with cp.async.bulk + mbarrier tx.
//   warp 5     MMA issuer   : one thread issues tcgen05.mma (SS for layer 0, TS for layers 1-2) and
//                            tcgen05.commit to release ring slots / publish accumulators.
// TMEM (512 columns): [0,128) A_hi, [128,256) A_lo, [256,384) D_a (layer 0 / layer 2), [384,512) D_b.
#include "swe_tc.cuh"

namespace swe {
namespace tc {

constexpr int GF = 64;                    // feature width this kernel is built for
constexpr int GH = 128;                   // hidden width 2F
constexpr int KC = 32;                    // K elements per chunk = one 128-byte swizzled row
constexpr int TILE_ROWS = 128;
constexpr int TILE16K = TILE_ROWS * 128;  // bytes of a [128 x 32] tf32 tile
constexpr int SLOT_BYTES = 2 * TILE16K;   // hi tile + lo tile
constexpr int A_STAGES = 3;
constexpr int W_STAGES = 3;
constexpr int N_ROW_THREADS = 128;
constexpr int N_THREADS = 192;
constexpr uint32_t COL_A_HI = 0, COL_A_LO = 128, COL_D_A = 256, COL_D_B = 384;

struct __align__(8) Barriers {
    uint64_t a_full[A_STAGES], a_empty[A_STAGES];
    uint64_t w_full[W_STAGES], w_empty[W_STAGES];
    uint64_t d_full[3];        // accumulators of layer 0/1/2 complete (tcgen05.commit)
    uint64_t a_ready[2];       // next layer's A operand written to TMEM (128 arrivals)
    uint64_t d_free;           // last accumulator drained, next tile may start (128 arrivals)
};

constexpr size_t GATE_TC_SMEM = 1024 /*align slack*/ + (size_t)(A_STAGES + W_STAGES) * SLOT_BYTES +
                                sizeof(float) * 320 + sizeof(int32_t) * 2 * TILE_ROWS + sizeof(Barriers) + 16;

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
};

__device__ __forceinline__ int l1_chunk_segment(int i, bool has_xd_dst) {
    // i-th ACTIVE chunk of layer 0 -> input segment (0 x_s[r], 1 x_s[c], 2 x_d[r], 3 x_d[c], 4 a_e)
    const int sg = i >> 1;
    return sg < 3 ? sg : ((sg == 3 && has_xd_dst) ? 3 : 4);
}

__global__ void __launch_bounds__(N_THREADS, 1) edge_gate_tc_kernel(const __grid_constant__ GateTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* a_ring = smem;
    unsigned char* w_ring = smem + (size_t)A_STAGES * SLOT_BYTES;
    float* s_bias = reinterpret_cast<float*>(w_ring + (size_t)W_STAGES * SLOT_BYTES);
    int32_t* s_src = reinterpret_cast<int32_t*>(s_bias + 320);
    int32_t* s_dst = s_src + TILE_ROWS;
    Barriers* bar = reinterpret_cast<Barriers*>(s_dst + TILE_ROWS);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool has_xd_dst = p.xd_dst != nullptr, has_a = p.a != nullptr;
    const int n_l1 = 2 * (3 + (has_xd_dst ? 1 : 0) + (has_a ? 1 : 0));      // active layer-0 chunks
    const long long n_tiles = (p.n_edges + TILE_ROWS - 1) / TILE_ROWS;

    if (threadIdx.x == 0) {
        for (int i = 0; i < A_STAGES; ++i) { mbar_init(&bar->a_full[i], N_ROW_THREADS); mbar_init(&bar->a_empty[i], 1); }
        for (int i = 0; i < W_STAGES; ++i) { mbar_init(&bar->w_full[i], 1); mbar_init(&bar->w_empty[i], 1); }
        for (int i = 0; i < 3; ++i) mbar_init(&bar->d_full[i], 1);
        for (int i = 0; i < 2; ++i) mbar_init(&bar->a_ready[i], N_ROW_THREADS);
        mbar_init(&bar->d_free, N_ROW_THREADS);
        fence_barrier_init();
    }
    {
        const float* gb = reinterpret_cast<const float*>(p.img + img_bias_off(p.n_l1_img));
        for (int i = threadIdx.x; i < 320; i += N_THREADS) s_bias[i] = gb[i];
    }
    if (warp == 4) tmem_alloc(tmem_holder, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_holder;

    if (warp < 4) {
        // =====================================================================================
        // row workers: producer of layer-0 A chunks, then epilogues
        // =====================================================================================
        const int row = threadIdx.x;                                  // TMEM lane / edge within the tile
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const float sl0 = (p.act[0] == SWE_ACT_PRELU && p.slope[0]) ? __ldg(p.slope[0]) : 0.f;
        const float sl1 = (p.act[1] == SWE_ACT_PRELU && p.slope[1]) ? __ldg(p.slope[1]) : 0.f;
        const float sl2 = (p.act[2] == SWE_ACT_PRELU && p.slope[2]) ? __ldg(p.slope[2]) : 0.f;
        uint32_t a_cnt = 0;                                           // A-ring uses so far
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const long long e0 = tile * TILE_ROWS;
            {   // edge endpoints of this tile (previous tile's gathers have all completed: they are
                // synchronous register loads issued before the a_full arrivals)
                long long e = e0 + row;
                if (e >= p.n_edges) e = p.n_edges - 1;
                asm volatile("bar.sync 1, 128;" ::: "memory");        // everyone done reading old ids
                s_src[row] = __ldg(p.src + e);
                s_dst[row] = __ldg(p.dst + e);
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            // ---------------- layer-0 input chunks
            const int piece = row & 7, r0 = row >> 3;
            for (int c = 0; c < n_l1; ++c, ++a_cnt) {
                const int sg = l1_chunk_segment(c, has_xd_dst);
                const int koff = (c & 1) * KC + piece * 4;
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = r0 + 16 * i;
                    const float* srcp;
                    if (sg == 4) {
                        long long e = e0 + r;
                        if (e >= p.n_edges) e = p.n_edges - 1;
                        srcp = p.a + e * GF;
                    } else {
                        const long long node = (sg & 1) ? s_dst[r] : s_src[r];
                        srcp = (sg < 2 ? p.xs : (sg == 2 ? p.xd_src : p.xd_dst)) + node * GF;
                    }
                    v[i] = (sg == 4) ? ldg4_stream(srcp + koff) : ldg4(srcp + koff);
                }
                const uint32_t slot = a_cnt % A_STAGES;
                mbar_wait(&bar->a_empty[slot], ((a_cnt / A_STAGES) & 1) ^ 1);
                unsigned char* hi_t = a_ring + (size_t)slot * SLOT_BYTES;
                unsigned char* lo_t = hi_t + TILE16K;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = r0 + 16 * i;
                    float4 h, l;
                    split_tf32(v[i].x, h.x, l.x); split_tf32(v[i].y, h.y, l.y);
                    split_tf32(v[i].z, h.z, l.z); split_tf32(v[i].w, h.w, l.w);
                    const uint32_t off = sw128_offset(r, piece * 4);
                    *reinterpret_cast<float4*>(hi_t + off) = h;
                    *reinterpret_cast<float4*>(lo_t + off) = l;
                }
                fence_proxy_async_smem();
                mbar_arrive(&bar->a_full[slot]);
            }
            const uint32_t ph = it & 1;
            // ---------------- epilogue of layer 0 and 1: D -> bias, activation -> hi/lo -> TMEM A operand
#pragma unroll 1
            for (int layer = 0; layer < 2; ++layer) {
                mbar_wait(&bar->d_full[layer], ph);
                tc_fence_after_sync();
                const uint32_t dcol = layer == 0 ? COL_D_A : COL_D_B;
                const float* bias = s_bias + layer * 128;
                const int act = p.act[layer];
                const float sl = layer == 0 ? sl0 : sl1;
#pragma unroll 1
                for (int cb = 0; cb < 4; ++cb) {
                    uint32_t v[32], hi[32], lo[32];
                    tmem_ld32(lane_addr + dcol + cb * 32, v);
                    tmem_wait_ld();
                    if (p.dbg && tile == 0) {
                        float* d = p.dbg + (size_t)layer * 128 * 128 + (size_t)row * 128 + cb * 32;
#pragma unroll
                        for (int j = 0; j < 32; ++j) d[j] = __uint_as_float(v[j]);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float y = act_apply(act, __uint_as_float(v[j]) + bias[cb * 32 + j], sl);
                        float h, l;
                        split_tf32(y, h, l);
                        hi[j] = __float_as_uint(h);
                        lo[j] = __float_as_uint(l);
                    }
                    tmem_st32(lane_addr + COL_A_HI + cb * 32, hi);
                    tmem_st32(lane_addr + COL_A_LO + cb * 32, lo);
                }
                tmem_wait_st();
                tc_fence_before_sync();
                mbar_arrive(&bar->a_ready[layer]);
            }
            // ---------------- final epilogue: bias, activation, L2 normalise, store s_ij
            {
                mbar_wait(&bar->d_full[2], ph);
                tc_fence_after_sync();
                uint32_t v0[32], v1[32];
                tmem_ld32(lane_addr + COL_D_A, v0);
                tmem_ld32(lane_addr + COL_D_A + 32, v1);
                tmem_wait_ld();
                tc_fence_before_sync();
                mbar_arrive(&bar->d_free);                            // accumulator drained
                if (p.dbg && tile == 0) {
                    float* d = p.dbg + (size_t)2 * 128 * 128 + (size_t)row * 64;
#pragma unroll
                    for (int j = 0; j < 32; ++j) { d[j] = __uint_as_float(v0[j]); d[32 + j] = __uint_as_float(v1[j]); }
                }
                const float* bias = s_bias + 256;
                float y[64];
                float ss = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    y[j] = act_apply(p.act[2], __uint_as_float(v0[j]) + bias[j], sl2);
                    y[32 + j] = act_apply(p.act[2], __uint_as_float(v1[j]) + bias[32 + j], sl2);
                }
                if (p.normalize) {
#pragma unroll
                    for (int j = 0; j < 64; ++j) ss = fmaf(y[j], y[j], ss);
                    const float nrm = sqrtf(ss);
#pragma unroll
                    for (int j = 0; j < 64; ++j) {
                        const float q = __fdiv_rn(y[j], nrm);         // s / ||s||   (gnn.py:425)
                        y[j] = (q != q) ? 0.f : q;                    // NaN -> 0    (gnn.py:426)
                    }
                }
                const long long e = e0 + row;
                if (e < p.n_edges) {
                    float* o = p.s_out + e * GF;
#pragma unroll
                    for (int j = 0; j < 64; j += 4) stg4(o + j, make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]));
                }
            }
        }
    } else if (warp == 4) {
        // =====================================================================================
        // weight loader
        // =====================================================================================
        if (lane == 0) {
            uint32_t w_cnt = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int c = 0; c < n_l1 + 8; ++c, ++w_cnt) {
                    const unsigned char* srcp;
                    uint32_t bytes = SLOT_BYTES;
                    if (c < n_l1) {
                        const int sg = l1_chunk_segment(c, has_xd_dst);
                        srcp = p.img + img_l1_off(2 * sg + (c & 1));
                    } else if (c < n_l1 + 4) {
                        srcp = p.img + img_l2_off(p.n_l1_img, c - n_l1);
                    } else {
                        srcp = p.img + img_l3_off(p.n_l1_img, c - n_l1 - 4);
                        bytes = SLOT_BYTES / 2;
                    }
                    const uint32_t slot = w_cnt % W_STAGES;
                    mbar_wait(&bar->w_empty[slot], ((w_cnt / W_STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bar->w_full[slot], bytes);
                    bulk_g2s(w_ring + (size_t)slot * SLOT_BYTES, srcp, bytes, &bar->w_full[slot]);
                }
            }
        }
    } else {
        // =====================================================================================
        // MMA issuer
        // =====================================================================================
        if (lane == 0) {
            const uint32_t idesc128 = make_idesc_tf32(128, 128), idesc64 = make_idesc_tf32(128, 64);
            const uint32_t a_ring_u32 = smem_u32(a_ring), w_ring_u32 = smem_u32(w_ring);
            uint32_t a_cnt = 0, w_cnt = 0, it = 0;
            for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                if (it > 0) { mbar_wait(&bar->d_free, (it - 1) & 1); tc_fence_after_sync(); }
                // ---------------- layer 0 (SS): D_a = Σ_chunks A_chunk · W_chunkᵀ
                for (int c = 0; c < n_l1; ++c, ++a_cnt, ++w_cnt) {
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
                        mma_tf32_ss(tmem_base + COL_D_A, dal, dwh, idesc128, (c | ks) ? 1u : 0u);
                        mma_tf32_ss(tmem_base + COL_D_A, dah, dwl, idesc128, 1u);
                        mma_tf32_ss(tmem_base + COL_D_A, dah, dwh, idesc128, 1u);
                    }
                    mma_commit(&bar->a_empty[sa]);
                    mma_commit(&bar->w_empty[sw]);
                }
                mma_commit(&bar->d_full[0]);
                // ---------------- layers 1 and 2 (TS): A from TMEM
#pragma unroll 1
                for (int layer = 1; layer < 3; ++layer) {
                    mbar_wait(&bar->a_ready[layer - 1], ph);
                    tc_fence_after_sync();
                    const uint32_t dcol = layer == 1 ? COL_D_B : COL_D_A;
                    const uint32_t idesc = layer == 1 ? idesc128 : idesc64;
                    const uint32_t lo_off = layer == 1 ? TILE16K : TILE16K / 2;       // lo tile follows hi tile
                    for (int c = 0; c < 4; ++c, ++w_cnt) {
                        const uint32_t sw = w_cnt % W_STAGES;
                        mbar_wait(&bar->w_full[sw], (w_cnt / W_STAGES) & 1);
                        tc_fence_after_sync();
                        const uint32_t w_hi = w_ring_u32 + sw * SLOT_BYTES, w_lo = w_hi + lo_off;
#pragma unroll
                        for (int ks = 0; ks < KC / 8; ++ks) {
                            const uint32_t kcol = c * KC + ks * 8;
                            const uint64_t dwh = make_desc_sw128(w_hi + ks * 32), dwl = make_desc_sw128(w_lo + ks * 32);
                            mma_tf32_ts(tmem_base + dcol, tmem_base + COL_A_LO + kcol, dwh, idesc, (c | ks) ? 1u : 0u);
                            mma_tf32_ts(tmem_base + dcol, tmem_base + COL_A_HI + kcol, dwl, idesc, 1u);
                            mma_tf32_ts(tmem_base + dcol, tmem_base + COL_A_HI + kcol, dwh, idesc, 1u);
                        }
                        mma_commit(&bar->w_empty[sw]);
                    }
                    mma_commit(&bar->d_full[layer]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace swe

using namespace swe;

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
    SWE_REQUIRE(xs && xd_src && src && dst && s_out && image && act3 && slope3 && n_edges >= 0, SWE_E_INVAL,
                "edge_gate_tc: bad arguments");
    SWE_REQUIRE(aligned16(xs) && aligned16(xd_src) && aligned16(s_out) && aligned16(image) && (!a || aligned16(a)) &&
                (!xd_dst || aligned16(xd_dst)), SWE_E_ALIGN, "edge_gate_tc: unaligned buffer");
    SWE_REQUIRE(k1 == (a ? 5 : 4) * tc::GF, SWE_E_UNSUPP, "edge_gate_tc: k1=%d does not match the inputs", k1);
    if (n_edges == 0) return 0;
    tc::GateTcParams p;
    p.xs = xs; p.xd_src = xd_src; p.xd_dst = xd_dst; p.a = a; p.src = src; p.dst = dst; p.n_edges = n_edges;
    p.img = (const unsigned char*)image; p.n_l1_img = k1 / tc::KC;
    for (int i = 0; i < 3; ++i) { p.act[i] = act3[i]; p.slope[i] = slope3[i]; }
    p.normalize = normalize; p.s_out = s_out; p.dbg = dbg;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc::edge_gate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)tc::GATE_TC_SMEM);
        if (e != cudaSuccess) { set_error("edge_gate_tc smem opt-in (%zu B): %s", tc::GATE_TC_SMEM, cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    const long long n_tiles = (n_edges + tc::TILE_ROWS - 1) / tc::TILE_ROWS;
    tc::edge_gate_tc_kernel<<<grid_for(n_tiles, 1), tc::N_THREADS, tc::GATE_TC_SMEM, (cudaStream_t)stream>>>(p);
    return check_launch("edge_gate_tc_fwd");
}
