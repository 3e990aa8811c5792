// Hop kernel with the F×F filter on the 5th-generation tensor cores (F = 64).  Replaces
// models/gnn.py:428-443 like hop_kernel<64,true>:
//     out[c] = act( o[c] + (Σ_p s_p ⊙ (o[c] − o[src_p]))·Wᵀ + addend[c] )
//
// Why: the hop is HBM-bound (1.28 KB per node), but on CUDA cores the 2·64² FLOP/node filter costs about
// as many issue slots as the node's memory time and the FFMA phase stalls the loads.  Here the CTA's
// threads do nothing but gather and aggregate; the filter is 24 tcgen05.mma per 128-node tile.
//
// One CTA = 256 threads, 128-node tiles, two CTAs per SM (their phases overlap each other):
//   1. the tile's rowptr and its contiguous slice of `src` are staged in shared memory (coalesced), so
//      the dependent chain per node is a single global round trip;
//   2. 16 lanes × 128 bit per node row, TWO nodes in flight per thread (up to 18 independent 16-B loads):
//      agg = Σ_p s_p ⊙ (o[c] − o[src_p]) in the reference's edge order (bit-identical to the FFMA kernel);
//   3. agg is split error-free into TF32 hi/lo and stored as the A operand (UMMA K-major SWIZZLE_128B);
//      one thread issues agg·Wᵀ as 3xTF32 (A_lo·W_hi + A_hi·W_lo + A_hi·W_hi), fp32 accumulation in TMEM;
//   4. epilogue: thread = (TMEM lane = node, 32 columns): tcgen05.ld, + o[c] (+ addend), activation, store.
#include "swe_tc.cuh"

namespace swe {
namespace tc {

constexpr int HF = 64;                         // feature width
constexpr int HOP_THREADS = 256;
constexpr int HOP_TILE = 128;
constexpr int HOP_KC = 32;                     // k elements per 128-byte swizzled row
constexpr int HOP_A_TILE = HOP_TILE * 128;     // bytes of one [128 x 32] tf32 tile
constexpr int HOP_W_TILE = HF * 128;           // bytes of one [64 x 32] tf32 tile
constexpr int HOP_SRC_CAP = 1024;              // staged src ids per tile (more -> read from global)
constexpr size_t HOP_W_IMAGE = 2 * 2 * (size_t)HOP_W_TILE;      // 2 chunks x (hi | lo) = 32 KB
constexpr size_t HOP_TC_SMEM = 1024 + 2 * 2 * (size_t)HOP_A_TILE + HOP_W_IMAGE +
                               sizeof(int32_t) * (HOP_TILE + 4 + HOP_SRC_CAP) + 64;

__global__ void hop_tc_pack_kernel(const float* __restrict__ w, unsigned char* __restrict__ img) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < HF * HF; idx += gridDim.x * blockDim.x) {
        const int n = idx / HF, k = idx % HF;
        float hi, lo;
        split_tf32(w[idx], hi, lo);
        const size_t base = (size_t)(k / HOP_KC) * 2 * HOP_W_TILE;
        const uint32_t off = sw128_offset(n, k % HOP_KC);
        *reinterpret_cast<float*>(img + base + off) = hi;
        *reinterpret_cast<float*>(img + base + HOP_W_TILE + off) = lo;
    }
}

struct HopTcParams {
    const float* o_src; const float* o_dst; const float* s;
    const int32_t* rowptr; const int32_t* src;
    int dst_lo, n_dst;
    const unsigned char* w_img;
    int with_gradient, upwind;
    const float* addend;
    int act; const float* slope;
    float* out;
    float* agg_out;
    long long* trace;           // optional [32 tiles][12 events] clock64 stamps of CTA 0 / thread 0 (profiling aid)
};

struct NodeAcc { float4 acc; };

// first (up to) 4 edges of two nodes in flight together, remaining edges (rare) sequentially
__device__ __forceinline__ void aggregate2(const HopTcParams& p, const int32_t* __restrict__ s_src, bool staged, int p_base,
                                           int pa0, int pa1, int pb0, int pb1, float4 oca, float4 ocb, int q,
                                           float4& acca, float4& accb) {
    auto src_of = [&](int pp) -> int { return staged ? s_src[pp - p_base] : __ldg(p.src + pp); };
    auto term = [&](const float4 oc, const float4 orow, const float4 sv) -> float4 {
        float4 t;
        if (p.with_gradient) {
            float4 d = make_float4(__fsub_rn(oc.x, orow.x), __fsub_rn(oc.y, orow.y), __fsub_rn(oc.z, orow.z), __fsub_rn(oc.w, orow.w));
            if (p.upwind) { d.x = fmaxf(d.x, 0.f); d.y = fmaxf(d.y, 0.f); d.z = fmaxf(d.z, 0.f); d.w = fmaxf(d.w, 0.f); }
            t = make_float4(__fmul_rn(d.x, sv.x), __fmul_rn(d.y, sv.y), __fmul_rn(d.z, sv.z), __fmul_rn(d.w, sv.w));
        } else {
            t = make_float4(__fmul_rn(sv.x, orow.x), __fmul_rn(sv.y, orow.y), __fmul_rn(sv.z, orow.z), __fmul_rn(sv.w, orow.w));
        }
        return t;
    };
    auto add = [](float4& a, const float4 t) {
        a.x = __fadd_rn(a.x, t.x); a.y = __fadd_rn(a.y, t.y); a.z = __fadd_rn(a.z, t.z); a.w = __fadd_rn(a.w, t.w);
    };
    int ra[4], rb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        ra[u] = (pa0 + u < pa1) ? src_of(pa0 + u) : -1;
        rb[u] = (pb0 + u < pb1) ? src_of(pb0 + u) : -1;
    }
    float4 oa[4], sa[4], ob[4], sb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (ra[u] >= 0) {
            oa[u] = ldg4(p.o_src + (long long)ra[u] * HF + 4 * q);
            sa[u] = ldg4_stream(p.s + (long long)(pa0 + u) * HF + 4 * q);
        }
        if (rb[u] >= 0) {
            ob[u] = ldg4(p.o_src + (long long)rb[u] * HF + 4 * q);
            sb[u] = ldg4_stream(p.s + (long long)(pb0 + u) * HF + 4 * q);
        }
    }
    acca = make_float4(0.f, 0.f, 0.f, 0.f);
    accb = acca;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (ra[u] >= 0) add(acca, term(oca, oa[u], sa[u]));
        if (rb[u] >= 0) add(accb, term(ocb, ob[u], sb[u]));
    }
    for (int pp = pa0 + 4; pp < pa1; ++pp) {
        const int r = src_of(pp);
        add(acca, term(oca, ldg4(p.o_src + (long long)r * HF + 4 * q), ldg4_stream(p.s + (long long)pp * HF + 4 * q)));
    }
    for (int pp = pb0 + 4; pp < pb1; ++pp) {
        const int r = src_of(pp);
        add(accb, term(ocb, ldg4(p.o_src + (long long)r * HF + 4 * q), ldg4_stream(p.s + (long long)pp * HF + 4 * q)));
    }
}

__global__ void __launch_bounds__(HOP_THREADS, 2) hop_tc_kernel(const __grid_constant__ HopTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    // 1 KB alignment by OFFSETTING the shared array (integer arithmetic on the pointer value would turn every
    // later access into a generic LD/ST instead of LDS/STS)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_tile = smem;                                  // chunk c: [hi 16 KB | lo 16 KB]
    unsigned char* w_tile = smem + 4 * (size_t)HOP_A_TILE;         // chunk c: [hi 8 KB | lo 8 KB]
    int32_t* s_rp = reinterpret_cast<int32_t*>(w_tile + HOP_W_IMAGE);          // [129 (+3 pad)]
    int32_t* s_src = s_rp + HOP_TILE + 4;                          // [HOP_SRC_CAP]
    uint64_t* d_full = reinterpret_cast<uint64_t*>(s_src + HOP_SRC_CAP);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(d_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(d_full, 1); fence_barrier_init(); }
    for (int i = threadIdx.x * 16; i < (int)HOP_W_IMAGE; i += HOP_THREADS * 16)
        *reinterpret_cast<float4*>(w_tile + i) = *reinterpret_cast<const float4*>(p.w_img + i);
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc(tmem_holder, 64);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_d = *tmem_holder;

    const float slope = (p.act == SWE_ACT_PRELU && p.slope) ? __ldg(p.slope) : 0.f;
    const int g = threadIdx.x >> 4, q = threadIdx.x & 15;          // aggregation: 16 groups x 16 lanes
    const int chunk = q >> 3, piece = q & 7;
    const int lq = warp & 3, hf = warp >> 2;                       // epilogue: TMEM lane quarter, column half
    const uint32_t idesc = make_idesc_tf32(HOP_TILE, HF);
    const uint32_t a_u32 = smem_u32(a_tile), w_u32 = smem_u32(w_tile);
    const int n_tiles = (p.n_dst + HOP_TILE - 1) / HOP_TILE;
    uint32_t phase = 0;

    const bool tr = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
    int t_i = 0;
#define SWE_STAMP(ev_) do { if (tr && t_i < 32) p.trace[t_i * 12 + (ev_)] = clock64(); } while (0)
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t_i) {
        const int row0 = tile * HOP_TILE;
        SWE_STAMP(0);
        const int rows = min(HOP_TILE, p.n_dst - row0);
        // ---- 1. stage the CSR slice of this tile
        if (threadIdx.x <= HOP_TILE) s_rp[threadIdx.x] = __ldg(p.rowptr + row0 + min((int)threadIdx.x, rows));
        __syncthreads();
        const int p_base = s_rp[0], p_cnt = s_rp[HOP_TILE] - p_base;
        const bool staged = p_cnt <= HOP_SRC_CAP;
        if (staged)
            for (int j = threadIdx.x; j < p_cnt; j += HOP_THREADS) s_src[j] = __ldg(p.src + p_base + j);
        __syncthreads();
        SWE_STAMP(1);
        // ---- 2./3. aggregate (two nodes per thread at a time) and write the A operand
#pragma unroll 1
        for (int jj = 0; jj < HOP_TILE / 16; jj += 2) {
            SWE_STAMP(2 + jj / 2);
            const int ia = g + 16 * jj, ib = ia + 16;
            float4 oca = make_float4(0.f, 0.f, 0.f, 0.f), ocb = oca;
            const int pa0 = s_rp[ia], pa1 = s_rp[ia + 1], pb0 = s_rp[ib], pb1 = s_rp[ib + 1];
            if (p.o_dst) {
                if (ia < rows) oca = ldg4(p.o_dst + ((long long)p.dst_lo + row0 + ia) * HF + 4 * q);
                if (ib < rows) ocb = ldg4(p.o_dst + ((long long)p.dst_lo + row0 + ib) * HF + 4 * q);
            }
            float4 acca, accb;
            aggregate2(p, s_src, staged, p_base, pa0, pa1, pb0, pb1, oca, ocb, q, acca, accb);
            if (p.agg_out) {
                if (ia < rows) stg4(p.agg_out + ((long long)p.dst_lo + row0 + ia) * HF + 4 * q, acca);
                if (ib < rows) stg4(p.agg_out + ((long long)p.dst_lo + row0 + ib) * HF + 4 * q, accb);
            }
            unsigned char* base = a_tile + (size_t)chunk * 2 * HOP_A_TILE;
            float4 hh, ll;
            split_tf32(acca.x, hh.x, ll.x); split_tf32(acca.y, hh.y, ll.y); split_tf32(acca.z, hh.z, ll.z); split_tf32(acca.w, hh.w, ll.w);
            uint32_t off = sw128_offset(ia, piece * 4);
            *reinterpret_cast<float4*>(base + off) = hh;
            *reinterpret_cast<float4*>(base + HOP_A_TILE + off) = ll;
            split_tf32(accb.x, hh.x, ll.x); split_tf32(accb.y, hh.y, ll.y); split_tf32(accb.z, hh.z, ll.z); split_tf32(accb.w, hh.w, ll.w);
            off = sw128_offset(ib, piece * 4);
            *reinterpret_cast<float4*>(base + off) = hh;
            *reinterpret_cast<float4*>(base + HOP_A_TILE + off) = ll;
        }
        SWE_STAMP(6);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        SWE_STAMP(7);
        // ---- filter on the tensor core: D[128 x 64] = agg · Wᵀ (3xTF32)
        if (threadIdx.x == 0) {
            tc_fence_after_sync();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t a_hi = a_u32 + c * 2 * HOP_A_TILE, a_lo = a_hi + HOP_A_TILE;
                const uint32_t w_hi = w_u32 + c * 2 * HOP_W_TILE, w_lo = w_hi + HOP_W_TILE;
#pragma unroll
                for (int ks = 0; ks < HOP_KC / 8; ++ks) {
                    const uint64_t dah = make_desc_sw128(a_hi + ks * 32), dal = make_desc_sw128(a_lo + ks * 32);
                    const uint64_t dwh = make_desc_sw128(w_hi + ks * 32), dwl = make_desc_sw128(w_lo + ks * 32);
                    mma_tf32_ss(tmem_d, dal, dwh, idesc, (c | ks) ? 1u : 0u);
                    mma_tf32_ss(tmem_d, dah, dwl, idesc, 1u);
                    mma_tf32_ss(tmem_d, dah, dwh, idesc, 1u);
                }
            }
            mma_commit(d_full);
        }
        // ---- 4. epilogue
        SWE_STAMP(8);
        mbar_wait(d_full, phase);
        phase ^= 1;
        tc_fence_after_sync();
        SWE_STAMP(9);
        {
            uint32_t v[32];
            tmem_ld32(tmem_d + ((uint32_t)(lq * 32) << 16) + hf * 32, v);
            tmem_wait_ld();
            const int i = lq * 32 + lane;
            if (i < rows) {
                const long long c = (long long)p.dst_lo + row0 + i;
                const float* od = p.o_dst ? p.o_dst + c * HF + hf * 32 : nullptr;
                const float* ad = p.addend ? p.addend + c * HF + hf * 32 : nullptr;
                float* o = p.out + c * HF + hf * 32;
                float4 tod[8];                           // all loads first: the stores below may alias for the compiler
#pragma unroll
                for (int j = 0; j < 8; ++j) tod[j] = od ? ldg4(od + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 r = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                           __uint_as_float(v[j + 3]));
                    if (od) { const float4 t = tod[j >> 2]; r.x = t.x + r.x; r.y = t.y + r.y; r.z = t.z + r.z; r.w = t.w + r.w; }
                    if (ad) { const float4 t = ldg4(ad + j); r.x += t.x; r.y += t.y; r.z += t.z; r.w += t.w; }
                    if (p.act != SWE_ACT_NONE) {
                        r.x = act_apply(p.act, r.x, slope); r.y = act_apply(p.act, r.y, slope);
                        r.z = act_apply(p.act, r.z, slope); r.w = act_apply(p.act, r.w, slope);
                    }
                    stg4(o + j, r);
                }
            }
        }
        tc_fence_before_sync();          // the next tile's MMA must not overwrite D before these loads retired
        SWE_STAMP(10);
    }
#undef SWE_STAMP
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 64);
}

}  // namespace tc
}  // namespace swe

using namespace swe;

extern "C" size_t swe_hop_tc_image_bytes(void) { return tc::HOP_W_IMAGE; }

extern "C" int swe_hop_tc_pack(const float* w, void* image, void* stream) {
    SWE_REQUIRE(w && image && aligned16(image), SWE_E_INVAL, "hop_tc_pack: bad arguments");
    tc::hop_tc_pack_kernel<<<16, 256, 0, (cudaStream_t)stream>>>(w, (unsigned char*)image);
    return check_launch("hop_tc_pack");
}

extern "C" int swe_propagate_hop_tc_fwd_traced(const float*, const float*, const float*, const int32_t*, const int32_t*, int32_t,
                                               int32_t, const void*, int32_t, int32_t, const float*, int32_t, const float*,
                                               float*, float*, long long*, void*);

extern "C" int swe_propagate_hop_tc_fwd(const float* o_src, const float* o_dst, const float* s, const int32_t* rowptr,
                                        const int32_t* src, int32_t dst_lo, int32_t n_dst, const void* w_image,
                                        int32_t with_gradient, int32_t upwind, const float* addend, int32_t act,
                                        const float* slope, float* agg_out, float* out, void* stream) {
    return swe_propagate_hop_tc_fwd_traced(o_src, o_dst, s, rowptr, src, dst_lo, n_dst, w_image, with_gradient, upwind,
                                           addend, act, slope, agg_out, out, nullptr, stream);
}

// + optional clock64 phase stamps of CTA 0 (profiling aid, not part of the ABI)
extern "C" int swe_propagate_hop_tc_fwd_traced(const float* o_src, const float* o_dst, const float* s, const int32_t* rowptr,
                                               const int32_t* src, int32_t dst_lo, int32_t n_dst, const void* w_image,
                                               int32_t with_gradient, int32_t upwind, const float* addend, int32_t act,
                                               const float* slope, float* agg_out, float* out, long long* trace,
                                               void* stream) {
    SWE_REQUIRE(o_src && s && rowptr && src && out && w_image && dst_lo >= 0 && n_dst >= 0, SWE_E_INVAL, "hop_tc: bad arguments");
    SWE_REQUIRE(!(with_gradient && !o_dst), SWE_E_INVAL, "hop_tc: with_gradient needs the destination rows");
    SWE_REQUIRE(aligned16(o_src) && aligned16(s) && aligned16(out) && aligned16(w_image) && (!o_dst || aligned16(o_dst)) &&
                (!addend || aligned16(addend)) && (!agg_out || aligned16(agg_out)), SWE_E_ALIGN, "hop_tc: unaligned buffer");
    SWE_REQUIRE(out != o_src && out != o_dst, SWE_E_INVAL, "hop_tc: output must not alias the hop input");
    if (n_dst == 0) return 0;
    tc::HopTcParams p;
    p.o_src = o_src; p.o_dst = o_dst; p.s = s; p.rowptr = rowptr; p.src = src; p.dst_lo = dst_lo; p.n_dst = n_dst;
    p.w_img = (const unsigned char*)w_image; p.with_gradient = with_gradient; p.upwind = upwind; p.addend = addend;
    p.act = act; p.slope = slope; p.out = out; p.agg_out = agg_out; p.trace = trace;
    cudaError_t e = cudaFuncSetAttribute(tc::hop_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::HOP_TC_SMEM);
    if (e != cudaSuccess) { set_error("hop_tc smem opt-in (%zu B): %s", tc::HOP_TC_SMEM, cudaGetErrorString(e)); return (int)e; }
    const int n_tiles = (n_dst + tc::HOP_TILE - 1) / tc::HOP_TILE;
    tc::hop_tc_kernel<<<grid_for(n_tiles, 2), tc::HOP_THREADS, tc::HOP_TC_SMEM, (cudaStream_t)stream>>>(p);
    return check_launch("propagate_hop_tc_fwd");
}
