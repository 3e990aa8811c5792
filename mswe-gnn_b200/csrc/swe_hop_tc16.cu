// Hop kernel with the F×F filter on the 5th-generation tensor cores (F = 64), fp16 hi/lo edition (kind::f16, K = 16 per
// instruction, 64-byte swizzle: 12 tcgen05.mma per 128-node tile instead of 24 (same cycles each), and an A operand of
// 32 KB instead of 64 KB — the shared-memory / L1 pipe this kernel's gathers live on gets that bandwidth and capacity
// back).  Same contract as swe_hop_tc.cu.  agg rows are scaled per row by a power of two (max |agg'| in [2^13, 2^14),
// exact) before the split, the filter per matrix; the epilogue undoes both.  Replaces
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
#include <stdlib.h>
#include "swe_tc.cuh"

namespace swe {
namespace tch {
using namespace swe::tc;

constexpr int HF = 64;                         // feature width
constexpr int HOP_TILE = 128;
constexpr int HOP_KC = 32;                     // k elements per 128-byte swizzled row
constexpr int HOP_A_TILE = HOP_TILE * 64;      // bytes of one [128 x 32] fp16 tile
constexpr int HOP_A_SLOT = 4 * HOP_A_TILE;     // 2 chunks x (hi | lo) = 32 KB
constexpr int HOP_W_TILE = HF * 64;            // bytes of one [64 x 32] fp16 tile
constexpr int HOP_GATHER_WARPS = 16;
constexpr int HOP_GATHER_THREADS = HOP_GATHER_WARPS * 32;      // 512: 32 groups of 16 lanes, 4 nodes per group per tile
constexpr int HOP_EPI_WARPS = 4;               // warps 16-19: one TMEM lane quarter each
constexpr int HOP_THREADS = HOP_GATHER_THREADS + HOP_EPI_WARPS * 32;        // 640 (20 warps: 96 registers each)
constexpr size_t HOP_W_IMAGE = 2 * 2 * (size_t)HOP_W_TILE;      // 2 chunks x (hi | lo) = 16 KB, followed by 4 floats (descale, pad)
constexpr size_t HOP_W_IMAGE_BYTES = HOP_W_IMAGE + 16;

constexpr int HOP_A_SLOTS = 1;                 // A-operand slots (64 KB each)

__global__ void hop_tc16_pack_kernel(const float* __restrict__ w, int ex, unsigned char* __restrict__ img) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < HF * HF; idx += gridDim.x * blockDim.x) {
        const int n = idx / HF, k = idx % HF;
        const float v = ldexpf(w[idx], ex);
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        const int kk = k % HOP_KC;
        const size_t off = (size_t)(k / HOP_KC) * 2 * HOP_W_TILE + sw64_piece_offset(n, kk >> 3) + (kk & 7) * 2;
        *reinterpret_cast<__half*>(img + off) = h;
        *reinterpret_cast<__half*>(img + off + HOP_W_TILE) = l;
    }
    if (blockIdx.x == 0 && threadIdx.x < 4)
        reinterpret_cast<float*>(img + HOP_W_IMAGE)[threadIdx.x] = threadIdx.x == 0 ? ldexpf(1.f, -ex) : 0.f;
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
    long long* trace;           // optional [3 roles][16 tiles][8 events] clock64 stamps of CTA 0 (profiling aid)
};

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src));
}

constexpr int HOP_EB = 3;       // edges of a node fetched in one batch (dual triangular meshes: in-degree <= 3 (+1 ghost edge))

template <bool WG, bool UP>
__device__ __forceinline__ float4 hop_term(const float4 oc, const float4 orow, const float4 sv) {
    if (WG) {                                          // (o[c] − o[r]) · s   (gnn.py:430-433)
        float4 d = make_float4(__fsub_rn(oc.x, orow.x), __fsub_rn(oc.y, orow.y), __fsub_rn(oc.z, orow.z), __fsub_rn(oc.w, orow.w));
        if (UP) { d.x = fmaxf(d.x, 0.f); d.y = fmaxf(d.y, 0.f); d.z = fmaxf(d.z, 0.f); d.w = fmaxf(d.w, 0.f); }
        return make_float4(__fmul_rn(d.x, sv.x), __fmul_rn(d.y, sv.y), __fmul_rn(d.z, sv.z), __fmul_rn(d.w, sv.w));
    }
    return make_float4(__fmul_rn(sv.x, orow.x), __fmul_rn(sv.y, orow.y), __fmul_rn(sv.z, orow.z), __fmul_rn(sv.w, orow.w));
}
__device__ __forceinline__ void hop_add(float4& a, const float4 t) {      // sequential, in edge order
    a.x = __fadd_rn(a.x, t.x); a.y = __fadd_rn(a.y, t.y); a.z = __fadd_rn(a.z, t.z); a.w = __fadd_rn(a.w, t.w);
}

// First HOP_EB edges of two nodes in flight together (up to 12 independent 16-byte loads + the two o[c]);
// further edges (rare) sequentially.  src ids come from `ids` (shared memory when staged, global otherwise),
// indexed by p - id_base.
template <bool WG, bool UP>
__device__ __forceinline__ void aggregate2(const float* __restrict__ o_src, const float* __restrict__ s,
                                           const int32_t* __restrict__ ids, int id_base,
                                           int pa0, int pa1, int pb0, int pb1, const float4 oca, const float4 ocb, int q4,
                                           float4& acca, float4& accb) {
    float4 oa[HOP_EB], sa[HOP_EB], ob[HOP_EB], sb[HOP_EB];
#pragma unroll
    for (int u = 0; u < HOP_EB; ++u) {
        if (pa0 + u < pa1) {
            oa[u] = ldg4(o_src + (long long)ids[pa0 + u - id_base] * HF + q4);
            sa[u] = ldg4_stream(s + (long long)(pa0 + u) * HF + q4);
        }
        if (pb0 + u < pb1) {
            ob[u] = ldg4(o_src + (long long)ids[pb0 + u - id_base] * HF + q4);
            sb[u] = ldg4_stream(s + (long long)(pb0 + u) * HF + q4);
        }
    }
    acca = make_float4(0.f, 0.f, 0.f, 0.f);
    accb = acca;
#pragma unroll
    for (int u = 0; u < HOP_EB; ++u) {
        if (pa0 + u < pa1) hop_add(acca, hop_term<WG, UP>(oca, oa[u], sa[u]));
        if (pb0 + u < pb1) hop_add(accb, hop_term<WG, UP>(ocb, ob[u], sb[u]));
    }
    for (int pp = pa0 + HOP_EB; pp < pa1; ++pp)
        hop_add(acca, hop_term<WG, UP>(oca, ldg4(o_src + (long long)ids[pp - id_base] * HF + q4), ldg4_stream(s + (long long)pp * HF + q4)));
    for (int pp = pb0 + HOP_EB; pp < pb1; ++pp)
        hop_add(accb, hop_term<WG, UP>(ocb, ldg4(o_src + (long long)ids[pp - id_base] * HF + q4), ldg4_stream(s + (long long)pp * HF + q4)));
}

// Persistent, one CTA per SM, warp-specialised:
//   warps 0-15  gather  : aggregate (every warp walks its own nodes, no CTA-wide barrier: the warps drift freely,
//                         which keeps the memory pipe uniformly busy), write the A operand; one tile later they
//                         also finish the output rows: D (from the shared-memory stage) + o[c] (+ addend) ->
//                         activation -> out, 16 lanes x 16 B per row, i.e. fully coalesced loads and stores
//   warps 16-19 epilogue: D (TMEM) -> shared-memory stage (thread = TMEM lane = node); lane 0 of warp 16 also
//                         issues the 24 tcgen05.mma of each tile (3xTF32); its commits release the A slot / publish D
// In steady state nothing but the gather warps' global loads is on the critical path.
constexpr int HOP_STAGE_LD = HF + 4;                                     // floats per staged row (272 B: conflict-free)
constexpr size_t HOP_STAGE_BYTES = (size_t)HOP_TILE * HOP_STAGE_LD * 4;   // 34,816 B

struct __align__(8) HopBarriers {
    uint64_t a_full[2], a_empty[2];    // A-operand slot written (512 arrivals) / consumed by the MMA (commit)
    uint64_t d_full[2], d_empty[2];    // accumulator slot complete (commit) / copied to the stage (128 arrivals)
    uint64_t st_full, st_empty;        // stage holds D of a tile (128 arrivals) / has been consumed (512 arrivals)
};

constexpr size_t HOP_TC_SMEM = 1024 + (size_t)HOP_A_SLOTS * HOP_A_SLOT + HOP_W_IMAGE + HOP_STAGE_BYTES + 2 * HOP_TILE * sizeof(float) +
                               sizeof(HopBarriers) + 16;

template <bool WG, bool UP, bool TRACE>
__global__ void __launch_bounds__(HOP_THREADS, 1) hop_tc16_kernel(const __grid_constant__ HopTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    // 1 KB alignment by OFFSETTING the shared array (integer arithmetic on the pointer value would turn every
    // later access into a generic LD/ST instead of LDS/STS)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_slots = smem;                                          // slot s: chunk c: [hi 16 KB | lo 16 KB]
    unsigned char* w_tile = smem + (size_t)HOP_A_SLOTS * HOP_A_SLOT;        // chunk c: [hi 8 KB | lo 8 KB]
    float* stage = reinterpret_cast<float*>(w_tile + HOP_W_IMAGE);          // [128][68] fp32
    float* s_inv = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(stage) + HOP_STAGE_BYTES);   // [tile parity][128] 2^-e of the row
    HopBarriers* bar = reinterpret_cast<HopBarriers*>(s_inv + 2 * HOP_TILE);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar->a_full[i], HOP_GATHER_THREADS); mbar_init(&bar->a_empty[i], 1);
            mbar_init(&bar->d_full[i], 1); mbar_init(&bar->d_empty[i], HOP_EPI_WARPS * 32);
        }
        mbar_init(&bar->st_full, HOP_EPI_WARPS * 32); mbar_init(&bar->st_empty, HOP_GATHER_THREADS);
        fence_barrier_init();
    }
    for (int i = threadIdx.x * 16; i < (int)HOP_W_IMAGE; i += HOP_THREADS * 16)
        *reinterpret_cast<float4*>(w_tile + i) = *reinterpret_cast<const float4*>(p.w_img + i);
    fence_proxy_async_smem();
    if (warp == HOP_GATHER_WARPS) tmem_alloc(tmem_holder, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_holder;
    const int n_tiles = (p.n_dst + HOP_TILE - 1) / HOP_TILE;
    // tiles are dealt round-robin: at any time the 148 CTAs work on one contiguous window of the mesh, so a row that
    // two tiles of different mesh lines both gather (ids ± one mesh line apart) is still in L2 when the second one
    // comes (a contiguous run per CTA re-read 0.26 GB per hop from DRAM: ncu r01d)
    const int n_my = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const bool tr0 = TRACE && p.trace != nullptr && blockIdx.x == 0 && lane == 0;
#define SWE_STAMP(role_, i_, ev_) do { if (TRACE && tr0 && (i_) < 16) p.trace[(role_) * 128 + (i_) * 8 + (ev_)] = clock64(); } while (0)

    if (warp < HOP_GATHER_WARPS) {
        // =====================================================================================
        // gather warps
        // =====================================================================================
        const int g = threadIdx.x >> 4, q = threadIdx.x & 15;      // 32 groups x 16 lanes; nodes g, g+32, g+64, g+96
        const int chunk = q >> 3, piece = q & 7, q4 = 4 * q;
        const bool trw = tr0 && warp == 0;
        const float slope = (p.act == SWE_ACT_PRELU && p.slope) ? __ldg(p.slope) : 0.f;
        // this lane's 4 columns = 8 bytes of fp16 at piece (q & 7) >> 1, half (q & 1) of the row's 64-byte chunk `chunk`
        const uint32_t a_sub = ((uint32_t)(q & 1)) * 8u;
        const uint32_t a_off0 = sw64_piece_offset(g, piece >> 1) + a_sub, a_off1 = sw64_piece_offset(g + 32, piece >> 1) + a_sub;
        const uint32_t a_off2 = sw64_piece_offset(g + 64, piece >> 1) + a_sub, a_off3 = sw64_piece_offset(g + 96, piece >> 1) + a_sub;
        const float w_descale = *reinterpret_cast<const float*>(p.w_img + HOP_W_IMAGE);
        // out rows of tile j: D (stage) + o[c] (+ addend) -> act -> store; all accesses are whole 256-byte rows
        auto finish_tile = [&](int j) {
            const int row0 = ((int)blockIdx.x + j * (int)gridDim.x) * HOP_TILE;
            const int rows = min(HOP_TILE, p.n_dst - row0);
            const long long base_row = (long long)p.dst_lo + row0;
            float4 oc[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = g + 32 * k;
                oc[k] = (p.o_dst && r < rows) ? ldg4(p.o_dst + (base_row + r) * HF + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.addend && r < rows) {
                    const float4 a4 = ldg4(p.addend + (base_row + r) * HF + q4);
                    oc[k].x += a4.x; oc[k].y += a4.y; oc[k].z += a4.z; oc[k].w += a4.w;
                }
            }
            mbar_wait(&bar->st_full, (uint32_t)j & 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = g + 32 * k;
                const float4 d = *reinterpret_cast<const float4*>(stage + r * HOP_STAGE_LD + q4);
                float4 rr = make_float4(oc[k].x + d.x, oc[k].y + d.y, oc[k].z + d.z, oc[k].w + d.w);
                if (p.act != SWE_ACT_NONE) {
                    rr.x = act_apply(p.act, rr.x, slope); rr.y = act_apply(p.act, rr.y, slope);
                    rr.z = act_apply(p.act, rr.z, slope); rr.w = act_apply(p.act, rr.w, slope);
                }
                if (r < rows) stg4(p.out + (base_row + r) * HF + q4, rr);
            }
            mbar_arrive(&bar->st_empty);
        };
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const int slot = i % HOP_A_SLOTS;
            const uint32_t u = (uint32_t)(i / HOP_A_SLOTS);
            const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * HOP_TILE;
            const int rows = min(HOP_TILE, p.n_dst - row0);
            const int32_t* rp = p.rowptr + row0;
            if (trw) SWE_STAMP(0, i, 0);
            unsigned char* base = a_slots + (size_t)slot * HOP_A_SLOT + (size_t)chunk * 2 * HOP_A_TILE;
            const float* orow_base = p.o_dst ? p.o_dst + ((long long)p.dst_lo + row0) * HF + q4 : nullptr;
#pragma unroll 1
            for (int jj = 0; jj < 2; ++jj) {
                const int ia = g + 64 * jj, ib = ia + 32;
                float4 oca = make_float4(0.f, 0.f, 0.f, 0.f), ocb = oca;
                const int pa0 = __ldg(rp + min(ia, rows)), pa1 = __ldg(rp + min(ia + 1, rows));
                const int pb0 = __ldg(rp + min(ib, rows)), pb1 = __ldg(rp + min(ib + 1, rows));
                if (orow_base) {
                    if (ia < rows) oca = ldg4(orow_base + ia * HF);
                    if (ib < rows) ocb = ldg4(orow_base + ib * HF);
                }
                float4 acca, accb;
                aggregate2<WG, UP>(p.o_src, p.s, p.src, 0, pa0, pa1, pb0, pb1, oca, ocb, q4, acca, accb);
                if (p.agg_out) {
                    if (ia < rows) stg4(p.agg_out + ((long long)p.dst_lo + row0 + ia) * HF + q4, acca);
                    if (ib < rows) stg4(p.agg_out + ((long long)p.dst_lo + row0 + ib) * HF + q4, accb);
                }
                if (jj == 0) {
                    mbar_wait(&bar->a_empty[slot], (u & 1) ^ 1);       // the MMA that last read this slot has completed
                    if (trw) SWE_STAMP(0, i, 1);
                }
                // per-row power-of-two scale (max |agg'| in [2^13, 2^14)): the row's 16 lanes agree on the maximum
                float ma = fmaxf(fmaxf(fabsf(acca.x), fabsf(acca.y)), fmaxf(fabsf(acca.z), fabsf(acca.w)));
                float mb = fmaxf(fmaxf(fabsf(accb.x), fabsf(accb.y)), fmaxf(fabsf(accb.z), fabsf(accb.w)));
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) {
                    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, off));
                    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, off));
                }
                uint32_t sba = 267u - (__float_as_uint(ma) >> 23), sbb = 267u - (__float_as_uint(mb) >> 23);
                sba = sba > 253u ? 253u : sba; sbb = sbb > 253u ? 253u : sbb;
                const float sca = __uint_as_float(sba << 23), scb = __uint_as_float(sbb << 23);
                if (q == 0) {
                    s_inv[(i & 1) * HOP_TILE + ia] = __uint_as_float((254u - sba) << 23) * w_descale;
                    s_inv[(i & 1) * HOP_TILE + ib] = __uint_as_float((254u - sbb) << 23) * w_descale;
                }
                const uint32_t offa = jj ? a_off2 : a_off0, offb = jj ? a_off3 : a_off1;
                uint2 hh, ll;
                split_f16x2(acca.x * sca, acca.y * sca, hh.x, ll.x); split_f16x2(acca.z * sca, acca.w * sca, hh.y, ll.y);
                *reinterpret_cast<uint2*>(base + offa) = hh;
                *reinterpret_cast<uint2*>(base + HOP_A_TILE + offa) = ll;
                split_f16x2(accb.x * scb, accb.y * scb, hh.x, ll.x); split_f16x2(accb.z * scb, accb.w * scb, hh.y, ll.y);
                *reinterpret_cast<uint2*>(base + offb) = hh;
                *reinterpret_cast<uint2*>(base + HOP_A_TILE + offb) = ll;
                if (trw) SWE_STAMP(0, i, 2 + jj);
                if (jj == 0 && i > 0) {                                  // the previous tile's D is in the stage by now
                    finish_tile(i - 1);
                    if (trw) SWE_STAMP(0, i, 4);
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(&bar->a_full[slot]);
        }
        if (n_my > 0) finish_tile(n_my - 1);
    } else {
        // =====================================================================================
        // epilogue warps: thread = one TMEM lane = one node; D -> stage
        // =====================================================================================
        const int lq = warp & 3;
        const bool trw = tr0 && lq == 0;
        const uint32_t idesc = make_idesc_f16(HOP_TILE, HF);
        const uint32_t a_u32 = smem_u32(a_slots), w_u32 = smem_u32(w_tile);
        float* my_row = stage + (lq * 32 + lane) * HOP_STAGE_LD;
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const int slot = i % HOP_A_SLOTS, dslot = i & 1;
            const uint32_t u = (uint32_t)(i / HOP_A_SLOTS), ud = (uint32_t)(i >> 1);
            if (warp == HOP_GATHER_WARPS) {
                if (lane == 0) {
                    SWE_STAMP(2, i, 0);
                    mbar_wait(&bar->d_empty[dslot], (ud & 1) ^ 1);     // accumulator of tile i-2 has been copied out
                    mbar_wait(&bar->a_full[slot], u & 1);
                    tc_fence_after_sync();
                    SWE_STAMP(2, i, 1);
                    const uint32_t d = tmem_base + dslot * 64;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint32_t a_hi = a_u32 + slot * HOP_A_SLOT + c * 2 * HOP_A_TILE, a_lo = a_hi + HOP_A_TILE;
                        const uint32_t w_hi = w_u32 + c * 2 * HOP_W_TILE, w_lo = w_hi + HOP_W_TILE;
#pragma unroll
                        for (int ks = 0; ks < HOP_KC / 16; ++ks) {
                            const uint64_t dah = make_desc_sw64(a_hi + ks * 32), dal = make_desc_sw64(a_lo + ks * 32);
                            const uint64_t dwh = make_desc_sw64(w_hi + ks * 32), dwl = make_desc_sw64(w_lo + ks * 32);
                            mma_f16_ss(d, dal, dwh, idesc, (c | ks) ? 1u : 0u);
                            mma_f16_ss(d, dah, dwl, idesc, 1u);
                            mma_f16_ss(d, dah, dwh, idesc, 1u);
                        }
                    }
                    mma_commit(&bar->a_empty[slot]);
                    mma_commit(&bar->d_full[dslot]);
                    SWE_STAMP(2, i, 2);
                }
                __syncwarp();
            }
            if (trw) SWE_STAMP(1, i, 0);
            mbar_wait(&bar->d_full[dslot], ud & 1);
            tc_fence_after_sync();
            mbar_wait(&bar->st_empty, ((uint32_t)i & 1) ^ 1);          // stage consumed by the gather warps (tile i-1)
            if (trw) SWE_STAMP(1, i, 1);
#pragma unroll 1
            const float inv = s_inv[(i & 1) * HOP_TILE + lq * 32 + lane];        // 2^-e of this row x the filter's 2^-f
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + dslot * 64 + hf * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(my_row + hf * 32 + j) =
                        make_float4(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv, __uint_as_float(v[j + 2]) * inv,
                                    __uint_as_float(v[j + 3]) * inv);
            }
            tc_fence_before_sync();
            mbar_arrive(&bar->d_empty[dslot]);
            mbar_arrive(&bar->st_full);
            if (trw) SWE_STAMP(1, i, 2);
        }
    }
#undef SWE_STAMP
    tc_fence_before_sync();
    __syncthreads();
    if (warp == HOP_GATHER_WARPS) tmem_dealloc(tmem_base, 128);
}


// =================================================================================================================
// s-ring edition (hop_tc16s_kernel): the gate rows s_p — 768 of the 1280 bytes a node costs — no longer travel as
// per-thread 16-byte loads behind the rowptr -> src -> row chain.  They are one contiguous run per block of nodes (CSR
// order), so every gather warp streams the run of ITS nodes into its own shared-memory buffers by cp.async.bulk, one
// tile ahead: the bytes are in flight whatever the warp is waiting for, they never pass through L1, and the registers
// the loads occupied are free.  Nothing couples the warps (the CTA-wide staging experiments lost exactly there):
//   * a warp owns 8 contiguous nodes of the tile, two passes of 4 nodes (half-warp = node pair a, b: 16 lanes x 16 B);
//   * per pass one buffer of SR_CAP edges (3 KB), refilled for the next tile as soon as the pass has read it;
//   * the warp's 9 rowptr values are fetched two tiles ahead, its src ids one tile ahead (cp.async, 4 B per lane),
//     so the only global round trip left on a pass's chain is the gather of the o rows;
//   * a pass whose rows exceed the staging (more than 12 edges on its four nodes, or more than 32 on the warp's eight:
//     never on a dual mesh) is not staged at all and runs one edge at a time from global memory; engine.py keeps edge
//     sets like that on hop_tc_kernel (plan.EdgeSet.max_block4);
//   * the four epilogue warps, idle for most of a tile, prefetch.global.L2 the rows tile i+3 will touch.
// Same arithmetic, same order: bit-identical to hop_tc16_kernel.
// =================================================================================================================
constexpr int SR_NODES = 8;                    // nodes per warp and tile
constexpr int SR_CAP = 12;                     // edges per pass buffer
constexpr int SR_BUF = SR_CAP * HF * 4;        // 3072 B
constexpr int SR_IDS = 32;                     // src ids staged per warp and tile
#ifndef SR_EARLY_FINISH
#define SR_EARLY_FINISH 0                      // 1: o[c] of the previous tile's output rows loaded together with pass 0's gathers (measured slower: spills)
#endif
struct __align__(16) SrWarp {
    uint64_t full[2];                          // pass buffer landed (expect_tx by lane 0)
    int32_t ptr[3][12];                        // rowptr[first node .. first node + 8] of tiles i, i+1, i+2 (mod 3)
    int32_t ids[2][SR_IDS];                    // src ids of the warp's edges, tile parity
};
constexpr size_t HOP_TC16S_SMEM = 1024 + (size_t)HOP_A_SLOT + HOP_W_IMAGE + HOP_STAGE_BYTES + 2 * HOP_TILE * sizeof(float) +
                                  (size_t)HOP_GATHER_WARPS * 2 * SR_BUF + HOP_GATHER_WARPS * sizeof(SrWarp) + sizeof(HopBarriers) + 16;
static_assert(HOP_GATHER_WARPS * SR_NODES == HOP_TILE, "8 nodes per gather warp");

#ifndef SR_PREFETCH
#define SR_PREFETCH 1
#endif
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

struct HopTc16sParams {
    HopTcParams h;
    unsigned zero_mask;                        // 0 (a value the compiler cannot fold: see the buffer release)
    int pf_dist, pf_l1;                        // prefetch distance in tiles (0: off), 1: into L1 instead of L2
};

__device__ __forceinline__ void bulk_g2s_u32(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <bool WG, bool UP, bool TRACE>
__global__ void __launch_bounds__(HOP_THREADS, 1) hop_tc16s_kernel(const __grid_constant__ HopTc16sParams pp) {
    const HopTcParams& p = pp.h;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_slots = smem;                                          // chunk c: [hi 8 KB | lo 8 KB]
    unsigned char* w_tile = smem + (size_t)HOP_A_SLOT;
    float* stage = reinterpret_cast<float*>(w_tile + HOP_W_IMAGE);          // [128][68] fp32
    float* s_inv = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(stage) + HOP_STAGE_BYTES);
    unsigned char* sring = reinterpret_cast<unsigned char*>(s_inv + 2 * HOP_TILE);          // [warp][pass][SR_BUF]
    SrWarp* ctl = reinterpret_cast<SrWarp*>(sring + (size_t)HOP_GATHER_WARPS * 2 * SR_BUF);
    HopBarriers* bar = reinterpret_cast<HopBarriers*>(ctl + HOP_GATHER_WARPS);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar->a_full[i], HOP_GATHER_THREADS); mbar_init(&bar->a_empty[i], 1);
            mbar_init(&bar->d_full[i], 1); mbar_init(&bar->d_empty[i], HOP_EPI_WARPS * 32);
        }
        mbar_init(&bar->st_full, HOP_EPI_WARPS * 32); mbar_init(&bar->st_empty, HOP_GATHER_THREADS);
        for (int w = 0; w < HOP_GATHER_WARPS; ++w) { mbar_init(&ctl[w].full[0], 1); mbar_init(&ctl[w].full[1], 1); }
        fence_barrier_init();
    }
    for (int i = threadIdx.x * 16; i < (int)HOP_W_IMAGE; i += HOP_THREADS * 16)
        *reinterpret_cast<float4*>(w_tile + i) = *reinterpret_cast<const float4*>(p.w_img + i);
    fence_proxy_async_smem();
    if (warp == HOP_GATHER_WARPS) tmem_alloc(tmem_holder, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_holder;
    const int n_tiles = (p.n_dst + HOP_TILE - 1) / HOP_TILE;
    const int n_my = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // round-robin tiles (see hop_tc16_kernel)

    if (warp < HOP_GATHER_WARPS) {
        // =====================================================================================
        // gather warps
        // =====================================================================================
        const int g = threadIdx.x >> 4, q = threadIdx.x & 15, h = lane >> 4;   // g: output-row group of finish_tile
        const int chunk = q >> 3, piece = q & 7, q4 = 4 * q;
        const float slope = (p.act == SWE_ACT_PRELU && p.slope) ? __ldg(p.slope) : 0.f;
        const uint32_t a_sub = ((uint32_t)(q & 1)) * 8u;
        const float w_descale = *reinterpret_cast<const float*>(p.w_img + HOP_W_IMAGE);
        SrWarp* cw = ctl + warp;
        unsigned char* sbuf = sring + (size_t)warp * 2 * SR_BUF;
        unsigned char* a_base = a_slots + (size_t)chunk * 2 * HOP_A_TILE;
        // out rows of tile j: D (stage) + o[c] (+ addend) -> act -> store; whole 256-byte rows.  Two halves: the loads go out
        // together with pass 0's gathers of the next tile (one round trip less on the warp's chain), the rest follows
        // pass 0's operand write
        float4 oc[4];
        auto finish_load = [&](int j) {
            const int row0 = ((int)blockIdx.x + j * (int)gridDim.x) * HOP_TILE;
            const int rows = min(HOP_TILE, p.n_dst - row0);
            const long long base_row = (long long)p.dst_lo + row0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = g + 32 * k;
                oc[k] = (p.o_dst && r < rows) ? ldg4(p.o_dst + (base_row + r) * HF + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto finish_store = [&](int j) {
            const int row0 = ((int)blockIdx.x + j * (int)gridDim.x) * HOP_TILE;
            const int rows = min(HOP_TILE, p.n_dst - row0);
            const long long base_row = (long long)p.dst_lo + row0;
            if (p.addend) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int r = g + 32 * k;
                    if (r < rows) {
                        const float4 a4 = ldg4(p.addend + (base_row + r) * HF + q4);
                        oc[k].x += a4.x; oc[k].y += a4.y; oc[k].z += a4.z; oc[k].w += a4.w;
                    }
                }
            }
            mbar_wait(&bar->st_full, (uint32_t)j & 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int r = g + 32 * k;
                const float4 d = *reinterpret_cast<const float4*>(stage + r * HOP_STAGE_LD + q4);
                float4 rr = make_float4(oc[k].x + d.x, oc[k].y + d.y, oc[k].z + d.z, oc[k].w + d.w);
                if (p.act != SWE_ACT_NONE) {
                    rr.x = act_apply(p.act, rr.x, slope); rr.y = act_apply(p.act, rr.y, slope);
                    rr.z = act_apply(p.act, rr.z, slope); rr.w = act_apply(p.act, rr.w, slope);
                }
                if (r < rows) stg4(p.out + (base_row + r) * HF + q4, rr);
            }
            mbar_arrive(&bar->st_empty);
        };
        // rowptr[first node + min(lane, 8)] of tile i (0 beyond this CTA's tiles: an empty block)
        auto load_ptr = [&](int i) -> int32_t {
            if (i >= n_my) return 0;
            const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * HOP_TILE;
            return __ldg(p.rowptr + min(row0 + SR_NODES * warp + min(lane, SR_NODES), p.n_dst));
        };
        // s rows of pass jj of tile i -> the pass buffer (lane 0); `zero` = 0, made of the values last read from it
        auto issue_s = [&](int i, int jj, uint32_t zero) {
            const int32_t* pt = cw->ptr[i % 3];
            const int P0 = pt[4 * jj], cnt = pt[4 * jj + 4] - P0;
            if (cnt > 0 && cnt <= SR_CAP && pt[SR_NODES] - pt[0] <= SR_IDS && lane == 0) {
                mbar_arrive_expect_tx(&cw->full[jj], (uint32_t)cnt * (HF * 4));
                bulk_g2s_u32(smem_u32(sbuf + (size_t)jj * SR_BUF) + zero, p.s + (long long)P0 * HF, (uint32_t)cnt * (HF * 4), &cw->full[jj]);
            }
        };
        auto issue_ids = [&](int i) {
            const int32_t* pt = cw->ptr[i % 3];
            const int PT0 = pt[0];
            if (PT0 + lane < pt[SR_NODES]) cp_async4(&cw->ids[i & 1][lane], p.src + PT0 + lane);
        };
        if (n_my > 0) {
            const int32_t r0 = load_ptr(0), r1 = load_ptr(1);
            if (lane <= SR_NODES) { cw->ptr[0][lane] = r0; cw->ptr[1][lane] = r1; }
            __syncwarp();
            issue_ids(0);
            issue_s(0, 0, 0u);
            issue_s(0, 1, 0u);
        }
        uint32_t ph = 0;                                                   // bit jj: parity of the pass buffer's next phase
        const bool trw = TRACE && p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
#define SR_STAMP(i_, ev_) do { if (TRACE && trw && (i_) < 16) p.trace[(i_) * 16 + (ev_)] = clock64(); } while (0)
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * HOP_TILE;
            const int rows = min(HOP_TILE, p.n_dst - row0);
            SR_STAMP(i, 0);
            cp_async_wait_all();                                           // this lane's ids of tile i and rowptr value of tile i+1
            __syncwarp();                                                  // ... and everybody else's
            SR_STAMP(i, 1);
            issue_ids(i + 1);
            // rowptr of tile i+2 straight into shared memory (no register to spill, nothing waits for it before the next tile)
            if (lane <= SR_NODES) {
                if (i + 2 < n_my)
                    cp_async4(&cw->ptr[(i + 2) % 3][lane],
                              p.rowptr + min(((int)blockIdx.x + (i + 2) * (int)gridDim.x) * HOP_TILE + SR_NODES * warp + lane, p.n_dst));
                else
                    cw->ptr[(i + 2) % 3][lane] = 0;
            }
            const int32_t* pt = cw->ptr[i % 3];
            const int32_t* idb = cw->ids[i & 1];
            const int PT0 = pt[0];
            const float* orow_base = p.o_dst ? p.o_dst + ((long long)p.dst_lo + row0) * HF + q4 : nullptr;
#pragma unroll 1
            for (int jj = 0; jj < 2; ++jj) {
                const int ia = SR_NODES * warp + 4 * jj + h, ib = ia + 2;  // rows of the tile
                const int P0 = pt[4 * jj], cnt = pt[4 * jj + 4] - P0;
                const int pa0 = pt[4 * jj + h], pa1 = pt[4 * jj + h + 1], pb0 = pt[4 * jj + h + 2], pb1 = pt[4 * jj + h + 3];
                // warp-uniform: the pass's s rows fit its buffer and the warp's ids the staging (always, on a dual mesh);
                // otherwise nothing was staged for this pass and it runs on global loads like hop_tc16_kernel
                const bool staged = cnt <= SR_CAP && pt[SR_NODES] - PT0 <= SR_IDS;
                float4 acca, accb;
                if (staged) {
                    // every shared-memory read of the pass BEFORE its first global load: a scoreboard is a counter, so an
                    // LDS issued behind an outstanding LDG that shares its scoreboard returns with it — the six id -> row
                    // pairs became six serial round trips (3 k cycles per pass in the clock64 trace)
                    int32_t ida[HOP_EB], idb2[HOP_EB];
#pragma unroll
                    for (int u = 0; u < HOP_EB; ++u) {
                        ida[u] = (pa0 + u < pa1) ? idb[pa0 + u - PT0] : 0;
                        idb2[u] = (pb0 + u < pb1) ? idb[pb0 + u - PT0] : 0;
                    }
                    float4 oca = make_float4(0.f, 0.f, 0.f, 0.f), ocb = oca;
                    if (orow_base) {
                        if (ia < rows) oca = ldg4(orow_base + ia * HF);
                        if (ib < rows) ocb = ldg4(orow_base + ib * HF);
                    }
                    float4 oa[HOP_EB], ob[HOP_EB];
#pragma unroll
                    for (int u = 0; u < HOP_EB; ++u) {
                        if (pa0 + u < pa1) oa[u] = ldg4(p.o_src + (long long)ida[u] * HF + q4);
                        if (pb0 + u < pb1) ob[u] = ldg4(p.o_src + (long long)idb2[u] * HF + q4);
                    }
                    if (SR_EARLY_FINISH && jj == 0 && i > 0) finish_load(i - 1);
                    SR_STAMP(i, jj ? 7 : 2);
                    if (cnt > 0) { mbar_wait(&cw->full[jj], (ph >> jj) & 1u); ph ^= 1u << jj; }
                    SR_STAMP(i, jj ? 8 : 3);
                    const unsigned char* sa = sbuf + (size_t)jj * SR_BUF + q * 16 + (pa0 - P0) * (HF * 4);
                    const unsigned char* sb = sbuf + (size_t)jj * SR_BUF + q * 16 + (pb0 - P0) * (HF * 4);
                    acca = make_float4(0.f, 0.f, 0.f, 0.f);
                    accb = acca;
#pragma unroll
                    for (int u = 0; u < HOP_EB; ++u) {
                        if (pa0 + u < pa1) hop_add(acca, hop_term<WG, UP>(oca, oa[u], *reinterpret_cast<const float4*>(sa + u * (HF * 4))));
                        if (pb0 + u < pb1) hop_add(accb, hop_term<WG, UP>(ocb, ob[u], *reinterpret_cast<const float4*>(sb + u * (HF * 4))));
                    }
                    for (int e = pa0 + HOP_EB; e < pa1; ++e)
                        hop_add(acca, hop_term<WG, UP>(oca, ldg4(p.o_src + (long long)idb[e - PT0] * HF + q4),
                                                       *reinterpret_cast<const float4*>(sa + (e - pa0) * (HF * 4))));
                    for (int e = pb0 + HOP_EB; e < pb1; ++e)
                        hop_add(accb, hop_term<WG, UP>(ocb, ldg4(p.o_src + (long long)idb[e - PT0] * HF + q4),
                                                       *reinterpret_cast<const float4*>(sb + (e - pb0) * (HF * 4))));
                } else {
                    // (rare: hubs.  One edge at a time — correct, few registers, no attempt at speed; graphs whose in-degrees
                    //  exceed the staging as a rule belong on hop_tc16_kernel / hop_tc_kernel)
                    float4 oca = make_float4(0.f, 0.f, 0.f, 0.f), ocb = oca;
                    if (orow_base) {
                        if (ia < rows) oca = ldg4(orow_base + ia * HF);
                        if (ib < rows) ocb = ldg4(orow_base + ib * HF);
                    }
                    acca = make_float4(0.f, 0.f, 0.f, 0.f);
                    accb = acca;
#pragma unroll 1
                    for (int e = pa0; e < pa1; ++e)
                        hop_add(acca, hop_term<WG, UP>(oca, ldg4(p.o_src + (long long)__ldg(p.src + e) * HF + q4), ldg4_stream(p.s + (long long)e * HF + q4)));
#pragma unroll 1
                    for (int e = pb0; e < pb1; ++e)
                        hop_add(accb, hop_term<WG, UP>(ocb, ldg4(p.o_src + (long long)__ldg(p.src + e) * HF + q4), ldg4_stream(p.s + (long long)e * HF + q4)));
                }
                // The buffer goes back to the bulk copy only when every lane's reads have RETURNED: the copy's address is
                // made of the sums (an instruction issued right behind the LDS could overtake them, swe_rowmlp_tc16.cu)
                {
                    uint32_t dep = __float_as_uint(acca.x) | __float_as_uint(acca.y) | __float_as_uint(acca.z) | __float_as_uint(acca.w) |
                                   __float_as_uint(accb.x) | __float_as_uint(accb.y) | __float_as_uint(accb.z) | __float_as_uint(accb.w);
                    uint32_t zero;
                    asm volatile("and.b32 %0, %1, %2;" : "=r"(zero) : "r"(dep), "r"(pp.zero_mask));
                    __syncwarp();
                    SR_STAMP(i, jj ? 9 : 4);
                    issue_s(i + 1, jj, zero);
                }
                if (p.agg_out) {
                    if (ia < rows) stg4(p.agg_out + ((long long)p.dst_lo + row0 + ia) * HF + q4, acca);
                    if (ib < rows) stg4(p.agg_out + ((long long)p.dst_lo + row0 + ib) * HF + q4, accb);
                }
                if (jj == 0) { mbar_wait(&bar->a_empty[0], ((uint32_t)i & 1) ^ 1); SR_STAMP(i, 5); }   // the MMAs of the previous tile have read the slot
                float ma = fmaxf(fmaxf(fabsf(acca.x), fabsf(acca.y)), fmaxf(fabsf(acca.z), fabsf(acca.w)));
                float mb = fmaxf(fmaxf(fabsf(accb.x), fabsf(accb.y)), fmaxf(fabsf(accb.z), fabsf(accb.w)));
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) {
                    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, off));
                    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, off));
                }
                uint32_t sba = 267u - (__float_as_uint(ma) >> 23), sbb = 267u - (__float_as_uint(mb) >> 23);
                sba = sba > 253u ? 253u : sba; sbb = sbb > 253u ? 253u : sbb;
                const float sca = __uint_as_float(sba << 23), scb = __uint_as_float(sbb << 23);
                if (q == 0) {
                    s_inv[(i & 1) * HOP_TILE + ia] = __uint_as_float((254u - sba) << 23) * w_descale;
                    s_inv[(i & 1) * HOP_TILE + ib] = __uint_as_float((254u - sbb) << 23) * w_descale;
                }
                const uint32_t offa = sw64_piece_offset((uint32_t)ia, (uint32_t)piece >> 1) + a_sub;
                const uint32_t offb = sw64_piece_offset((uint32_t)ib, (uint32_t)piece >> 1) + a_sub;
                uint2 hh, ll;
                split_f16x2(acca.x * sca, acca.y * sca, hh.x, ll.x); split_f16x2(acca.z * sca, acca.w * sca, hh.y, ll.y);
                *reinterpret_cast<uint2*>(a_base + offa) = hh;
                *reinterpret_cast<uint2*>(a_base + HOP_A_TILE + offa) = ll;
                split_f16x2(accb.x * scb, accb.y * scb, hh.x, ll.x); split_f16x2(accb.z * scb, accb.w * scb, hh.y, ll.y);
                *reinterpret_cast<uint2*>(a_base + offb) = hh;
                *reinterpret_cast<uint2*>(a_base + HOP_A_TILE + offb) = ll;
                if (jj == 0 && i > 0) { if (!SR_EARLY_FINISH) finish_load(i - 1); finish_store(i - 1); SR_STAMP(i, 6); }   // the previous tile's D is in the stage by now
            }
            fence_proxy_async_smem();
            mbar_arrive(&bar->a_full[0]);
            SR_STAMP(i, 10);
        }
        if (n_my > 0) { finish_load(n_my - 1); finish_store(n_my - 1); }
#undef SR_STAMP
    } else {
        // =====================================================================================
        // epilogue warps: thread = one TMEM lane = one node; D -> stage (as in hop_tc16_kernel)
        // =====================================================================================
        const int lq = warp & 3;
        const uint32_t idesc = make_idesc_f16(HOP_TILE, HF);
        const uint32_t a_u32 = smem_u32(a_slots), w_u32 = smem_u32(w_tile);
        float* my_row = stage + (lq * 32 + lane) * HOP_STAGE_LD;
        // These warps sleep for most of a tile: they pull the rows a later tile will gather (its own o[c] rows, its sources,
        // its addend rows) into L2 — every pass of the gather warps waited for at least one first-touch row from DRAM
        auto prefetch_tile = [&](int t) {
            if (t >= n_my) return;
            const int row0n = ((int)blockIdx.x + t * (int)gridDim.x) * HOP_TILE;
            const int rowsn = min(HOP_TILE, p.n_dst - row0n);
            const int tid = lq * 32 + lane;
            if (tid < rowsn) {
                if (p.o_dst) {
                    const float* r = p.o_dst + ((long long)p.dst_lo + row0n + tid) * HF;
                    if (pp.pf_l1) { prefetch_l1(r); prefetch_l1(r + 32); } else { prefetch_l2(r); prefetch_l2(r + 32); }
                }
                if (p.addend) {
                    const float* r = p.addend + ((long long)p.dst_lo + row0n + tid) * HF;
                    prefetch_l2(r); prefetch_l2(r + 32);
                }
            }
            const int e0 = __ldg(p.rowptr + row0n), e1 = __ldg(p.rowptr + row0n + rowsn);
            for (int e = e0 + tid; e < e1; e += HOP_EPI_WARPS * 32) {
                const float* r = p.o_src + (long long)__ldg(p.src + e) * HF;
                if (pp.pf_l1) { prefetch_l1(r); prefetch_l1(r + 32); } else { prefetch_l2(r); prefetch_l2(r + 32); }
            }
        };
#if SR_PREFETCH
        for (int t = 1; t < pp.pf_dist; ++t) prefetch_tile(t);           // (tile 0 is being gathered already; the loop takes over at pf_dist)
#endif
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const int dslot = i & 1;
            const uint32_t ud = (uint32_t)(i >> 1);
            if (warp == HOP_GATHER_WARPS) {
                if (lane == 0) {
                    mbar_wait(&bar->d_empty[dslot], (ud & 1) ^ 1);     // accumulator of tile i-2 has been copied out
                    mbar_wait(&bar->a_full[0], (uint32_t)i & 1);
                    tc_fence_after_sync();
                    const uint32_t d = tmem_base + dslot * 64;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const uint32_t a_hi = a_u32 + c * 2 * HOP_A_TILE, a_lo = a_hi + HOP_A_TILE;
                        const uint32_t w_hi = w_u32 + c * 2 * HOP_W_TILE, w_lo = w_hi + HOP_W_TILE;
#pragma unroll
                        for (int ks = 0; ks < HOP_KC / 16; ++ks) {
                            const uint64_t dah = make_desc_sw64(a_hi + ks * 32), dal = make_desc_sw64(a_lo + ks * 32);
                            const uint64_t dwh = make_desc_sw64(w_hi + ks * 32), dwl = make_desc_sw64(w_lo + ks * 32);
                            mma_f16_ss(d, dal, dwh, idesc, (c | ks) ? 1u : 0u);
                            mma_f16_ss(d, dah, dwl, idesc, 1u);
                            mma_f16_ss(d, dah, dwh, idesc, 1u);
                        }
                    }
                    mma_commit(&bar->a_empty[0]);
                    mma_commit(&bar->d_full[dslot]);
                }
                __syncwarp();
            }
            mbar_wait(&bar->d_full[dslot], ud & 1);
            tc_fence_after_sync();
            mbar_wait(&bar->st_empty, ((uint32_t)i & 1) ^ 1);          // stage consumed by the gather warps (tile i-1)
            const float inv = s_inv[(i & 1) * HOP_TILE + lq * 32 + lane];        // 2^-e of this row x the filter's 2^-f
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + dslot * 64 + hf * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(my_row + hf * 32 + j) =
                        make_float4(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv, __uint_as_float(v[j + 2]) * inv,
                                    __uint_as_float(v[j + 3]) * inv);
            }
            tc_fence_before_sync();
            mbar_arrive(&bar->d_empty[dslot]);
            mbar_arrive(&bar->st_full);
#if SR_PREFETCH
            if (pp.pf_dist > 0) prefetch_tile(i + pp.pf_dist);
#endif
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == HOP_GATHER_WARPS) tmem_dealloc(tmem_base, 128);
}

}  // namespace tch
}  // namespace swe

using namespace swe;

extern "C" size_t swe_hop_tc16_image_bytes(void) { return tch::HOP_W_IMAGE_BYTES; }

// wmax = max |w| (host value): the matrix is scaled by the power of two that puts it in [2^13, 2^14)
extern "C" int swe_hop_tc16_pack(const float* w, float wmax, void* image, void* stream) {
    SWE_REQUIRE(w && image && aligned16(image), SWE_E_INVAL, "hop_tc16_pack: bad arguments");
    int e = 0;
    if (wmax > 0.f && wmax < 3.0e38f) { (void)frexpf(wmax, &e); e = 14 - e; }
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    tch::hop_tc16_pack_kernel<<<16, 256, 0, (cudaStream_t)stream>>>(w, e, (unsigned char*)image);
    return check_launch("hop_tc16_pack");
}

extern "C" int swe_propagate_hop_tc16_fwd(const float* o_src, const float* o_dst, const float* s, const int32_t* rowptr,
                                          const int32_t* src, int32_t dst_lo, int32_t n_dst, const void* w_image,
                                          int32_t with_gradient, int32_t upwind, const float* addend, int32_t act,
                                          const float* slope, float* agg_out, float* out, void* stream) {
    SWE_REQUIRE(o_src && s && rowptr && src && out && w_image && dst_lo >= 0 && n_dst >= 0, SWE_E_INVAL, "hop_tc16: bad arguments");
    SWE_REQUIRE(!(with_gradient && !o_dst), SWE_E_INVAL, "hop_tc16: with_gradient needs the destination rows");
    SWE_REQUIRE(aligned16(o_src) && aligned16(s) && aligned16(out) && aligned16(w_image) && (!o_dst || aligned16(o_dst)) &&
                (!addend || aligned16(addend)) && (!agg_out || aligned16(agg_out)), SWE_E_ALIGN, "hop_tc16: unaligned buffer");
    SWE_REQUIRE(out != o_src && out != o_dst, SWE_E_INVAL, "hop_tc16: output must not alias the hop input");
    if (n_dst == 0) return 0;
    tch::HopTcParams p;
    p.o_src = o_src; p.o_dst = o_dst; p.s = s; p.rowptr = rowptr; p.src = src; p.dst_lo = dst_lo; p.n_dst = n_dst;
    p.w_img = (const unsigned char*)w_image; p.with_gradient = with_gradient; p.upwind = upwind; p.addend = addend;
    p.act = act; p.slope = slope; p.out = out; p.agg_out = agg_out; p.trace = nullptr;
    void (*kern)(const tch::HopTcParams) = with_gradient ? (upwind ? tch::hop_tc16_kernel<true, true, false> : tch::hop_tc16_kernel<true, false, false>)
                                                          : tch::hop_tc16_kernel<false, false, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tch::HOP_TC_SMEM);
    if (e != cudaSuccess) { set_error("hop_tc16 smem opt-in (%zu B): %s", tch::HOP_TC_SMEM, cudaGetErrorString(e)); return (int)e; }
    const int n_tiles = (n_dst + tch::HOP_TILE - 1) / tch::HOP_TILE;
    kern<<<grid_for(n_tiles, 1), tch::HOP_THREADS, tch::HOP_TC_SMEM, (cudaStream_t)stream>>>(p);
    return check_launch("propagate_hop_tc16_fwd");
}

static long long* g_tc16s_trace = nullptr;     // profiling aid: consumed by the next launch (tools/bench_hop.py)
extern "C" void swe_hop_tc16s_set_trace(long long* t) { g_tc16s_trace = t; }

// s-ring edition (hop_tc16s_kernel): same contract and image as swe_propagate_hop_tc16_fwd, bit-identical results
extern "C" int swe_propagate_hop_tc16s_fwd(const float* o_src, const float* o_dst, const float* s, const int32_t* rowptr,
                                           const int32_t* src, int32_t dst_lo, int32_t n_dst, const void* w_image,
                                           int32_t with_gradient, int32_t upwind, const float* addend, int32_t act,
                                           const float* slope, float* agg_out, float* out, void* stream) {
    SWE_REQUIRE(o_src && s && rowptr && src && out && w_image && dst_lo >= 0 && n_dst >= 0, SWE_E_INVAL, "hop_tc16s: bad arguments");
    SWE_REQUIRE(!(with_gradient && !o_dst), SWE_E_INVAL, "hop_tc16s: with_gradient needs the destination rows");
    SWE_REQUIRE(aligned16(o_src) && aligned16(s) && aligned16(out) && aligned16(w_image) && (!o_dst || aligned16(o_dst)) &&
                (!addend || aligned16(addend)) && (!agg_out || aligned16(agg_out)), SWE_E_ALIGN, "hop_tc16s: unaligned buffer");
    SWE_REQUIRE(out != o_src && out != o_dst, SWE_E_INVAL, "hop_tc16s: output must not alias the hop input");
    if (n_dst == 0) return 0;
    tch::HopTc16sParams pp;
    tch::HopTcParams& p = pp.h;
    p.o_src = o_src; p.o_dst = o_dst; p.s = s; p.rowptr = rowptr; p.src = src; p.dst_lo = dst_lo; p.n_dst = n_dst;
    p.w_img = (const unsigned char*)w_image; p.with_gradient = with_gradient; p.upwind = upwind; p.addend = addend;
    p.act = act; p.slope = slope; p.out = out; p.agg_out = agg_out;
    p.trace = g_tc16s_trace; g_tc16s_trace = nullptr;
    pp.zero_mask = 0u;
    { static int pd = -1, pl = 0; if (pd < 0) { const char* e1 = getenv("MSWE_HOP_PF_DIST"); pd = e1 ? atoi(e1) : 3; const char* e2 = getenv("MSWE_HOP_PF_L1"); pl = e2 ? atoi(e2) : 0; }
      pp.pf_dist = pd; pp.pf_l1 = pl; }
    void (*kern)(const tch::HopTc16sParams) = with_gradient ? (upwind ? tch::hop_tc16s_kernel<true, true, false> : tch::hop_tc16s_kernel<true, false, false>)
                                                             : tch::hop_tc16s_kernel<false, false, false>;
    if (p.trace && with_gradient && !upwind) kern = tch::hop_tc16s_kernel<true, false, true>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tch::HOP_TC16S_SMEM);
    if (e != cudaSuccess) { set_error("hop_tc16s smem opt-in (%zu B): %s", tch::HOP_TC16S_SMEM, cudaGetErrorString(e)); return (int)e; }
    const int n_tiles = (n_dst + tch::HOP_TILE - 1) / tch::HOP_TILE;
    kern<<<grid_for(n_tiles, 1), tch::HOP_THREADS, tch::HOP_TC16S_SMEM, (cudaStream_t)stream>>>(pp);
    return check_launch("propagate_hop_tc16s_fwd");
}
