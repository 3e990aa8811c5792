// Row linear  out[r] = x[r] · Wᵀ  (F = 64, no bias, no activation) as a streaming kernel: the first step of every SWEGNN
// call, o_0 = x_d W_0ᵀ (reference models/gnn.py:401-402), on rows [row_lo, row_lo + n_rows).
//
// swe_row_mlp_tc.cu runs this shape as a one-tile-at-a-time pipeline whose hand-overs (row warps -> tensor core -> four
// epilogue warps -> stage -> row warps) take ~5 k cycles per 128 rows against ~0.7 k cycles of HBM time.  Here the input
// never passes through registers on its way in: a loader thread streams whole 128-row tiles (32 KB, contiguous in global
// memory) into a 3-deep shared-memory ring with cp.async.bulk + mbarrier, so three tiles of loads are always in flight per
// SM; eight converter warps turn a landed tile into the fp16 hi/lo A operand (per-row power-of-two scale, as the fp16 hop
// and gate do; two operand slots), one warp issues the 12 tcgen05.mma (kind::f16, 3 products, fp32 accumulation, two
// accumulator slots in TMEM), four epilogue warps move D to a padded stage, and the converter warps write the stage out
// as whole rows one tile later.  Weight image: swe_hop_tc16_pack (same 64 x 64 filter layout and per-matrix scale).
#include <stdlib.h>
#include "swe_tc.cuh"

namespace swe {
namespace rowlin {
using namespace swe::tc;

constexpr int F = 64, TILE = 128, KC = 32;
constexpr int A_TILE = TILE * 64;               // [128 x 32] fp16 = 8 KB
constexpr int A_SLOT = 4 * A_TILE;              // 2 chunks x (hi | lo) = 32 KB
constexpr int W_TILE = F * 64;                  // [64 x 32] fp16 = 4 KB
constexpr size_t W_IMAGE = 4 * (size_t)W_TILE;  // 16 KB (+ 16 B: descale) — the layout of swe_hop_tc16_pack
constexpr int IN_STAGES = 3;
constexpr size_t IN_TILE = (size_t)TILE * F * 4;             // 32 KB of fp32 rows, as they lie in global memory
constexpr int STAGE_LD = F + 4;
constexpr size_t STAGE_BYTES = (size_t)TILE * STAGE_LD * 4;
constexpr int CONV_WARPS = 8, CONV_THREADS = 256, EPI_WARPS = 4;
constexpr int THREADS = CONV_THREADS + EPI_WARPS * 32 + 64;  // + MMA issuer warp + loader warp = 448

struct __align__(8) Bar {
    uint64_t in_full[IN_STAGES], in_empty[IN_STAGES];   // tile landed (tx bytes) / converted (256 arrivals)
    uint64_t a_full[2], a_empty[2];                     // operand slot written (256) / consumed (commit)
    uint64_t d_full[2], d_empty[2];                     // accumulator slot complete (commit) / staged (128)
    uint64_t st_full, st_empty;                         // stage holds a tile (128) / written out (256)
};

constexpr size_t SMEM = 1024 + IN_STAGES * IN_TILE + 2 * (size_t)A_SLOT + W_IMAGE + STAGE_BYTES + 2 * TILE * 4 + sizeof(Bar) + 16;

struct Params {
    const float* x; float* out;
    uint32_t zero_mask;                      // 0 at run time, opaque to the compiler (see the ring-slot release)
    long long row_lo, n_rows;
    const unsigned char* w_img;
};

__global__ void __launch_bounds__(THREADS, 1) row_linear_tc16_kernel(const __grid_constant__ Params p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* a_slots = smem;                                       // slot s: chunk c: [hi 8 KB | lo 8 KB]
    unsigned char* w_tile = a_slots + 2 * (size_t)A_SLOT;
    float* in_ring = reinterpret_cast<float*>(w_tile + W_IMAGE);          // [IN_STAGES][128][64]
    float* stage = in_ring + IN_STAGES * (IN_TILE / 4);                   // [128][68]
    float* s_inv = stage + TILE * STAGE_LD;                               // [2][128] row 2^-e x filter 2^-f
    Bar* bar = reinterpret_cast<Bar*>(s_inv + 2 * TILE);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < IN_STAGES; ++i) { mbar_init(&bar->in_full[i], 1); mbar_init(&bar->in_empty[i], CONV_THREADS); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar->a_full[i], CONV_THREADS); mbar_init(&bar->a_empty[i], 1);
            mbar_init(&bar->d_full[i], 1); mbar_init(&bar->d_empty[i], EPI_WARPS * 32);
        }
        mbar_init(&bar->st_full, EPI_WARPS * 32); mbar_init(&bar->st_empty, CONV_THREADS);
        fence_barrier_init();
    }
    for (int i = threadIdx.x * 16; i < (int)W_IMAGE; i += THREADS * 16)
        *reinterpret_cast<float4*>(w_tile + i) = *reinterpret_cast<const float4*>(p.w_img + i);
    fence_proxy_async_smem();
    if (warp == CONV_WARPS) tmem_alloc(tmem_holder, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_holder;
    const long long n_tiles = (p.n_rows + TILE - 1) / TILE;
    const int n_my = (int)((n_tiles - (long long)blockIdx.x + gridDim.x - 1) / gridDim.x);     // tiles bid, bid + grid, ...
    auto tile_row0 = [&](int i) { return ((long long)blockIdx.x + (long long)i * gridDim.x) * TILE; };

    if (warp < CONV_WARPS) {
        // =====================================================================================
        // converter warps: ring -> per-row scale -> fp16 hi/lo operand; one tile later: stage -> out rows
        // =====================================================================================
        const int g = threadIdx.x >> 4, q = threadIdx.x & 15, q4 = 4 * q;          // 16 groups x 16 lanes; rows g + 16 k
        const int chunk = q >> 3, piece = q & 7;
        const uint32_t a_sub = ((uint32_t)(q & 1)) * 8u;
        const float w_descale = *reinterpret_cast<const float*>(p.w_img + W_IMAGE);
        auto write_out = [&](int j) {
            const long long r0 = tile_row0(j);
            mbar_wait(&bar->st_full, (uint32_t)j & 1);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int r = g + 16 * k;
                const float4 d = *reinterpret_cast<const float4*>(stage + r * STAGE_LD + q4);
                if (r0 + r < p.n_rows) stg4(p.out + (p.row_lo + r0 + r) * F + q4, d);
            }
            mbar_arrive(&bar->st_empty);
        };
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const int s = i % IN_STAGES, slot = i & 1;
            const long long r0 = tile_row0(i);
            const int rows = (int)min((long long)TILE, p.n_rows - r0);
            mbar_wait(&bar->in_full[s], (uint32_t)(i / IN_STAGES) & 1);
            const float* src = in_ring + (size_t)s * (IN_TILE / 4);
            float4 x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int r = g + 16 * k;
                x[k] = r < rows ? *reinterpret_cast<const float4*>(src + r * F + q4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // the slot goes back to the loader only when the reads have RETURNED: the arrive's address depends on the values
            // (an arrive issued right behind the LDS could overtake them; see swe_rowmlp_tc16.cu)
            {
                uint32_t dep = 0u, zero;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    dep |= __float_as_uint(x[k].x) | __float_as_uint(x[k].y) | __float_as_uint(x[k].z) | __float_as_uint(x[k].w);
                asm volatile("and.b32 %0, %1, %2;" : "=r"(zero) : "r"(dep), "r"(p.zero_mask));
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar->in_empty[s]) + zero) : "memory");
            }
            // per-row power-of-two scale (max |x'| in [2^13, 2^14)); the row's 16 lanes agree on the maximum
            float sc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float m = fmaxf(fmaxf(fabsf(x[k].x), fabsf(x[k].y)), fmaxf(fabsf(x[k].z), fabsf(x[k].w)));
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
                uint32_t sb = 267u - (__float_as_uint(m) >> 23);
                sb = sb > 253u ? 253u : sb;
                sc[k] = __uint_as_float(sb << 23);
                if (q == 0) s_inv[slot * TILE + g + 16 * k] = __uint_as_float((254u - sb) << 23) * w_descale;
            }
            mbar_wait(&bar->a_empty[slot], (((uint32_t)i >> 1) & 1) ^ 1);   // the MMAs that last read this slot have completed
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
            if (i > 0) write_out(i - 1);
        }
        if (n_my > 0) write_out(n_my - 1);
    } else if (warp < CONV_WARPS + EPI_WARPS) {
        // =====================================================================================
        // epilogue warps: thread = TMEM lane = row; D x (row 2^-e, filter 2^-f) -> stage
        // =====================================================================================
        const int lq = warp & 3;
        float* my_row = stage + (lq * 32 + lane) * STAGE_LD;
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const int dslot = i & 1;
            const uint32_t ud = (uint32_t)i >> 1;
            mbar_wait(&bar->d_full[dslot], ud & 1);
            tc_fence_after_sync();
            const float inv = s_inv[dslot * TILE + lq * 32 + lane];
            mbar_wait(&bar->st_empty, ((uint32_t)i & 1) ^ 1);           // stage written out (tile i - 1)
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
        }
    } else if (warp == CONV_WARPS + EPI_WARPS) {
        // =====================================================================================
        // MMA issuer (whole warp converged, the issuing lane is elected inside the asm block)
        // =====================================================================================
        const uint32_t idesc = make_idesc_f16(TILE, F);
        const uint64_t desc_hi = make_desc_sw64(0) & 0xFFFFFFFF00000000ull;
        const uint32_t desc_lo0 = (uint32_t)make_desc_sw64(0);
        auto dsc = [&](uint32_t lo) { return desc_hi | (uint64_t)lo; };
        const uint32_t a_d = desc_lo0 + (smem_u32(a_slots) >> 4), w_d = desc_lo0 + (smem_u32(w_tile) >> 4);
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const int slot = i & 1;
            const uint32_t u = (uint32_t)i >> 1;
            mbar_wait(&bar->d_empty[slot], (u & 1) ^ 1);               // accumulator of tile i - 2 has been staged
            mbar_wait(&bar->a_full[slot], u & 1);
            tc_fence_after_sync();
            const uint32_t d = tmem_base + slot * 64;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t a_hi = a_d + ((uint32_t)slot * A_SLOT + (uint32_t)c * 2 * A_TILE) / 16, a_lo = a_hi + A_TILE / 16;
                const uint32_t w_hi = w_d + ((uint32_t)c * 2 * W_TILE) / 16, w_lo = w_hi + W_TILE / 16;
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                    mma_f16_ss_warp(d, dsc(a_lo + 2 * ks), dsc(w_hi + 2 * ks), idesc, (c | ks) ? 1u : 0u);
                    mma_f16_ss_warp(d, dsc(a_hi + 2 * ks), dsc(w_lo + 2 * ks), idesc, 1u);
                    mma_f16_ss_warp(d, dsc(a_hi + 2 * ks), dsc(w_hi + 2 * ks), idesc, 1u);
                }
            }
            mma_commit_warp(&bar->a_empty[slot]);
            mma_commit_warp(&bar->d_full[slot]);
        }
    } else if (lane == 0) {
        // =====================================================================================
        // loader: whole tiles (contiguous rows) into the ring, three in flight
        // =====================================================================================
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            const int s = i % IN_STAGES;
            const long long r0 = tile_row0(i);
            const uint32_t bytes = (uint32_t)(min((long long)TILE, p.n_rows - r0) * F * 4);
            mbar_wait(&bar->in_empty[s], (((uint32_t)(i / IN_STAGES)) & 1) ^ 1);
            mbar_arrive_expect_tx(&bar->in_full[s], bytes);
            bulk_g2s(in_ring + (size_t)s * (IN_TILE / 4), p.x + (p.row_lo + r0) * F, bytes, &bar->in_full[s]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == CONV_WARPS) tmem_dealloc(tmem_base, 128);
}

}  // namespace rowlin
}  // namespace swe

using namespace swe;

// out[row_lo + r, :] = x[row_lo + r, :] · Wᵀ for r < n_rows; x, out: [*, 64] fp32; w_image: swe_hop_tc16_pack of W [64, 64]
extern "C" int swe_row_linear_tc16(const float* x, int64_t row_lo, int64_t n_rows, const void* w_image, float* out, void* stream) {
    SWE_REQUIRE(x && out && w_image && row_lo >= 0 && n_rows >= 0, SWE_E_INVAL, "row_linear_tc16: bad arguments");
    SWE_REQUIRE(aligned16(x) && aligned16(out) && aligned16(w_image), SWE_E_ALIGN, "row_linear_tc16: unaligned buffer");
    if (n_rows == 0) return 0;
    rowlin::Params p;
    p.x = x; p.out = out; p.zero_mask = 0u; p.row_lo = row_lo; p.n_rows = n_rows; p.w_img = (const unsigned char*)w_image;
    cudaError_t e = cudaFuncSetAttribute(rowlin::row_linear_tc16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rowlin::SMEM);
    if (e != cudaSuccess) { set_error("row_linear_tc16 smem opt-in (%zu B): %s", rowlin::SMEM, cudaGetErrorString(e)); return (int)e; }
    const long long n_tiles = (n_rows + rowlin::TILE - 1) / rowlin::TILE;
    rowlin::row_linear_tc16_kernel<<<grid_for(n_tiles, 1), rowlin::THREADS, rowlin::SMEM, (cudaStream_t)stream>>>(p);
    return check_launch("row_linear_tc16");
}
