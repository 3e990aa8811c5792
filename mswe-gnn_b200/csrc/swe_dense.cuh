// CUDA-core FP32 tile GEMM engine used by every MLP-bearing kernel (encoders, edge gate, filter
// matrices, decoder) in the exact-fp32 path.
//
// Tile: TM = 128 rows × NO output columns per CTA of 256 threads.  Thread (ty, tx) with
// tx = t % TX, ty = t / TX, TX = NO/8, owns RM = TM/(256/TX) consecutive rows and 8 columns
// split in two groups of 4: [4·tx, 4·tx+4) and [NO/2 + 4·tx, NO/2 + 4·tx + 4) so that the
// 128-bit shared-memory reads of W are contiguous across the lanes of a quarter-warp
// (bank-conflict free) and A reads are warp-broadcasts.
//   A (activations): shared memory, row-major, leading dimension lda (≡ 4 mod 32 floats → padded)
//   W (weights)    : shared memory, k-major [K][NO] — i.e. the packed transposed Linear weight
#pragma once
#include "swe_common.cuh"

namespace swe {

template <int NO>
struct DenseCfg {
    static_assert(NO % 8 == 0 && NO >= 16 && NO <= 128, "NO must be 16..128, multiple of 8");
    static constexpr int TX = NO / 8;
    static constexpr int TY = NT / TX;
    static constexpr int RM = TM / TY;
    static_assert(RM >= 1, "tile too small for this width");
};

// acc[i][0..3] ↔ columns 4·tx + j ; acc[i][4..7] ↔ columns NO/2 + 4·tx + j
template <int NO>
__device__ __forceinline__ void dense_zero(float (&acc)[DenseCfg<NO>::RM][8]) {
#pragma unroll
    for (int i = 0; i < DenseCfg<NO>::RM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
}

// acc += A[rows, 0:K] · W[0:K, :]   (K multiple of 4; A and W already resident in shared memory)
template <int NO>
__device__ __forceinline__ void dense_acc(float (&acc)[DenseCfg<NO>::RM][8], const float* __restrict__ A,
                                          int lda, const float* __restrict__ W, int K) {
    using C = DenseCfg<NO>;
    const int tx = threadIdx.x % C::TX;
    const int ty = threadIdx.x / C::TX;
    const float* a_base = A + (ty * C::RM) * lda;
    const float* w_lo = W + 4 * tx;
    const float* w_hi = W + NO / 2 + 4 * tx;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 a[C::RM];
#pragma unroll
        for (int i = 0; i < C::RM; ++i) a[i] = *reinterpret_cast<const float4*>(a_base + i * lda + k);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 b0 = *reinterpret_cast<const float4*>(w_lo + (k + j) * NO);
            const float4 b1 = *reinterpret_cast<const float4*>(w_hi + (k + j) * NO);
#pragma unroll
            for (int i = 0; i < C::RM; ++i) {
                const float av = j == 0 ? a[i].x : j == 1 ? a[i].y : j == 2 ? a[i].z : a[i].w;
                acc[i][0] = fmaf(av, b0.x, acc[i][0]);
                acc[i][1] = fmaf(av, b0.y, acc[i][1]);
                acc[i][2] = fmaf(av, b0.z, acc[i][2]);
                acc[i][3] = fmaf(av, b0.w, acc[i][3]);
                acc[i][4] = fmaf(av, b1.x, acc[i][4]);
                acc[i][5] = fmaf(av, b1.y, acc[i][5]);
                acc[i][6] = fmaf(av, b1.z, acc[i][6]);
                acc[i][7] = fmaf(av, b1.w, acc[i][7]);
            }
        }
    }
}

// bias + activation in registers
template <int NO>
__device__ __forceinline__ void dense_bias_act(float (&acc)[DenseCfg<NO>::RM][8], const float* __restrict__ bias,
                                               int act, float slope) {
    using C = DenseCfg<NO>;
    const int tx = threadIdx.x % C::TX;
    float b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = 0.f;
    if (bias) {
        const float4 b0 = ldg4(bias + 4 * tx), b1 = ldg4(bias + NO / 2 + 4 * tx);
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
    }
#pragma unroll
    for (int i = 0; i < C::RM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = act_apply(act, acc[i][j] + b[j], slope);
}

// registers → shared memory tile (row-major, leading dimension ldy)
template <int NO>
__device__ __forceinline__ void dense_store_smem(const float (&acc)[DenseCfg<NO>::RM][8], float* __restrict__ Y, int ldy) {
    using C = DenseCfg<NO>;
    const int tx = threadIdx.x % C::TX;
    const int ty = threadIdx.x / C::TX;
#pragma unroll
    for (int i = 0; i < C::RM; ++i) {
        float* y = Y + (ty * C::RM + i) * ldy;
        stg4(y + 4 * tx, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        stg4(y + NO / 2 + 4 * tx, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
    }
}

// registers → global rows [row0, row0+TM) ∩ [0, n_rows) of a [*, NO] row-major matrix
template <int NO>
__device__ __forceinline__ void dense_store_global(const float (&acc)[DenseCfg<NO>::RM][8], float* __restrict__ Y,
                                                   long long row0, long long n_rows) {
    using C = DenseCfg<NO>;
    const int tx = threadIdx.x % C::TX;
    const int ty = threadIdx.x / C::TX;
#pragma unroll
    for (int i = 0; i < C::RM; ++i) {
        const long long r = row0 + ty * C::RM + i;
        if (r < n_rows) {
            float* y = Y + r * NO;
            stg4(y + 4 * tx, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
            stg4(y + NO / 2 + 4 * tx, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
        }
    }
}

// sum over the NO columns of each owned row (the TX threads sharing a row are adjacent lanes)
template <int NO>
__device__ __forceinline__ void dense_row_reduce_sum(float (&v)[DenseCfg<NO>::RM]) {
    using C = DenseCfg<NO>;
#pragma unroll
    for (int off = C::TX / 2; off >= 1; off >>= 1)
#pragma unroll
        for (int i = 0; i < C::RM; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], off);
}

// async copy of a contiguous block of `n_floats` (multiple of 4) global → shared, whole CTA
__device__ __forceinline__ void block_cp_async(float* __restrict__ dst, const float* __restrict__ src, int n_floats) {
    for (int i = threadIdx.x * 4; i < n_floats; i += NT * 4) cp_async16(dst + i, src + i);
}

}  // namespace swe
