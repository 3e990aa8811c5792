// tcgen05 issue / throughput microbenchmark, second edition (one CTA or CTA pair per SM):
//   * kind::tf32 (K = 8) vs kind::f16 (K = 16), SS vs TS, N in {64, 128, 256}, accumulating into ONE TMEM region or
//     alternating between two (does the accumulator dependency pace the pipe?)
//   * cta_group::2 (M = 256 over a CTA pair) for the same kinds
//   * tcgen05.ld / tcgen05.st throughput with 4 and 8 warps
// Build (binary goes to /tmp, never into the tree):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mswe-gnn_b200/csrc -o /tmp/mma_rate2 tools/microbench/mma_rate2.cu
#include <cstdio>
#include "swe_tc.cuh"
namespace swe { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }
using namespace swe::tc;

// c = F32, a = b = F16 (format 0), K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_tf32_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_tf32_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// cta_group::1
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(160, 1) rate1_kernel(int f16, int ts, int n_cols, int alt, int iters, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0 pairs
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (warp == 4) tmem_alloc(&holder, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = holder;
    if (warp == 4 && lane == 0) {
        const uint32_t idesc = f16 ? make_idesc_f16(128, n_cols) : make_idesc_tf32(128, n_cols);
        const uint32_t a = smem_u32(smem), b = a + 32768;            // B: up to 256 rows x 128 B = 32 KB (+ks*32)
        const uint32_t d_alt = (n_cols <= 128) ? 128u : 0u;          // second accumulator region (N = 256: none, same D)
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t da = make_desc_sw128(a + ks * 32), db = make_desc_sw128(b + ks * 32);
                const uint32_t d = tm + 256 - ((alt && (ks & 1)) ? d_alt : 0u);
                if (f16) { if (ts) mma_f16_ts(d, tm + ks * 8, db, idesc, 1u); else mma_f16_ss(d, da, db, idesc, 1u); }
                else     { if (ts) mma_tf32_ts(d, tm + ks * 8, db, idesc, 1u); else mma_tf32_ss(d, da, db, idesc, 1u); }
            }
        }
        mma_commit(&bar);
        const long long t1 = clock64();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// cta_group::2 (M = 256 over the pair; B's N rows split between the two CTAs' shared memories)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1) rate2_kernel(int f16, int ts, int n_cols, int iters, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before_sync();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc_fence_after_sync();
    const uint32_t tm = holder;
    if (warp == 4 && lane == 0) {
        if (cta_rank == 0) {
            const uint32_t idesc = f16 ? make_idesc_f16(256, n_cols) : make_idesc_tf32(256, n_cols);
            const uint32_t a = smem_u32(smem), b = a + 32768;
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t da = make_desc_sw128(a + ks * 32), db = make_desc_sw128(b + ks * 32);
                    const uint32_t d = tm + 256;
                    if (f16) { if (ts) mma2_f16_ts(d, tm + ks * 8, db, idesc, 1u); else mma2_f16_ss(d, da, db, idesc, 1u); }
                    else     { if (ts) mma2_tf32_ts(d, tm + ks * 8, db, idesc, 1u); else mma2_tf32_ss(d, da, db, idesc, 1u); }
                }
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
            const long long t1 = clock64();
            mbar_wait(&bar, 0);
            const long long t2 = clock64();
            if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        } else {
            mbar_wait(&bar, 0);
        }
    }
    tc_fence_before_sync();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// TMEM load / store throughput: n_warps (4 or 8) warps, each moving 128 columns of its lane quarter per iteration
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(288, 1) tmem_rw_kernel(int store, int n_warps, int iters, long long* out, float* sink) {
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    if (warp == 8) tmem_alloc(&holder, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = holder;
    if (warp < n_warps) {
        const uint32_t addr = tm + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = threadIdx.x + j;
        const long long t0 = clock64();
        float acc = 0.f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (store) tmem_st32(addr + c * 32, v);
                else { tmem_ld32(addr + c * 32, v); }
            }
            if (store) tmem_wait_st(); else { tmem_wait_ld(); acc += __uint_as_float(v[it & 31]); }
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = t1 - t0;
        if (acc == 123.456f) sink[0] = acc;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tm, 512);
}

int main() {
    long long* out; float* sink;
    cudaMalloc(&out, 16); cudaMalloc(&sink, 16);
    const size_t smem = 1024 + 96 * 1024;
    cudaFuncSetAttribute(rate1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 2000;
    long long h[2];
    for (int f16 = 0; f16 < 2; ++f16)
        for (int ts = 0; ts < 2; ++ts)
            for (int n : {256, 128, 64})
                for (int alt = 0; alt < 2; ++alt) {
                    if (alt && n == 256) continue;
                    rate1_kernel<<<148, 160, smem>>>(f16, ts, n, alt, iters, out);
                    cudaError_t e = cudaDeviceSynchronize();
                    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                    printf("cta1 %s %s N=%3d %s: issue %.1f, complete %.1f cycles/MMA (%s)\n", f16 ? "f16 " : "tf32", ts ? "TS" : "SS", n,
                           alt ? "alt-D " : "same-D", (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), cudaGetErrorString(e));
                    if (e != cudaSuccess) return 1;
                }
    for (int f16 = 0; f16 < 2; ++f16)
        for (int ts = 0; ts < 2; ++ts)
            for (int n : {256, 128, 64}) {
                rate2_kernel<<<148, 160, smem>>>(f16, ts, n, iters, out);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                printf("cta2 %s %s M=256 N=%3d: issue %.1f, complete %.1f cycles/MMA (%s)\n", f16 ? "f16 " : "tf32", ts ? "TS" : "SS", n,
                       (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), cudaGetErrorString(e));
                if (e != cudaSuccess) return 1;
            }
    for (int store = 0; store < 2; ++store)
        for (int nw : {4, 8}) {
            tmem_rw_kernel<<<148, 288>>>(store, nw, 1000, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)nw * 32 * 128 * 4 * 1000;
            printf("tmem %s %d warps: %.1f cycles per 128-column sweep per warp, %.1f B/cycle/SM (%s)\n", store ? "st" : "ld", nw,
                   (double)h[0] / 1000, bytes / (double)h[0], cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
