// tcgen05.mma kind::tf32 issue/throughput microbenchmark (one CTA per SM): cycles per MMA for
//   SS (A and B from shared memory) vs TS (A from TMEM), N = 128 vs 64, with and without a second warp group hammering
//   shared memory with 16-byte stores (what the gate kernel's row workers do while the MMAs run).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../mswe-gnn_b200/csrc -o /tmp/mma_rate mma_rate.cu
#include <cstdio>
#include "swe_tc.cuh"
namespace swe { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }
using namespace swe::tc;

__global__ void __launch_bounds__(288, 1) mma_rate_kernel(int mode, int n_cols, int iters, int hammer, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
    fence_proxy_async_smem();
    if (warp == 8) tmem_alloc(&holder, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = holder;
    if (warp == 8) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_tf32(128, n_cols);
            const uint32_t a = smem_u32(smem), b = a + 32768;
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t da = make_desc_sw128(a + ks * 32), db = make_desc_sw128(b + ks * 32);
                    if (mode == 0) mma_tf32_ss(tm + 256, da, db, idesc, 1u);   // D at columns [256, 256 + N)
                    else mma_tf32_ts(tm + 256, tm + ks * 8, db, idesc, 1u);
                }
            }
            mma_commit(&bar);
            const long long t1 = clock64();
            mbar_wait(&bar, 0);
            const long long t2 = clock64();
            if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
            stop = 1;
        }
    } else if (hammer) {
        float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
        unsigned char* dst = smem + 65536 + threadIdx.x * 16;
        while (!stop) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(smem_u32(dst + j * 4096)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tm, 512);
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    const size_t smem = 1024 + 65536 + 65536;
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 2000;
    for (int hammer = 0; hammer < 2; ++hammer)
        for (int mode = 0; mode < 2; ++mode)
            for (int n : {256, 128, 64}) {
                mma_rate_kernel<<<148, 288, smem>>>(mode, n, iters, hammer, out);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2];
                cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                printf("%s N=%d hammer=%d: issue %.1f cycles/MMA, complete %.1f cycles/MMA (%s)\n", mode ? "TS" : "SS", n, hammer,
                       (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), cudaGetErrorString(e));
            }
    return 0;
}
