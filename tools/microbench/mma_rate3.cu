// tcgen05.mma rate, third edition: kind / operand source / swizzle are COMPILE-TIME parameters and the issue loop is the
// tight one of mma_interf.cu.  (mma_rate2.cu selects them with runtime flags inside the loop: its 105-115 cycles per
// instruction, identical for every kind and N, turned out to be the branchy issue loop, not the tensor pipe.)
//   kind::tf32 (K = 8) / kind::f16 (K = 16), SS / TS, 128-byte / 64-byte swizzle, N in {128, 64}, M = 128,
//   issued by one thread inside `if (lane == 0)` or by the converged warp (elect.sync inside the asm block).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mswe-gnn_b200/csrc -o /tmp/mma_rate3 tools/microbench/mma_rate3.cu
#include <cstdio>
#include "swe_tc.cuh"
namespace swe { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }
using namespace swe::tc;

__device__ __forceinline__ void mma_tf32_ss_warp(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_tf32_ts_warp(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int F16, int TS, int SW64, int CONV>
__global__ void __launch_bounds__(160, 1) rate_kernel(int n_cols, int iters, long long* out) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = F16 ? 0x3c003c00u : 0x3f800000u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (warp == 4) tmem_alloc(&holder, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = holder;
    if (warp == 4 && (CONV || lane == 0)) {
        const uint32_t idesc = F16 ? make_idesc_f16(128, n_cols) : make_idesc_tf32(128, n_cols);
        const uint32_t a = smem_u32(smem), b = a + 32768;
        // 64-byte rows hold two 32-byte K slices, the next pair lives in the next [rows x 64 B] chunk
        uint64_t da[4], db[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const uint32_t off = SW64 ? (uint32_t)((ks & 1) * 32 + (ks >> 1) * 8192) : (uint32_t)(ks * 32);
            da[ks] = SW64 ? make_desc_sw64(a + off) : make_desc_sw128(a + off);
            db[ks] = SW64 ? make_desc_sw64(b + off) : make_desc_sw128(b + off);
        }
        const uint32_t d = tm + 256;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                if (CONV) {
                    if (F16) { if (TS) mma_f16_ts_warp(d, tm + ks * 8, db[ks], idesc, 1u); else mma_f16_ss_warp(d, da[ks], db[ks], idesc, 1u); }
                    else     { if (TS) mma_tf32_ts_warp(d, tm + ks * 8, db[ks], idesc, 1u); else mma_tf32_ss_warp(d, da[ks], db[ks], idesc, 1u); }
                } else {
                    if (F16) { if (TS) mma_f16_ts(d, tm + ks * 8, db[ks], idesc, 1u); else mma_f16_ss(d, da[ks], db[ks], idesc, 1u); }
                    else     { if (TS) mma_tf32_ts(d, tm + ks * 8, db[ks], idesc, 1u); else mma_tf32_ss(d, da[ks], db[ks], idesc, 1u); }
                }
            }
        }
        if (CONV) mma_commit_warp(&bar); else mma_commit(&bar);
        const long long t1 = clock64();
        mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tm, 512);
}

template <int F16, int TS, int SW64, int CONV>
static int run(long long* out) {
    const size_t smem = 1024 + 96 * 1024;
    cudaFuncSetAttribute(rate_kernel<F16, TS, SW64, CONV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 2000;
    for (int n : {128, 64}) {
        rate_kernel<F16, TS, SW64, CONV><<<148, 160, smem>>>(n, iters, out);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        printf("%s %s %s %s N=%3d: issue %.1f, complete %.1f cycles/MMA (%s)\n", CONV ? "warp-converged" : "lane-0 branch ", SW64 ? "sw64 " : "sw128",
               F16 ? "f16 " : "tf32", TS ? "TS" : "SS", n, (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
    }
    return 0;
}

int main() {
    long long* out;
    cudaMalloc(&out, 16);
    int r = 0;
#define ALL(conv) r |= run<0, 0, 0, conv>(out) | run<0, 1, 0, conv>(out) | run<1, 0, 0, conv>(out) | run<1, 1, 0, conv>(out) | \
                       run<0, 0, 1, conv>(out) | run<0, 1, 1, conv>(out) | run<1, 0, 1, conv>(out) | run<1, 1, 1, conv>(out)
    ALL(0);
    ALL(1);
    return r;
}
