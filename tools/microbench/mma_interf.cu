// What slows tcgen05.mma (kind::f16, M = 128, N = 128, K = 16, SS) down inside a busy CTA?  One CTA per SM, warp 16 issues
// the gate kernel's pattern (A_lo·W_hi, A_hi·W_lo, A_hi·W_hi per K step, operands rotating through 3-slot rings, 64-byte
// swizzle) while warps 0-15 do nothing / tcgen05.ld+st / 8-byte shared-memory stores / mbarrier polling / global loads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mswe-gnn_b200/csrc -o /tmp/mma_interf tools/microbench/mma_interf.cu
#include <cstdio>
#include "swe_tc.cuh"
namespace swe { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } }
using namespace swe::tc;

__global__ void __launch_bounds__(544, 1) interf_kernel(int mode, int rotate, int n_cols, int iters, const float* gsrc, long long* out, float* sink) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar, never;
    __shared__ uint32_t holder;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&never, 1); fence_barrier_init(); stop = 0; }
    fence_proxy_async_smem();
    if (warp == 16) tmem_alloc(&holder, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = holder;
    if (warp == 16) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(128, n_cols);
            const uint32_t a0 = smem_u32(smem), b0 = a0 + 49152;       // A ring: 3 x 16 KB (hi | lo), B ring: 3 x 16 KB
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                const uint32_t sl = rotate ? (uint32_t)(it % 3) : 0u;
                const uint32_t ah = a0 + sl * 16384, al = ah + 8192, wh = b0 + sl * 16384, wl = wh + 8192;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    mma_f16_ss(tm + 256, make_desc_sw64(al + ks * 32), make_desc_sw64(wh + ks * 32), idesc, 1u);
                    mma_f16_ss(tm + 256, make_desc_sw64(ah + ks * 32), make_desc_sw64(wl + ks * 32), idesc, 1u);
                    mma_f16_ss(tm + 256, make_desc_sw64(ah + ks * 32), make_desc_sw64(wh + ks * 32), idesc, 1u);
                }
            }
            mma_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t2 = clock64();
            if (blockIdx.x == 0) out[0] = t2 - t0;
            stop = 1;
        }
    } else if (mode == 1) {                  // TMEM traffic: ld 64 columns, st 32 columns (an epilogue's mix), all 16 warps
        const uint32_t addr = tm + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 32;
        uint32_t v[32];
        float acc = 0.f;
        while (!stop) {
            tmem_ld32(addr, v); tmem_wait_ld(); acc += __uint_as_float(v[3]);
            tmem_ld32(addr, v); tmem_wait_ld(); acc += __uint_as_float(v[5]);
            tmem_st32(addr + 128, v); tmem_wait_st();
        }
        if (acc == 1.2345f) sink[0] = acc;
    } else if (mode == 2) {                  // 8-byte shared-memory stores (the gather's operand stores)
        unsigned char* dst = smem + 98304 + threadIdx.x * 8;
        while (!stop) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(smem_u32(dst + j * 4096)), "r"(j), "r"(j + 1) : "memory");
            fence_proxy_async_smem();
        }
    } else if (mode == 3) {                  // mbarrier polling with the kernels' back-off
        while (!stop) { if (!mbar_try_wait(&never, 0)) __nanosleep(64); }
    } else if (mode == 4) {                  // gathered global loads (L1 / LSU traffic)
        float acc = 0.f;
        size_t i = (size_t)(blockIdx.x * 512 + threadIdx.x) * 4;
        while (!stop) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float4 x = swe::ldg4(gsrc + ((i + j * 77777 * 4) & ((1u << 26) - 4))); acc += x.x; }
            i += 4096 * 13;
        }
        if (acc == 1.2345f) sink[0] = acc;
    } else if (mode == 5) {                  // FP32 / conversion ALU work
        float a = threadIdx.x, b = 1.0001f; uint32_t h = 0, l = 0;
        while (!stop) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { split_f16x2(a, b, h, l); a = fmaf(a, 1.0001f, __uint_as_float(h)); b += __uint_as_float(l); }
        }
        if (a == 1.2345f) sink[0] = a + b;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tm, 512);
}

int main() {
    long long* out; float* sink; float* gsrc;
    cudaMalloc(&out, 16); cudaMalloc(&sink, 16); cudaMalloc(&gsrc, (size_t)1 << 28); cudaMemset(gsrc, 0, (size_t)1 << 28);
    const size_t smem = 1024 + 160 * 1024;
    cudaFuncSetAttribute(interf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 1500;
    const char* names[] = {"idle", "tmem ld/st", "sts.64 + proxy fence", "mbarrier polling", "global gathers", "alu / cvt"};
    for (int rotate = 0; rotate < 2; ++rotate)
        for (int mode = 0; mode < 6; ++mode) {
            interf_kernel<<<148, 544, smem>>>(mode, rotate, 128, iters, gsrc, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            long long h;
            cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("SS f16 N=128 %s operands, 16 warps %-22s: %.1f cycles/MMA (%s)\n", rotate ? "rotating" : "fixed   ", names[mode],
                   (double)h / (iters * 6), cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
        }
    return 0;
}
