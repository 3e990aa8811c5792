// Shared device helpers for the mSWE-GNN sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/swe_gnn_b200.h"

namespace swe {

constexpr int NT = 256;          // threads per CTA in every tiled kernel
constexpr int TM = 128;          // rows (edges or nodes) per tile
constexpr int NUM_SMS = 148;     // B200

// ------------------------------------------------------------------------------------------
// error plumbing (host)
// ------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int  check_launch(const char* what);

#define SWE_REQUIRE(cond, code, ...)                   \
    do { if (!(cond)) { ::swe::set_error(__VA_ARGS__); return (code); } } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
// keeps gathered rows in L1 as well (re-used by neighbouring edges of the same tile)
__device__ __forceinline__ void cp_async16_ca(void* smem_dst, const void* gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming (read-once) 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float act_apply(int act, float v, float slope) {
    switch (act) {
        case SWE_ACT_PRELU:     return v >= 0.f ? v : slope * v;
        case SWE_ACT_RELU:      return fmaxf(v, 0.f);
        case SWE_ACT_TANH:      return tanhf(v);
        case SWE_ACT_LEAKYRELU: return v >= 0.f ? v : 0.1f * v;
        case SWE_ACT_ELU:       return v > 0.f ? v : expm1f(v);
        case SWE_ACT_SWISH:     return v / (1.f + expf(-v));
        case SWE_ACT_SIGMOID:   return 1.f / (1.f + expf(-v));
        default:                return v;
    }
}
// d act(v) / d v expressed with the pre-activation v (backward kernels)
__device__ __forceinline__ float act_grad(int act, float v, float slope) {
    switch (act) {
        case SWE_ACT_PRELU:     return v > 0.f ? 1.f : slope;           // torch: input > 0 ? grad : weight * grad
        case SWE_ACT_RELU:      return v > 0.f ? 1.f : 0.f;
        case SWE_ACT_TANH:      { float t = tanhf(v); return 1.f - t * t; }
        case SWE_ACT_LEAKYRELU: return v > 0.f ? 1.f : 0.1f;
        case SWE_ACT_ELU:       return v > 0.f ? 1.f : expf(v);
        case SWE_ACT_SWISH:     { float sg = 1.f / (1.f + expf(-v)); return sg * (1.f + v * (1.f - sg)); }
        case SWE_ACT_SIGMOID:   { float sg = 1.f / (1.f + expf(-v)); return sg * (1.f - sg); }
        default:                return 1.f;
    }
}
__device__ __forceinline__ float load_slope(const swe_layer_t& L) {
    return (L.act == SWE_ACT_PRELU && L.slope) ? __ldg(L.slope) : 0.f;
}

inline int grid_for(long long n_tiles, int ctas_per_sm) {
    long long cap = (long long)NUM_SMS * ctas_per_sm;
    if (n_tiles < 1) n_tiles = 1;
    return (int)(n_tiles < cap ? n_tiles : cap);
}

}  // namespace swe
