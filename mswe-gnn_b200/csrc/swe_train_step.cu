// The rest of the training step on the device: the wet-cell loss with its gradient, global-norm
// gradient clipping and AdamW over ONE flat parameter buffer — so that forward + backward + update is a fixed kernel
// sequence without host reads (capturable in a CUDA graph).
//   loss      : /root/reference/training/loss.py:76-118 (conservation = 0): per variable RMSE (or MAE) over the finest-scale
//               rows where pred - real has a non-zero entry (`mask_on_water`, loss.py:30-36), velocity term weighted
//   clip      : torch.nn.utils.clip_grad_norm_(params, max_norm)   (Lightning gradient_clip_val, /root/reference/main.py:109)
//   optimizer : torch.optim.AdamW                                   (/root/reference/training/train.py:147-155)
// All reductions are per-CTA partial sums combined in CTA order: bit-reproducible.
#include "swe_common.cuh"

namespace swe {

constexpr int TS_THREADS = 256;
constexpr int TS_MAX_PARTS = 2 * NUM_SMS;

// ---------------------------------------------------------------------------------------------
// loss
// ---------------------------------------------------------------------------------------------
// partials[cta] = { sum_0, sum_1, count, - }:  sum_v = Σ diff_v² (RMSE) or Σ |diff_v| (MAE) over the kept rows
__global__ void __launch_bounds__(TS_THREADS) loss_partials_kernel(const float* __restrict__ pred, const float* __restrict__ real,
                                                                   long long real_stride, const unsigned char* __restrict__ rows,
                                                                   long long n, int only_water, int mae, float4* __restrict__ partials) {
    float s0 = 0.f, s1 = 0.f, c = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (rows && !rows[i]) continue;
        const float2 p = *reinterpret_cast<const float2*>(pred + 2 * i);
        const float d0 = p.x - real[i * real_stride], d1 = p.y - real[i * real_stride + real_stride / 2];
        if (only_water && d0 == 0.f && d1 == 0.f) continue;
        s0 += mae ? fabsf(d0) : d0 * d0;
        s1 += mae ? fabsf(d1) : d1 * d1;
        c += 1.f;
    }
    __shared__ float sh[3][TS_THREADS / 32];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, off); s1 += __shfl_xor_sync(0xffffffffu, s1, off); c += __shfl_xor_sync(0xffffffffu, c, off);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s0; sh[1][threadIdx.x >> 5] = s1; sh[2][threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f, k = 0.f;
        for (int w = 0; w < TS_THREADS / 32; ++w) { a += sh[0][w]; b += sh[1][w]; k += sh[2][w]; }
        partials[blockIdx.x] = make_float4(a, b, k, 0.f);
    }
}

// loss (+)= scale · Σ_v w_v err_v / Σ_v w_v ;  dpred = d loss / d pred  (zero outside the kept rows)
__global__ void __launch_bounds__(TS_THREADS) loss_grad_kernel(const float* __restrict__ pred, const float* __restrict__ real,
                                                               long long real_stride, const unsigned char* __restrict__ rows,
                                                               long long n, int only_water, int mae, const float4* __restrict__ partials,
                                                               int n_parts, float w0, float w1, float scale, int accumulate,
                                                               float* __restrict__ loss, float* __restrict__ dpred) {
    __shared__ float s_c[2];
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0, k = 0.0;
        for (int i = 0; i < n_parts; ++i) { const float4 q = partials[i]; a += q.x; b += q.y; k += q.z; }
        const float cnt = (float)k, wsum = w0 + w1;
        float e0, e1, c0, c1;
        if (mae) {                                               // err_v = S_v / cnt ; d err_v / d diff = sign / cnt
            e0 = (float)a / cnt; e1 = (float)b / cnt;
            c0 = scale * w0 / (wsum * cnt); c1 = scale * w1 / (wsum * cnt);
        } else {                                                 // err_v = sqrt(S_v / cnt) ; d err_v / d diff = diff / (cnt err_v)
            e0 = sqrtf((float)a / cnt); e1 = sqrtf((float)b / cnt);
            c0 = e0 > 0.f ? scale * w0 / (wsum * cnt * e0) : 0.f;
            c1 = e1 > 0.f ? scale * w1 / (wsum * cnt * e1) : 0.f;
        }
        s_c[0] = c0; s_c[1] = c1;
        if (blockIdx.x == 0) {
            const float l = scale * (w0 * e0 + w1 * e1) / wsum;  // cnt == 0: NaN, as the reference's mean over no rows
            *loss = accumulate ? *loss + l : l;
        }
    }
    __syncthreads();
    const float c0 = s_c[0], c1 = s_c[1];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float2 g = make_float2(0.f, 0.f);
        if (!rows || rows[i]) {
            const float2 p = *reinterpret_cast<const float2*>(pred + 2 * i);
            const float d0 = p.x - real[i * real_stride], d1 = p.y - real[i * real_stride + real_stride / 2];
            if (!(only_water && d0 == 0.f && d1 == 0.f)) {
                if (mae) { g.x = c0 * (d0 > 0.f ? 1.f : (d0 < 0.f ? -1.f : 0.f)); g.y = c1 * (d1 > 0.f ? 1.f : (d1 < 0.f ? -1.f : 0.f)); }
                else { g.x = c0 * d0; g.y = c1 * d1; }
            }
        }
        *reinterpret_cast<float2*>(dpred + 2 * i) = g;
    }
}

// ---------------------------------------------------------------------------------------------
// clip + AdamW over the flat parameter buffer
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TS_THREADS) sumsq_partials_kernel(const float* __restrict__ g, long long n, float* __restrict__ partials) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s = fmaf(g[i], g[i], s);
    __shared__ float sh[TS_THREADS / 32];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int w = 0; w < TS_THREADS / 32; ++w) a += sh[w];
        partials[blockIdx.x] = a;
    }
}

// state[0] = step count (float), state[1] = last total gradient norm, state[2] = last clip coefficient
__global__ void __launch_bounds__(TS_THREADS) adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                           float* __restrict__ v, long long n, const float* __restrict__ lr_ptr,
                                                           float b1, float b2, float eps, float wd, float max_norm,
                                                           const float* __restrict__ partials, int n_parts, float* __restrict__ state) {
    __shared__ float s_k[3];
    if (threadIdx.x == 0) {
        double a = 0.0;
        for (int i = 0; i < n_parts; ++i) a += partials[i];
        const float norm = (float)sqrt(a);
        float coef = 1.f;
        if (max_norm > 0.f) { coef = max_norm / (norm + 1e-6f); coef = coef > 1.f ? 1.f : coef; }     // clip_grad_norm_
        const double t = (double)state[0] + 1.0;
        const double bc1 = 1.0 - pow((double)b1, t), bc2 = 1.0 - pow((double)b2, t);
        s_k[0] = coef; s_k[1] = (float)((double)(*lr_ptr) / bc1); s_k[2] = (float)sqrt(bc2);
        if (blockIdx.x == 0) { state[1] = norm; state[2] = coef; }
    }
    __syncthreads();
    const float coef = s_k[0], step_size = s_k[1], bc2_sqrt = s_k[2], lr = *lr_ptr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * coef;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = m[i] + (gi - m[i]) * (1.f - b1);                  // lerp, as torch's exp_avg.lerp_(grad, 1 - beta1)
        const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= step_size * (mi / denom);
        p[i] = pi; m[i] = mi; v[i] = vi; g[i] = gi;                         // (the clipped gradient is what clip_grad_norm_ leaves)
    }
}
__global__ void adamw_step_advance_kernel(float* state) { state[0] += 1.f; }

}  // namespace swe

using namespace swe;

static int parts_for(long long n) {
    long long b = (n + TS_THREADS - 1) / TS_THREADS;
    return (int)(b < 1 ? 1 : (b > TS_MAX_PARTS ? TS_MAX_PARTS : b));
}

extern "C" size_t swe_train_step_ws_bytes(void) { return (size_t)TS_MAX_PARTS * sizeof(float4); }

// pred [n, 2]; real: element (i, v) at real[i * real_stride + v * real_stride / 2] (a [n, 2, T] target sliced at one time step
// has real_stride = 2 T; a contiguous [n, 2] target real_stride = 2); rows: optional [n] bytes, non-zero = row takes part
extern "C" int swe_loss_fwd_bwd(const float* pred, const float* real, int64_t real_stride, const unsigned char* rows, int64_t n,
                                int32_t only_where_water, int32_t mae, float w0, float w1, float scale, int32_t accumulate,
                                float* loss, float* dpred, void* ws, void* stream) {
    SWE_REQUIRE(pred && real && loss && dpred && ws && n >= 0 && real_stride >= 2 && (real_stride & 1) == 0, SWE_E_INVAL,
                "loss_fwd_bwd: bad arguments");
    SWE_REQUIRE((reinterpret_cast<uintptr_t>(pred) & 7u) == 0 && (reinterpret_cast<uintptr_t>(dpred) & 7u) == 0 && aligned16(ws),
                SWE_E_ALIGN, "loss_fwd_bwd: unaligned buffer");
    const int parts = parts_for(n);
    loss_partials_kernel<<<parts, TS_THREADS, 0, (cudaStream_t)stream>>>(pred, real, real_stride, rows, n, only_where_water, mae, (float4*)ws);
    loss_grad_kernel<<<parts, TS_THREADS, 0, (cudaStream_t)stream>>>(pred, real, real_stride, rows, n, only_where_water, mae,
                                                                      (const float4*)ws, parts, w0, w1, scale, accumulate, loss, dpred);
    return check_launch("loss_fwd_bwd");
}

// One optimizer step on the flat buffers (params, grads, exp_avg, exp_avg_sq: [n] fp32): global-norm clipping with max_norm
// (<= 0: none) and AdamW; lr is read from device memory (a scheduler may change it between replays of a captured step)
extern "C" int swe_clip_adamw_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr,
                                   float beta1, float beta2, float eps, float weight_decay, float max_norm, float* state,
                                   void* ws, void* stream) {
    SWE_REQUIRE(params && grads && exp_avg && exp_avg_sq && lr && state && ws && n >= 0, SWE_E_INVAL, "clip_adamw_step: bad arguments");
    if (n == 0) return 0;
    const int parts = parts_for(n);
    sumsq_partials_kernel<<<parts, TS_THREADS, 0, (cudaStream_t)stream>>>(grads, n, (float*)ws);
    adamw_kernel<<<parts, TS_THREADS, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                  weight_decay, max_norm, (const float*)ws, parts, state);
    adamw_step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
    return check_launch("clip_adamw_step");
}
