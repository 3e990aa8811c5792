// Training path of the mSWE-GNN hot path: forward pieces that keep pre-activations, and the
// hand-derived backward kernels (the reference has no explicit backward, it
// is what torch.autograd derives for models/gnn.py:387-445 inside training/train.py:125-145).
// Exact-fp32 CUDA-core arithmetic; every reduction has a fixed order (no atomics).
#include "swe_dense.cuh"

namespace swe {

constexpr int LDT = TM + 4;                // leading dimension of a transposed [n][row] tile

__device__ __forceinline__ int round4(int v) { return (v + 3) & ~3; }

// ---------------------------------------------------------------------------------------------
// row provider: dst[r][0:w_pad) = act(X_seg[row0 + r, :]) (zero beyond the segment width / n_rows)
// ---------------------------------------------------------------------------------------------
// columns [c0, c0 + w_pad) of the segment, rows [row0, row0 + n_tile_rows)
__device__ __forceinline__ void load_seg_tile(float* __restrict__ dst, int ldd, const swe_seg_t& sg, long long row0,
                                              long long n_rows, int w_pad, int c0 = 0, int n_tile_rows = TM) {
    const float slope = (sg.act == SWE_ACT_PRELU && sg.slope) ? __ldg(sg.slope) : 0.f;
    const bool vec = (sg.ld % 4 == 0) && (sg.width % 4 == 0) && ((reinterpret_cast<uintptr_t>(sg.base) & 15u) == 0);
    if (vec) {
        const int qpr = w_pad / 4;
        for (int idx = threadIdx.x; idx < n_tile_rows * qpr; idx += NT) {
            const int r = idx / qpr, q = idx % qpr;
            const long long g = row0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g < n_rows && c0 + 4 * q < sg.width) {
                const long long sr = sg.idx ? (long long)__ldg(sg.idx + g) : g;
                v = ldg4(sg.base + sr * sg.ld + c0 + 4 * q);
                if (sg.act != SWE_ACT_NONE) {
                    v.x = act_apply(sg.act, v.x, slope); v.y = act_apply(sg.act, v.y, slope);
                    v.z = act_apply(sg.act, v.z, slope); v.w = act_apply(sg.act, v.w, slope);
                }
            }
            stg4(dst + r * ldd + 4 * q, v);
        }
    } else {
        for (int idx = threadIdx.x; idx < n_tile_rows * w_pad; idx += NT) {
            const int r = idx / w_pad, c = idx % w_pad;
            const long long g = row0 + r;
            float v = 0.f;
            if (g < n_rows && c0 + c < sg.width) {
                const long long sr = sg.idx ? (long long)__ldg(sg.idx + g) : g;
                v = act_apply(sg.act, __ldg(sg.base + sr * sg.ld + c0 + c), slope);
            }
            dst[r * ldd + c] = v;
        }
    }
}

constexpr int KSUB = 64;          // reduction extent staged at a time: keeps a CTA under 70 KB of shared memory (3 CTAs per SM)

// ---------------------------------------------------------------------------------------------
// pre = X · Wᵀ + b
// ---------------------------------------------------------------------------------------------
template <int NO>
__global__ void __launch_bounds__(NT) mlp_layer_fwd_kernel(const __grid_constant__ swe_rows_t X, long long n_rows,
                                                           const float* __restrict__ wt, const float* __restrict__ bias,
                                                           float* __restrict__ pre) {
    extern __shared__ __align__(16) float smem[];
    float* A = smem;                       // [TM][KSUB + 4]
    float* W = smem + TM * (KSUB + 4);     // [KSUB][NO]
    const long long n_tiles = (n_rows + TM - 1) / TM;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * TM;
        float acc[DenseCfg<NO>::RM][8];
        dense_zero<NO>(acc);
        int koff = 0;
        for (int j = 0; j < X.n_seg; ++j) {
            const int wp = round4(X.seg[j].width);
            for (int c0 = 0; c0 < wp; c0 += KSUB) {
                const int wc = min(KSUB, wp - c0);
                load_seg_tile(A, KSUB + 4, X.seg[j], row0, n_rows, wc, c0);
                block_cp_async(W, wt + (long long)(koff + c0) * NO, wc * NO);
                cp_async_commit();
                cp_async_wait<0>();
                __syncthreads();
                dense_acc<NO>(acc, A, KSUB + 4, W, wc);
                __syncthreads();
            }
            koff += wp;
        }
        dense_bias_act<NO>(acc, bias, SWE_ACT_NONE, 0.f);
        dense_store_global<NO>(acc, pre, row0, n_rows);
    }
}

// ---------------------------------------------------------------------------------------------
// delta = dh ⊙ act'(pre);  dx (+)= delta · W[:, k_off : k_off+KO);  per-CTA partial sums of delta
// ---------------------------------------------------------------------------------------------
template <int KO>
__global__ void __launch_bounds__(NT) mlp_layer_bwd_dx_kernel(
    float* __restrict__ dh, const float* __restrict__ pre, int act, const float* __restrict__ slope_p, long long n_rows,
    int n, const float* __restrict__ w, int w_ld, int k_off, int k_valid, float* __restrict__ dx, int accumulate,
    int write_delta, float* __restrict__ part) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LDD = KSUB + 4;
    float* D = smem;                       // [TM][KSUB + 4]: delta columns [c0, c0 + KSUB)
    float* Wb = smem + TM * LDD;           // [KSUB][KO]:     W rows     [c0, c0 + KSUB), columns k_off..
    __shared__ double red[NT / 32];
    const float slope = (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
    float db_acc0 = 0.f, db_acc1 = 0.f;                    // bias-gradient partials of columns t and KSUB + t (n <= 2 KSUB)
    // PReLU-slope gradient: ONE scalar summed over rows x columns with heavy cancellation.  Per thread in fp32 (fp64 here
    // cost two F2F.F64 per float4 on the quarter-rate pipe: +2 ms per cfg2-train step, and did not move the error, which
    // comes from the forward's rounding at the PReLU kinks); the cross-thread and cross-CTA sums are fp64.
    float ds_acc = 0.f;
    const long long n_tiles = (n_rows + TM - 1) / TM;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * TM;
        float acc[DenseCfg<KO>::RM][8];
        dense_zero<KO>(acc);
        for (int c0 = 0; c0 < n; c0 += KSUB) {
            const int nc = min(KSUB, n - c0), qpr = nc / 4;
            if (dx) {
                for (int idx = threadIdx.x; idx < nc * KO; idx += NT) {
                    const int nn = idx / KO, k = idx % KO;
                    Wb[idx] = (k < k_valid) ? __ldg(w + (long long)(c0 + nn) * w_ld + k_off + k) : 0.f;
                }
            }
            for (int idx = threadIdx.x; idx < TM * qpr; idx += NT) {
                const int r = idx / qpr, q = idx % qpr;
                const long long g = row0 + r;
                float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
                if (g < n_rows) {
                    d = *reinterpret_cast<const float4*>(dh + g * n + c0 + 4 * q);
                    if (pre) {
                        const float4 p = ldg4(pre + g * n + c0 + 4 * q);
                        if (act == SWE_ACT_PRELU) {
                            ds_acc += (p.x > 0.f ? 0.f : d.x * p.x) + (p.y > 0.f ? 0.f : d.y * p.y) +
                                      (p.z > 0.f ? 0.f : d.z * p.z) + (p.w > 0.f ? 0.f : d.w * p.w);
                        }
                        d.x *= act_grad(act, p.x, slope); d.y *= act_grad(act, p.y, slope);
                        d.z *= act_grad(act, p.z, slope); d.w *= act_grad(act, p.w, slope);
                        if (write_delta) *reinterpret_cast<float4*>(dh + g * n + c0 + 4 * q) = d;
                    }
                }
                stg4(D + r * LDD + 4 * q, d);
            }
            __syncthreads();
            if (part && (int)threadIdx.x < nc) {
                float t = 0.f;
                for (int r = 0; r < TM; ++r) t += D[r * LDD + threadIdx.x];
                if (c0 == 0) db_acc0 += t; else db_acc1 += t;
            }
            if (dx) dense_acc<KO>(acc, D, LDD, Wb, nc);
            __syncthreads();
        }
        if (dx) {
            using C = DenseCfg<KO>;
            const int tx = threadIdx.x % C::TX, ty = threadIdx.x / C::TX;
#pragma unroll
            for (int i = 0; i < C::RM; ++i) {
                const long long r = row0 + ty * C::RM + i;
                if (r < n_rows) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float* y = dx + r * KO + h * (KO / 2) + 4 * tx;
                        float4 v = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
                        if (accumulate) {
                            const float4 o = *reinterpret_cast<const float4*>(y);
                            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                        }
                        stg4(y, v);
                    }
                }
            }
        }
    }
    if (part) {
        float* my = part + (long long)blockIdx.x * (n + 1);
        const int t = (int)threadIdx.x;                       // (signed: n - KSUB is negative for narrow layers)
        if (t < min(n, KSUB)) my[t] = db_acc0;
        if (t < n - KSUB) my[KSUB + t] = db_acc1;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) ds_acc += __shfl_xor_sync(0xffffffffu, ds_acc, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ds_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < NT / 32; ++i) t += red[i];
            my[n] = (float)t;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// part[cta][n_i * KO + k] = Σ_rows delta[r, n_i] · X[r, k]
// ---------------------------------------------------------------------------------------------
template <int KO>
__global__ void __launch_bounds__(NT) mlp_layer_bwd_dw_kernel(const float* __restrict__ delta, long long n_rows, int n,
                                                              const __grid_constant__ swe_rows_t X,
                                                              float* __restrict__ part) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LDR = KSUB + 4;          // KSUB rows (the reduction extent) are staged at a time
    float* Dt = smem;                      // [TM (n index)][KSUB + 4 (row index)]
    float* Xs = smem + TM * LDR;           // [KSUB rows][KO]
    for (int idx = threadIdx.x; idx < TM * LDR; idx += NT) Dt[idx] = 0.f;
    float acc[DenseCfg<KO>::RM][8];
    dense_zero<KO>(acc);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qpr = n / 4;
    const long long n_tiles = (n_rows + KSUB - 1) / KSUB;
    __syncthreads();
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * KSUB;
        // transposed load: lane = row (conflict-free shared stores), loop over 16-byte pieces of the row
        for (int it = warp; it < (KSUB / 32) * qpr; it += NT / 32) {
            const int rb = it / qpr, q = it % qpr;
            const int r = rb * 32 + lane;
            const long long g = row0 + r;
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g < n_rows) d = ldg4(delta + g * n + 4 * q);
            Dt[(4 * q + 0) * LDR + r] = d.x; Dt[(4 * q + 1) * LDR + r] = d.y;
            Dt[(4 * q + 2) * LDR + r] = d.z; Dt[(4 * q + 3) * LDR + r] = d.w;
        }
        load_seg_tile(Xs, KO, X.seg[0], row0, n_rows, KO, 0, KSUB);
        __syncthreads();
        dense_acc<KO>(acc, Dt, LDR, Xs, KSUB);
        __syncthreads();
    }
    using C = DenseCfg<KO>;
    const int tx = threadIdx.x % C::TX, ty = threadIdx.x / C::TX;
    float* my = part + (long long)blockIdx.x * n * KO;
#pragma unroll
    for (int i = 0; i < C::RM; ++i) {
        const int nn = ty * C::RM + i;
        if (nn < n) {
            stg4(my + nn * KO + 4 * tx, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
            stg4(my + nn * KO + KO / 2 + 4 * tx, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
        }
    }
}

__global__ void reduce_partials_kernel(const float* __restrict__ part, int n_parts, long long part_stride, int item_off,
                                       int n_items, int ko, int k_valid, float* __restrict__ out, int ld_out, int k_off) {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_items; j += gridDim.x * blockDim.x) {
        const int k = j % ko;
        if (k >= k_valid) continue;
        // per-CTA partials in CTA order, fp32 with two interleaved accumulators (fp64 here made the kernel conversion-bound:
        // 3.3 instead of 1.5 ms per cfg2-train step, for no measurable change of the gradients)
        // eight interleaved accumulators (a fixed order: partial c goes to accumulator c mod 8): the loop is a chain of dependent
        // loads, and with two accumulators a launch over ~150-300 partials was pure latency (~100 such launches per training step)
        float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float* pj = part + item_off + j;
        int c = 0;
        for (; c + 7 < n_parts; c += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = pj[(long long)(c + u) * part_stride];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] += v[u];
        }
        for (int u = 0; c < n_parts; ++c, ++u) t[u] += pj[(long long)c * part_stride];
        out[(long long)(j / ko) * ld_out + k_off + k] += ((t[0] + t[1]) + (t[2] + t[3])) + ((t[4] + t[5]) + (t[6] + t[7]));
    }
}

// ---------------------------------------------------------------------------------------------
// gate normalisation
// ---------------------------------------------------------------------------------------------
template <int F>
__global__ void __launch_bounds__(NT) gate_norm_fwd_kernel(const float* __restrict__ pre3, int act,
                                                           const float* __restrict__ slope_p, int normalize,
                                                           long long n_edges, float* __restrict__ s_out) {
    constexpr int QPR = F / 4, NG = NT / QPR;
    const float slope = (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
    const int q = threadIdx.x % QPR;
    const long long n_it = (n_edges + NG - 1) / NG;
    for (long long it = blockIdx.x; it < n_it; it += gridDim.x) {
        const long long e = it * NG + threadIdx.x / QPR;
        float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < n_edges) {
            u = ldg4_stream(pre3 + e * F + 4 * q);
            u.x = act_apply(act, u.x, slope); u.y = act_apply(act, u.y, slope);
            u.z = act_apply(act, u.z, slope); u.w = act_apply(act, u.w, slope);
        }
        if (normalize) {
            float ss = u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
#pragma unroll
            for (int off = QPR / 2; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
            const float nrm = sqrtf(ss);
            u.x = __fdiv_rn(u.x, nrm); u.y = __fdiv_rn(u.y, nrm); u.z = __fdiv_rn(u.z, nrm); u.w = __fdiv_rn(u.w, nrm);
            u.x = (u.x != u.x) ? 0.f : u.x; u.y = (u.y != u.y) ? 0.f : u.y;
            u.z = (u.z != u.z) ? 0.f : u.z; u.w = (u.w != u.w) ? 0.f : u.w;
        }
        if (e < n_edges) stg4(s_out + e * F + 4 * q, u);
    }
}

template <int F>
__global__ void __launch_bounds__(NT) gate_norm_bwd_kernel(float* __restrict__ ds, const float* __restrict__ pre3, int act,
                                                           const float* __restrict__ slope_p, long long n_edges) {
    constexpr int QPR = F / 4, NG = NT / QPR;
    const float slope = (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
    const int q = threadIdx.x % QPR;
    const long long n_it = (n_edges + NG - 1) / NG;
    for (long long it = blockIdx.x; it < n_it; it += gridDim.x) {
        const long long e = it * NG + threadIdx.x / QPR;
        float4 u = make_float4(0.f, 0.f, 0.f, 0.f), g = u;
        if (e < n_edges) {
            u = ldg4_stream(pre3 + e * F + 4 * q);
            g = *reinterpret_cast<const float4*>(ds + e * F + 4 * q);
            u.x = act_apply(act, u.x, slope); u.y = act_apply(act, u.y, slope);
            u.z = act_apply(act, u.z, slope); u.w = act_apply(act, u.w, slope);
        }
        float ss = u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
        float ug = u.x * g.x + u.y * g.y + u.z * g.z + u.w * g.w;
#pragma unroll
        for (int off = QPR / 2; off >= 1; off >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, off);
            ug += __shfl_xor_sync(0xffffffffu, ug, off);
        }
        if (e < n_edges) {
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ss > 0.f) {
                const float inv = 1.f / sqrtf(ss);
                const float c = ug / ss;                                // (s·ds)/||u|| with s = u/||u||
                r.x = (g.x - u.x * c) * inv; r.y = (g.y - u.y * c) * inv;
                r.z = (g.z - u.z * c) * inv; r.w = (g.w - u.w * c) * inv;
            }
            *reinterpret_cast<float4*>(ds + e * F + 4 * q) = r;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// elementwise activation over a row range
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) act_fwd_kernel(const float* __restrict__ x, long long n4, int act,
                                                     const float* __restrict__ slope_p, float* __restrict__ y) {
    const float slope = (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
        float4 v = ldg4(x + 4 * i);
        v.x = act_apply(act, v.x, slope); v.y = act_apply(act, v.y, slope);
        v.z = act_apply(act, v.z, slope); v.w = act_apply(act, v.w, slope);
        stg4(y + 4 * i, v);
    }
}

__global__ void __launch_bounds__(NT) act_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x, long long n4,
                                                     int act, const float* __restrict__ slope_p, float* __restrict__ gx,
                                                     float* __restrict__ slope_part) {
    __shared__ double red[NT / 32];
    const float slope = (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
    float ds_acc = 0.f;                                      // per thread fp32, across threads / CTAs fp64 (see mlp_layer_bwd_dx)
    for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n4; i += (long long)gridDim.x * NT) {
        const float4 p = ldg4(x + 4 * i);
        float4 d = ldg4(g + 4 * i);
        if (act == SWE_ACT_PRELU)
            ds_acc += (p.x > 0.f ? 0.f : d.x * p.x) + (p.y > 0.f ? 0.f : d.y * p.y) +
                      (p.z > 0.f ? 0.f : d.z * p.z) + (p.w > 0.f ? 0.f : d.w * p.w);
        d.x *= act_grad(act, p.x, slope); d.y *= act_grad(act, p.y, slope);
        d.z *= act_grad(act, p.z, slope); d.w *= act_grad(act, p.w, slope);
        stg4(gx + 4 * i, d);
    }
    if (slope_part) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) ds_acc += __shfl_xor_sync(0xffffffffu, ds_acc, off);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ds_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < NT / 32; ++i) t += red[i];
            slope_part[blockIdx.x] = (float)t;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// hop backward
// ---------------------------------------------------------------------------------------------
template <int F>
__global__ void __launch_bounds__(NT) row_flags_kernel(const float* __restrict__ o, int row_lo, int n_rows,
                                                       uint8_t* __restrict__ flags) {
    constexpr int QPR = F / 4, NG = NT / QPR;
    const int q = threadIdx.x % QPR;
    const int n_it = (n_rows + NG - 1) / NG;
    for (int it = blockIdx.x; it < n_it; it += gridDim.x) {
        const int i = it * NG + threadIdx.x / QPR;
        float t = 0.f;
        if (i < n_rows) {
            const float4 v = ldg4(o + ((long long)row_lo + i) * F + 4 * q);
            t = (v.x + v.y) + (v.z + v.w);
        }
#pragma unroll
        for (int off = QPR / 2; off >= 1; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (i < n_rows && q == 0) flags[row_lo + i] = (t != 0.f) ? 1 : 0;
    }
}

template <int F>
__global__ void __launch_bounds__(NT) hop_bwd_dst_kernel(
    const float* __restrict__ da, const float* __restrict__ o_src, const float* __restrict__ o_dst,
    const float* __restrict__ s, float* __restrict__ ds, int accumulate_ds, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ src, const uint8_t* __restrict__ wet_src, const uint8_t* __restrict__ wet_dst,
    int dst_lo, int n_dst, int with_gradient, const float* __restrict__ g_next, float* __restrict__ g_part) {
    constexpr int QPR = F / 4, NG = NT / QPR;
    const int q = threadIdx.x % QPR;
    for (long long i = (long long)blockIdx.x * NG + threadIdx.x / QPR; i < n_dst; i += (long long)gridDim.x * NG) {
        const long long c = (long long)dst_lo + i;
        const float4 dac = ldg4(da + c * F + 4 * q);
        float4 oc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (o_dst) oc = ldg4(o_dst + c * F + 4 * q);
        const int wc = wet_dst ? wet_dst[c] : 0;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int p0 = __ldg(rowptr + i), p1 = __ldg(rowptr + i + 1);
        for (int p = p0; p < p1; ++p) {
            const int r = __ldg(src + p);
            const float a = (wc | wet_src[r]) ? 1.f : 0.f;
            const float4 orow = ldg4(o_src + (long long)r * F + 4 * q);
            float4 d;
            if (with_gradient) {
                d = make_float4(oc.x - orow.x, oc.y - orow.y, oc.z - orow.z, oc.w - orow.w);
                const float4 sv = ldg4_stream(s + (long long)p * F + 4 * q);
                acc.x += a * sv.x; acc.y += a * sv.y; acc.z += a * sv.z; acc.w += a * sv.w;
            } else {
                d = orow;
            }
            float4 t = make_float4(a * dac.x * d.x, a * dac.y * d.y, a * dac.z * d.z, a * dac.w * d.w);
            float* dsp = ds + (long long)p * F + 4 * q;
            if (accumulate_ds) {
                const float4 o = *reinterpret_cast<const float4*>(dsp);
                t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
            }
            stg4(dsp, t);
        }
        if (with_gradient) {
            const float4 gn = ldg4(g_next + c * F + 4 * q);
            stg4(g_part + c * F + 4 * q, make_float4(gn.x + dac.x * acc.x, gn.y + dac.y * acc.y,
                                                     gn.z + dac.z * acc.z, gn.w + dac.w * acc.w));
        }
    }
}

template <int F>
__global__ void __launch_bounds__(NT) hop_bwd_src_kernel(
    const float* __restrict__ da, const float* __restrict__ s, const int32_t* __restrict__ t_rowptr,
    const int32_t* __restrict__ t_pos, const int32_t* __restrict__ dst, const uint8_t* __restrict__ wet_src,
    const uint8_t* __restrict__ wet_dst, int src_lo, int n_src, int with_gradient, int accumulate,
    float* __restrict__ g_io) {
    constexpr int QPR = F / 4, NG = NT / QPR;
    const int q = threadIdx.x % QPR;
    for (long long i = (long long)blockIdx.x * NG + threadIdx.x / QPR; i < n_src; i += (long long)gridDim.x * NG) {
        const long long n = (long long)src_lo + i;
        const int wn = wet_src[n];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int q0 = __ldg(t_rowptr + i), q1 = __ldg(t_rowptr + i + 1);
        for (int qq = q0; qq < q1; ++qq) {
            const int p = __ldg(t_pos + qq);
            const int c = __ldg(dst + p);
            const float a = (wn | (wet_dst ? wet_dst[c] : 0)) ? 1.f : 0.f;
            const float4 sv = ldg4(s + (long long)p * F + 4 * q);
            const float4 dc = ldg4(da + (long long)c * F + 4 * q);
            acc.x += a * sv.x * dc.x; acc.y += a * sv.y * dc.y; acc.z += a * sv.z * dc.z; acc.w += a * sv.w * dc.w;
        }
        float* gp = g_io + n * F + 4 * q;
        float4 v;
        if (with_gradient) {
            const float4 o = *reinterpret_cast<const float4*>(gp);
            v = make_float4(o.x - acc.x, o.y - acc.y, o.z - acc.z, o.w - acc.w);
        } else if (accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(gp);
            v = make_float4(o.x + acc.x, o.y + acc.y, o.z + acc.z, o.w + acc.w);
        } else {
            v = acc;
        }
        stg4(gp, v);
    }
}

template <int F>
__global__ void __launch_bounds__(NT) edge_to_node_sum_kernel(const float* __restrict__ e, const int32_t* __restrict__ rowptr,
                                                              const int32_t* __restrict__ pos, int node_lo, int n_nodes,
                                                              float* __restrict__ out, int accumulate) {
    constexpr int QPR = F / 4, NG = NT / QPR;
    const int q = threadIdx.x % QPR;
    for (long long i = (long long)blockIdx.x * NG + threadIdx.x / QPR; i < n_nodes; i += (long long)gridDim.x * NG) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int p0 = __ldg(rowptr + i), p1 = __ldg(rowptr + i + 1);
        for (int p = p0; p < p1; ++p) {
            const long long pp = pos ? __ldg(pos + p) : p;
            const float4 v = ldg4_stream(e + pp * F + 4 * q);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float* op = out + ((long long)node_lo + i) * F + 4 * q;
        if (accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(op);
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        stg4(op, acc);
    }
}

template <int F>
__global__ void __launch_bounds__(NT) pool_mean_bwd_kernel(const float* __restrict__ g, const int32_t* __restrict__ f_rowptr,
                                                           const int32_t* __restrict__ coarse, int fine_lo, int n_fine,
                                                           const int32_t* __restrict__ pool_rowptr, int coarse_lo,
                                                           float* __restrict__ dx, int accumulate) {
    constexpr int QPR = F / 4, NG = NT / QPR;
    const int q = threadIdx.x % QPR;
    for (long long i = (long long)blockIdx.x * NG + threadIdx.x / QPR; i < n_fine; i += (long long)gridDim.x * NG) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int p0 = __ldg(f_rowptr + i), p1 = __ldg(f_rowptr + i + 1);
        for (int p = p0; p < p1; ++p) {
            const int c = __ldg(coarse + p);
            const int cnt = max(__ldg(pool_rowptr + (c - coarse_lo) + 1) - __ldg(pool_rowptr + (c - coarse_lo)), 1);
            const float4 v = ldg4(g + (long long)c * F + 4 * q);
            const float fc = (float)cnt;
            acc.x += __fdiv_rn(v.x, fc); acc.y += __fdiv_rn(v.y, fc); acc.z += __fdiv_rn(v.z, fc); acc.w += __fdiv_rn(v.w, fc);
        }
        float* op = dx + ((long long)fine_lo + i) * F + 4 * q;
        if (accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(op);
            acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
        }
        stg4(op, acc);
    }
}

// ---------------------------------------------------------------------------------------------
// encoder inputs / head
// ---------------------------------------------------------------------------------------------
__global__ void static_inputs_fwd_kernel(const float* __restrict__ x, int n_cols, const int32_t* __restrict__ perm,
                                         int n_nodes, int n_static_raw, int with_wl, float* __restrict__ xin, int ks) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)n_nodes * ks;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / ks), c = (int)(idx % ks);
        const long long srow = (long long)(perm ? perm[i] : i) * n_cols;
        float v = 0.f;
        if (c < n_static_raw) v = x[srow + c];
        else if (c == n_static_raw && with_wl) v = x[srow + n_static_raw - 1] + x[srow + n_cols - 2];
        xin[idx] = v;
    }
}

__global__ void node_inputs_bwd_kernel(const float* __restrict__ dxs, int ks, const float* __restrict__ dxd, int kd,
                                       int n_cols, const int32_t* __restrict__ perm, int n_nodes, int n_dyn_rows,
                                       int n_static_raw, int with_wl, float* __restrict__ dx) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
        float* row = dx + (long long)(perm ? perm[i] : i) * n_cols;
        for (int c = 0; c < n_static_raw; ++c) row[c] += dxs[(long long)i * ks + c];
        if (with_wl) {
            const float w = dxs[(long long)i * ks + n_static_raw];
            row[n_static_raw - 1] += w;
            row[n_cols - 2] += w;
        }
        if (i < n_dyn_rows)
            for (int c = 0; c < n_cols - n_static_raw; ++c) row[n_static_raw + c] += dxd[(long long)i * kd + c];
    }
}

__device__ __forceinline__ float head_residual(const float* xr, int n_static_raw, int n_cols, int previous_t, int res_mode,
                                               const float* __restrict__ res_w, int j) {
    float res = 0.f;
    if (res_mode == 1) {
        for (int t = 0; t < previous_t; ++t) res = fmaf(xr[n_static_raw + 2 * t + j], __ldg(res_w + t), res);
    } else if (res_mode == 2) {
        for (int t = 0; t < previous_t; ++t) res = fmaf(xr[n_static_raw + 2 * t + j], __ldg(res_w + 2 * t + j), res);
    } else if (res_mode == 3) {
        res = xr[n_cols - 2 + j];
    }
    return res;
}

__global__ void __launch_bounds__(NT) head_fwd_kernel(const float* __restrict__ pre3, int ldp, int act,
                                                      const float* __restrict__ slope_p, const float* __restrict__ x0,
                                                      int n_cols, const int32_t* __restrict__ perm, int n_nodes,
                                                      int previous_t, int res_mode, const float* __restrict__ res_w,
                                                      float eps, float* __restrict__ pred) {
    const float slope = (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
    const int n_static_raw = n_cols - 2 * previous_t;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
        const long long orow = perm ? perm[i] : i;
        const float* xr = x0 + orow * n_cols;
        float y[2];
        for (int j = 0; j < 2; ++j) {
            const float v = act_apply(act, pre3[(long long)i * ldp + j], slope);
            y[j] = fmaxf(v + head_residual(xr, n_static_raw, n_cols, previous_t, res_mode, res_w, j), 0.f);
        }
        pred[orow * 2 + 0] = (fabsf(y[0]) > eps) ? y[0] : 0.f;
        pred[orow * 2 + 1] = (y[0] != 0.f) ? y[1] : 0.f;
    }
}

// res_part[cta][2*previous_t] laid out [t][var]
__global__ void __launch_bounds__(NT) head_bwd_kernel(const float* __restrict__ dpred, const float* __restrict__ pre3, int ldp,
                                                      int act, const float* __restrict__ slope_p,
                                                      const float* __restrict__ x0, int n_cols,
                                                      const int32_t* __restrict__ perm, int n_nodes, int previous_t,
                                                      int res_mode, const float* __restrict__ res_w, float eps,
                                                      float* __restrict__ dh3, float* __restrict__ dx0,
                                                      float* __restrict__ res_part) {
    __shared__ float red[NT / 32][16];
    const float slope = (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
    (void)slope;
    const int n_static_raw = n_cols - 2 * previous_t;
    float racc[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) racc[t] = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += gridDim.x * blockDim.x) {
        const long long orow = perm ? perm[i] : i;
        const float* xr = x0 + orow * n_cols;
        float z[2], y[2];
        for (int j = 0; j < 2; ++j) {
            const float v = act_apply(act, pre3[(long long)i * ldp + j], (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f);
            z[j] = v + head_residual(xr, n_static_raw, n_cols, previous_t, res_mode, res_w, j);
            y[j] = fmaxf(z[j], 0.f);
        }
        float dz[2];
        dz[0] = (fabsf(y[0]) > eps && z[0] > 0.f) ? dpred[orow * 2 + 0] : 0.f;
        dz[1] = (y[0] != 0.f && z[1] > 0.f) ? dpred[orow * 2 + 1] : 0.f;
        for (int c = 0; c < ldp; ++c) dh3[(long long)i * ldp + c] = c < 2 ? dz[c] : 0.f;
        if (res_mode == 1 || res_mode == 2) {
            for (int t = 0; t < previous_t && t < 8; ++t)
                for (int j = 0; j < 2; ++j) {
                    racc[2 * t + j] += xr[n_static_raw + 2 * t + j] * dz[j];
                    if (dx0) dx0[orow * n_cols + n_static_raw + 2 * t + j] += dz[j] * __ldg(res_w + (res_mode == 1 ? t : 2 * t + j));
                }
        } else if (res_mode == 3 && dx0) {
            dx0[orow * n_cols + n_cols - 2] += dz[0];
            dx0[orow * n_cols + n_cols - 1] += dz[1];
        }
    }
    if (res_part) {
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            float v = racc[t];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][t] = v;
        }
        __syncthreads();
        if (threadIdx.x < 16) {
            float t = 0.f;
            for (int w = 0; w < NT / 32; ++w) t += red[w][threadIdx.x];
            res_part[(long long)blockIdx.x * 16 + threadIdx.x] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host dispatch helpers
// ---------------------------------------------------------------------------------------------
template <typename K>
static int opt_in(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(%zu B): %s", bytes, cudaGetErrorString(e)); return (int)e; }
    return 0;
}

#define SWE_DISPATCH_W(W_, ...)                                   \
    switch (W_) {                                                 \
        case 16:  { constexpr int WW = 16;  __VA_ARGS__; } break; \
        case 32:  { constexpr int WW = 32;  __VA_ARGS__; } break; \
        case 64:  { constexpr int WW = 64;  __VA_ARGS__; } break; \
        case 128: { constexpr int WW = 128; __VA_ARGS__; } break; \
        default: set_error("unsupported tile width %d (16, 32, 64, 128)", W_); return SWE_E_UNSUPP; \
    }
#define SWE_DISPATCH_FB(F_, ...)                                \
    switch (F_) {                                               \
        case 16: { constexpr int FF = 16; __VA_ARGS__; } break; \
        case 32: { constexpr int FF = 32; __VA_ARGS__; } break; \
        case 64: { constexpr int FF = 64; __VA_ARGS__; } break; \
        default: set_error("unsupported feature width F=%d (16, 32, 64)", F_); return SWE_E_UNSUPP; \
    }

static int check_rows(const swe_rows_t* X, const char* what) {
    SWE_REQUIRE(X && X->n_seg >= 1 && X->n_seg <= SWE_MAX_SEGS, SWE_E_INVAL, "%s: bad segment count", what);
    for (int j = 0; j < X->n_seg; ++j) {
        const swe_seg_t& sg = X->seg[j];
        SWE_REQUIRE(sg.base && sg.width >= 1 && sg.width <= 128 && sg.ld >= sg.width, SWE_E_INVAL,
                    "%s: segment %d base/width/ld invalid (width %d, ld %d)", what, j, sg.width, sg.ld);
    }
    return 0;
}

static int rows_grid(long long n_rows, int per_sm) { return grid_for((n_rows + TM - 1) / TM, per_sm); }

}  // namespace swe

using namespace swe;

extern "C" int swe_mlp_layer_fwd(const swe_rows_t* X, int64_t n_rows, const float* wt, const float* bias,
                                 int32_t n_out, float* pre, void* stream) {
    if (int r = check_rows(X, "mlp_layer_fwd")) return r;
    SWE_REQUIRE(wt && pre && n_rows >= 0, SWE_E_INVAL, "mlp_layer_fwd: bad arguments");
    SWE_REQUIRE(aligned16(wt) && aligned16(pre) && (!bias || aligned16(bias)), SWE_E_ALIGN, "mlp_layer_fwd: unaligned buffer");
    if (n_rows == 0) return 0;
    SWE_DISPATCH_W(n_out, {
        auto k = mlp_layer_fwd_kernel<WW>;
        const size_t bytes = sizeof(float) * (size_t)(TM * (KSUB + 4) + KSUB * WW);
        if (int r = opt_in(k, bytes)) return r;
        k<<<rows_grid(n_rows, 3), NT, bytes, (cudaStream_t)stream>>>(*X, n_rows, wt, bias, pre);
    });
    return check_launch("mlp_layer_fwd");
}

static int dw_grid(long long n_rows) { return grid_for((n_rows + KSUB - 1) / KSUB, 3); }
extern "C" int swe_mlp_layer_bwd_dx_grid(int64_t n_rows) { return rows_grid(n_rows, 3); }
extern "C" int swe_mlp_layer_bwd_dw_grid(int64_t n_rows) { return dw_grid(n_rows); }

extern "C" int swe_mlp_layer_bwd_dx(float* dh, const float* pre, int32_t act, const float* slope, int64_t n_rows,
                                    int32_t n, const float* w, int32_t w_ld, int32_t k_off, int32_t k_valid, int32_t ko,
                                    float* dx, int32_t accumulate, int32_t write_delta, float* part, int32_t* grid_out,
                                    void* stream) {
    SWE_REQUIRE(dh && n_rows >= 0 && n >= 4 && n <= 128 && n % 4 == 0, SWE_E_INVAL, "mlp_layer_bwd_dx: bad arguments (n=%d)", n);
    SWE_REQUIRE(!dx || (w && w_ld >= 1 && k_off >= 0 && k_valid >= 1 && k_valid <= ko), SWE_E_INVAL,
                "mlp_layer_bwd_dx: bad weight block");
    SWE_REQUIRE(aligned16(dh) && (!pre || aligned16(pre)) && (!dx || aligned16(dx)), SWE_E_ALIGN, "mlp_layer_bwd_dx: unaligned buffer");
    const int grid = rows_grid(n_rows, 3);
    if (grid_out) *grid_out = grid;
    if (n_rows == 0) return 0;
    SWE_DISPATCH_W(ko, {
        auto k = mlp_layer_bwd_dx_kernel<WW>;
        const size_t bytes = sizeof(float) * (size_t)(TM * (KSUB + 4) + KSUB * WW);
        if (int r = opt_in(k, bytes)) return r;
        k<<<grid, NT, bytes, (cudaStream_t)stream>>>(dh, pre, act, slope, n_rows, n, w, w_ld, k_off, k_valid, dx,
                                                      accumulate, write_delta, part);
    });
    return check_launch("mlp_layer_bwd_dx");
}

extern "C" int swe_mlp_layer_bwd_dw(const float* delta, int64_t n_rows, int32_t n, const swe_rows_t* X, int32_t ko,
                                    float* part, int32_t* grid_out, void* stream) {
    if (int r = check_rows(X, "mlp_layer_bwd_dw")) return r;
    SWE_REQUIRE(X->n_seg == 1 && X->seg[0].width <= ko, SWE_E_INVAL, "mlp_layer_bwd_dw: needs one segment no wider than ko");
    SWE_REQUIRE(delta && part && n_rows >= 0 && n >= 4 && n <= 128 && n % 4 == 0, SWE_E_INVAL, "mlp_layer_bwd_dw: bad arguments");
    SWE_REQUIRE(aligned16(delta) && aligned16(part), SWE_E_ALIGN, "mlp_layer_bwd_dw: unaligned buffer");
    const int grid = dw_grid(n_rows);
    if (grid_out) *grid_out = grid;
    SWE_DISPATCH_W(ko, {
        auto k = mlp_layer_bwd_dw_kernel<WW>;
        const size_t bytes = sizeof(float) * (size_t)(TM * (KSUB + 4) + KSUB * WW);
        if (int r = opt_in(k, bytes)) return r;
        k<<<grid, NT, bytes, (cudaStream_t)stream>>>(delta, n_rows, n, *X, part);
    });
    return check_launch("mlp_layer_bwd_dw");
}

extern "C" int swe_reduce_partials(const float* part, int32_t n_parts, int64_t part_stride, int32_t item_off,
                                   int32_t n_items, int32_t ko, int32_t k_valid, float* out, int32_t ld_out,
                                   int32_t k_off, void* stream) {
    SWE_REQUIRE(part && out && n_parts >= 0 && n_items >= 0 && ko >= 1, SWE_E_INVAL, "reduce_partials: bad arguments");
    if (n_items == 0 || n_parts == 0) return 0;
    reduce_partials_kernel<<<(n_items + 255) / 256, 256, 0, (cudaStream_t)stream>>>(part, n_parts, part_stride, item_off,
                                                                                    n_items, ko, k_valid, out, ld_out, k_off);
    return check_launch("reduce_partials");
}

extern "C" int swe_gate_norm_fwd(const float* pre3, int32_t act, const float* slope, int32_t normalize, int64_t n_edges,
                                 float* s_out, int32_t F, void* stream) {
    SWE_REQUIRE(pre3 && s_out && n_edges >= 0, SWE_E_INVAL, "gate_norm_fwd: bad arguments");
    SWE_REQUIRE(aligned16(pre3) && aligned16(s_out), SWE_E_ALIGN, "gate_norm_fwd: unaligned buffer");
    if (n_edges == 0) return 0;
    SWE_DISPATCH_FB(F, {
        constexpr int NG = NT / (FF / 4);
        gate_norm_fwd_kernel<FF><<<grid_for((n_edges + NG - 1) / NG, 8), NT, 0, (cudaStream_t)stream>>>(
            pre3, act, slope, normalize, n_edges, s_out);
    });
    return check_launch("gate_norm_fwd");
}

extern "C" int swe_gate_norm_bwd(float* ds, const float* pre3, int32_t act, const float* slope, int32_t normalize,
                                 int64_t n_edges, int32_t F, void* stream) {
    SWE_REQUIRE(ds && pre3 && n_edges >= 0, SWE_E_INVAL, "gate_norm_bwd: bad arguments");
    SWE_REQUIRE(aligned16(pre3) && aligned16(ds), SWE_E_ALIGN, "gate_norm_bwd: unaligned buffer");
    if (n_edges == 0 || !normalize) return 0;
    SWE_DISPATCH_FB(F, {
        constexpr int NG = NT / (FF / 4);
        gate_norm_bwd_kernel<FF><<<grid_for((n_edges + NG - 1) / NG, 8), NT, 0, (cudaStream_t)stream>>>(
            ds, pre3, act, slope, n_edges);
    });
    return check_launch("gate_norm_bwd");
}

extern "C" int swe_act_fwd(const float* x, int32_t row_lo, int32_t n_rows, int32_t act, const float* slope, float* y,
                           int32_t F, void* stream) {
    SWE_REQUIRE(x && y && row_lo >= 0 && n_rows >= 0 && F % 4 == 0, SWE_E_INVAL, "act_fwd: bad arguments");
    if (n_rows == 0) return 0;
    const long long n4 = (long long)n_rows * F / 4;
    act_fwd_kernel<<<grid_for((n4 + NT - 1) / NT, 8), NT, 0, (cudaStream_t)stream>>>(x + (long long)row_lo * F, n4, act, slope,
                                                                                    y + (long long)row_lo * F);
    return check_launch("act_fwd");
}

extern "C" int swe_act_bwd(const float* g, const float* x, int32_t row_lo, int32_t n_rows, int32_t act,
                           const float* slope, float* gx, float* slope_part, int32_t* grid_out, int32_t F, void* stream) {
    SWE_REQUIRE(g && x && gx && row_lo >= 0 && n_rows >= 0 && F % 4 == 0, SWE_E_INVAL, "act_bwd: bad arguments");
    const long long n4 = (long long)n_rows * F / 4;
    const int grid = grid_for((n4 + NT - 1) / NT, 4);
    if (grid_out) *grid_out = grid;
    if (n_rows == 0) return 0;
    const long long off = (long long)row_lo * F;
    act_bwd_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(g + off, x + off, n4, act, slope, gx + off, slope_part);
    return check_launch("act_bwd");
}

extern "C" int swe_row_flags(const float* o, int32_t row_lo, int32_t n_rows, uint8_t* flags, int32_t F, void* stream) {
    SWE_REQUIRE(o && flags && row_lo >= 0 && n_rows >= 0, SWE_E_INVAL, "row_flags: bad arguments");
    if (n_rows == 0) return 0;
    SWE_DISPATCH_FB(F, {
        constexpr int NG = NT / (FF / 4);
        row_flags_kernel<FF><<<grid_for((n_rows + NG - 1) / NG, 8), NT, 0, (cudaStream_t)stream>>>(o, row_lo, n_rows, flags);
    });
    return check_launch("row_flags");
}

extern "C" int swe_hop_bwd_dst(const float* da, const float* o_src, const float* o_dst, const float* s, float* ds,
                               int32_t accumulate_ds, const int32_t* rowptr, const int32_t* src, const uint8_t* wet_src,
                               const uint8_t* wet_dst, int32_t dst_lo, int32_t n_dst, int32_t with_gradient,
                               const float* g_next, float* g_part, int32_t F, void* stream) {
    SWE_REQUIRE(da && o_src && s && ds && rowptr && src && wet_src && dst_lo >= 0 && n_dst >= 0, SWE_E_INVAL,
                "hop_bwd_dst: bad arguments");
    SWE_REQUIRE(!with_gradient || (o_dst && g_next && g_part), SWE_E_INVAL, "hop_bwd_dst: with_gradient needs o_dst, g_next, g_part");
    if (n_dst == 0) return 0;
    SWE_DISPATCH_FB(F, {
        constexpr int NG = NT / (FF / 4);
        hop_bwd_dst_kernel<FF><<<grid_for((n_dst + NG - 1) / NG, 8), NT, 0, (cudaStream_t)stream>>>(
            da, o_src, o_dst, s, ds, accumulate_ds, rowptr, src, wet_src, wet_dst, dst_lo, n_dst, with_gradient, g_next, g_part);
    });
    return check_launch("hop_bwd_dst");
}

extern "C" int swe_hop_bwd_src(const float* da, const float* s, const int32_t* t_rowptr, const int32_t* t_pos,
                               const int32_t* dst, const uint8_t* wet_src, const uint8_t* wet_dst, int32_t src_lo,
                               int32_t n_src, int32_t with_gradient, int32_t accumulate, float* g_io, int32_t F,
                               void* stream) {
    SWE_REQUIRE(da && s && t_rowptr && t_pos && dst && wet_src && g_io && src_lo >= 0 && n_src >= 0, SWE_E_INVAL,
                "hop_bwd_src: bad arguments");
    if (n_src == 0) return 0;
    SWE_DISPATCH_FB(F, {
        constexpr int NG = NT / (FF / 4);
        hop_bwd_src_kernel<FF><<<grid_for((n_src + NG - 1) / NG, 8), NT, 0, (cudaStream_t)stream>>>(
            da, s, t_rowptr, t_pos, dst, wet_src, wet_dst, src_lo, n_src, with_gradient, accumulate, g_io);
    });
    return check_launch("hop_bwd_src");
}

extern "C" int swe_edge_to_node_sum(const float* e, const int32_t* rowptr, const int32_t* pos, int32_t node_lo,
                                    int32_t n_nodes, float* out, int32_t accumulate, int32_t F, void* stream) {
    SWE_REQUIRE(e && rowptr && out && node_lo >= 0 && n_nodes >= 0, SWE_E_INVAL, "edge_to_node_sum: bad arguments");
    if (n_nodes == 0) return 0;
    SWE_DISPATCH_FB(F, {
        constexpr int NG = NT / (FF / 4);
        edge_to_node_sum_kernel<FF><<<grid_for((n_nodes + NG - 1) / NG, 8), NT, 0, (cudaStream_t)stream>>>(
            e, rowptr, pos, node_lo, n_nodes, out, accumulate);
    });
    return check_launch("edge_to_node_sum");
}

extern "C" int swe_pool_mean_bwd(const float* g, const int32_t* f_rowptr, const int32_t* coarse, int32_t fine_lo,
                                 int32_t n_fine, const int32_t* pool_rowptr, int32_t coarse_lo, float* dx,
                                 int32_t accumulate, int32_t F, void* stream) {
    SWE_REQUIRE(g && f_rowptr && coarse && pool_rowptr && dx && fine_lo >= 0 && n_fine >= 0, SWE_E_INVAL,
                "pool_mean_bwd: bad arguments");
    if (n_fine == 0) return 0;
    SWE_DISPATCH_FB(F, {
        constexpr int NG = NT / (FF / 4);
        pool_mean_bwd_kernel<FF><<<grid_for((n_fine + NG - 1) / NG, 8), NT, 0, (cudaStream_t)stream>>>(
            g, f_rowptr, coarse, fine_lo, n_fine, pool_rowptr, coarse_lo, dx, accumulate);
    });
    return check_launch("pool_mean_bwd");
}

extern "C" int swe_static_inputs_fwd(const float* x, int32_t n_cols, const int32_t* perm, int32_t n_nodes,
                                     int32_t n_static_raw, int32_t with_wl, float* xin_s, int32_t ks, void* stream) {
    SWE_REQUIRE(x && xin_s && n_nodes >= 0 && n_static_raw >= 1 && ks >= n_static_raw + (with_wl ? 1 : 0), SWE_E_INVAL,
                "static_inputs_fwd: bad arguments");
    if (n_nodes == 0) return 0;
    const long long total = (long long)n_nodes * ks;
    static_inputs_fwd_kernel<<<grid_for((total + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(
        x, n_cols, perm, n_nodes, n_static_raw, with_wl, xin_s, ks);
    return check_launch("static_inputs_fwd");
}

extern "C" int swe_node_inputs_bwd(const float* dxin_s, int32_t ks, const float* dxin_d, int32_t kd, int32_t n_cols,
                                   const int32_t* perm, int32_t n_nodes, int32_t n_dyn_rows, int32_t n_static_raw,
                                   int32_t with_wl, float* dx, void* stream) {
    SWE_REQUIRE(dxin_s && dxin_d && dx && n_nodes >= 0 && n_dyn_rows <= n_nodes, SWE_E_INVAL, "node_inputs_bwd: bad arguments");
    if (n_nodes == 0) return 0;
    node_inputs_bwd_kernel<<<grid_for((n_nodes + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(
        dxin_s, ks, dxin_d, kd, n_cols, perm, n_nodes, n_dyn_rows, n_static_raw, with_wl, dx);
    return check_launch("node_inputs_bwd");
}

extern "C" int swe_head_fwd(const float* pre3, int32_t ldp, int32_t act, const float* slope, const float* x0,
                            int32_t n_cols, const int32_t* perm, int32_t n_nodes, int32_t previous_t, int32_t res_mode,
                            const float* res_w, float eps, float* pred, void* stream) {
    SWE_REQUIRE(pre3 && x0 && pred && ldp >= 2 && n_nodes >= 0 && previous_t >= 1 && n_cols > 2 * previous_t, SWE_E_INVAL,
                "head_fwd: bad arguments");
    SWE_REQUIRE(res_mode >= 0 && res_mode <= 3 && (res_mode == 0 || res_mode == 3 || res_w), SWE_E_INVAL, "head_fwd: bad residual mode");
    if (n_nodes == 0) return 0;
    head_fwd_kernel<<<grid_for((n_nodes + NT - 1) / NT, 8), NT, 0, (cudaStream_t)stream>>>(
        pre3, ldp, act, slope, x0, n_cols, perm, n_nodes, previous_t, res_mode, res_w, eps, pred);
    return check_launch("head_fwd");
}

extern "C" int swe_head_bwd(const float* dpred, const float* pre3, int32_t ldp, int32_t act, const float* slope,
                            const float* x0, int32_t n_cols, const int32_t* perm, int32_t n_nodes, int32_t previous_t,
                            int32_t res_mode, const float* res_w, float eps, float* dh3, float* dx0, float* res_part,
                            int32_t* grid_out, void* stream) {
    SWE_REQUIRE(dpred && pre3 && x0 && dh3 && ldp >= 2 && n_nodes >= 0 && previous_t >= 1 && previous_t <= 8 &&
                n_cols > 2 * previous_t, SWE_E_INVAL, "head_bwd: bad arguments (previous_t <= 8)");
    SWE_REQUIRE(res_mode >= 0 && res_mode <= 3 && (res_mode == 0 || res_mode == 3 || res_w), SWE_E_INVAL, "head_bwd: bad residual mode");
    const int grid = grid_for((n_nodes + NT - 1) / NT, 2);
    if (grid_out) *grid_out = grid;
    if (n_nodes == 0) return 0;
    head_bwd_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(dpred, pre3, ldp, act, slope, x0, n_cols, perm, n_nodes,
                                                          previous_t, res_mode, res_w, eps, dh3, dx0, res_part);
    return check_launch("head_bwd");
}
