// Forward kernels of the mSWE-GNN hot path (exact-fp32 CUDA-core path) and their C-ABI entries.
// Reference spans each kernel replaces are cited in include/swe_gnn_b200.h.
#include "swe_dense.cuh"

namespace swe {

// =============================================================================================
// helpers shared by the row-MLP kernels (encoders / decoder)
// =============================================================================================

// Runs layers [l0, l1) of `m`, all of output width F, on a TM-row tile held in shared memory.
// X: input tile (leading dimension ldx, K = layer[l0].k_in columns).  Ha/Hb: two [TM][F+4]
// buffers, Wb: weight staging buffer of max(F,32)*F floats.  Returns the buffer holding the
// result (X itself when l0 == l1).  All threads of the CTA must call it.
template <int F>
__device__ float* mlp_chain(const swe_mlp_t& m, int l0, int l1, float* X, int ldx, float* Ha, float* Hb, float* Wb) {
    constexpr int LDH = F + 4;
    float* cur = X;
    int ld = ldx;
    for (int l = l0; l < l1; ++l) {
        const swe_layer_t& L = m.layer[l];
        block_cp_async(Wb, L.wt, L.k_in * F);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        float acc[DenseCfg<F>::RM][8];
        dense_zero<F>(acc);
        dense_acc<F>(acc, cur, ld, Wb, L.k_in);
        dense_bias_act<F>(acc, L.bias, L.act, load_slope(L));
        float* dst = (cur == Ha) ? Hb : Ha;
        dense_store_smem<F>(acc, dst, LDH);
        __syncthreads();
        cur = dst;
        ld = LDH;
    }
    return cur;
}

template <int F>
__device__ __forceinline__ void store_tile_rows(const float* __restrict__ H, float* __restrict__ out,
                                                long long row0, long long n_rows) {
    constexpr int LDH = F + 4, QPR = F / 4;
    for (int idx = threadIdx.x; idx < TM * QPR; idx += NT) {
        const int r = idx / QPR, q = idx % QPR;
        const long long g = row0 + r;
        if (g < n_rows) stg4(out + g * F + 4 * q, *reinterpret_cast<const float4*>(H + r * LDH + 4 * q));
    }
}

constexpr int X0_LD = 36;     // raw-input tile: up to 32 columns (+4 pad)

template <int F>
struct RowMlpSmem {
    static constexpr int LDH = F + 4;
    static constexpr int WB = (F > 32 ? F : 32) * F;
    static constexpr size_t bytes = sizeof(float) * (size_t)(TM * X0_LD + 2 * TM * LDH + WB);
};

// =============================================================================================
// node encoders
// =============================================================================================
template <int F>
__global__ void __launch_bounds__(NT) node_encode_kernel(
    const float* __restrict__ x, int n_cols, const int32_t* __restrict__ perm, int n_nodes, int n_static_raw,
    int with_wl, int n_dyn_rows, const __grid_constant__ swe_mlp_t ms, const __grid_constant__ swe_mlp_t md,
    float* __restrict__ xs_out, float* __restrict__ xd_out) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LDH = F + 4;
    float* X0 = smem;
    float* Ha = X0 + TM * X0_LD;
    float* Hb = Ha + TM * LDH;
    float* Wb = Hb + TM * LDH;
    const int n_tiles = (n_nodes + TM - 1) / TM;
    const int n_dyn = n_cols - n_static_raw;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * TM;
        // ---- static branch: [x[:, :n_static_raw], WL]
        const int ks = ms.layer[0].k_in;
        for (int idx = threadIdx.x; idx < TM * ks; idx += NT) {
            const int r = idx / ks, c = idx % ks;
            const int node = row0 + r;
            float v = 0.f;
            if (node < n_nodes) {
                const long long src = (long long)(perm ? perm[node] : node) * n_cols;
                if (c < n_static_raw) v = __ldg(x + src + c);
                else if (c == n_static_raw && with_wl)
                    v = __ldg(x + src + n_static_raw - 1) + __ldg(x + src + n_cols - 2);
            }
            X0[r * X0_LD + c] = v;
        }
        __syncthreads();
        float* res = mlp_chain<F>(ms, 0, ms.n_layers, X0, X0_LD, Ha, Hb, Wb);
        store_tile_rows<F>(res, xs_out, row0, n_nodes);
        __syncthreads();
        // ---- dynamic branch (only the rows that are ever consumed)
        if (row0 < n_dyn_rows) {
            const int kd = md.layer[0].k_in;
            for (int idx = threadIdx.x; idx < TM * kd; idx += NT) {
                const int r = idx / kd, c = idx % kd;
                const int node = row0 + r;
                float v = 0.f;
                if (node < n_dyn_rows && c < n_dyn) {
                    const long long src = (long long)(perm ? perm[node] : node) * n_cols;
                    v = __ldg(x + src + n_static_raw + c);
                }
                X0[r * X0_LD + c] = v;
            }
            __syncthreads();
            res = mlp_chain<F>(md, 0, md.n_layers, X0, X0_LD, Ha, Hb, Wb);
            store_tile_rows<F>(res, xd_out, row0, n_dyn_rows);
            __syncthreads();
        }
    }
}

// =============================================================================================
// edge encoder (rows = edges in CSR order)
// =============================================================================================
template <int F>
__global__ void __launch_bounds__(NT) edge_encode_kernel(
    const float* __restrict__ edge_attr, int n_feat, const int32_t* __restrict__ eid, long long n_edges,
    const __grid_constant__ swe_mlp_t m, float* __restrict__ a_out) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LDH = F + 4;
    float* X0 = smem;
    float* Ha = X0 + TM * X0_LD;
    float* Hb = Ha + TM * LDH;
    float* Wb = Hb + TM * LDH;
    const long long n_tiles = (n_edges + TM - 1) / TM;
    const int k0 = m.layer[0].k_in;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * TM;
        for (int idx = threadIdx.x; idx < TM * k0; idx += NT) {
            const int r = idx / k0, c = idx % k0;
            const long long p = row0 + r;
            float v = 0.f;
            if (p < n_edges && c < n_feat) v = __ldg(edge_attr + (long long)(eid ? eid[p] : p) * n_feat + c);
            X0[r * X0_LD + c] = v;
        }
        __syncthreads();
        float* res = mlp_chain<F>(m, 0, m.n_layers, X0, X0_LD, Ha, Hb, Wb);
        store_tile_rows<F>(res, a_out, row0, n_edges);
        __syncthreads();
    }
}

// =============================================================================================
// edge gate:  s_e = normalise( MLP([x_s[r] | x_s[c] | x_d[r] | x_d[c] | a_e]) )
// =============================================================================================
template <int F>
struct GateSmem {
    static constexpr int LDS = F + 4;           // gathered segment tile
    static constexpr int LDH = 2 * F + 4;       // hidden tile
    static constexpr int SEG = TM * LDS;
    static constexpr int WBUF = F * 2 * F;      // one K=F block of a [*, 2F] weight
    static constexpr int H = TM * LDH;
    static_assert(2 * SEG >= H, "h2 aliases the two segment buffers");
    static constexpr size_t bytes = sizeof(float) * (size_t)(2 * SEG + 2 * WBUF + H) + sizeof(int32_t) * 2 * TM;
};

// gather one F-wide segment of the layer-0 input for the TM edges of this tile (cp.async)
template <int F>
__device__ __forceinline__ void gate_gather(float* __restrict__ dst, int seg, const float* __restrict__ xs,
                                            const float* __restrict__ xd_src, const float* __restrict__ xd_dst,
                                            const float* __restrict__ a,
                                            const int32_t* __restrict__ s_src, const int32_t* __restrict__ s_dst,
                                            long long e0, long long n_edges) {
    constexpr int QPR = F / 4, LDS = F + 4;
    for (int idx = threadIdx.x; idx < TM * QPR; idx += NT) {
        const int r = idx / QPR, q = idx % QPR;
        const float* srcp;
        if (seg == 4) {
            long long e = e0 + r;
            if (e >= n_edges) e = n_edges - 1;
            srcp = a + e * F;
        } else {
            const long long node = (seg & 1) ? s_dst[r] : s_src[r];
            srcp = ((seg < 2) ? xs : (seg == 2 ? xd_src : xd_dst)) + node * F;
        }
        cp_async16_ca(dst + r * LDS + 4 * q, srcp + 4 * q);
    }
}

template <int F, int NO>
__device__ __forceinline__ void gate_final(float (&acc)[DenseCfg<NO>::RM][8], const swe_layer_t& L, int normalize,
                                           float* __restrict__ s_out, long long e0, long long n_edges) {
    using C = DenseCfg<NO>;
    dense_bias_act<NO>(acc, L.bias, L.act, load_slope(L));
    if (normalize) {
        float ss[C::RM];
#pragma unroll
        for (int i = 0; i < C::RM; ++i) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) t = fmaf(acc[i][j], acc[i][j], t);
            ss[i] = t;
        }
        dense_row_reduce_sum<NO>(ss);
#pragma unroll
        for (int i = 0; i < C::RM; ++i) {
            const float nrm = sqrtf(ss[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float v = __fdiv_rn(acc[i][j], nrm);      // s / ||s||  (gnn.py:425)
                acc[i][j] = (v != v) ? 0.f : v;            // NaN -> 0   (gnn.py:426)
            }
        }
    }
    dense_store_global<NO>(acc, s_out, e0, n_edges);
}

template <int F>
__global__ void __launch_bounds__(NT, 1) edge_gate_kernel(
    const float* __restrict__ xs, const float* __restrict__ xd_src, const float* __restrict__ xd_dst,
    const float* __restrict__ a, const int32_t* __restrict__ src, const int32_t* __restrict__ dst, long long n_edges,
    const __grid_constant__ swe_mlp_t m, int normalize, float* __restrict__ s_out) {
    using S = GateSmem<F>;
    extern __shared__ __align__(16) float smem[];
    float* seg0 = smem;
    float* wb0 = smem + 2 * S::SEG;
    float* h1 = wb0 + 2 * S::WBUF;
    float* h2 = smem;                               // aliases seg buffers (free once layer 0 is done)
    int32_t* s_src = reinterpret_cast<int32_t*>(h1 + S::H);
    int32_t* s_dst = s_src + TM;
    // active input segments: x_s[r], x_s[c], x_d[r], then x_d[c] unless known-zero, then a_e if any
    const int nseg = 3 + (xd_dst ? 1 : 0) + (a ? 1 : 0);
    const int L = m.n_layers;
    const long long n_tiles = (n_edges + TM - 1) / TM;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long e0 = tile * TM;
        if (threadIdx.x < TM) {
            long long e = e0 + threadIdx.x;
            if (e >= n_edges) e = n_edges - 1;
            s_src[threadIdx.x] = src[e];
            s_dst[threadIdx.x] = dst[e];
        }
        __syncthreads();
        const swe_layer_t& L0 = m.layer[0];
        const int no0 = L0.n_out;                  // 2F, or F when the MLP has a single layer

        // ---------------- layer 0: Σ over segments, double-buffered gather + weight block
        auto prefetch = [&](int it) {
            const int sg = it < 3 ? it : ((it == 3 && xd_dst) ? 3 : 4);
            gate_gather<F>(seg0 + (it & 1) * S::SEG, sg, xs, xd_src, xd_dst, a, s_src, s_dst, e0, n_edges);
            block_cp_async(wb0 + (it & 1) * S::WBUF, L0.wt + (long long)sg * F * no0, F * no0);
            cp_async_commit();
        };
        if (L == 1) {
            float acc[DenseCfg<F>::RM][8];
            dense_zero<F>(acc);
            prefetch(0);
            for (int it = 0; it < nseg; ++it) {
                if (it + 1 < nseg) { prefetch(it + 1); cp_async_wait<1>(); } else cp_async_wait<0>();
                __syncthreads();
                dense_acc<F>(acc, seg0 + (it & 1) * S::SEG, S::LDS, wb0 + (it & 1) * S::WBUF, F);
                __syncthreads();
            }
            gate_final<F, F>(acc, L0, normalize, s_out, e0, n_edges);
            continue;
        }
        {
            float acc[DenseCfg<2 * F>::RM][8];
            dense_zero<2 * F>(acc);
            prefetch(0);
            for (int it = 0; it < nseg; ++it) {
                if (it + 1 < nseg) { prefetch(it + 1); cp_async_wait<1>(); } else cp_async_wait<0>();
                __syncthreads();
                dense_acc<2 * F>(acc, seg0 + (it & 1) * S::SEG, S::LDS, wb0 + (it & 1) * S::WBUF, F);
                __syncthreads();
            }
            dense_bias_act<2 * F>(acc, L0.bias, L0.act, load_slope(L0));
            dense_store_smem<2 * F>(acc, h1, S::LDH);
        }
        float* cur = h1;
        // ---------------- hidden layers 2F -> 2F
        for (int l = 1; l < L - 1; ++l) {
            const swe_layer_t& Ll = m.layer[l];
            block_cp_async(wb0, Ll.wt, F * 2 * F);
            block_cp_async(wb0 + S::WBUF, Ll.wt + (long long)F * 2 * F, F * 2 * F);
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();                        // also publishes `cur`
            float acc[DenseCfg<2 * F>::RM][8];
            dense_zero<2 * F>(acc);
            dense_acc<2 * F>(acc, cur, S::LDH, wb0, F);
            dense_acc<2 * F>(acc, cur + F, S::LDH, wb0 + S::WBUF, F);
            dense_bias_act<2 * F>(acc, Ll.bias, Ll.act, load_slope(Ll));
            float* nxt = (cur == h1) ? h2 : h1;
            dense_store_smem<2 * F>(acc, nxt, S::LDH);
            cur = nxt;
            __syncthreads();
        }
        // ---------------- last layer 2F -> F, normalise, store
        {
            const swe_layer_t& Ll = m.layer[L - 1];
            block_cp_async(wb0, Ll.wt, 2 * F * F);          // [2F][F] contiguous
            cp_async_commit();
            cp_async_wait<0>();
            __syncthreads();
            float acc[DenseCfg<F>::RM][8];
            dense_zero<F>(acc);
            dense_acc<F>(acc, cur, S::LDH, wb0, 2 * F);
            gate_final<F, F>(acc, Ll, normalize, s_out, e0, n_edges);
        }
        __syncthreads();                            // smem is recycled by the next tile
    }
}

// =============================================================================================
// hop:  out[c] = act( o[c] + (Σ_p s_p ⊙ (o[c] − o[src_p]))·Wᵀ + addend[c] )
// =============================================================================================
template <int F>
struct HopSmem {
    static constexpr int LDA = F + 4;
    static constexpr size_t bytes = sizeof(float) * (size_t)(TM * LDA + F * F);
};

template <int F>
__device__ __forceinline__ float4 hop_aggregate(const float* __restrict__ o_src, const float4 oc,
                                                const float* __restrict__ s, const int32_t* __restrict__ src,
                                                int p0, int p1, int q, int with_gradient, int upwind) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = p0; p < p1; p += 4) {
        int32_t r[4];
        float4 orow[4], sv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = (p + u < p1) ? __ldg(src + p + u) : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r[u] >= 0) {
                orow[u] = ldg4(o_src + (long long)r[u] * F + 4 * q);
                sv[u] = ldg4_stream(s + (long long)(p + u) * F + 4 * q);
            }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (r[u] >= 0) {
                float4 t;
                if (with_gradient) {                       // (o[c] − o[r]) · s   (gnn.py:430-433)
                    float4 d = make_float4(__fsub_rn(oc.x, orow[u].x), __fsub_rn(oc.y, orow[u].y),
                                           __fsub_rn(oc.z, orow[u].z), __fsub_rn(oc.w, orow[u].w));
                    if (upwind) { d.x = d.x < 0.f ? 0.f : d.x; d.y = d.y < 0.f ? 0.f : d.y;
                                  d.z = d.z < 0.f ? 0.f : d.z; d.w = d.w < 0.f ? 0.f : d.w; }
                    t = make_float4(__fmul_rn(d.x, sv[u].x), __fmul_rn(d.y, sv[u].y),
                                    __fmul_rn(d.z, sv[u].z), __fmul_rn(d.w, sv[u].w));
                } else {                                   // s · o[r]            (gnn.py:435)
                    t = make_float4(__fmul_rn(sv[u].x, orow[u].x), __fmul_rn(sv[u].y, orow[u].y),
                                    __fmul_rn(sv[u].z, orow[u].z), __fmul_rn(sv[u].w, orow[u].w));
                }
                // sequential in-segment sum in original edge order == CPU scatter_add_ order
                acc.x = __fadd_rn(acc.x, t.x); acc.y = __fadd_rn(acc.y, t.y);
                acc.z = __fadd_rn(acc.z, t.z); acc.w = __fadd_rn(acc.w, t.w);
            }
    }
    return acc;
}

template <int F, bool FILTER>
__global__ void __launch_bounds__(NT, FILTER ? 3 : 4) hop_kernel(
    const float* __restrict__ o_src, const float* __restrict__ o_dst, const float* __restrict__ s,
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, int dst_lo, int n_dst,
    const float* __restrict__ wt, int with_gradient, int upwind, const float* __restrict__ addend,
    int act, const float* __restrict__ slope_p, float* __restrict__ out, float* __restrict__ agg_out) {
    extern __shared__ __align__(16) float smem[];
    constexpr int QPR = F / 4, NG = NT / QPR, LDA = F + 4;
    float* agg = smem;
    float* W = smem + TM * LDA;
    const float slope = (act == SWE_ACT_PRELU && slope_p) ? __ldg(slope_p) : 0.f;
    if (FILTER) {
        block_cp_async(W, wt, F * F);
        cp_async_commit();
    }
    const int g = threadIdx.x / QPR, q = threadIdx.x % QPR;
    const int n_tiles = (n_dst + TM - 1) / TM;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int i = g; i < TM; i += NG) {
            const int c_rel = tile * TM + i;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), oc = acc;
            if (c_rel < n_dst) {
                const long long c = (long long)dst_lo + c_rel;
                if (o_dst) oc = ldg4(o_dst + c * F + 4 * q);
                acc = hop_aggregate<F>(o_src, oc, s, src, __ldg(rowptr + c_rel), __ldg(rowptr + c_rel + 1), q,
                                       with_gradient, upwind);
                if (!FILTER) {
                    float4 v = make_float4(__fadd_rn(oc.x, acc.x), __fadd_rn(oc.y, acc.y),
                                           __fadd_rn(oc.z, acc.z), __fadd_rn(oc.w, acc.w));
                    if (addend) {
                        const float4 ad = ldg4(addend + c * F + 4 * q);
                        v.x += ad.x; v.y += ad.y; v.z += ad.z; v.w += ad.w;
                    }
                    v.x = act_apply(act, v.x, slope); v.y = act_apply(act, v.y, slope);
                    v.z = act_apply(act, v.z, slope); v.w = act_apply(act, v.w, slope);
                    stg4(out + c * F + 4 * q, v);
                }
            }
            if (FILTER) {
                stg4(agg + i * LDA + 4 * q, acc);
                if (agg_out && c_rel < n_dst) stg4(agg_out + ((long long)dst_lo + c_rel) * F + 4 * q, acc);
            }
        }
        if (FILTER) {
            cp_async_wait<0>();
            __syncthreads();
            using C = DenseCfg<F>;
            float acc[C::RM][8];
            dense_zero<F>(acc);
            dense_acc<F>(acc, agg, LDA, W, F);
            const int tx = threadIdx.x % C::TX, ty = threadIdx.x / C::TX;
#pragma unroll
            for (int i = 0; i < C::RM; ++i) {
                const int c_rel = tile * TM + ty * C::RM + i;
                if (c_rel < n_dst) {
                    const long long c = (long long)dst_lo + c_rel;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int col = h * (F / 2) + 4 * tx;
                        float4 v = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
                        if (o_dst) {
                            const float4 oc = ldg4(o_dst + c * F + col);
                            v.x = oc.x + v.x; v.y = oc.y + v.y; v.z = oc.z + v.z; v.w = oc.w + v.w;
                        }
                        if (addend) {
                            const float4 ad = ldg4(addend + c * F + col);
                            v.x += ad.x; v.y += ad.y; v.z += ad.z; v.w += ad.w;
                        }
                        v.x = act_apply(act, v.x, slope); v.y = act_apply(act, v.y, slope);
                        v.z = act_apply(act, v.z, slope); v.w = act_apply(act, v.w, slope);
                        stg4(out + c * F + col, v);
                    }
                }
            }
            __syncthreads();
        }
    }
    if (FILTER) cp_async_wait<0>();
}

// out[row_lo + i] = x[row_lo + i] · Wᵀ
template <int F>
__global__ void __launch_bounds__(NT, 3) node_linear_kernel(const float* __restrict__ x, int row_lo, int n_rows,
                                                            const float* __restrict__ wt, float* __restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    constexpr int QPR = F / 4, LDA = F + 4;
    float* A = smem;
    float* W = smem + TM * LDA;
    block_cp_async(W, wt, F * F);
    cp_async_commit();
    const int n_tiles = (n_rows + TM - 1) / TM;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int idx = threadIdx.x; idx < TM * QPR; idx += NT) {
            const int r = idx / QPR, q = idx % QPR;
            const int rr = tile * TM + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rr < n_rows) v = ldg4_stream(x + ((long long)row_lo + rr) * F + 4 * q);
            stg4(A + r * LDA + 4 * q, v);
        }
        cp_async_wait<0>();
        __syncthreads();
        float acc[DenseCfg<F>::RM][8];
        dense_zero<F>(acc);
        dense_acc<F>(acc, A, LDA, W, F);
        dense_store_global<F>(acc, out + (long long)row_lo * F, (long long)tile * TM, n_rows);
        __syncthreads();
    }
    cp_async_wait<0>();
}

// =============================================================================================
// mean pooling onto the coarser scale
// =============================================================================================
template <int F>
__global__ void __launch_bounds__(NT) pool_mean_kernel(const float* __restrict__ x, const int32_t* __restrict__ rowptr,
                                                       const int32_t* __restrict__ fine, int coarse_lo, int n_coarse,
                                                       float* __restrict__ out) {
    constexpr int QPR = F / 4, NG = NT / QPR;
    const int q = threadIdx.x % QPR;
    for (long long i = (long long)blockIdx.x * NG + threadIdx.x / QPR; i < n_coarse; i += (long long)gridDim.x * NG) {
        const int p0 = __ldg(rowptr + i), p1 = __ldg(rowptr + i + 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = p0; p < p1; ++p) {
            const float4 v = ldg4_stream(x + (long long)__ldg(fine + p) * F + 4 * q);
            acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
            acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
        }
        const float cnt = (float)max(p1 - p0, 1);                     // count.clamp(min=1)
        acc.x = __fdiv_rn(acc.x, cnt); acc.y = __fdiv_rn(acc.y, cnt);
        acc.z = __fdiv_rn(acc.z, cnt); acc.w = __fdiv_rn(acc.w, cnt);
        stg4(out + ((long long)coarse_lo + i) * F + 4 * q, acc);
    }
}

// =============================================================================================
// decoder head (+ residual, ReLU, dry mask, optional window shift)
// =============================================================================================
template <int F>
__global__ void __launch_bounds__(NT) decode_head_kernel(
    const float* __restrict__ h, int act_in, const float* __restrict__ slope_in, const __grid_constant__ swe_mlp_t dec,
    const float* x0, int n_cols, const int32_t* __restrict__ perm, int n_nodes, int previous_t, int res_mode,
    const float* __restrict__ res_w, float eps, float* pred, const int32_t* __restrict__ step_ptr,
    long long pred_step_stride, float* x_next) {
    extern __shared__ __align__(16) float smem[];
    constexpr int LDH = F + 4, QPR = F / 4;
    float* X0 = smem;                      // unused raw-input slot keeps the layout of RowMlpSmem
    float* Ha = X0 + TM * X0_LD;
    float* Hb = Ha + TM * LDH;
    float* Wb = Hb + TM * LDH;
    const float sl_in = (act_in == SWE_ACT_PRELU && slope_in) ? __ldg(slope_in) : 0.f;
    const int n_tiles = (n_nodes + TM - 1) / TM;
    const int L = dec.n_layers;
    const int n_static_raw = n_cols - 2 * previous_t;
    if (step_ptr) pred += (long long)(*step_ptr) * pred_step_stride;
    const swe_layer_t& LL = dec.layer[L - 1];
    const float sl_last = load_slope(LL);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * TM;
        for (int idx = threadIdx.x; idx < TM * QPR; idx += NT) {
            const int r = idx / QPR, q = idx % QPR;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < n_nodes) {
                v = ldg4_stream(h + (long long)(row0 + r) * F + 4 * q);
                v.x = act_apply(act_in, v.x, sl_in); v.y = act_apply(act_in, v.y, sl_in);
                v.z = act_apply(act_in, v.z, sl_in); v.w = act_apply(act_in, v.w, sl_in);
            }
            stg4(Ha + r * LDH + 4 * q, v);
        }
        __syncthreads();
        const float* cur = mlp_chain<F>(dec, 0, L - 1, Ha, LDH, Ha, Hb, Wb);
        // last layer F -> 2: thread = (row, var)
        const int r = threadIdx.x >> 1, j = threadIdx.x & 1;
        const int node = row0 + r;
        float y = 0.f;
        for (int k = 0; k < F; k += 4) {
            const float4 a = *reinterpret_cast<const float4*>(cur + r * LDH + k);
            y = fmaf(a.x, __ldg(LL.wt + (k + 0) * 2 + j), y);
            y = fmaf(a.y, __ldg(LL.wt + (k + 1) * 2 + j), y);
            y = fmaf(a.z, __ldg(LL.wt + (k + 2) * 2 + j), y);
            y = fmaf(a.w, __ldg(LL.wt + (k + 3) * 2 + j), y);
        }
        if (LL.bias) y += __ldg(LL.bias + j);
        y = act_apply(LL.act, y, sl_last);
        const long long orow = (node < n_nodes) ? (long long)(perm ? perm[node] : node) : 0;
        const float* xr = x0 + orow * n_cols;
        float res = 0.f;
        if (node < n_nodes) {
            if (res_mode == 1) {
                for (int t = 0; t < previous_t; ++t) res = fmaf(xr[n_static_raw + 2 * t + j], __ldg(res_w + t), res);
            } else if (res_mode == 2) {
                for (int t = 0; t < previous_t; ++t) res = fmaf(xr[n_static_raw + 2 * t + j], __ldg(res_w + 2 * t + j), res);
            } else if (res_mode == 3) {
                res = xr[n_cols - 2 + j];
            }
        }
        y = fmaxf(y + res, 0.f);
        const float hval = __shfl_sync(0xffffffffu, y, (threadIdx.x & 31) & ~1);
        float o;
        if (j == 0) o = (fabsf(y) > eps) ? y : 0.f;                    // h · [|h| > eps]
        else        o = (hval != 0.f) ? y : 0.f;                       // q · [h != 0] (un-thresholded h)
        const float o_h = __shfl_sync(0xffffffffu, o, (threadIdx.x & 31) & ~1);
        const float o_q = __shfl_sync(0xffffffffu, o, (threadIdx.x & 31) | 1);
        __syncwarp();
        if (node < n_nodes) {
            pred[orow * 2 + j] = o;
            if (x_next && j == 0) {
                float* xn = x_next + orow * n_cols;
                for (int c = 0; c < n_static_raw; ++c) xn[c] = xr[c];
                for (int c = n_static_raw; c < n_cols - 2; ++c) xn[c] = xr[c + 2];
                xn[n_cols - 2] = o_h;
                xn[n_cols - 1] = o_q;
            }
        }
        __syncthreads();
    }
}

__global__ void apply_bc_kernel(float* __restrict__ x, int n_cols, int n_static_raw, int previous_t, int type_bc,
                                const int64_t* __restrict__ node_bc, int n_bc, const float* __restrict__ bc,
                                int n_steps_total, const int32_t* __restrict__ step_ptr) {
    const int step = step_ptr ? *step_ptr : 0;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_bc * previous_t; idx += gridDim.x * blockDim.x) {
        const int b = idx / previous_t, t = idx % previous_t;
        x[(long long)node_bc[b] * n_cols + n_static_raw + 2 * t + (type_bc - 1)] =
            bc[((long long)b * previous_t + t) * n_steps_total + step];
    }
}

// dst[i, :] = src[idx[i], :]  (rows of `width` floats; halo send-buffer packing)
__global__ void pack_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ idx, long long n, int width,
                                 float* __restrict__ dst) {
    if ((width & 3) == 0) {
        const int qpr = width >> 2;
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n * qpr; t += (long long)gridDim.x * blockDim.x) {
            const long long i = t / qpr;
            const int q = (int)(t % qpr);
            stg4(dst + i * width + 4 * q, ldg4(src + (long long)__ldg(idx + i) * width + 4 * q));
        }
    } else {
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n * width; t += (long long)gridDim.x * blockDim.x) {
            const long long i = t / width;
            dst[t] = __ldg(src + (long long)__ldg(idx + i) * width + (t % width));
        }
    }
}

__global__ void step_advance_kernel(int32_t* step_ptr) { if (threadIdx.x == 0 && blockIdx.x == 0) *step_ptr += 1; }

// =============================================================================================
// host-side dispatch
// =============================================================================================
static int check_mlp_rows(const swe_mlp_t* m, int F, int last_out, const char* what) {
    SWE_REQUIRE(m && m->n_layers >= 1 && m->n_layers <= SWE_MAX_LAYERS, SWE_E_INVAL, "%s: bad layer count", what);
    for (int l = 0; l < m->n_layers; ++l) {
        const swe_layer_t& L = m->layer[l];
        const int want_out = (l == m->n_layers - 1) ? last_out : F;
        SWE_REQUIRE(L.wt && aligned16(L.wt), SWE_E_ALIGN, "%s: layer %d weight null/unaligned", what, l);
        SWE_REQUIRE(L.n_out == want_out, SWE_E_UNSUPP, "%s: layer %d n_out=%d, expected %d", what, l, L.n_out, want_out);
        SWE_REQUIRE(L.k_in % 4 == 0 && L.k_in >= 4, SWE_E_INVAL, "%s: layer %d k_in=%d not padded to 4", what, l, L.k_in);
        if (l == 0) SWE_REQUIRE(L.k_in <= 32 || L.k_in == F, SWE_E_UNSUPP, "%s: first layer k_in=%d > 32", what, L.k_in);
        else SWE_REQUIRE(L.k_in == F, SWE_E_UNSUPP, "%s: layer %d k_in=%d != F", what, l, L.k_in);
        SWE_REQUIRE(!L.bias || aligned16(L.bias), SWE_E_ALIGN, "%s: layer %d bias unaligned", what, l);
    }
    return 0;
}

template <typename K>
static int opt_in_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(%zu B): %s", bytes, cudaGetErrorString(e)); return (int)e; }
    return 0;
}

#define SWE_DISPATCH_F(F_, ...)                                 \
    switch (F_) {                                               \
        case 16: { constexpr int FF = 16; __VA_ARGS__; } break; \
        case 32: { constexpr int FF = 32; __VA_ARGS__; } break; \
        case 64: { constexpr int FF = 64; __VA_ARGS__; } break; \
        default: set_error("unsupported feature width F=%d (16, 32, 64)", F_); return SWE_E_UNSUPP; \
    }

}  // namespace swe

using namespace swe;

extern "C" int swe_node_encode_fwd(const float* x, int32_t n_cols, const int32_t* perm, int32_t n_nodes,
                                   int32_t n_static_raw, int32_t with_wl, int32_t n_dyn_rows,
                                   const swe_mlp_t* ms, const swe_mlp_t* md, float* xs_out, float* xd_out,
                                   int32_t F, void* stream) {
    SWE_REQUIRE(x && xs_out && xd_out && n_nodes >= 0 && n_dyn_rows >= 0 && n_dyn_rows <= n_nodes, SWE_E_INVAL,
                "node_encode: bad arguments");
    SWE_REQUIRE(aligned16(xs_out) && aligned16(xd_out), SWE_E_ALIGN, "node_encode: outputs unaligned");
    SWE_REQUIRE(n_static_raw >= 1 && n_cols > n_static_raw, SWE_E_INVAL, "node_encode: bad column split");
    if (int r = check_mlp_rows(ms, F, F, "static_node_encoder")) return r;
    if (int r = check_mlp_rows(md, F, F, "dynamic_node_encoder")) return r;
    SWE_REQUIRE(ms->layer[0].k_in >= n_static_raw + (with_wl ? 1 : 0), SWE_E_INVAL, "static encoder k_in too small");
    SWE_REQUIRE(md->layer[0].k_in >= n_cols - n_static_raw, SWE_E_INVAL, "dynamic encoder k_in too small");
    SWE_REQUIRE(ms->layer[0].k_in <= 32 && md->layer[0].k_in <= 32, SWE_E_UNSUPP, "encoder input wider than 32");
    if (n_nodes == 0) return 0;
    SWE_DISPATCH_F(F, {
        auto k = node_encode_kernel<FF>;
        if (int r = opt_in_smem(k, RowMlpSmem<FF>::bytes)) return r;
        k<<<grid_for((n_nodes + TM - 1) / TM, 2), NT, RowMlpSmem<FF>::bytes, (cudaStream_t)stream>>>(
            x, n_cols, perm, n_nodes, n_static_raw, with_wl, n_dyn_rows, *ms, *md, xs_out, xd_out);
    });
    return check_launch("node_encode_fwd");
}

extern "C" int swe_edge_encode_fwd(const float* edge_attr, int32_t n_feat, const int32_t* eid, int64_t n_edges,
                                   const swe_mlp_t* m, float* a_out, int32_t F, void* stream) {
    SWE_REQUIRE(edge_attr && a_out && n_edges >= 0 && n_feat >= 1, SWE_E_INVAL, "edge_encode: bad arguments");
    SWE_REQUIRE(aligned16(a_out), SWE_E_ALIGN, "edge_encode: output unaligned");
    if (int r = check_mlp_rows(m, F, F, "edge_encoder")) return r;
    SWE_REQUIRE(m->layer[0].k_in >= n_feat && m->layer[0].k_in <= 32, SWE_E_UNSUPP, "edge encoder input width");
    if (n_edges == 0) return 0;
    SWE_DISPATCH_F(F, {
        auto k = edge_encode_kernel<FF>;
        if (int r = opt_in_smem(k, RowMlpSmem<FF>::bytes)) return r;
        k<<<grid_for((n_edges + TM - 1) / TM, 2), NT, RowMlpSmem<FF>::bytes, (cudaStream_t)stream>>>(
            edge_attr, n_feat, eid, n_edges, *m, a_out);
    });
    return check_launch("edge_encode_fwd");
}

extern "C" int swe_edge_gate_fwd(const float* xs, const float* xd_src, const float* xd_dst, const float* a,
                                 const int32_t* src, const int32_t* dst, int64_t n_edges, const swe_mlp_t* m,
                                 int32_t normalize, float* s_out, int32_t F, void* stream) {
    SWE_REQUIRE(xs && xd_src && src && dst && s_out && n_edges >= 0, SWE_E_INVAL, "edge_gate: bad arguments");
    SWE_REQUIRE(aligned16(xs) && aligned16(xd_src) && aligned16(s_out) && (!a || aligned16(a)) &&
                (!xd_dst || aligned16(xd_dst)), SWE_E_ALIGN, "edge_gate: unaligned buffer");
    SWE_REQUIRE(m && m->n_layers >= 1 && m->n_layers <= SWE_MAX_LAYERS, SWE_E_INVAL, "edge_gate: bad layer count");
    const int nseg = a ? 5 : 4, L = m->n_layers;
    for (int l = 0; l < L; ++l) {
        const swe_layer_t& Ly = m->layer[l];
        const int kin = (l == 0) ? nseg * F : 2 * F;
        const int nout = (l == L - 1) ? F : 2 * F;
        SWE_REQUIRE(Ly.wt && aligned16(Ly.wt) && (!Ly.bias || aligned16(Ly.bias)), SWE_E_ALIGN,
                    "edge_gate: layer %d weight/bias null or unaligned", l);
        SWE_REQUIRE(Ly.k_in == kin && Ly.n_out == nout, SWE_E_UNSUPP,
                    "edge_gate: layer %d is %d->%d, expected %d->%d", l, Ly.k_in, Ly.n_out, kin, nout);
    }
    if (n_edges == 0) return 0;
    SWE_DISPATCH_F(F, {
        auto k = edge_gate_kernel<FF>;
        if (int r = opt_in_smem(k, GateSmem<FF>::bytes)) return r;
        const int per_sm = GateSmem<FF>::bytes > 110 * 1024 ? 1 : 2;
        k<<<grid_for((n_edges + TM - 1) / TM, per_sm), NT, GateSmem<FF>::bytes, (cudaStream_t)stream>>>(
            xs, xd_src, xd_dst, a, src, dst, n_edges, *m, normalize, s_out);
    });
    return check_launch("edge_gate_fwd");
}

extern "C" int swe_node_linear_fwd(const float* x, int32_t row_lo, int32_t n_rows, const float* wt, float* out,
                                   int32_t F, void* stream) {
    SWE_REQUIRE(x && wt && out && row_lo >= 0 && n_rows >= 0, SWE_E_INVAL, "node_linear: bad arguments");
    SWE_REQUIRE(aligned16(x) && aligned16(wt) && aligned16(out), SWE_E_ALIGN, "node_linear: unaligned buffer");
    if (n_rows == 0) return 0;
    SWE_DISPATCH_F(F, {
        auto k = node_linear_kernel<FF>;
        if (int r = opt_in_smem(k, HopSmem<FF>::bytes)) return r;
        k<<<grid_for((n_rows + TM - 1) / TM, 3), NT, HopSmem<FF>::bytes, (cudaStream_t)stream>>>(x, row_lo, n_rows, wt, out);
    });
    return check_launch("node_linear_fwd");
}

static int hop_launch(const float* o_src, const float* o_dst, const float* s, const int32_t* rowptr,
                      const int32_t* src, int32_t dst_lo, int32_t n_dst, const float* wt, int32_t with_gradient,
                      int32_t upwind, const float* addend, int32_t act, const float* slope, float* out,
                      float* agg_out, int32_t F, void* stream) {
    SWE_REQUIRE(o_src && s && rowptr && src && out && dst_lo >= 0 && n_dst >= 0, SWE_E_INVAL, "hop: bad arguments");
    SWE_REQUIRE(!(with_gradient && !o_dst), SWE_E_INVAL, "hop: with_gradient needs the destination rows");
    SWE_REQUIRE(aligned16(o_src) && aligned16(s) && aligned16(out) && (!o_dst || aligned16(o_dst)) &&
                (!wt || aligned16(wt)) && (!addend || aligned16(addend)) && (!agg_out || aligned16(agg_out)),
                SWE_E_ALIGN, "hop: unaligned buffer");
    SWE_REQUIRE(out != o_src && out != o_dst, SWE_E_INVAL, "hop: output must not alias the hop input");
    if (n_dst == 0) return 0;
    SWE_DISPATCH_F(F, {
        if (wt) {
            auto k = hop_kernel<FF, true>;
            if (int r = opt_in_smem(k, HopSmem<FF>::bytes)) return r;
            k<<<grid_for((n_dst + TM - 1) / TM, 3), NT, HopSmem<FF>::bytes, (cudaStream_t)stream>>>(
                o_src, o_dst, s, rowptr, src, dst_lo, n_dst, wt, with_gradient, upwind, addend, act, slope, out, agg_out);
        } else {
            auto k = hop_kernel<FF, false>;
            k<<<grid_for((n_dst + TM - 1) / TM, 4), NT, 0, (cudaStream_t)stream>>>(
                o_src, o_dst, s, rowptr, src, dst_lo, n_dst, nullptr, with_gradient, upwind, addend, act, slope, out, nullptr);
        }
    });
    return check_launch("propagate_hop_fwd");
}

extern "C" int swe_propagate_hop_fwd(const float* o_src, const float* o_dst, const float* s, const int32_t* rowptr,
                                     const int32_t* src, int32_t dst_lo, int32_t n_dst, const float* wt,
                                     int32_t with_gradient, int32_t upwind, const float* addend, int32_t act,
                                     const float* slope, float* out, int32_t F, void* stream) {
    return hop_launch(o_src, o_dst, s, rowptr, src, dst_lo, n_dst, wt, with_gradient, upwind, addend, act, slope, out,
                      nullptr, F, stream);
}

extern "C" int swe_propagate_hop_train_fwd(const float* o_src, const float* o_dst, const float* s, const int32_t* rowptr,
                                           const int32_t* src, int32_t dst_lo, int32_t n_dst, const float* wt,
                                           int32_t with_gradient, int32_t upwind, const float* addend, float* agg_out,
                                           float* out, int32_t F, void* stream) {
    return hop_launch(o_src, o_dst, s, rowptr, src, dst_lo, n_dst, wt, with_gradient, upwind, addend, SWE_ACT_NONE,
                      nullptr, out, agg_out, F, stream);
}

extern "C" int swe_pool_mean_fwd(const float* x, const int32_t* rowptr, const int32_t* fine, int32_t coarse_lo,
                                 int32_t n_coarse, float* out, int32_t F, void* stream) {
    SWE_REQUIRE(x && rowptr && fine && out && coarse_lo >= 0 && n_coarse >= 0, SWE_E_INVAL, "pool: bad arguments");
    SWE_REQUIRE(aligned16(x) && aligned16(out), SWE_E_ALIGN, "pool: unaligned buffer");
    if (n_coarse == 0) return 0;
    SWE_DISPATCH_F(F, {
        constexpr int NG = NT / (FF / 4);
        pool_mean_kernel<FF><<<grid_for((n_coarse + NG - 1) / NG, 8), NT, 0, (cudaStream_t)stream>>>(
            x, rowptr, fine, coarse_lo, n_coarse, out);
    });
    return check_launch("pool_mean_fwd");
}

extern "C" int swe_decode_head_fwd(const float* h, int32_t act_in, const float* slope_in, const swe_mlp_t* dec,
                                   const float* x0, int32_t n_cols, const int32_t* perm, int32_t n_nodes,
                                   int32_t previous_t, int32_t res_mode, const float* res_w, float eps,
                                   float* pred, const int32_t* step_ptr, int64_t pred_step_stride,
                                   float* x_next, int32_t F, void* stream) {
    SWE_REQUIRE(h && x0 && pred && n_nodes >= 0 && previous_t >= 1 && n_cols > 2 * previous_t, SWE_E_INVAL,
                "decode_head: bad arguments");
    SWE_REQUIRE(aligned16(h), SWE_E_ALIGN, "decode_head: unaligned buffer");
    SWE_REQUIRE(res_mode >= 0 && res_mode <= 3 && (res_mode == 0 || res_mode == 3 || res_w), SWE_E_INVAL,
                "decode_head: bad residual mode");
    if (int r = check_mlp_rows(dec, F, 2, "node_decoder")) return r;
    SWE_REQUIRE(dec->layer[0].k_in == F, SWE_E_UNSUPP, "decoder input width must be F");
    if (n_nodes == 0) return 0;
    SWE_DISPATCH_F(F, {
        auto k = decode_head_kernel<FF>;
        if (int r = opt_in_smem(k, RowMlpSmem<FF>::bytes)) return r;
        k<<<grid_for((n_nodes + TM - 1) / TM, 2), NT, RowMlpSmem<FF>::bytes, (cudaStream_t)stream>>>(
            h, act_in, slope_in, *dec, x0, n_cols, perm, n_nodes, previous_t, res_mode, res_w, eps, pred,
            step_ptr, (long long)pred_step_stride, x_next);
    });
    return check_launch("decode_head_fwd");
}

extern "C" int swe_apply_bc(float* x, int32_t n_cols, int32_t n_static_raw, int32_t previous_t, int32_t type_bc,
                            const int64_t* node_bc, int32_t n_bc, const float* bc, int32_t n_steps_total,
                            const int32_t* step_ptr, void* stream) {
    SWE_REQUIRE(x && node_bc && bc && n_bc >= 0 && previous_t >= 1, SWE_E_INVAL, "apply_bc: bad arguments");
    SWE_REQUIRE(type_bc == 1 || type_bc == 2, SWE_E_INVAL, "BC_type=%d is not a valid input (1: depth, 2: discharge)", type_bc);
    if (n_bc == 0) return 0;
    apply_bc_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(x, n_cols, n_static_raw, previous_t, type_bc, node_bc, n_bc,
                                                        bc, n_steps_total, step_ptr);
    return check_launch("apply_bc");
}

extern "C" int swe_pack_rows(const float* src, const int32_t* idx, int64_t n_rows, int32_t width, float* dst, void* stream) {
    SWE_REQUIRE(src && idx && dst && n_rows >= 0 && width >= 1, SWE_E_INVAL, "pack_rows: bad arguments");
    SWE_REQUIRE((width & 3) != 0 || (aligned16(src) && aligned16(dst)), SWE_E_ALIGN, "pack_rows: unaligned buffer");
    if (n_rows == 0) return 0;
    const long long work = n_rows * ((width & 3) == 0 ? width / 4 : width);
    pack_rows_kernel<<<grid_for((work + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(src, idx, n_rows, width, dst);
    return check_launch("pack_rows");
}

extern "C" int swe_step_advance(int32_t* step_ptr, void* stream) {
    SWE_REQUIRE(step_ptr, SWE_E_INVAL, "step_advance: null pointer");
    step_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(step_ptr);
    return check_launch("step_advance");
}
