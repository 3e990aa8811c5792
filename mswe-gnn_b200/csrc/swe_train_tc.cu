// The two backward GEMMs of a make_mlp Linear (models/models.py:133-145, differentiated) on the 5th-generation
// tensor cores (tcgen05 + TMEM) for the wide edge-MLP layers of the training step (training/train.py:125-145):
//
//   dx[r, :]  (+)= delta[r, :] · W[:, k_off : k_off + ko)                 rows x small weight   (mlp_dx_tc_kernel)
//   dW[n, k]   =  Σ_r delta[r, n] · X[r, k]                               reduction over rows   (mlp_dw_tc_kernel)
//
// Precision: as in swe_gate_tc.cu every product is 3 TF32 MMAs on error-free hi/lo splits, fp32 accumulation in
// TMEM (relative error ≈ 1e-6, the exact-fp32 CUDA-core kernels in swe_backward.cu stay the small-shape path).
//
// Shared-memory layout of every operand tile: [rows x 32 features] panels, 128-byte swizzle, 8-row atoms of 1 KB.
//   * dx reads the delta panels K-major (M = 128 rows, K = features) against a weight image built once per CTA;
//   * dW reads [rows x 32 features] panels MN-major (M/N = features, K = rows; the 128B-swizzle-with-32B-base layout,
//     the only one the tensor core reads 32-bit operands transposed from): the rows are stored as they arrive from
//     global memory — no transposed staging; the 32-feature panels are the MN atoms (leading byte offset = panel).
#include "swe_tc.cuh"

namespace swe {
namespace tc {

constexpr int TR_ROW_THREADS = 256;        // warps 0-7: staging + epilogue
constexpr int TR_THREADS = 288;            // + warp 8: MMA issuer

// MN-major tf32 operands exist in ONE shared-memory layout (cute::UMMA::Layout_MN_SW128_32B_Atom, layout type
// SWIZZLE_128B_BASE32B): rows of 128 B = 32 consecutive M/N elements, 4-row atoms (one row per K index), the four
// 32-byte units of a row XOR-swizzled with (row & 3).  LBO = stride between 32-element MN atoms, SBO = stride between
// 4-row K groups; one K = 8 MMA reads two groups.
__device__ __forceinline__ uint64_t make_desc_mn_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                               // SWIZZLE_128B_BASE32B
    return d;
}
// byte offset of the 16-byte chunk c16 (4 elements) of row `row` inside a [rows x 32] MN-major panel
__host__ __device__ constexpr uint32_t mn32b_offset(uint32_t row, uint32_t c16) {
    return row * 128u + ((((c16 >> 1) ^ row) & 3u) << 5) + ((c16 & 1u) << 4);
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) { return make_idesc_tf32(M, N) | (1u << 15) | (1u << 16); }

__device__ __forceinline__ float leaky_slope_of(int act, const float* slope_p) {
    switch (act) {
        case SWE_ACT_PRELU:     return slope_p ? __ldg(slope_p) : 0.25f;
        case SWE_ACT_RELU:      return 0.f;
        case SWE_ACT_LEAKYRELU: return 0.1f;
        default:                return 1.f;          // SWE_ACT_NONE (other activations are rejected at launch)
    }
}

struct SegS {                               // a provider segment, resolved into shared memory
    const float* base; const int32_t* idx;
    int ld, col0, width; float slope;       // leaky-family activation on load: v > 0 ? v : slope * v
};
struct ProvS { SegS seg[SWE_MAX_SEGS]; int seg_of_panel[8]; int n_seg; int _pad; };
__device__ __forceinline__ void resolve_segs(const swe_rows_t& X, ProvS* P) {
    int c = 0;
    for (int j = 0; j < X.n_seg; ++j) {
        const swe_seg_t& g = X.seg[j];
        SegS& s = P->seg[j];
        s.base = g.base; s.idx = g.idx; s.ld = g.ld; s.col0 = c; s.width = g.width;
        s.slope = leaky_slope_of(g.act, g.slope);
        for (int q = c / 32; q < (c + g.width) / 32 && q < 8; ++q) P->seg_of_panel[q] = j;
        c += g.width;
    }
    P->n_seg = X.n_seg;
}
// address of 4 consecutive columns starting at `col` of provider row g (rows beyond the end are clamped: the
// caller zeroes them when it stores).  Split from the load so that a thread's loads are all issued back to back.
__device__ __forceinline__ const float* provider_addr(const ProvS* P, long long g, long long n_rows, int col) {
    const SegS& sg = P->seg[P->seg_of_panel[col >> 5]];
    if (g >= n_rows) g = n_rows - 1;
    long long r = g;
    if (sg.idx) r = (long long)__ldg(sg.idx + g);
    return sg.base + r * sg.ld + (col - sg.col0);
}
__device__ __forceinline__ float4 provider_finish(const ProvS* P, float4 v, long long g, long long n_rows, int col) {
    const float sl = g < n_rows ? P->seg[P->seg_of_panel[col >> 5]].slope : 0.f;
    const float keep = g < n_rows ? 1.f : 0.f;
    v.x = keep * fmaxf(v.x, 0.f) + sl * fminf(v.x, 0.f); v.y = keep * fmaxf(v.y, 0.f) + sl * fminf(v.y, 0.f);
    v.z = keep * fmaxf(v.z, 0.f) + sl * fminf(v.z, 0.f); v.w = keep * fmaxf(v.w, 0.f) + sl * fminf(v.w, 0.f);
    return v;
}
__device__ __forceinline__ void split_store4(unsigned char* hi_p, unsigned char* lo_p, const float4& v) {
    float4 hh, ll;
    split_tf32(v.x, hh.x, ll.x); split_tf32(v.y, hh.y, ll.y); split_tf32(v.z, hh.z, ll.z); split_tf32(v.w, hh.w, ll.w);
    *reinterpret_cast<float4*>(hi_p) = hh;
    *reinterpret_cast<float4*>(lo_p) = ll;
}

// =============================================================================================
// dW: D[m, q] = Σ_rows P[r, m] · Q[r, q];  P 128 features wide, Q up to 256
// =============================================================================================
constexpr int DW_RS = 32;                       // rows per stage = 4 MMA K-steps
constexpr int DW_PANEL = DW_RS * 128;           // bytes of a [32 rows x 32 features] panel
constexpr int DW_STAGES = 2;
constexpr int DW_ROW_THREADS = 512;             // warps 0-15: staging + epilogue (the staging is instruction/latency bound)
constexpr int DW_THREADS = DW_ROW_THREADS + 32; // + warp 16: MMA issuer
constexpr int DW_P_BYTES = 4 * DW_PANEL;        // 128 features (hi or lo)
constexpr int DW_Q_BYTES = 8 * DW_PANEL;        // 256 features (hi or lo)
constexpr int DW_STAGE_BYTES = 2 * DW_P_BYTES + 2 * DW_Q_BYTES;      // 96 KB

struct DwBarriers { uint64_t full[DW_STAGES], empty[DW_STAGES], d_full; };
constexpr size_t DW_TC_SMEM = 1024 + (size_t)DW_STAGES * DW_STAGE_BYTES + 2 * sizeof(ProvS) + sizeof(DwBarriers) + 32;

struct DwTcParams {
    swe_rows_t P, Q;            // P: total width pw (64 or 128); Q: total width qw (multiple of 32, <= 256)
    int pw, qw;
    long long n_rows;
    int n;                      // rows of dW (= width of delta)
    int swapped;                // 0: P = delta, Q = X;  1: P = X, Q = delta
    float* part;                // [grid][Σ_seg n * w_seg], segment-major, [n][w_seg] inside
    int debug;                  // bit 0: skip the MMAs, bit 1: skip the global loads (profiling aid)
};

__global__ void __launch_bounds__(DW_THREADS, 1) mlp_dw_tc_kernel(const __grid_constant__ DwTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* ring = smem;
    ProvS* prP = reinterpret_cast<ProvS*>(ring + (size_t)DW_STAGES * DW_STAGE_BYTES);
    ProvS* prQ = prP + 1;
    DwBarriers* bar = reinterpret_cast<DwBarriers*>(prQ + 1);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_stg_all = (p.n_rows + DW_RS - 1) / DW_RS;
    const int n_my = (int)((n_stg_all - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int i = 0; i < DW_STAGES; ++i) { mbar_init(&bar->full[i], DW_ROW_THREADS); mbar_init(&bar->empty[i], 1); }
        mbar_init(&bar->d_full, 1);
        fence_barrier_init();
        resolve_segs(p.P, prP);
        resolve_segs(p.Q, prQ);
    }
    const int qw = p.qw, pw = p.pw;
    if (pw < 128) {
        // a 64-wide P occupies two of the four 32-feature panels; the MMA still reads M = 128: zero the rest once
        for (int st = 0; st < DW_STAGES; ++st)
            for (int half = 0; half < 2; ++half) {
                float4* z = reinterpret_cast<float4*>(ring + (size_t)st * DW_STAGE_BYTES + half * DW_P_BYTES + 2 * DW_PANEL);
                for (int i = threadIdx.x; i < 2 * DW_PANEL / 16; i += DW_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        fence_proxy_async_smem();
    }
    if (warp == DW_ROW_THREADS / 32) tmem_alloc(tmem_holder, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_holder;

    if (warp < DW_ROW_THREADS / 32) {
        const int t = threadIdx.x;
        const int nq = qw / 64;                                  // Q chunks of 16 B per thread and stage
        const int qpr = qw / 4;                                  // 16-B chunks per Q row
        const int npc = pw / 64, ppr = pw / 4;                   // the same for P
        // everything about a thread's chunks that does not depend on the stage, once: source segment (pointer to its
        // column, row index map, row pitch, slope), row inside the stage, byte offset in the stage
        // 512 threads cover a stage in passes of 512 chunks; the row pitch in chunks (16, 32 or 64) divides 512, so a
        // thread keeps its column — hence its source segment — over the passes and only its row advances.  Everything
        // that does not depend on the stage is computed once: segment column pointer, row map, pitch, slope, first
        // row and its byte offset in the stage.
        const int prow = t / ppr, pch = t % ppr, pstep = DW_ROW_THREADS / ppr;      // rows per pass
        const int qrow = t / qpr, qch = t % qpr, qstep = DW_ROW_THREADS / qpr;
        const SegS& sgp = prP->seg[prP->seg_of_panel[(pch * 4) >> 5]];
        const SegS& sgq = prQ->seg[prQ->seg_of_panel[(qch * 4) >> 5]];
        const float* const pb = sgp.base + (pch * 4 - sgp.col0);
        const float* const qb = sgq.base + (qch * 4 - sgq.col0);
        const int32_t* const pi = sgp.idx; const int32_t* const qi = sgq.idx;
        const int pld = sgp.ld, qld = sgq.ld;
        const float psl = sgp.slope, qsl = sgq.slope;
        const uint32_t poff = (uint32_t)((pch >> 3) * DW_PANEL) + mn32b_offset(prow, pch & 7);
        const uint32_t qoff = (uint32_t)(2 * DW_P_BYTES + (qch >> 3) * DW_PANEL) + mn32b_offset(qrow, qch & 7);
        // (a row step that is a multiple of 4 keeps the swizzle phase: the offset advances by 128 B per row)
#define SWE_CHUNK_ADDR(base_, idx_, ld_, row_, row0_, out_) do {                             \
            long long g_ = (row0_) + (row_);                                                  \
            if (g_ >= p.n_rows) g_ = p.n_rows - 1;             /* clamped; zeroed when stored */ \
            long long r_ = g_;                                                                \
            if (idx_) r_ = (long long)__ldg((idx_) + g_);                                     \
            out_ = (base_) + r_ * (ld_);                                                      \
        } while (0)
#define SWE_CHUNK_FINISH(slope_, row_, v_, row0_) do {                                        \
            const bool in_ = (row0_) + (row_) < p.n_rows;                                     \
            const float keep_ = in_ ? 1.f : 0.f, sl_ = in_ ? (slope_) : 0.f;                  \
            v_.x = keep_ * fmaxf(v_.x, 0.f) + sl_ * fminf(v_.x, 0.f); v_.y = keep_ * fmaxf(v_.y, 0.f) + sl_ * fminf(v_.y, 0.f); \
            v_.z = keep_ * fmaxf(v_.z, 0.f) + sl_ * fminf(v_.z, 0.f); v_.w = keep_ * fmaxf(v_.w, 0.f) + sl_ * fminf(v_.w, 0.f); \
        } while (0)
        float4 cp[2], cq[4], np[2], nq4[4];
        auto issue = [&](int i, float4 (&vp)[2], float4 (&vq)[4]) {
            const long long row0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * DW_RS;
            const float* ap[2]; const float* aq[4];
#pragma unroll
            for (int j = 0; j < 2; ++j) SWE_CHUNK_ADDR(pb, pi, pld, prow + (j < npc ? j : 0) * pstep, row0, ap[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) SWE_CHUNK_ADDR(qb, qi, qld, qrow + (j < nq ? j : 0) * qstep, row0, aq[j]);
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (j < npc) vp[j] = ldg4(ap[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < nq) vq[j] = ldg4(aq[j]);
        };
        if (p.debug & 2) {
#pragma unroll
            for (int j = 0; j < 2; ++j) cp[j] = np[j] = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
            for (int j = 0; j < 4; ++j) cq[j] = nq4[j] = make_float4(1.f, 2.f, 3.f, 4.f);
        }
        auto store = [&](int i, const float4 (&vp)[2], const float4 (&vq)[4]) {
            const uint32_t slot = i % DW_STAGES;
            mbar_wait(&bar->empty[slot], ((i / DW_STAGES) & 1) ^ 1);
            unsigned char* st = ring + (size_t)slot * DW_STAGE_BYTES;
            const long long row0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * DW_RS;
#pragma unroll
            for (int j = 0; j < 2; ++j)
                if (j < npc) {
                    float4 v = vp[j];
                    SWE_CHUNK_FINISH(psl, prow + j * pstep, v, row0);
                    unsigned char* d = st + poff + j * pstep * 128;
                    split_store4(d, d + DW_P_BYTES, v);
                }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < nq) {
                    float4 v = vq[j];
                    SWE_CHUNK_FINISH(qsl, qrow + j * qstep, v, row0);
                    unsigned char* d = st + qoff + j * qstep * 128;
                    split_store4(d, d + DW_Q_BYTES, v);
                }
            fence_proxy_async_smem();
            mbar_arrive(&bar->full[slot]);
        };
        // two register sets, two stages of loads in flight: a set is re-issued (stage i + 2) as soon as it has
        // been stored (stage i), while the other set's loads (stage i + 1) are still on their way
        const bool ld = !(p.debug & 2);
        if (ld && n_my > 0) issue(0, cp, cq);
        if (ld && n_my > 1) issue(1, np, nq4);
#pragma unroll 1
        for (int i = 0; i < n_my; i += 2) {
            store(i, cp, cq);
            if (ld && i + 2 < n_my) issue(i + 2, cp, cq);
            if (i + 1 < n_my) {
                store(i + 1, np, nq4);
                if (ld && i + 3 < n_my) issue(i + 3, np, nq4);
            }
        }
        // ---- epilogue: TMEM lane m = feature of P, columns = features of Q -> per-CTA partial of dW
#undef SWE_CHUNK_ADDR
#undef SWE_CHUNK_FINISH
        if (n_my > 0) {
            mbar_wait(&bar->d_full, 0);
            tc_fence_after_sync();
            const int qd = warp & 3, hf = warp >> 2;                  // lane quadrant, column-piece phase (0..3)
            const int m = qd * 32 + lane;
            const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
            float* my = p.part + (long long)blockIdx.x * p.n * (p.swapped ? 128 : qw);
            for (int pc = hf; pc < qw / 32 && m < pw; pc += 4) {
                uint32_t v[32];
                tmem_ld32(lane_addr + pc * 32, v);
                tmem_wait_ld();
                if (!p.swapped) {
                    // P = delta (m = row of dW), Q = X: the 32 columns lie inside one segment of X
                    const SegS& sq = prQ->seg[prQ->seg_of_panel[pc]];
                    float* o = my + (long long)sq.col0 * p.n + (long long)m * sq.width + (pc * 32 - sq.col0);
#pragma unroll
                    for (int u = 0; u < 32; u += 4)
                        stg4(o + u, make_float4(__uint_as_float(v[u]), __uint_as_float(v[u + 1]), __uint_as_float(v[u + 2]),
                                                __uint_as_float(v[u + 3])));
                } else {
                    // P = X (m = column of dW), Q = delta (columns = rows of dW)
                    const SegS& sp = prP->seg[prP->seg_of_panel[m >> 5]];
                    float* o = my + (long long)sp.col0 * p.n + (m - sp.col0);
                    const int w = sp.width;
#pragma unroll
                    for (int u = 0; u < 32; ++u) o[(long long)(pc * 32 + u) * w] = __uint_as_float(v[u]);
                }
            }
            tc_fence_before_sync();
        }
    } else if (lane == 0) {
        // ---- MMA issuer
        const uint32_t idesc = make_idesc_tf32_mn(128, qw);
        const uint32_t ring_u32 = smem_u32(ring);
        for (int i = 0; i < n_my; ++i) {
            const uint32_t slot = i % DW_STAGES;
            mbar_wait(&bar->full[slot], (i / DW_STAGES) & 1);
            tc_fence_after_sync();
            const uint32_t p_hi = ring_u32 + slot * DW_STAGE_BYTES, p_lo = p_hi + DW_P_BYTES;
            const uint32_t q_hi = p_hi + 2 * DW_P_BYTES, q_lo = q_hi + DW_Q_BYTES;
#pragma unroll
            for (int ks = 0; ks < DW_RS / 8; ++ks) {
                if (p.debug & 1) break;
                const uint64_t dph = make_desc_mn_sw128_32b(p_hi + ks * 1024, DW_PANEL, 512);
                const uint64_t dpl = make_desc_mn_sw128_32b(p_lo + ks * 1024, DW_PANEL, 512);
                const uint64_t dqh = make_desc_mn_sw128_32b(q_hi + ks * 1024, DW_PANEL, 512);
                const uint64_t dql = make_desc_mn_sw128_32b(q_lo + ks * 1024, DW_PANEL, 512);
                mma_tf32_ss(tmem_base, dpl, dqh, idesc, (i | ks) ? 1u : 0u);
                mma_tf32_ss(tmem_base, dph, dql, idesc, 1u);
                mma_tf32_ss(tmem_base, dph, dqh, idesc, 1u);
            }
            mma_commit(&bar->empty[slot]);
        }
        if (n_my > 0) mma_commit(&bar->d_full);
    }
    __syncthreads();
    if (warp == DW_ROW_THREADS / 32) tmem_dealloc(tmem_base, 256);
}

// =============================================================================================
// dx: out[r, 0:ko) (+)= delta[r, 0:n) · W[0:n, k_off : k_off + ko)
// =============================================================================================
constexpr int DX_A_STAGES = 3;
constexpr int DX_SLOT = 2 * 128 * 128;             // [128 rows x 32] hi + lo = 32 KB

struct DxBarriers { uint64_t a_full[DX_A_STAGES], a_empty[DX_A_STAGES], d_full[2], d_empty[2]; };
constexpr size_t DX_TC_SMEM = 1024 + 2 * 128 * 128 * 4 /* weight image */ + (size_t)DX_A_STAGES * DX_SLOT +
                              sizeof(DxBarriers) + 16;

struct DxTcParams {
    const float* delta; long long n_rows; int n;
    const float* w; int w_ld, k_off, k_valid, ko;
    float* out0; float* out1; int split, acc0, acc1;      // columns [0, split) -> out0 [*, split], the rest -> out1
    // FUSED: delta = dh ⊙ act'(pre) is formed while the rows are staged (written back over dh for the dW kernel) and
    // the per-CTA partial sums of the bias / PReLU-slope gradients are produced on the way: part[cta][n + 1]
    float* dh; const float* pre; const float* slope_p; int act; float* part;
};

template <bool FUSED>
__global__ void __launch_bounds__(TR_THREADS, 1) mlp_dx_tc_kernel(const __grid_constant__ DxTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* w_img = smem;                                    // hi panels, then lo panels
    unsigned char* a_ring = smem + 2 * 128 * 128 * 4;
    DxBarriers* bar = reinterpret_cast<DxBarriers*>(a_ring + (size_t)DX_A_STAGES * DX_SLOT);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = p.n, ko = p.ko, n_chunks = n / 32;
    const uint32_t w_panel = (uint32_t)ko * 128u;                   // bytes of a [ko x 32] weight panel
    const uint32_t w_lo_off = (uint32_t)n_chunks * w_panel;
    const long long n_tiles_all = (p.n_rows + 127) / 128;
    const int n_my = (int)((n_tiles_all - blockIdx.x + gridDim.x - 1) / gridDim.x);

    if (threadIdx.x == 0) {
        for (int i = 0; i < DX_A_STAGES; ++i) { mbar_init(&bar->a_full[i], TR_ROW_THREADS); mbar_init(&bar->a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bar->d_full[i], 1); mbar_init(&bar->d_empty[i], TR_ROW_THREADS); }
        fence_barrier_init();
    }
    // weight image: B[N = output column, K = delta column] K-major, element (nn, k) = W[k, k_off + nn]
    for (int idx = threadIdx.x; idx < n * ko; idx += TR_THREADS) {
        const int k = idx / ko, nn = idx % ko;
        const float v = nn < p.k_valid ? __ldg(p.w + (long long)k * p.w_ld + p.k_off + nn) : 0.f;
        float hi, lo;
        split_tf32(v, hi, lo);
        const uint32_t off = (uint32_t)(k >> 5) * w_panel + sw128_offset(nn, k & 31);
        *reinterpret_cast<float*>(w_img + off) = hi;
        *reinterpret_cast<float*>(w_img + w_lo_off + off) = lo;
    }
    fence_proxy_async_smem();
    if (warp == 8) tmem_alloc(tmem_holder, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_holder;

    if (warp < 8) {
        const int piece = threadIdx.x & 7, r0 = threadIdx.x >> 3;
        const uint32_t g_off0 = sw128_offset(r0, piece * 4);
        uint32_t a_cnt = 0;
        const float fslope = FUSED ? leaky_slope_of(p.act, p.slope_p) : 1.f;
        const bool prelu = FUSED && p.act == SWE_ACT_PRELU;
        float bacc[4][4];                                          // FUSED: Σ_rows delta of this thread's 4 columns per chunk
        float sacc = 0.f;                                          //        Σ dh · pre over pre <= 0 (PReLU slope gradient): per thread
                                                                   //        fp32, across threads fp64 (see swe_backward.cu)
#pragma unroll
        for (int c = 0; c < 4; ++c) { bacc[c][0] = bacc[c][1] = bacc[c][2] = bacc[c][3] = 0.f; }
        auto stage = [&](int i) {
            const long long row0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * 128;
            auto issue = [&](int c, float4 (&v)[4], float4 (&pv)[4]) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const long long g = row0 + r0 + 32 * j;
                    const long long at = g * n + c * 32 + piece * 4;
                    if (FUSED) {
                        v[j] = g < p.n_rows ? *reinterpret_cast<const float4*>(p.dh + at) : make_float4(0.f, 0.f, 0.f, 0.f);
                        pv[j] = g < p.n_rows ? ldg4_stream(p.pre + at) : make_float4(1.f, 1.f, 1.f, 1.f);
                    } else {
                        v[j] = g < p.n_rows ? ldg4_stream(p.delta + at) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            };
            float4 cur[4], nxt[4], pnx[4];
            issue(0, nxt, pnx);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c < n_chunks) {
                    // raw loads of chunk c -> delta (+ partial sums, + write-back)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 d = nxt[j];
                        if (FUSED) {
                            const float4 q = pnx[j];
                            if (prelu)
                                sacc += (q.x > 0.f ? 0.f : d.x * q.x) + (q.y > 0.f ? 0.f : d.y * q.y) +
                                        (q.z > 0.f ? 0.f : d.z * q.z) + (q.w > 0.f ? 0.f : d.w * q.w);
                            d.x *= q.x > 0.f ? 1.f : fslope; d.y *= q.y > 0.f ? 1.f : fslope;
                            d.z *= q.z > 0.f ? 1.f : fslope; d.w *= q.w > 0.f ? 1.f : fslope;
                            bacc[c][0] += d.x; bacc[c][1] += d.y; bacc[c][2] += d.z; bacc[c][3] += d.w;
                            const long long g = row0 + r0 + 32 * j;
                            if (g < p.n_rows) stg4(p.dh + g * n + c * 32 + piece * 4, d);
                        }
                        cur[j] = d;
                    }
                    if (c + 1 < n_chunks) issue(c + 1, nxt, pnx);
                    const uint32_t slot = a_cnt % DX_A_STAGES;
                    mbar_wait(&bar->a_empty[slot], ((a_cnt / DX_A_STAGES) & 1) ^ 1);
                    unsigned char* hi_t = a_ring + (size_t)slot * DX_SLOT + g_off0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) split_store4(hi_t + j * 4096, hi_t + DX_SLOT / 2 + j * 4096, cur[j]);
                    fence_proxy_async_smem();
                    mbar_arrive(&bar->a_full[slot]);
                    ++a_cnt;
                }
            }
        };
        const int qd = warp & 3, hf = warp >> 2;
        const int row = qd * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(qd * 32) << 16);
        auto epilogue = [&](int i) {
            const int buf = i & 1;
            mbar_wait(&bar->d_full[buf], (uint32_t)(i >> 1) & 1);
            tc_fence_after_sync();
            const long long g = ((long long)blockIdx.x + (long long)i * gridDim.x) * 128 + row;
            for (int cb = 0; cb < ko / 64; ++cb) {
                const int col = hf * (ko / 2) + cb * 32;
                uint32_t v[32];
                tmem_ld32(lane_addr + buf * 128 + col, v);
                tmem_wait_ld();
                if (g < p.n_rows) {
                    const bool first = col < p.split;
                    float* o = first ? p.out0 + g * p.split + col : p.out1 + g * (ko - p.split) + (col - p.split);
                    const bool acc = first ? p.acc0 : p.acc1;
#pragma unroll
                    for (int u = 0; u < 32; u += 4) {
                        float4 r = make_float4(__uint_as_float(v[u]), __uint_as_float(v[u + 1]), __uint_as_float(v[u + 2]),
                                               __uint_as_float(v[u + 3]));
                        if (acc) { const float4 old = *reinterpret_cast<const float4*>(o + u); r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w; }
                        stg4(o + u, r);
                    }
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&bar->d_empty[buf]);
        };
        if (n_my > 0) stage(0);
#pragma unroll 1
        for (int i = 0; i < n_my; ++i) {
            if (i + 1 < n_my) stage(i + 1);
            epilogue(i);
        }
        if (FUSED && p.part) {
            // every MMA has completed (the last epilogue waited for it): the operand ring is free and serves as the
            // scratch of a fixed-order reduction over the 32 threads that share a column
            float* red = reinterpret_cast<float*>(a_ring);
            asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int u = 0; u < 4; ++u) red[r0 * 128 + c * 32 + piece * 4 + u] = bacc[c][u];
            double* redd = reinterpret_cast<double*>(red + 32 * 128);
            redd[threadIdx.x] = sacc;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            float* my = p.part + (long long)blockIdx.x * (n + 1);
            if ((int)threadIdx.x < n) {
                float t = 0.f;
                for (int r = 0; r < 32; ++r) t += red[r * 128 + threadIdx.x];
                my[threadIdx.x] = t;
            }
            if (threadIdx.x == 0) {
                double t = 0.0;
                for (int r = 0; r < TR_ROW_THREADS; ++r) t += redd[r];
                my[n] = (float)t;
            }
        }
    } else if (lane == 0) {
        const uint32_t idesc = make_idesc_tf32(128, ko);
        const uint32_t a_u32 = smem_u32(a_ring), w_u32 = smem_u32(w_img);
        uint32_t a_cnt = 0;
        for (int i = 0; i < n_my; ++i) {
            const int buf = i & 1;
            mbar_wait(&bar->d_empty[buf], (((uint32_t)(i >> 1)) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d = tmem_base + buf * 128;
            for (int c = 0; c < n_chunks; ++c, ++a_cnt) {
                const uint32_t slot = a_cnt % DX_A_STAGES;
                mbar_wait(&bar->a_full[slot], (a_cnt / DX_A_STAGES) & 1);
                tc_fence_after_sync();
                const uint32_t a_hi = a_u32 + slot * DX_SLOT, a_lo = a_hi + DX_SLOT / 2;
                const uint32_t w_hi = w_u32 + c * w_panel, w_lo = w_hi + w_lo_off;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint64_t dah = make_desc_sw128(a_hi + ks * 32), dal = make_desc_sw128(a_lo + ks * 32);
                    const uint64_t dwh = make_desc_sw128(w_hi + ks * 32), dwl = make_desc_sw128(w_lo + ks * 32);
                    mma_tf32_ss(d, dal, dwh, idesc, (c | ks) ? 1u : 0u);
                    mma_tf32_ss(d, dah, dwl, idesc, 1u);
                    mma_tf32_ss(d, dah, dwh, idesc, 1u);
                }
                mma_commit(&bar->a_empty[slot]);
            }
            mma_commit(&bar->d_full[buf]);
        }
    }
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 256);
}

}  // namespace tc
}  // namespace swe

using namespace swe;

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
static bool leaky_family(int act) {
    return act == SWE_ACT_NONE || act == SWE_ACT_PRELU || act == SWE_ACT_RELU || act == SWE_ACT_LEAKYRELU;
}
static int provider_width_tc(const swe_rows_t* X) {       // total width, or -1 when the tensor-core path cannot read it
    if (!X || X->n_seg < 1 || X->n_seg > SWE_MAX_SEGS) return -1;
    int w = 0;
    for (int j = 0; j < X->n_seg; ++j) {
        const swe_seg_t& g = X->seg[j];
        if (!g.base || g.width % 32 != 0 || g.width <= 0 || g.ld % 4 != 0 || !aligned16(g.base) || !leaky_family(g.act)) return -1;
        w += g.width;
    }
    return w;
}

static int g_train_tc_debug = 0;
extern "C" void swe_train_tc_set_debug(int v) { g_train_tc_debug = v; }

extern "C" int swe_mlp_layer_bwd_dw_tc_grid(int64_t n_rows) {
    const long long n_stg = (n_rows + tc::DW_RS - 1) / tc::DW_RS;
    return grid_for(n_stg, 1);
}

extern "C" int swe_mlp_layer_bwd_dw_tc(const float* delta, int64_t n_rows, int32_t n, const swe_rows_t* X, float* part,
                                       int32_t* grid_out, void* stream) {
    SWE_REQUIRE(delta && X && part && n_rows >= 0, SWE_E_INVAL, "mlp_layer_bwd_dw_tc: bad arguments");
    SWE_REQUIRE(aligned16(delta) && aligned16(part), SWE_E_ALIGN, "mlp_layer_bwd_dw_tc: unaligned buffer");
    const int xw = provider_width_tc(X);
    SWE_REQUIRE(xw > 0, SWE_E_UNSUPP, "mlp_layer_bwd_dw_tc: provider segments must be 32-column multiples with a leaky-family activation");
    SWE_REQUIRE((n == 128 || n == 64) && (xw == 64 || xw == 128 || xw == 256), SWE_E_UNSUPP,
                "mlp_layer_bwd_dw_tc: unsupported shape n=%d, provider width %d", n, xw);
    if (grid_out) *grid_out = 0;
    if (n_rows == 0) return 0;
    tc::DwTcParams p;
    memset(&p, 0, sizeof(p));
    swe_rows_t D;
    memset(&D, 0, sizeof(D));
    D.n_seg = 1; D.seg[0].base = delta; D.seg[0].idx = nullptr; D.seg[0].slope = nullptr; D.seg[0].ld = n; D.seg[0].width = n;
    D.seg[0].act = SWE_ACT_NONE;
    p.swapped = n == 64 && xw == 128;          // keep M = 128 fully used
    p.P = p.swapped ? *X : D;
    p.Q = p.swapped ? D : *X;
    p.pw = p.swapped ? xw : n;
    p.qw = p.swapped ? n : xw;
    p.n_rows = n_rows; p.n = n; p.part = part; p.debug = g_train_tc_debug;
    cudaError_t e = cudaFuncSetAttribute(tc::mlp_dw_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::DW_TC_SMEM);
    if (e != cudaSuccess) { set_error("mlp_layer_bwd_dw_tc smem opt-in (%zu B): %s", tc::DW_TC_SMEM, cudaGetErrorString(e)); return (int)e; }
    const int grid = swe_mlp_layer_bwd_dw_tc_grid(n_rows);
    tc::mlp_dw_tc_kernel<<<grid, tc::DW_THREADS, tc::DW_TC_SMEM, (cudaStream_t)stream>>>(p);
    if (grid_out) *grid_out = grid;
    return check_launch("mlp_layer_bwd_dw_tc");
}

static int dx_tc_launch(tc::DxTcParams& p, bool fused, void* stream, const char* what) {
    void (*kern)(const tc::DxTcParams) = fused ? tc::mlp_dx_tc_kernel<true> : tc::mlp_dx_tc_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::DX_TC_SMEM);
    if (e != cudaSuccess) { set_error("%s smem opt-in (%zu B): %s", what, tc::DX_TC_SMEM, cudaGetErrorString(e)); return (int)e; }
    const long long n_tiles = (p.n_rows + 127) / 128;
    kern<<<grid_for(n_tiles, 1), tc::TR_THREADS, tc::DX_TC_SMEM, (cudaStream_t)stream>>>(p);
    return check_launch(what);
}

extern "C" int swe_mlp_layer_bwd_dx_tc_grid(int64_t n_rows) { return grid_for((n_rows + 127) / 128, 1); }

extern "C" int swe_mlp_layer_bwd_dx_tc(const float* delta, int64_t n_rows, int32_t n, const float* w, int32_t w_ld,
                                       int32_t k_off, int32_t k_valid, int32_t ko, float* dx0, int32_t accumulate0,
                                       float* dx1, int32_t accumulate1, int32_t split, void* stream) {
    SWE_REQUIRE(delta && w && dx0 && n_rows >= 0, SWE_E_INVAL, "mlp_layer_bwd_dx_tc: bad arguments");
    SWE_REQUIRE((n == 64 || n == 128) && (ko == 64 || ko == 128), SWE_E_UNSUPP, "mlp_layer_bwd_dx_tc: n=%d ko=%d", n, ko);
    SWE_REQUIRE(split == ko || (split == 64 && ko == 128 && dx1), SWE_E_INVAL, "mlp_layer_bwd_dx_tc: split=%d ko=%d", split, ko);
    SWE_REQUIRE(k_valid >= 0 && k_valid <= ko && k_off >= 0 && k_off + k_valid <= w_ld, SWE_E_INVAL, "mlp_layer_bwd_dx_tc: column range");
    SWE_REQUIRE(aligned16(delta) && aligned16(dx0) && (!dx1 || aligned16(dx1)), SWE_E_ALIGN, "mlp_layer_bwd_dx_tc: unaligned buffer");
    if (n_rows == 0) return 0;
    tc::DxTcParams p;
    memset(&p, 0, sizeof(p));
    p.delta = delta; p.n_rows = n_rows; p.n = n; p.w = w; p.w_ld = w_ld; p.k_off = k_off; p.k_valid = k_valid; p.ko = ko;
    p.out0 = dx0; p.out1 = dx1; p.split = split; p.acc0 = accumulate0; p.acc1 = accumulate1;
    return dx_tc_launch(p, false, stream, "mlp_layer_bwd_dx_tc");
}

extern "C" int swe_mlp_layer_bwd_dx_tc_fused(float* dh, const float* pre, int32_t act, const float* slope, int64_t n_rows,
                                             int32_t n, const float* w, int32_t w_ld, int32_t k_off, int32_t k_valid,
                                             int32_t ko, float* dx0, int32_t accumulate0, float* dx1, int32_t accumulate1,
                                             int32_t split, float* part, int32_t* grid_out, void* stream) {
    SWE_REQUIRE(dh && pre && w && dx0 && n_rows >= 0, SWE_E_INVAL, "mlp_layer_bwd_dx_tc_fused: bad arguments");
    SWE_REQUIRE((n == 64 || n == 128) && (ko == 64 || ko == 128), SWE_E_UNSUPP, "mlp_layer_bwd_dx_tc_fused: n=%d ko=%d", n, ko);
    SWE_REQUIRE(leaky_family(act), SWE_E_UNSUPP, "mlp_layer_bwd_dx_tc_fused: activation %d", act);
    SWE_REQUIRE(split == ko || (split == 64 && ko == 128 && dx1), SWE_E_INVAL, "mlp_layer_bwd_dx_tc_fused: split=%d ko=%d", split, ko);
    SWE_REQUIRE(k_valid >= 0 && k_valid <= ko && k_off >= 0 && k_off + k_valid <= w_ld, SWE_E_INVAL, "mlp_layer_bwd_dx_tc_fused: column range");
    SWE_REQUIRE(aligned16(dh) && aligned16(pre) && aligned16(dx0) && (!dx1 || aligned16(dx1)), SWE_E_ALIGN,
                "mlp_layer_bwd_dx_tc_fused: unaligned buffer");
    if (grid_out) *grid_out = 0;
    if (n_rows == 0) return 0;
    tc::DxTcParams p;
    memset(&p, 0, sizeof(p));
    p.delta = dh; p.n_rows = n_rows; p.n = n; p.w = w; p.w_ld = w_ld; p.k_off = k_off; p.k_valid = k_valid; p.ko = ko;
    p.out0 = dx0; p.out1 = dx1; p.split = split; p.acc0 = accumulate0; p.acc1 = accumulate1;
    p.dh = dh; p.pre = pre; p.slope_p = slope; p.act = act; p.part = part;
    if (grid_out) *grid_out = swe_mlp_layer_bwd_dx_tc_grid(n_rows);
    return dx_tc_launch(p, true, stream, "mlp_layer_bwd_dx_tc_fused");
}
