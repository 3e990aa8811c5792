// Dataset-side helpers of the rollout / training loops on the device:
//   swe_temporal_window : one temporal sample of a simulation — what /root/reference/utils/dataset.py:410-471 (`to_temporal`)
//                         materialises as a Python list of Data objects, one per start time — cut out of the resident
//                         simulation by one kernel: x (static columns + previous_t (depth, discharge) pairs with the
//                         dry-bed prefix), y (rollout targets) and the boundary-condition window
//   swe_rollout_metrics : confusion matrix per time step and threshold (get_rollout_confusion_matrix,
//                         /root/reference/utils/miscellaneous.py:123-151) and the error sums behind get_rollout_loss
//                         (miscellaneous.py:177-199) in one pass over [N, 2, T] predictions — CSI / F1 / RMSE / MAE are a
//                         handful of divisions on the [T, ...] result
// Reductions: per-CTA partial sums combined in CTA order (bit-reproducible).
#include "swe_common.cuh"

namespace swe {

// WD, V: [n, t_sim] (row-major), BC: [n_bc, t_bc]; the reference prepends previous_t - 1 zero columns ("dry bed") to
// WD, V and BC and appends BC's last column once more (dataset.py:426-428): read here through index arithmetic
__global__ void temporal_window_kernel(const float* __restrict__ xs, int n_static, const float* __restrict__ wd,
                                       const float* __restrict__ v, long long n, int t_sim, const float* __restrict__ bc,
                                       int n_bc, int t_bc, int init_time, int previous_t, int rollout_steps,
                                       float* __restrict__ x, float* __restrict__ y, float* __restrict__ bc_out) {
    const int n_cols = n_static + 2 * previous_t;
    const int pad = previous_t - 1;
    const long long work_x = n * n_cols, work_y = n * 2 * rollout_steps, work_b = (long long)n_bc * previous_t * (rollout_steps + 1);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < work_x + work_y + work_b;
         t += (long long)gridDim.x * blockDim.x) {
        if (t < work_x) {
            const long long i = t / n_cols;
            const int c = (int)(t - i * n_cols);
            float val;
            if (c < n_static) val = xs[i * n_static + c];
            else {
                const int k = c - n_static, step = init_time + (k >> 1) - pad;        // column of the un-padded series
                val = step < 0 ? 0.f : ((k & 1) ? v : wd)[i * t_sim + step];
            }
            x[t] = val;
        } else if (t < work_x + work_y) {
            const long long u = t - work_x;
            const long long i = u / (2 * rollout_steps);
            const int r = (int)(u - i * 2 * rollout_steps), var = r / rollout_steps, k = r - var * rollout_steps;
            const int step = init_time + previous_t + k - pad;
            y[u] = step < 0 ? 0.f : (var ? v : wd)[i * t_sim + step];
        } else {
            const long long u = t - work_x - work_y;                                   // [n_bc, previous_t, rollout_steps + 1]
            const int per = previous_t * (rollout_steps + 1);
            const int b = (int)(u / per), rem = (int)(u - (long long)b * per), p = rem / (rollout_steps + 1), r = rem - p * (rollout_steps + 1);
            int step = init_time + r + p - pad;                                        // padded series: [0]*pad | BC | BC[-1]
            step = step > t_bc - 1 ? t_bc - 1 : step;
            bc_out[u] = step < 0 ? 0.f : bc[(long long)b * t_bc + step];
        }
    }
}

// per (cta, t): [0..4K) TP, TN, FP, FN per threshold; then Σd0², Σd1², Σ|d0|, Σ|d1| (all rows), the same four over rows with a
// non-zero difference, and the number of those rows
constexpr int METRIC_MAX_THR = 4;
constexpr int METRIC_COLS = 4 * METRIC_MAX_THR + 9;

__global__ void __launch_bounds__(256) rollout_metrics_partials_kernel(const float* __restrict__ pred, const float* __restrict__ real,
                                                                       long long n, int T, const float* __restrict__ thr, int n_thr,
                                                                       double* __restrict__ partials) {
    // one warp handles one time step at a time over this CTA's share of the rows: lanes stride over rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const long long rows_per_cta = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * rows_per_cta, hi = min(n, lo + rows_per_cta);
    for (int t = warp; t < T; t += n_warps) {
        double acc[METRIC_COLS];
#pragma unroll
        for (int c = 0; c < METRIC_COLS; ++c) acc[c] = 0.0;
        for (long long i = lo + lane; i < hi; i += 32) {
            const float p0 = pred[(i * 2) * T + t], p1 = pred[(i * 2 + 1) * T + t];
            const float r0 = real[(i * 2) * T + t], r1 = real[(i * 2 + 1) * T + t];
#pragma unroll
            for (int k = 0; k < METRIC_MAX_THR; ++k)
                if (k < n_thr) {
                    const bool pf = p0 > thr[k], rf = r0 > thr[k];
                    acc[4 * k + (pf ? (rf ? 0 : 2) : (rf ? 3 : 1))] += 1.0;
                }
            const float d0 = p0 - r0, d1 = p1 - r1;
            double* a = acc + 4 * METRIC_MAX_THR;
            a[0] += (double)d0 * d0; a[1] += (double)d1 * d1; a[2] += fabsf(d0); a[3] += fabsf(d1);
            if (d0 != 0.f || d1 != 0.f) { a[4] += (double)d0 * d0; a[5] += (double)d1 * d1; a[6] += fabsf(d0); a[7] += fabsf(d1); a[8] += 1.0; }
        }
#pragma unroll
        for (int c = 0; c < METRIC_COLS; ++c) {
            double s = acc[c];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) partials[((long long)blockIdx.x * T + t) * METRIC_COLS + c] = s;
        }
    }
}

__global__ void rollout_metrics_reduce_kernel(const double* __restrict__ partials, int n_parts, int T, double* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * METRIC_COLS) return;
    double s = 0.0;
    for (int pi = 0; pi < n_parts; ++pi) s += partials[(long long)pi * T * METRIC_COLS + idx];
    out[idx] = s;
}

}  // namespace swe

using namespace swe;

extern "C" int swe_temporal_window(const float* x_static, int32_t n_static, const float* wd, const float* v, int64_t n, int32_t t_sim,
                                   const float* bc, int32_t n_bc, int32_t t_bc, int32_t init_time, int32_t previous_t,
                                   int32_t rollout_steps, float* x, float* y, float* bc_out, void* stream) {
    SWE_REQUIRE(wd && v && x && y && n >= 0 && n_static >= 0 && (n_static == 0 || x_static) && t_sim >= 1 && previous_t >= 1 &&
                rollout_steps >= 1 && init_time >= 0 && (n_bc == 0 || (bc && bc_out && t_bc >= 1)), SWE_E_INVAL,
                "temporal_window: bad arguments");
    SWE_REQUIRE(init_time + rollout_steps <= t_sim, SWE_E_INVAL, "temporal_window: the window [%d, %d) leaves the %d-step simulation",
                init_time, init_time + rollout_steps, t_sim);
    const long long work = n * (n_static + 2 * previous_t) + n * 2 * rollout_steps + (long long)n_bc * previous_t * (rollout_steps + 1);
    if (work == 0) return 0;
    temporal_window_kernel<<<grid_for((work + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(
        x_static, n_static, wd, v, n, t_sim, bc, n_bc, t_bc, init_time, previous_t, rollout_steps, x, y, bc_out);
    return check_launch("temporal_window");
}

extern "C" int32_t swe_rollout_metrics_cols(void) { return METRIC_COLS; }
extern "C" size_t swe_rollout_metrics_ws_bytes(int32_t T) { return (size_t)NUM_SMS * T * METRIC_COLS * sizeof(double); }

// pred, real: [n, 2, T] fp32; thr: DEVICE array of n_thr (<= 4) water-depth thresholds; out: [T, swe_rollout_metrics_cols()]
// doubles (per threshold TP, TN, FP, FN; then the squared / absolute error sums over all rows and over the wet rows, and
// the number of wet rows)
extern "C" int swe_rollout_metrics(const float* pred, const float* real, int64_t n, int32_t T, const float* thr, int32_t n_thr,
                                   double* out, void* ws, void* stream) {
    SWE_REQUIRE(pred && real && out && ws && n >= 0 && T >= 1 && n_thr >= 0 && n_thr <= METRIC_MAX_THR && (n_thr == 0 || thr),
                SWE_E_INVAL, "rollout_metrics: bad arguments");
    const int parts = (int)(n < NUM_SMS ? (n < 1 ? 1 : n) : NUM_SMS);
    rollout_metrics_partials_kernel<<<parts, 256, 0, (cudaStream_t)stream>>>(pred, real, n, T, thr, n_thr, (double*)ws);
    rollout_metrics_reduce_kernel<<<(T * METRIC_COLS + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const double*)ws, parts, T, out);
    return check_launch("rollout_metrics");
}
