// Plan construction kernels (integer work): stable destination-CSR build, weight packing, and the
// library's error plumbing.  Results are bit-exact against oracle/plan_oracle.py.
#include <cub/device/device_scan.cuh>
#include <stdarg.h>
#include "swe_common.cuh"

namespace swe {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

__global__ void pack_linear_kernel(const float* __restrict__ w, int n_out, int k_in, int k_pad, float* __restrict__ wt) {
    const int total = k_pad * n_out;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int k = idx / n_out, n = idx % n_out;
        wt[idx] = (k < k_in) ? w[(long long)n * k_in + k] : 0.f;
    }
}

__device__ __forceinline__ int map_id(const int32_t* __restrict__ node_map, long long v) {
    return node_map ? node_map[v] : (int)v;
}

__global__ void csr_count_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col, long long n_edges,
                                 const int32_t* __restrict__ node_map, int dst_lo, int n_dst, int src_lo, int src_hi,
                                 int by_row, int32_t* __restrict__ counts, int32_t* __restrict__ err) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (long long)gridDim.x * blockDim.x) {
        const int key = map_id(node_map, by_row ? row[e] : col[e]);
        const int oth = map_id(node_map, by_row ? col[e] : row[e]);
        if (key < dst_lo || key >= dst_lo + n_dst || oth < src_lo || oth >= src_hi) { atomicAdd(err, 1); continue; }
        atomicAdd(counts + (key - dst_lo), 1);
    }
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col, long long n_edges,
                                const int32_t* __restrict__ node_map, int dst_lo, int n_dst, int src_lo, int src_hi,
                                int by_row, int32_t* __restrict__ cursor, int32_t* __restrict__ eid) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (long long)gridDim.x * blockDim.x) {
        const int key = map_id(node_map, by_row ? row[e] : col[e]);
        const int oth = map_id(node_map, by_row ? col[e] : row[e]);
        if (key < dst_lo || key >= dst_lo + n_dst || oth < src_lo || oth >= src_hi) continue;
        const int pos = atomicAdd(cursor + (key - dst_lo), 1);
        eid[pos] = (int32_t)e;
    }
}

// one thread per destination: order its segment by original edge id (=> stable), emit src/dst
__global__ void csr_finish_kernel(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                                  const int32_t* __restrict__ node_map, int dst_lo, int n_dst, int by_row,
                                  const int32_t* __restrict__ rowptr, int32_t* __restrict__ eid,
                                  int32_t* __restrict__ src, int32_t* __restrict__ dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_dst; i += gridDim.x * blockDim.x) {
        const int p0 = rowptr[i], p1 = rowptr[i + 1];
        for (int a = p0 + 1; a < p1; ++a) {                 // insertion sort: in-degrees are tiny on meshes
            const int32_t v = eid[a];
            int b = a - 1;
            while (b >= p0 && eid[b] > v) { eid[b + 1] = eid[b]; --b; }
            eid[b + 1] = v;
        }
        for (int p = p0; p < p1; ++p) {
            const long long e = eid[p];
            src[p] = map_id(node_map, by_row ? col[e] : row[e]);
            dst[p] = dst_lo + i;
        }
    }
}

}  // namespace swe

using namespace swe;

extern "C" int swe_abi_version(void) { return SWE_ABI_VERSION; }
extern "C" const char* swe_last_error(void) { return g_err; }
extern "C" const char* swe_build_arch(void) { return "sm_100a"; }

extern "C" int swe_pack_linear(const float* w, int32_t n_out, int32_t k_in, int32_t k_pad, float* wt, void* stream) {
    SWE_REQUIRE(w && wt && n_out >= 1 && k_in >= 1 && k_pad >= k_in, SWE_E_INVAL, "pack_linear: bad arguments");
    const int total = k_pad * n_out;
    pack_linear_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, n_out, k_in, k_pad, wt);
    return check_launch("pack_linear");
}

static size_t scan_temp_bytes(int n_items) {
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, n_items);
    return bytes;
}

extern "C" size_t swe_csr_build_ws_bytes(int64_t n_edges, int32_t n_dst) {
    (void)n_edges;
    const size_t a = ((size_t)(n_dst + 1) * sizeof(int32_t) + 255) & ~(size_t)255;
    return 2 * a + scan_temp_bytes(n_dst + 1) + 256;
}

extern "C" int swe_csr_build(const int64_t* row, const int64_t* col, int64_t n_edges, const int32_t* node_map,
                             int32_t dst_lo, int32_t n_dst, int32_t src_lo, int32_t src_hi, int32_t by_row,
                             int32_t* rowptr, int32_t* src, int32_t* dst, int32_t* eid, int32_t* err_flag,
                             void* ws, size_t ws_bytes, void* stream) {
    SWE_REQUIRE(rowptr && err_flag && ws && n_edges >= 0 && n_dst >= 0 && dst_lo >= 0, SWE_E_INVAL, "csr_build: bad arguments");
    SWE_REQUIRE(n_edges == 0 || (row && col && src && dst && eid), SWE_E_INVAL, "csr_build: null edge buffers");
    SWE_REQUIRE(n_edges < (1ll << 31), SWE_E_UNSUPP, "csr_build: more than 2^31-1 edges in one edge set");
    SWE_REQUIRE(ws_bytes >= swe_csr_build_ws_bytes(n_edges, n_dst), SWE_E_INVAL, "csr_build: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t a = ((size_t)(n_dst + 1) * sizeof(int32_t) + 255) & ~(size_t)255;
    int32_t* counts = reinterpret_cast<int32_t*>(ws);
    int32_t* cursor = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws) + a);
    void* scan_tmp = reinterpret_cast<char*>(ws) + 2 * a;
    size_t scan_bytes = scan_temp_bytes(n_dst + 1);
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)(n_dst + 1) * sizeof(int32_t), st);
    if (e != cudaSuccess) { set_error("csr_build memset: %s", cudaGetErrorString(e)); return (int)e; }
    const int blocks = (int)((n_edges + 255) / 256 < 148 * 16 ? (n_edges + 255) / 256 + 1 : 148 * 16);
    if (n_edges > 0)
        csr_count_kernel<<<blocks, 256, 0, st>>>(row, col, n_edges, node_map, dst_lo, n_dst, src_lo, src_hi, by_row, counts, err_flag);
    e = cub::DeviceScan::ExclusiveSum(scan_tmp, scan_bytes, counts, rowptr, n_dst + 1, st);
    if (e != cudaSuccess) { set_error("csr_build scan: %s", cudaGetErrorString(e)); return (int)e; }
    if (n_edges > 0) {
        e = cudaMemcpyAsync(cursor, rowptr, (size_t)(n_dst + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) { set_error("csr_build copy: %s", cudaGetErrorString(e)); return (int)e; }
        csr_fill_kernel<<<blocks, 256, 0, st>>>(row, col, n_edges, node_map, dst_lo, n_dst, src_lo, src_hi, by_row, cursor, eid);
        const int nb = (n_dst + 255) / 256 < 148 * 16 ? (n_dst + 255) / 256 + 1 : 148 * 16;
        csr_finish_kernel<<<nb, 256, 0, st>>>(row, col, node_map, dst_lo, n_dst, by_row, rowptr, eid, src, dst);
    }
    return check_launch("csr_build");
}
