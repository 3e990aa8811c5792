// Halo exchange of a partitioned mesh over peer memory (NVLink / NVSwitch): one kernel per exchange writes this rank's
// boundary rows straight into the halo rows of its neighbours' arrays (the arrays live in IPC-mapped arenas), signals
// them, and waits for their signals — no pack buffer, no NCCL call, no host involvement, so the whole partitioned
// rollout step is one fixed kernel sequence (captured in a CUDA graph).  The reference has no distributed code
// (SURVEY.md §2.1); the contract is SURVEY.md §8(e): owned rows bit-identical to the single-GPU result.
//
// Protocol (one monotone sequence number per rank, kept in device memory so that graph replays advance it):
//   push : every CTA copies its share of the rows, then __threadfence_system(); the LAST CTA to finish (device-scope
//          counter) stores seq into flag[me] on every neighbour (st.release.sys)
//   wait : one thread per neighbour spins on flag[neighbour] >= seq (ld.acquire.sys) with a time-out trap
// The neighbour sets are symmetric (union of send and receive peers of the scale), every exchange is awaited before its
// rows are consumed, and no local kernel writes halo rows (the plan's destinations are owned rows only), so a rank can be
// at most one exchange ahead of a neighbour and never overwrites rows that neighbour still reads.
#include "swe_common.cuh"

namespace swe {

constexpr int HALO_MAX_PEERS = 16;

struct HaloPeer {
    const int32_t* send_idx;        // local rows to send, in the order the neighbour's halo range expects them
    long long n_send;
    float* remote_rows;             // neighbour's array, first halo row that belongs to this rank (peer-mapped)
    unsigned int* remote_flag;      // neighbour's flag slot of this rank (peer-mapped)
    const unsigned int* local_flag; // this rank's flag slot of the neighbour
};

struct HaloParams {
    const float* arr; int width;
    int n_peers;
    HaloPeer peer[HALO_MAX_PEERS];
    unsigned int* seq;              // this rank's exchange counter (device memory)
    unsigned int* done;             // CTA completion counter of the push (device memory, zero between launches)
    int do_push, do_wait;
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) halo_push_wait_kernel(const __grid_constant__ HaloParams p) {
    __shared__ int s_last;
    const unsigned int seq = *p.seq + 1u;                       // the exchange this launch performs
    if (p.do_push) {
        const int lanes = p.width >> 2;                         // 16-byte pieces per row
        long long total = 0;
        for (int q = 0; q < p.n_peers; ++q) total += p.peer[q].n_send;
        const long long work = total * lanes;
        for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < work; t += (long long)gridDim.x * blockDim.x) {
            long long i = t / lanes;
            const int piece = (int)(t - i * lanes);
            int q = 0;
            while (i >= p.peer[q].n_send) { i -= p.peer[q].n_send; ++q; }
            const HaloPeer& pe = p.peer[q];
            const float4 v = ldg4(p.arr + (long long)__ldg(pe.send_idx + i) * p.width + 4 * piece);
            *reinterpret_cast<float4*>(pe.remote_rows + i * p.width + 4 * piece) = v;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(p.done, 1u) == gridDim.x - 1) ? 1 : 0;
        __syncthreads();
        if (!s_last) return;
        __threadfence_system();                                  // the other CTAs' rows are ordered before the flags below
        if (threadIdx.x == 0) *p.done = 0u;
        if ((int)threadIdx.x < p.n_peers) st_release_sys(p.peer[threadIdx.x].remote_flag, seq);
    } else if (blockIdx.x != 0) {
        return;
    }
    if (p.do_wait) {
        if ((int)threadIdx.x < p.n_peers) {
            const unsigned int* f = p.peer[threadIdx.x].local_flag;
            const long long t0 = clock64();
            while ((int)(ld_acquire_sys(f) - seq) < 0) {
                if (clock64() - t0 > 20000000000ll) {            // ~10 s: a neighbour died or the schedules diverged
                    printf("swe halo: rank waits for exchange %u of peer slot %d (flag %u)\n", seq, (int)threadIdx.x, ld_acquire_sys(f));
                    __trap();
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) *p.seq = seq;
    }
}

}  // namespace swe

using namespace swe;

extern "C" int swe_ipc_alloc(size_t bytes, void** ptr_out, unsigned char* handle64) {
    SWE_REQUIRE(ptr_out && handle64 && bytes > 0, SWE_E_INVAL, "ipc_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { set_error("ipc_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return (int)e; }
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), p);
    if (e != cudaSuccess) { set_error("ipc_alloc: %s", cudaGetErrorString(e)); cudaFree(p); return (int)e; }
    *ptr_out = p;
    return 0;
}

extern "C" int swe_ipc_open(const unsigned char* handle64, void** ptr_out) {
    SWE_REQUIRE(handle64 && ptr_out, SWE_E_INVAL, "ipc_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("ipc_open: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int swe_ipc_close(void* ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) { set_error("ipc_close: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int swe_ipc_free(void* ptr) {
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) { set_error("ipc_free: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int swe_halo_exchange(const float* arr, int32_t width, int32_t n_peers, const int32_t* const* send_idx,
                                 const int64_t* n_send, float* const* remote_rows, uint32_t* const* remote_flags,
                                 const uint32_t* const* local_flags, uint32_t* seq, uint32_t* done, int32_t do_push,
                                 int32_t do_wait, void* stream) {
    SWE_REQUIRE(arr && seq && done && n_peers >= 0 && n_peers <= HALO_MAX_PEERS && width >= 4 && (width & 3) == 0, SWE_E_INVAL,
                "halo_exchange: bad arguments");
    SWE_REQUIRE(aligned16(arr), SWE_E_ALIGN, "halo_exchange: unaligned array");
    if (n_peers == 0 || (!do_push && !do_wait)) return 0;
    HaloParams p;
    memset(&p, 0, sizeof(p));
    p.arr = arr; p.width = width; p.n_peers = n_peers; p.seq = seq; p.done = done; p.do_push = do_push; p.do_wait = do_wait;
    long long total = 0;
    for (int q = 0; q < n_peers; ++q) {
        SWE_REQUIRE(n_send[q] >= 0 && (n_send[q] == 0 || (send_idx[q] && remote_rows[q])) && remote_flags[q] && local_flags[q],
                    SWE_E_INVAL, "halo_exchange: peer %d incomplete", q);
        p.peer[q].send_idx = send_idx[q]; p.peer[q].n_send = n_send[q]; p.peer[q].remote_rows = remote_rows[q];
        p.peer[q].remote_flag = remote_flags[q]; p.peer[q].local_flag = local_flags[q];
        total += n_send[q];
    }
    const long long work = total * (width >> 2);
    int grid = 1;
    if (do_push) grid = (int)((work + 255) / 256 < 1 ? 1 : ((work + 255) / 256 > 2 * NUM_SMS ? 2 * NUM_SMS : (work + 255) / 256));
    halo_push_wait_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    return check_launch("halo_exchange");
}
