// tmvb_comm.cuh -- peer-memory exchange between the ranks of one NVSwitch box (one process per GPU).
//
// The reference has no multi-device path; SURVEY.md 8(e) adds one exchange per outer iteration (the K x V sufficient
// statistics + a K-vector).  Instead of NCCL all-reduce calls followed by separate normalisation kernels, every rank maps
// its peers' buffers (CUDA IPC handles, exchanged once by the host driver) and ONE kernel per iteration does
// reduce-scatter -> column sums -> normalise -> all-gather with plain loads/stores on the mapped peer pointers over NVLink
// (lda_exchange_mstep_kernel, tmvb_lda.cu).  This header holds the model-independent part: the handle blob, the mapped
// pointer table, and the device-side barriers (CTA-grid barrier; cross-GPU flag barrier with a bounded spin).
#pragma once

#include "tmvb_common.cuh"

namespace tmvb {

constexpr int kMaxPeers = 8;
constexpr int kCommBufs = 5;          // stats | table[0] | table[1] | small | ctl
constexpr int kCtlBytes = 16384;      // per-rank control block (flags, grid barrier, partial sums)
constexpr int kCtlPartOff = 1024;     // double part[2][kCtlPartLen]: column-sum partials | elbo_w partial
constexpr int kCtlPartLen = 264;      // >= max K_ld (256) + 1
constexpr int kSpinTimeoutMsDefault = 4000;  // a peer that does not arrive in time: give up (status flag), never hang; TMVB_COMM_TIMEOUT_MS

struct Comm {
    int rank = 0, world = 1;
    bool connected = false;
    unsigned char *d_ctl = nullptr;          // local control block (cudaMalloc, zeroed)
    double *d_small_red = nullptr;           // [kCtlPartLen + 2] local: `small` summed over ranks
    void *peer[kCommBufs][kMaxPeers] = {};   // mapped pointers; [.][rank] = local
    void *local[kCommBufs] = {};
    unsigned long long epoch = 0;            // three barrier epochs are consumed per exchange
    unsigned long long calls = 0;
    int timeout_ms = kSpinTimeoutMsDefault;  // bounded spin of the device barriers (env TMVB_COMM_TIMEOUT_MS at connect time)
};

// blob = kCommBufs cudaIpcMemHandle_t (64 B each) for the local buffers
int comm_export(Comm *c, void *const local_bufs[kCommBufs], void *blob, size_t blob_bytes);
int comm_connect(Comm *c, int rank, int world, const void *blobs, size_t blob_bytes);
void comm_free(Comm *c);

#ifdef __CUDACC__

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long global_ns()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// control block layout
struct CtlView {
    unsigned long long *flag;   // [kMaxPeers] flag[r] = last epoch rank r has signalled to this rank
    unsigned *grid_count;       // monotonically increasing arrival counter of the local grid barrier
    unsigned *status;           // != 0: a spin timed out
    double *part;               // [2][kCtlPartLen]
};
__device__ __forceinline__ CtlView ctl_view(void *ctl)
{
    unsigned char *b = static_cast<unsigned char *>(ctl);
    CtlView v;
    v.flag = reinterpret_cast<unsigned long long *>(b);
    v.grid_count = reinterpret_cast<unsigned *>(b + 128);
    v.status = reinterpret_cast<unsigned *>(b + 132);
    v.part = reinterpret_cast<double *>(b + kCtlPartOff);
    return v;
}

// All CTAs of a co-resident (cooperative) grid; `target` = gridDim.x * (number of barriers passed so far + 1).
__device__ __forceinline__ void grid_barrier(unsigned *count, unsigned target, unsigned *status, long long timeout_ns)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();   // this CTA's writes (local and peer) before the arrival
        atomicAdd(count, 1u);
        const long long t0 = global_ns();
        while (ld_acquire_gpu(count) < target) {
            if (global_ns() - t0 > timeout_ns) {
                atomicExch(status, 2u);
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
}

// Cross-GPU barrier, executed by CTA 0 between two grid barriers: thread r < world signals rank r and waits for it.
__device__ __forceinline__ void peer_barrier(void *const *peer_ctl, void *my_ctl, int rank, int world, unsigned long long epoch, long long timeout_ns)
{
    const int r = threadIdx.x;
    if (r < world && r != rank) {
        __threadfence_system();
        st_release_sys(ctl_view(peer_ctl[r]).flag + rank, epoch);
        const CtlView me = ctl_view(my_ctl);
        const long long t0 = global_ns();
        while (ld_acquire_sys(me.flag + r) < epoch) {
            if (global_ns() - t0 > timeout_ns) {
                atomicExch(me.status, 1u);
                break;
            }
        }
    }
    __syncthreads();
}

#endif  // __CUDACC__

}  // namespace tmvb
