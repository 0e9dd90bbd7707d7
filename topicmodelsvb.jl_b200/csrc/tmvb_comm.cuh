// tmvb_comm.cuh -- peer-memory exchange between the ranks of one NVSwitch box (one process per GPU).
//
// The reference has no multi-device path; SURVEY.md 8(e) adds one exchange per outer iteration (the K x V sufficient
// statistics + a K-vector).  Instead of NCCL all-reduce calls followed by separate normalisation kernels, every rank maps
// its peers' buffers (CUDA IPC handles, exchanged once by the host driver) and ONE kernel per iteration does
// reduce-scatter -> column sums -> normalise -> all-gather with plain loads/stores on the mapped peer pointers over NVLink
// (lda_exchange_mstep_kernel, tmvb_lda.cu).  This header holds the model-independent part: the handle blob, the mapped
// pointer table, the control block (flag words with bounded spins, arrival counters, epoch) and the memory-order helpers.
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
    int timeout_ms = kSpinTimeoutMsDefault;  // bounded spin of the device barriers (env TMVB_COMM_TIMEOUT_MS at connect time)
};

// blob = kCommBufs cudaIpcMemHandle_t (64 B each) for the local buffers
int comm_export(Comm *c, void *const local_bufs[kCommBufs], void *blob, size_t blob_bytes);
int comm_connect(Comm *c, int rank, int world, const void *blobs, size_t blob_bytes);
void comm_free(Comm *c);

// generic in-place all-reduce over peer memory (tmvb_peer.cu): up to kPeerFloatBufs float buffers (element counts multiples of
// four) and one fp64 vector, summed over the ranks; blob layout = float buffers | small | ctl
constexpr int kPeerFloatBufs = kCommBufs - 2;
struct PeerReduce {
    float *f[kPeerFloatBufs] = {};
    long long nf[kPeerFloatBufs] = {};
    double *small = nullptr;
    long long n_small = 0;
};
int peer_export(Comm *c, const PeerReduce &b, void *blob, size_t blob_bytes);
int peer_allreduce(Comm *c, const PeerReduce &b, cudaStream_t stream, int n_sm);
// reads and clears the sticky time-out flag of the device-side spins (synchronises the stream); non-zero -> fail()
int peer_status(Comm *c, cudaStream_t stream, int *status);

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
    unsigned long long *flag;        // [kMaxPeers] flag[r] = last epoch rank r has signalled to this rank
    unsigned *cnt_b, *cnt_c;         // arrival counters of the local CTAs (phases B and C), reset by CTA 0 at the end of a call
    unsigned *status;                // != 0: a spin timed out
    unsigned long long *epoch;       // three flag epochs are consumed per exchange
    unsigned long long *calls;       // exchanges so far (parity selects the partial-sum buffer)
    unsigned long long *alpha_done;  // progress of the update_alpha! CTA within the current exchange
    double *part;                    // [2][kCtlPartLen]
};
__device__ __forceinline__ CtlView ctl_view(void *ctl)
{
    unsigned char *b = static_cast<unsigned char *>(ctl);
    CtlView v;
    v.flag = reinterpret_cast<unsigned long long *>(b);
    v.cnt_b = reinterpret_cast<unsigned *>(b + 128);
    v.status = reinterpret_cast<unsigned *>(b + 132);
    v.epoch = reinterpret_cast<unsigned long long *>(b + 136);
    v.calls = reinterpret_cast<unsigned long long *>(b + 144);
    v.alpha_done = reinterpret_cast<unsigned long long *>(b + 152);
    v.cnt_c = reinterpret_cast<unsigned *>(b + 160);
    v.part = reinterpret_cast<double *>(b + kCtlPartOff);
    return v;
}

#endif  // __CUDACC__

}  // namespace tmvb
