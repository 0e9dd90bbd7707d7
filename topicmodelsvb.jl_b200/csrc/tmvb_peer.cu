// tmvb_peer.cu -- a generic in-place all-reduce (sum) over CUDA-IPC peer memory for the model families whose M-step needs
// nothing but the summed statistics (gpuCTM, gpuCTPF, gpufLDA, gpufCTM; gpuLDA has its own fused exchange + M-step kernel in
// tmvb_lda.cu).  ONE kernel per outer iteration instead of one NCCL all-reduce per buffer: rank r sums slice r of every
// buffer over all ranks with loads from the mapped peer buffers (all peer loads of an element in flight at once, fixed rank
// order, so every rank ends with bit-identical sums) and stores the sum into slice r of EVERY rank's buffer.  Nobody but
// rank r ever reads slice r, so the stores need no second buffer.  Two flag rounds over the control blocks (tmvb_comm.cuh):
// "my statistics are complete" before the first load, "my slice has landed everywhere" before the kernel ends; bounded
// spins with a sticky status word, no grid barrier, no cooperative launch (only the last CTA to finish waits for the peers).
#include <algorithm>

#include "tmvb_comm.cuh"

namespace tmvb {

struct PeerARArgs {
    int rank, world;
    long long timeout_ns;
    float *f[kPeerFloatBufs][kMaxPeers];
    long long nf4[kPeerFloatBufs];   // float4 elements per buffer (0: unused)
    double *small[kMaxPeers];
    long long n_small;
    void *ctl[kMaxPeers];
};

static __device__ __forceinline__ void peer_wait_flags(const CtlView &me, int rank, int world, unsigned long long target, long long timeout_ns)
{
    const int r = threadIdx.x;
    if (r < world && r != rank) {
        const long long t0 = global_ns();
        while (ld_acquire_sys(me.flag + r) < target) {
            if (global_ns() - t0 > timeout_ns) {
                atomicExch(me.status, 1u);
                break;
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const PeerARArgs x)
{
    __shared__ unsigned long long s_epoch;
    __shared__ int s_last;
    const int tid = threadIdx.x, G = gridDim.x;
    const CtlView me = ctl_view(x.ctl[x.rank]);
    if (tid == 0) s_epoch = *me.epoch;
    __syncthreads();
    const unsigned long long epoch = s_epoch;

    // ---- A: every rank's E-step has finished (its statistics are complete)
    if (blockIdx.x == 0 && tid < x.world && tid != x.rank) {
        __threadfence_system();
        st_release_sys(ctl_view(x.ctl[tid]).flag + x.rank, epoch + 1);
    }
    peer_wait_flags(me, x.rank, x.world, epoch + 1, x.timeout_ns);

    // ---- B: reduce slice `rank` of every buffer over the ranks, store the sum into every rank's copy
#pragma unroll
    for (int b = 0; b < kPeerFloatBufs; b++) {
        const long long n4 = x.nf4[b];
        if (n4 <= 0) continue;
        const long long lo = n4 * x.rank / x.world, hi = n4 * (x.rank + 1) / x.world;
        for (long long q = lo + (long long)blockIdx.x * blockDim.x + tid; q < hi; q += (long long)G * blockDim.x) {
            float4 v[kMaxPeers];
#pragma unroll
            for (int pr = 0; pr < kMaxPeers; pr++)
                v[pr] = pr < x.world ? __ldcg(reinterpret_cast<const float4 *>(x.f[b][pr]) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int pr = 0; pr < kMaxPeers; pr++) {
                acc.x += v[pr].x;
                acc.y += v[pr].y;
                acc.z += v[pr].z;
                acc.w += v[pr].w;
            }
            for (int pr = 0; pr < x.world; pr++) reinterpret_cast<float4 *>(x.f[b][pr])[q] = acc;
        }
    }
    {
        const long long lo = x.n_small * x.rank / x.world, hi = x.n_small * (x.rank + 1) / x.world;
        for (long long q = lo + (long long)blockIdx.x * blockDim.x + tid; q < hi; q += (long long)G * blockDim.x) {
            double v[kMaxPeers];
#pragma unroll
            for (int pr = 0; pr < kMaxPeers; pr++) v[pr] = pr < x.world ? __ldcg(x.small[pr] + q) : 0.0;
            double acc = 0.0;
#pragma unroll
            for (int pr = 0; pr < kMaxPeers; pr++) acc += v[pr];
            for (int pr = 0; pr < x.world; pr++) x.small[pr][q] = acc;
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence_system();   // this CTA's stores into the peers' buffers before its arrival
        s_last = (atomicAdd(me.cnt_b, 1u) == (unsigned)(G - 1));
    }
    __syncthreads();
    if (!s_last) return;

    // ---- C (the last CTA of this rank): my slice has landed everywhere; wait until everybody's has
    if (tid < x.world && tid != x.rank) {
        __threadfence_system();
        st_release_sys(ctl_view(x.ctl[tid]).flag + x.rank, epoch + 2);
    }
    peer_wait_flags(me, x.rank, x.world, epoch + 2, x.timeout_ns);
    if (tid == 0) {
        *me.cnt_b = 0u;
        *me.epoch = epoch + 2;
        __threadfence();
    }
}

int peer_export(Comm *c, const PeerReduce &b, void *blob, size_t blob_bytes)
{
    static_assert(kPeerFloatBufs + 2 == kCommBufs, "blob layout: float buffers | small | ctl");
    TMVB_CHECK_ARG(b.f[0] != nullptr && b.small != nullptr, "peer reduce needs at least one statistics buffer and the small vector");
    void *bufs[kCommBufs];
    for (int k = 0; k < kPeerFloatBufs; k++) bufs[k] = b.f[k] ? (void *)b.f[k] : (void *)b.f[0];   // unused slots alias the first buffer
    bufs[kPeerFloatBufs] = b.small;
    bufs[kCommBufs - 1] = nullptr;
    return comm_export(c, bufs, blob, blob_bytes);
}

int peer_allreduce(Comm *c, const PeerReduce &b, cudaStream_t stream, int n_sm)
{
    TMVB_CHECK_ARG(c->connected, "comm_connect has not been called");
    PeerARArgs x;
    memset(&x, 0, sizeof(x));
    x.rank = c->rank;
    x.world = c->world;
    x.timeout_ns = (long long)c->timeout_ms * 1000000ll;
    long long work = 0;
    for (int k = 0; k < kPeerFloatBufs; k++) {
        TMVB_CHECK_ARG(b.nf[k] % 4 == 0, "statistics buffers must hold a multiple of four floats");
        x.nf4[k] = b.f[k] ? b.nf[k] / 4 : 0;
        work = std::max(work, x.nf4[k] / c->world);
        for (int r = 0; r < c->world; r++) x.f[k][r] = (float *)c->peer[k][r];
    }
    x.n_small = b.n_small;
    for (int r = 0; r < c->world; r++) {
        x.small[r] = (double *)c->peer[kPeerFloatBufs][r];
        x.ctl[r] = c->peer[kCommBufs - 1][r];
    }
    const int grid = (int)std::max<long long>(1, std::min<long long>(n_sm, (work + 255) / 256));
    peer_allreduce_kernel<<<grid, 256, 0, stream>>>(x);
    TMVB_CUDA(cudaGetLastError());
    return 0;
}

int peer_status(Comm *c, cudaStream_t stream, int *status)
{
    *status = 0;
    if (!c->connected || !c->d_ctl) return 0;
    unsigned st = 0;
    TMVB_CUDA(cudaMemcpyAsync(&st, c->d_ctl + 132, 4, cudaMemcpyDeviceToHost, stream));
    TMVB_CUDA(cudaStreamSynchronize(stream));
    if (st) {
        cudaMemsetAsync(c->d_ctl + 132, 0, 4, stream);
        *status = (int)st;
        return fail(900 + (int)st, "peer all-reduce timed out: a rank did not reach the exchange within %d ms; the statistics of this iteration are incomplete",
                    c->timeout_ms);
    }
    return 0;
}

}  // namespace tmvb
