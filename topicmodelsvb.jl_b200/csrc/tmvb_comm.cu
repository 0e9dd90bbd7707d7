// tmvb_comm.cu -- host side of the peer-memory exchange: CUDA IPC export / import of the per-rank buffers.
#include <stdlib.h>
#include <string.h>

#include "tmvb_comm.cuh"

namespace tmvb {

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "blob layout assumes 64-byte IPC handles");

static int env_int_or(const char *name, int dflt)
{
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

int comm_export(Comm *c, void *const local_bufs[kCommBufs], void *blob, size_t blob_bytes)
{
    TMVB_CHECK_ARG(blob != nullptr && blob_bytes >= kCommBufs * sizeof(cudaIpcMemHandle_t), "blob too small");
    if (!c->d_ctl) {
        TMVB_CUDA(cudaMalloc((void **)&c->d_ctl, kCtlBytes));
        TMVB_CUDA(cudaMemset(c->d_ctl, 0, kCtlBytes));
        TMVB_CUDA(cudaMalloc((void **)&c->d_small_red, (kCtlPartLen + 2) * 8));
        TMVB_CUDA(cudaMemset(c->d_small_red, 0, (kCtlPartLen + 2) * 8));
    }
    memset(blob, 0, blob_bytes);
    for (int b = 0; b < kCommBufs; b++) {
        c->local[b] = (b == kCommBufs - 1) ? (void *)c->d_ctl : local_bufs[b];
        TMVB_CHECK_ARG(c->local[b] != nullptr, "buffer not allocated");
        cudaIpcMemHandle_t hnd;
        TMVB_CUDA(cudaIpcGetMemHandle(&hnd, c->local[b]));
        memcpy(static_cast<unsigned char *>(blob) + b * sizeof(hnd), &hnd, sizeof(hnd));
    }
    return 0;
}

int comm_connect(Comm *c, int rank, int world, const void *blobs, size_t blob_bytes)
{
    TMVB_CHECK_ARG(world >= 1 && world <= kMaxPeers, "world size must be 1..8");
    TMVB_CHECK_ARG(rank >= 0 && rank < world, "rank out of range");
    TMVB_CHECK_ARG(c->d_ctl != nullptr, "comm_export must precede comm_connect");
    TMVB_CHECK_ARG(blobs != nullptr && blob_bytes >= kCommBufs * sizeof(cudaIpcMemHandle_t), "blobs missing");
    c->rank = rank;   // set first: comm_free closes whatever has been mapped even if a later handle fails to open
    c->world = world;
    c->timeout_ms = env_int_or("TMVB_COMM_TIMEOUT_MS", kSpinTimeoutMsDefault);
    if (c->timeout_ms < 1) c->timeout_ms = kSpinTimeoutMsDefault;
    for (int r = 0; r < world; r++) {
        for (int b = 0; b < kCommBufs; b++) {
            if (r == rank) {
                c->peer[b][r] = c->local[b];
                continue;
            }
            cudaIpcMemHandle_t hnd;
            memcpy(&hnd, static_cast<const unsigned char *>(blobs) + (size_t)r * blob_bytes + b * sizeof(hnd), sizeof(hnd));
            void *p = nullptr;
            TMVB_CUDA(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
            c->peer[b][r] = p;
        }
    }
    c->connected = true;
    return 0;
}

void comm_free(Comm *c)
{
    for (int r = 0; r < c->world && r < kMaxPeers; r++)
        for (int b = 0; b < kCommBufs; b++)
            if (r != c->rank && c->peer[b][r]) cudaIpcCloseMemHandle(c->peer[b][r]);
    cudaFree(c->d_ctl);
    cudaFree(c->d_small_red);
    *c = Comm();
}

}  // namespace tmvb
