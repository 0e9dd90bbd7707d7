// tmvb_sort.cu -- per-topic vocabulary ranking on the device (the `model.topics = [reverse(sortperm(...))]`
// epilogue of every train!, gpuLDA.jl:374 / gpuCTM.jl:517 / gpuCTPF.jl:707), so that the host does not
// spend K argsorts of V floats after each call.  Plumbing, not a hot kernel: CUB's stable segmented sort.
#include <cub/cub.cuh>

#include "tmvb_common.cuh"

namespace tmvb {

// keys[i*V + j] = mat[j*ld + i] * scale_i ; vals[i*V + j] = j
__global__ void topics_gather_kernel(const float *__restrict__ mat, const float *__restrict__ scale, int K, int ld, int V,
                                     float *__restrict__ keys, int *__restrict__ vals)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < (long long)K * V; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q / V), j = (int)(q - (long long)i * V);
        keys[q] = mat[(size_t)j * ld + i] * (scale ? scale[i] : 1.0f);
        vals[q] = j;
    }
}
// out[i*V + r] = 1 + vals_sorted[i*V + (V-1-r)]   (reverse of the ascending stable order == reverse(sortperm(.)), 1-based)
__global__ void topics_reverse_kernel(const int *__restrict__ vals, int K, int V, int *__restrict__ out)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < (long long)K * V; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q / V), r = (int)(q - (long long)i * V);
        out[q] = 1 + vals[(size_t)i * V + (V - 1 - r)];
    }
}

// d_mat: [V][ld] floats (topic i of term j at j*ld+i).  d_out: int[K*V] device.  ws/ws_bytes: caller-kept workspace.
int topics_argsort(const float *d_mat, const float *d_scale, int K, int ld, int V, int *d_out, void **ws, size_t *ws_bytes,
                   cudaStream_t stream, int n_sm)
{
    if (K == 0 || V == 0) return 0;
    const size_t n = (size_t)K * V;
    size_t cub_bytes = 0;
    float *kin = nullptr, *kout = nullptr;
    int *vin = nullptr, *vout = nullptr, *offs = nullptr;
    TMVB_CUDA(cub::DeviceSegmentedSort::StableSortPairs(nullptr, cub_bytes, kin, kout, vin, vout, (int)n, K, offs, offs + 1, stream));
    const size_t need = 4 * n * 4 + (K + 1) * 4 + 256 + cub_bytes;
    if (need > *ws_bytes) {
        if (*ws) TMVB_CUDA(cudaFree(*ws));
        *ws = nullptr;
        *ws_bytes = 0;
        TMVB_CUDA(cudaMalloc(ws, need));
        *ws_bytes = need;
    }
    char *base = (char *)*ws;
    kin = (float *)base;
    kout = kin + n;
    vin = (int *)(kout + n);
    vout = vin + n;
    offs = vout + n;
    void *cub_ws = (void *)(((uintptr_t)(offs + K + 1) + 255) & ~(uintptr_t)255);
    std::vector<int> h_offs(K + 1);
    for (int i = 0; i <= K; i++) h_offs[i] = i * V;
    TMVB_CUDA(cudaMemcpyAsync(offs, h_offs.data(), (K + 1) * 4, cudaMemcpyHostToDevice, stream));
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)n_sm * 16);
    topics_gather_kernel<<<grid, 256, 0, stream>>>(d_mat, d_scale, K, ld, V, kin, vin);
    TMVB_CUDA(cudaGetLastError());
    TMVB_CUDA(cub::DeviceSegmentedSort::StableSortPairs(cub_ws, cub_bytes, kin, kout, vin, vout, (int)n, K, offs, offs + 1, stream));
    topics_reverse_kernel<<<grid, 256, 0, stream>>>(vout, K, V, d_out);
    TMVB_CUDA(cudaGetLastError());
    TMVB_CUDA(cudaStreamSynchronize(stream));  // h_offs staging
    return 0;
}

}  // namespace tmvb
