// tmvb_sort.cu -- per-topic vocabulary ranking on the device (the `model.topics = [reverse(sortperm(...))]`
// epilogue of every train!, gpuLDA.jl:374 / gpuCTM.jl:517 / gpuCTPF.jl:707), so that the host does not
// spend K argsorts of V floats after each call.  Plumbing, not a hot kernel: CUB's stable segmented sort.
#include <cub/cub.cuh>

#include "tmvb_common.cuh"
#include "tmvb_recs.cuh"

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

// ---- complete rankings of the CTPF recommendation step (gpuCTPF.jl:716-729) -------------------------------------------
__global__ void rank_iota_kernel(int *__restrict__ vals, int *__restrict__ seg_begin, int *__restrict__ seg_end, int nseg, int len, int ld)
{
    const long long n = (long long)nseg * ld;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int seg = (int)(q / ld), j = (int)(q - (long long)seg * ld);
        vals[q] = j;
        if (j == 0) {
            seg_begin[seg] = seg * ld;
            seg_end[seg] = seg * ld + len;
        }
    }
}
// out[off[seg] + r] = 1 + vals_sorted[seg][len - 1 - r], r < len - nmask[seg]: the reverse of the ascending stable order is
// findall(.)[reverse(sortperm(.))] of the reference (1-based); the masked entries (key -inf) sorted first, i.e. come last here
__global__ void rank_compact_kernel(const int *__restrict__ vals, const int *__restrict__ nmask, const long long *__restrict__ off, int nseg, int len,
                                    int ld, int *__restrict__ out)
{
    for (int seg = blockIdx.x; seg < nseg; seg += gridDim.x) {
        const int keep = len - nmask[seg];
        const long long o = off[seg];
        for (int r = threadIdx.x; r < keep; r += blockDim.x) out[o + r] = 1 + vals[(size_t)seg * ld + (len - 1 - r)];
    }
}

// d_keys: [nseg][ld] scores (masked entries -inf), overwritten.  d_out: concatenated rankings at d_out_off[seg] (device).
int segmented_rank(float *d_keys, int nseg, int len, int ld, const int *d_nmask, const long long *d_out_off, int *d_out, void **ws, size_t *ws_bytes,
                   cudaStream_t stream, int n_sm)
{
    if (nseg == 0 || len == 0) return 0;
    const size_t n = (size_t)nseg * ld;
    if (n >= (1ull << 31)) return fail(-2, "recommendation matrix too large for the device ranking (M * U >= 2^31)");
    size_t cub_bytes = 0;
    float *kout = nullptr;
    int *vin = nullptr, *vout = nullptr, *sb = nullptr, *se = nullptr;
    TMVB_CUDA(cub::DeviceSegmentedSort::StableSortPairs(nullptr, cub_bytes, d_keys, kout, vin, vout, (int)n, nseg, sb, se, stream));
    const size_t need = 3 * n * 4 + 2 * (size_t)nseg * 4 + 512 + cub_bytes;
    if (need > *ws_bytes) {
        if (*ws) TMVB_CUDA(cudaFree(*ws));
        *ws = nullptr;
        *ws_bytes = 0;
        TMVB_CUDA(cudaMalloc(ws, need));
        *ws_bytes = need;
    }
    char *base = (char *)*ws;
    kout = (float *)base;
    vin = (int *)(kout + n);
    vout = vin + n;
    sb = vout + n;
    se = sb + nseg;
    void *cub_ws = (void *)(((uintptr_t)(se + nseg) + 255) & ~(uintptr_t)255);
    const int grid = (int)std::min<long long>((long long)(n + 255) / 256, (long long)n_sm * 16);
    rank_iota_kernel<<<grid, 256, 0, stream>>>(vin, sb, se, nseg, len, ld);
    TMVB_CUDA(cudaGetLastError());
    TMVB_CUDA(cub::DeviceSegmentedSort::StableSortPairs(cub_ws, cub_bytes, d_keys, kout, vin, vout, (int)n, nseg, sb, se, stream));
    rank_compact_kernel<<<std::min(nseg, n_sm * 16), 256, 0, stream>>>(vout, d_nmask, d_out_off, nseg, len, ld, d_out);
    TMVB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace tmvb
