// tmvb_shard.cu -- implementation of the shard plumbing shared by the LDA / CTM / CTPF handles.
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <string>

#include "tmvb_shard.cuh"

namespace tmvb {

// ------------------------------------------------------------------ lane layouts ------------
int row_stride(int CH, int lpt)
{
    int r = CH;
    if (lpt < 8)
        while (r % (2 * lpt) != lpt) r++;
    return 4 * r;
}

int pick_layout(int K_ld, int *RS_out)
{
    const int CH = K_ld / 4;
    int best = -1;
    double best_cost = 0.0;
    const int force_lpt = env_int("TMVB_LPT", 0);
    for (int k = 0; k < kNumLaneLayouts; k++) {
        const LaneLayout &l = kLaneLayouts[k];
        if (l.lpt * l.cpl < CH) continue;
        if (force_lpt && l.lpt != force_lpt) continue;
        const int S = 32 / l.lpt, RS = row_stride(CH, l.lpt);
        const double instr_tok = (l.cpl * 9.0 + 2.0 * log2((double)l.lpt) + 12.0) / S;
        const double fixed = (l.cpl + ((K_ld + 31) / 32) * 2.0 * S) / 80.0;
        const double cost = (instr_tok + fixed) * sqrt((double)RS / K_ld);
        if (best < 0 || cost < best_cost - 1e-9) {
            best = k;
            best_cost = cost;
            *RS_out = RS;
        }
    }
    return best;
}

int env_int(const char *name, int dflt)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

int grid_for(long long work, int block, int n_sm)
{
    long long g = (work + block - 1) / block;
    long long cap = (long long)n_sm * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ------------------------------------------------------------------ layout kernels ----------
// dst[p][0..K_ld) = src[perm ? perm[p] : p][0..K) , zero padded
__global__ void pad_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, const int *__restrict__ perm,
                                long long rows, int K, int K_ld)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < rows * K_ld; q += (long long)gridDim.x * blockDim.x) {
        const long long r = q / K_ld;
        const int i = (int)(q - r * K_ld);
        const long long sr = perm ? perm[r] : r;
        dst[q] = (i < K) ? src[sr * K + i] : 0.0f;
    }
}
// dst[perm ? perm[p] : p][0..K) = src[p][0..K)
__global__ void unpad_rows_kernel(const float *__restrict__ src, float *__restrict__ dst, const int *__restrict__ perm,
                                  long long rows, int K, int K_ld)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < rows * K; q += (long long)gridDim.x * blockDim.x) {
        const long long r = q / K;
        const int i = (int)(q - r * K);
        const long long dr = perm ? perm[r] : r;
        dst[dr * K + i] = src[r * K_ld + i];
    }
}
// check_model invariants that need a pass over the data (modelutils.jl:264-273 and the gpuCTM/gpuCTPF
// counterparts), evaluated on the device copy: bit0 non-finite, bit1 sign violation
// (what = 0: x >= 0; 1: x <= 0; 2: x > 0; 3: finite only)
__global__ void validate_kernel(const float *__restrict__ x, long long n, int what, int *__restrict__ err)
{
    int e = 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const float v = x[q];
        if (!isfinite(v)) e |= 1;
        // what == 1 (Elogtheta <= 0): psi(gamma_i) - psi(sum gamma) evaluated in fp32 can come out a few ulp above zero when one topic
        // holds all the mass (K = 1: the two arguments are equal), and such a state must survive a second train! call
        if ((what == 0 && v < 0.f) || (what == 1 && v > 1e-6f) || (what == 2 && !(v > 0.f))) e |= 2;
    }
    if (e) atomicOr(err, e << (2 * what));
}
// CSR re-layout: internal document p takes the tokens of caller document perm[p]; Int64 -> int32 / float
template <typename IntT>
__global__ void pack_corpus_kernel(const IntT *__restrict__ terms64, const IntT *__restrict__ counts64,
                                   const long long *__restrict__ src_off, const long long *__restrict__ dst_off, long long M,
                                   int V, int *__restrict__ terms, float *__restrict__ counts, int *__restrict__ err,
                                   float *__restrict__ doc_c)
{
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (long long d = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); d < M; d += (long long)gridDim.x * wpb) {
        const long long so = src_off[d], o = dst_off[d];
        const int Nd = (int)(dst_off[d + 1] - o);
        float csum = 0.0f;
        for (int n = lane; n < Nd; n += 32) {
            const long long t = terms64[so + n], c = counts64[so + n];
            if (t < 0 || t >= V) atomicOr(err, 1);
            if (c <= 0) atomicOr(err, 2);
            terms[o + n] = (int)t;
            counts[o + n] = (float)c;
            csum += (float)c;
        }
        if (doc_c) {   // C_d: integer-valued, so the fp32 sum is exact below 2^24 whatever the order
            csum = warp_sum(csum);
            if (lane == 0) doc_c[d] = csum;
        }
    }
}
// rowsum_i = sum_j stats[j][i]   (the `sum(beta_temp, dims=2)` of LDA.jl:123)
__global__ void colsum_kernel(const float *__restrict__ stats, int V, int K_ld, double *__restrict__ rowsum)
{
    extern __shared__ double sh[];
    const int R = blockDim.x / K_ld;
    const int r = threadIdx.x / K_ld, i = threadIdx.x - r * K_ld;
    double acc = 0.0;
    if (r < R)
        for (int j = blockIdx.x * R + r; j < V; j += gridDim.x * R) acc += (double)stats[(size_t)j * K_ld + i];
    sh[threadIdx.x] = (r < R) ? acc : 0.0;
    __syncthreads();
    if (threadIdx.x < K_ld) {
        double a = 0.0;
        for (int q = 0; q < R; q++) a += sh[q * K_ld + threadIdx.x];
        if (a != 0.0) atomicAdd(rowsum + threadIdx.x, a);
    }
}
// beta_new = stats ./ rowsum ; stats <- 0 ; elbo_w += sum stats * ln(beta_new + eps)
// (LDA.jl:121-125 / CTM.jl:114-118 and the Elogpw term LDA.jl:64-67 / CTM.jl:69-73 rewritten over the statistics).
// beta_old != NULL: elbo_w += sum stats * [ln(beta_new + eps) - ln(beta_old + eps)] -- the second term is the part of
// sum_d sum_n c_n H(phi_n) (LDA.jl:76-79) that is linear in the statistics (ln u_ni = ln beta_old_i,w + Elogtheta_old_i),
// which spares the E-step kernel one logarithm per (token, topic).
__global__ void normalize_kernel(float *__restrict__ stats, float *__restrict__ beta_new, const float *__restrict__ beta_old,
                                 const double *__restrict__ rowsum, int V, int K, int K_ld, double *__restrict__ elbo_w, int want_elbo)
{
    // thread (rl, c): rows rl + RPP * n, the 16-byte chunk c of each -- the four column reciprocals are per-thread constants
    // (one fp64 division per thread instead of one per element: the element-wise version ran at 0.5 TB/s on L2-resident data)
    const int CH = K_ld >> 2, RPP = blockDim.x / CH;
    const int rl = threadIdx.x / CH, c = threadIdx.x - rl * CH;
    double acc = 0.0;
    if (rl < RPP) {
        double inv[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double rs = (4 * c + j < K) ? rowsum[4 * c + j] : 0.0;
            inv[j] = rs > 0.0 ? 1.0 / rs : 0.0;
        }
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = blockIdx.x * RPP + rl; r < V; r += gridDim.x * RPP) {
            const size_t q = (size_t)r * K_ld + 4 * c;
            const float4 S = *reinterpret_cast<const float4 *>(stats + q);
            float4 b;
            b.x = (float)((double)S.x * inv[0]);
            b.y = (float)((double)S.y * inv[1]);
            b.z = (float)((double)S.z * inv[2]);
            b.w = (float)((double)S.w * inv[3]);
            if (want_elbo) {
                float4 l = make_float4(logf(b.x + TMVB_EPS), logf(b.y + TMVB_EPS), logf(b.z + TMVB_EPS), logf(b.w + TMVB_EPS));
                if (beta_old) {
                    const float4 bo = *reinterpret_cast<const float4 *>(beta_old + q);
                    l.x -= logf(bo.x + TMVB_EPS);
                    l.y -= logf(bo.y + TMVB_EPS);
                    l.z -= logf(bo.z + TMVB_EPS);
                    l.w -= logf(bo.w + TMVB_EPS);
                }
                if (4 * c + 0 < K) acc += (double)(S.x * l.x);
                if (4 * c + 1 < K) acc += (double)(S.y * l.y);
                if (4 * c + 2 < K) acc += (double)(S.z * l.z);
                if (4 * c + 3 < K) acc += (double)(S.w * l.w);
            }
            *reinterpret_cast<float4 *>(beta_new + q) = b;
            *reinterpret_cast<float4 *>(stats + q) = z4;
        }
    }
    if (want_elbo) {
        acc = warp_sum_d(acc);
        if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(elbo_w, acc);
    }
}

// ------------------------------------------------------------------ shard ------------------
int shard_create(Shard *s, int64_t K, int64_t M, int64_t V, int device, void *stream, size_t pinned_doubles)
{
    TMVB_CHECK_ARG(K > 0, "number of topics must be a positive integer");  // gpuLDA.jl:47
    TMVB_CHECK_ARG(M >= 0 && V >= 0, "M and V must be nonnegative");
    TMVB_CHECK_ARG(M < (1ll << 31) && V < (1ll << 31), "M and V must fit in int32");
    if (K > 256) return fail(-2, "K=%lld is not supported (K <= 256)", (long long)K);
    const int K_ld = (int)((K + 7) / 8 * 8);
    int RS = 0;
    const int li = pick_layout(K_ld, &RS);
    if (li < 0) return fail(-2, "internal: no lane layout for K=%lld", (long long)K);
    int ndev = 0;
    TMVB_TRY(tmvb_device_count(&ndev));
    if (ndev == 0) return fail(-3, "no CUDA device: libtmvb has no CPU fallback");
    if (device < 0) TMVB_CUDA(cudaGetDevice(&device));
    TMVB_CHECK_ARG(device < ndev, "device index out of range");
    TMVB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TMVB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(-3, "device %d is sm_%d%d; libtmvb is built for sm_100a only", device, prop.major, prop.minor);

    s->device = device;
    s->n_sm = prop.multiProcessorCount;
    s->smem_optin = prop.sharedMemPerBlockOptin;
    s->K = K;
    s->M = M;
    s->V = V;
    s->K_ld = K_ld;
    s->RS = RS;
    s->layout = li;
    s->lpt = kLaneLayouts[li].lpt;
    s->cpl = kLaneLayouts[li].cpl;
    if (stream) {
        s->stream = (cudaStream_t)stream;
    } else {
        TMVB_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        s->own_stream = true;
    }
    const size_t kv = (size_t)std::max<int64_t>(V, 1) * K_ld;
    for (int b = 0; b < 2; b++) {
        TMVB_CUDA(cudaMalloc((void **)&s->d_beta[b], kv * 4));
        TMVB_CUDA(cudaMemsetAsync(s->d_beta[b], 0, kv * 4, s->stream));
    }
    TMVB_CUDA(cudaMalloc((void **)&s->d_stats, kv * 4));
    TMVB_CUDA(cudaMemsetAsync(s->d_stats, 0, kv * 4, s->stream));
    TMVB_CUDA(cudaMalloc((void **)&s->d_rowchk, 256 * 8));
    TMVB_CUDA(cudaMalloc((void **)&s->d_counters, 64 * 4));
    TMVB_CUDA(cudaMemsetAsync(s->d_counters, 0, 64 * 4, s->stream));
    s->pinned_doubles = pinned_doubles;
    TMVB_CUDA(cudaMallocHost((void **)&s->h_pinned, pinned_doubles * 8));
    for (auto &e : s->ev) TMVB_CUDA(cudaEventCreate(&e));
    s->n_streams = std::min(1 + Shard::kAux, std::max(1, env_int("TMVB_STREAMS", 4)));
    s->use_graphs = env_int("TMVB_GRAPH", 1) != 0;
    for (int a = 0; a + 1 < s->n_streams; a++) {
        TMVB_CUDA(cudaStreamCreateWithFlags(&s->aux[a], cudaStreamNonBlocking));
        TMVB_CUDA(cudaEventCreateWithFlags(&s->ev_join[a], cudaEventDisableTiming));
    }
    TMVB_CUDA(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
    return 0;
}

void shard_free(Shard *s)
{
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    shard_drop_graphs(s);
    cudaFree(s->d_doc_off);
    cudaFree(s->d_src_off);
    cudaFree(s->d_terms);
    cudaFree(s->d_perm);
    cudaFree(s->d_counts);
    cudaFree(s->d_doc_c);
    cudaFree(s->d_beta[0]);
    cudaFree(s->d_beta[1]);
    cudaFree(s->d_stats);
    cudaFree(s->d_rowchk);
    cudaFree(s->d_counters);
    cudaFree(s->d_scratch);
    cudaFree(s->d_sort_ws);
    if (s->h_pinned) cudaFreeHost(s->h_pinned);
    for (auto &e : s->ev)
        if (e) cudaEventDestroy(e);
    for (auto &a : s->aux)
        if (a) cudaStreamDestroy(a);
    for (auto &e : s->ev_join)
        if (e) cudaEventDestroy(e);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    *s = Shard();
}

int shard_scratch(Shard *s, size_t bytes)
{
    if (bytes <= s->scratch_bytes) return 0;
    if (s->d_scratch) TMVB_CUDA(cudaFree(s->d_scratch));
    s->d_scratch = nullptr;
    s->scratch_bytes = 0;
    TMVB_CUDA(cudaMalloc(&s->d_scratch, bytes));
    s->scratch_bytes = bytes;
    return 0;
}

// Split the length-sorted documents into launches whose shared-memory tile capacity ("cap", in tokens) fits the
// longest document of the launch.
static int plan_buckets(Shard *s, size_t fixed_bytes)
{
    s->buckets.clear();
    const std::vector<int> &len_sorted = s->len_sorted;
    const int M = (int)len_sorted.size();
    if (M == 0) return 0;
    const size_t per_tok = (size_t)s->RS * 4 + 8 + s->per_tok_extra;
    if (fixed_bytes + 16 * per_tok > s->smem_optin) return fail(-4, "internal: E-step working set does not fit in shared memory");
    int cap_max = 16;
    while (fixed_bytes + (size_t)(cap_max + 16) * per_tok <= s->smem_optin) cap_max += 16;
    // developer knob: a smaller tile trades shared-memory reads for L2 reads of the overflow rows and raises the occupancy
    if (s->tile_cap_max > 0) cap_max = std::min(cap_max, s->tile_cap_max);
    cap_max = std::max(16, std::min(cap_max, env_int("TMVB_TILE_CAP_MAX", cap_max)));
    std::vector<int> caps;
    for (int c = 16; c < cap_max; c = (c < 128) ? c + 16 : (c < 256 ? c + 32 : c + c / 4 / 16 * 16)) caps.push_back(c);
    caps.push_back(cap_max);
    int begin = 0;  // documents are sorted by length, longest first
    for (int ci = (int)caps.size() - 1; ci >= 0 && begin < M; ci--) {
        const int lo = (ci == 0) ? -1 : caps[ci - 1];  // this launch takes lengths in (lo, caps[ci]] (+ overflow for the largest)
        int end = begin;
        while (end < M && len_sorted[end] > lo) end++;
        if (end == begin) continue;
        Bucket b;
        b.doc_begin = begin;
        b.doc_end = end;
        b.cap = std::min(caps[ci], std::max(16, (len_sorted[begin] + 15) / 16 * 16));
        if (b.cap > cap_max) b.cap = cap_max;
        b.cap2 = 0;
        b.warps = 1;
        b.nr = 0;
        b.hyb = 0;
        b.smem = fixed_bytes + (size_t)b.cap * per_tok;
        b.grid = 0;
        s->buckets.push_back(b);
        begin = end;
    }
    if ((int)s->buckets.size() > kMaxBuckets) return fail(-1, "internal: too many launch buckets");
    return 0;
}

int shard_set_corpus(Shard *s, const int64_t *N_cumsum, const void *terms, const void *counts, size_t fixed_bytes, int elem_bytes)
{
    TMVB_CHECK_ARG(elem_bytes == 8 || elem_bytes == 4, "token arrays must be Int64 or Int32");
    TMVB_CHECK_ARG(N_cumsum != nullptr, "N_cumsum is NULL");
    TMVB_CUDA(cudaSetDevice(s->device));
    const int64_t M = s->M;
    TMVB_CHECK_ARG(N_cumsum[0] == 0, "N_cumsum[0] must be 0");
    const int64_t nnz = N_cumsum[M];
    TMVB_CHECK_ARG(nnz >= 0, "N_cumsum must be nondecreasing");
    TMVB_CHECK_ARG(nnz == 0 || (terms != nullptr && counts != nullptr), "terms/counts are NULL");

    // the token arrays do not depend on the document order: start their host -> device copies first so that they
    // overlap the host-side sort below (the caller's arrays stay alive until the synchronize at the end of this call)
    if (nnz > 0) {
        TMVB_TRY(shard_scratch(s, (size_t)nnz * 16));
        unsigned char *t_in = (unsigned char *)s->d_scratch, *c_in = t_in + (size_t)nnz * elem_bytes;
        TMVB_CUDA(cudaMemcpyAsync(t_in, terms, (size_t)nnz * elem_bytes, cudaMemcpyHostToDevice, s->stream));
        TMVB_CUDA(cudaMemcpyAsync(c_in, counts, (size_t)nnz * elem_bytes, cudaMemcpyHostToDevice, s->stream));
        s->st.h2d_bytes += nnz * 2 * elem_bytes;
    }

    // host: O(M) counting sort of the documents by length (descending, stable)
    std::vector<int> len(M);
    int maxlen = 0;
    for (int64_t d = 0; d < M; d++) {
        const int64_t l = N_cumsum[d + 1] - N_cumsum[d];
        if (l < 0 || l > (1 << 30)) {
            cudaStreamSynchronize(s->stream);  // the copies above still read the caller's arrays
            return fail(-1, "invalid argument: N_cumsum must be nondecreasing (document %lld)", (long long)d);
        }
        len[d] = (int)l;
        maxlen = std::max(maxlen, (int)l);
    }
    std::vector<int64_t> start((size_t)maxlen + 2, 0);
    for (int64_t d = 0; d < M; d++) start[maxlen - len[d] + 1]++;
    for (int l = 0; l <= maxlen; l++) start[l + 1] += start[l];
    s->h_perm.assign(M, 0);
    for (int64_t d = 0; d < M; d++) s->h_perm[start[maxlen - len[d]]++] = (int)d;
    std::vector<long long> src_off(std::max<int64_t>(M, 1)), dst_off(M + 1);
    s->len_sorted.assign(M, 0);
    dst_off[0] = 0;
    for (int64_t p = 0; p < M; p++) {
        const int d = s->h_perm[p];
        src_off[p] = N_cumsum[d];
        s->len_sorted[p] = len[d];
        dst_off[p + 1] = dst_off[p] + len[d];
    }
    if (int rc = plan_buckets(s, fixed_bytes)) {
        cudaStreamSynchronize(s->stream);
        return rc;
    }

    const size_t nz = (size_t)std::max<int64_t>(nnz, 1);
    if (!s->d_doc_off) {
        TMVB_CUDA(cudaMalloc((void **)&s->d_doc_off, (M + 1) * 8));
        TMVB_CUDA(cudaMalloc((void **)&s->d_src_off, std::max<int64_t>(M, 1) * 8));
        TMVB_CUDA(cudaMalloc((void **)&s->d_perm, std::max<int64_t>(M, 1) * 4));
        TMVB_CUDA(cudaMalloc((void **)&s->d_doc_c, std::max<int64_t>(M, 1) * 4));
    }
    if (nz > s->nnz_cap) {  // token arrays are reused across calls while they fit
        cudaFree(s->d_terms);
        cudaFree(s->d_counts);
        s->d_terms = nullptr;
        s->d_counts = nullptr;
        s->nnz_cap = 0;
        TMVB_CUDA(cudaMalloc((void **)&s->d_terms, nz * 4));
        TMVB_CUDA(cudaMalloc((void **)&s->d_counts, nz * 4));
        s->nnz_cap = nz;
    }
    TMVB_CUDA(cudaMemcpyAsync(s->d_doc_off, dst_off.data(), (M + 1) * 8, cudaMemcpyHostToDevice, s->stream));
    if (M > 0) {
        TMVB_CUDA(cudaMemcpyAsync(s->d_src_off, src_off.data(), M * 8, cudaMemcpyHostToDevice, s->stream));
        TMVB_CUDA(cudaMemcpyAsync(s->d_perm, s->h_perm.data(), M * 4, cudaMemcpyHostToDevice, s->stream));
    }
    s->st.h2d_bytes += (M + 1) * 8 + M * 12;
    if (nnz > 0) {
        unsigned char *t_in = (unsigned char *)s->d_scratch, *c_in = t_in + (size_t)nnz * elem_bytes;
        TMVB_CUDA(cudaMemsetAsync(s->d_counters + 63, 0, 4, s->stream));
        const int grid = grid_for(M * 32, 256, s->n_sm);
        if (elem_bytes == 8)
            pack_corpus_kernel<long long><<<grid, 256, 0, s->stream>>>((const long long *)t_in, (const long long *)c_in, s->d_src_off, s->d_doc_off, M,
                                                                       (int)s->V, s->d_terms, s->d_counts, s->d_counters + 63, s->d_doc_c);
        else
            pack_corpus_kernel<int><<<grid, 256, 0, s->stream>>>((const int *)t_in, (const int *)c_in, s->d_src_off, s->d_doc_off, M, (int)s->V,
                                                                 s->d_terms, s->d_counts, s->d_counters + 63, s->d_doc_c);
        s->st.kernel_launches++;
        TMVB_CUDA(cudaGetLastError());
        int err = 0;
        TMVB_CUDA(cudaMemcpyAsync(&err, s->d_counters + 63, 4, cudaMemcpyDeviceToHost, s->stream));
        TMVB_CUDA(cudaStreamSynchronize(s->stream));  // also keeps the host vectors alive long enough
        if (err & 1) return fail(-1, "invalid argument: terms must lie in [0, V)");
        if (err & 2) return fail(-1, "invalid argument: all counts must be positive integers");  // Corpus.jl:43
    } else {
        TMVB_CUDA(cudaStreamSynchronize(s->stream));
    }
    s->nnz = nnz;
    s->corpus_set = true;
    return 0;
}

int shard_upload_rows(Shard *s, const float *host, float *d_dst, int64_t rows, const int *d_perm, int validate)
{
    if (!host || rows == 0) return 0;
    const size_t n = (size_t)rows * s->K;
    TMVB_TRY(shard_scratch(s, n * 4));
    TMVB_CUDA(cudaMemcpyAsync(s->d_scratch, host, n * 4, cudaMemcpyHostToDevice, s->stream));
    if (validate >= 0) {
        validate_kernel<<<grid_for(n, 256, s->n_sm), 256, 0, s->stream>>>((const float *)s->d_scratch, n, validate, s->d_counters + 62);
        s->st.kernel_launches++;
    }
    pad_rows_kernel<<<grid_for(rows * s->K_ld, 256, s->n_sm), 256, 0, s->stream>>>((const float *)s->d_scratch, d_dst, d_perm, rows, (int)s->K, s->K_ld);
    TMVB_CUDA(cudaGetLastError());
    s->st.kernel_launches++;
    s->st.h2d_bytes += n * 4;
    return 0;
}

__global__ void rowsum_check_kernel(const double *__restrict__ rowsum, int K, int *__restrict__ err)
{
    const int i = threadIdx.x;
    // isapprox(sum, 1) with Julia's default rtol = sqrt(eps(Float32)) (utils.jl:150-160 isstochastic on a Float32 matrix)
    if (i < K && !(fabs(rowsum[i] - 1.0) <= 3.4526698e-4)) atomicOr(err, 1 << 14);
}

int shard_check_stochastic(Shard *s, const float *d_table)
{
    if (s->V == 0) return 0;
    TMVB_CUDA(cudaMemsetAsync(s->d_rowchk, 0, 256 * 8, s->stream));
    const int R = std::max(1, 256 / s->K_ld);
    const int threads = std::max(R * s->K_ld, s->K_ld);
    const int grid = (int)std::min<int64_t>((s->V + R - 1) / R, (int64_t)s->n_sm * 8);
    colsum_kernel<<<grid, threads, threads * 8, s->stream>>>(d_table, (int)s->V, s->K_ld, s->d_rowchk);
    rowsum_check_kernel<<<1, 256, 0, s->stream>>>(s->d_rowchk, (int)s->K, s->d_counters + 62);
    TMVB_CUDA(cudaGetLastError());
    s->st.kernel_launches += 2;
    return 0;
}

int shard_validation(Shard *s, int *mask)
{
    *mask = 0;
    TMVB_CUDA(cudaMemcpyAsync(mask, s->d_counters + 62, 4, cudaMemcpyDeviceToHost, s->stream));
    TMVB_CUDA(cudaStreamSynchronize(s->stream));
    TMVB_CUDA(cudaMemsetAsync(s->d_counters + 62, 0, 4, s->stream));
    return 0;
}

int shard_download_rows(Shard *s, const float *d_src, float *host, int64_t rows, const int *d_perm)
{
    if (!host || rows == 0) return 0;
    const size_t n = (size_t)rows * s->K;
    TMVB_TRY(shard_scratch(s, n * 4));
    unpad_rows_kernel<<<grid_for(n, 256, s->n_sm), 256, 0, s->stream>>>(d_src, (float *)s->d_scratch, d_perm, rows, (int)s->K, s->K_ld);
    TMVB_CUDA(cudaGetLastError());
    s->st.kernel_launches++;
    TMVB_CUDA(cudaMemcpyAsync(host, s->d_scratch, n * 4, cudaMemcpyDeviceToHost, s->stream));
    TMVB_CUDA(cudaStreamSynchronize(s->stream));
    s->st.d2h_bytes += n * 4;
    return 0;
}

const void *pick_by_warps(const Bucket &b, const void *ctx) { return static_cast<const void *const *>(ctx)[b.warps - 1]; }

// the launches of one E-step, enqueued on s->stream and its auxiliary streams (fork / join through events)
int shard_enqueue_buckets(Shard *s, BucketKernelFn pick, const void *ctx, void *dev_struct)
{
    TMVB_CUDA(cudaMemsetAsync(s->d_counters, 0, kMaxBuckets * 4, s->stream));
    const int ns = (s->buckets.size() > 1) ? s->n_streams : 1;
    if (ns > 1) {
        TMVB_CUDA(cudaEventRecord(s->ev_fork, s->stream));
        for (int a = 0; a + 1 < ns; a++) TMVB_CUDA(cudaStreamWaitEvent(s->aux[a], s->ev_fork, 0));
    }
    for (size_t bi = 0; bi < s->buckets.size(); bi++) {
        Bucket &b = s->buckets[bi];
        const void *fn = pick(b, ctx);
        int *counter = s->d_counters + bi;
        void *args[] = {dev_struct, (void *)&b.doc_begin, (void *)&b.doc_end, (void *)&b.cap, (void *)&b.cap2, (void *)&counter};
        cudaStream_t st = (bi % ns == 0) ? s->stream : s->aux[bi % ns - 1];
        TMVB_CUDA(cudaLaunchKernel(fn, dim3(b.grid), dim3(32 * b.warps), args, b.smem, st));
    }
    for (int a = 0; a + 1 < ns; a++) {
        TMVB_CUDA(cudaEventRecord(s->ev_join[a], s->aux[a]));
        TMVB_CUDA(cudaStreamWaitEvent(s->stream, s->ev_join[a], 0));
    }
    return 0;
}

void shard_drop_graphs(Shard *s)
{
    for (auto &g : s->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    s->graphs.clear();
}

// One E-step is up to ~10 short launches on 4 streams; on a multi-GPU shard each lasts 20-130 us, so the ~100 us the
// host needs to issue them (8 processes competing for the host cores) would bound the E-step.  The launch sequence is
// therefore captured once per distinct (kernel set, by-value parameter block) into a CUDA graph and replayed with a
// single cudaGraphLaunch; the by-value struct changes only with beta's double-buffer parity, want_elbo and the
// train! keywords, so a handful of graphs serves a whole training run.  TMVB_GRAPH=0 disables the capture.
int shard_launch_key(Shard *s, BucketKernelFn pick, const void *ctx, const void *dev_struct, size_t dev_struct_bytes, std::string *key_out)
{
    std::string key((const char *)dev_struct, dev_struct_bytes);
    for (Bucket &b : s->buckets) {
        const void *fn = pick(b, ctx);
        if (!fn) return fail(-4, "internal: no E-step kernel for a launch bucket (warps=%d nr=%d)", b.warps, b.nr);
        if (b.grid == 0) {
            int occ = 0;
            TMVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 32 * b.warps, b.smem));
            if (occ < 1) return fail(-4, "internal: E-step kernel does not fit (cap=%d warps=%d smem=%zu)", b.cap, b.warps, b.smem);
            b.grid = std::min(b.doc_end - b.doc_begin, occ * s->n_sm);
            // short launches (multi-GPU shards): fewer CTAs that each draw several documents, so that the launches of one E-step
            // share the machine instead of each paying a CTA start per document (NSF K=50, one rank's shard of 8 / 4 / 2 / 1:
            // E-step 0.248 / 0.418 / 0.729 / 1.357 ms with one document per CTA, 0.225 / 0.395 / 0.722 / 1.344 ms with four)
            const int dpc = env_int("TMVB_DOCS_PER_CTA", 4);
            if (dpc > 1) b.grid = std::max(1, std::min(b.grid, (b.doc_end - b.doc_begin + dpc - 1) / dpc));
        }
        // the launch geometry is part of the key, so a re-planned corpus never replays a stale graph
        const long long geo[8] = {(long long)(size_t)fn, b.doc_begin, b.doc_end, b.cap, b.cap2, b.grid, (long long)b.smem, b.warps + 64 * b.nr + 4096 * b.hyb};
        key.append((const char *)geo, sizeof(geo));
    }
    *key_out = key;
    return 0;
}

int shard_launch(Shard *s, BucketKernelFn pick, const void *ctx, void *dev_struct, size_t dev_struct_bytes)
{
    if (s->buckets.empty()) return 0;
    std::string key;
    TMVB_TRY(shard_launch_key(s, pick, ctx, dev_struct, dev_struct_bytes, &key));
    s->st.kernel_launches += (int64_t)s->buckets.size();
    if (!s->use_graphs) return shard_enqueue_buckets(s, pick, ctx, dev_struct);

    for (auto &g : s->graphs)
        if (g.key == key) {
            TMVB_CUDA(cudaGraphLaunch(g.exec, s->stream));
            return 0;
        }
    if (s->graphs.size() >= 16) shard_drop_graphs(s);
    cudaGraph_t graph = nullptr;
    TMVB_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = shard_enqueue_buckets(s, pick, ctx, dev_struct);
    const cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
    if (rc != 0 || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        s->use_graphs = false;  // capture is not possible here (e.g. the caller's stream is already capturing): launch directly
        return shard_enqueue_buckets(s, pick, ctx, dev_struct);
    }
    Shard::LaunchGraph lg;
    lg.key = key;
    const cudaError_t ei = cudaGraphInstantiate(&lg.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) {
        cudaGetLastError();
        s->use_graphs = false;
        return shard_enqueue_buckets(s, pick, ctx, dev_struct);
    }
    s->graphs.push_back(lg);
    TMVB_CUDA(cudaGraphLaunch(lg.exec, s->stream));
    return 0;
}

int shard_normalize(Shard *s, double *d_acc, bool want_elbo, bool entropy_term)
{
    if (s->V == 0) return 0;
    TMVB_CUDA(cudaMemsetAsync(d_acc, 0, (2 * s->K_ld) * 8, s->stream));
    const int R = std::max(1, 256 / s->K_ld);
    const int threads = std::max(R * s->K_ld, s->K_ld);
    const int grid = (int)std::min<int64_t>((s->V + R - 1) / R, (int64_t)s->n_sm * 8);
    colsum_kernel<<<grid, threads, threads * 8, s->stream>>>(s->d_stats, (int)s->V, s->K_ld, d_acc);
    TMVB_CUDA(cudaGetLastError());
    {
        const int CH = s->K_ld / 4, RPP = std::max(1, 256 / CH);
        const int ngrid = (int)std::min<int64_t>((s->V + RPP - 1) / RPP, (int64_t)s->n_sm * 8);
        normalize_kernel<<<ngrid, std::max(256, CH), 0, s->stream>>>(s->d_stats, s->d_beta[s->cur ^ 1], (want_elbo && entropy_term) ? s->d_beta[s->cur] : nullptr,
                                                                   d_acc, (int)s->V, (int)s->K, s->K_ld, d_acc + s->K_ld, want_elbo ? 1 : 0);
    }
    TMVB_CUDA(cudaGetLastError());
    s->st.kernel_launches += 2;
    s->cur ^= 1;  // beta_old <- beta ; beta <- new  (LDA.jl:122-123)
    return 0;
}

int shard_topics(Shard *s, const float *d_mat, const float *d_scale, int32_t *out)
{
    if (s->V == 0) return 0;
    const size_t bytes = (size_t)s->K * s->V * 4;
    TMVB_TRY(shard_scratch(s, bytes));
    TMVB_TRY(topics_argsort(d_mat, d_scale, (int)s->K, s->K_ld, (int)s->V, (int *)s->d_scratch, &s->d_sort_ws, &s->sort_ws_bytes, s->stream,
                            s->n_sm));
    s->st.kernel_launches += 3;
    TMVB_CUDA(cudaMemcpyAsync(out, s->d_scratch, bytes, cudaMemcpyDeviceToHost, s->stream));
    TMVB_CUDA(cudaStreamSynchronize(s->stream));
    s->st.d2h_bytes += bytes;
    return 0;
}

int shard_pack_aux(Shard *s, const int64_t *cumsum, const int64_t *ids, const int64_t *vals, int64_t id_limit, long long **d_off, int **d_ids,
                   float **d_vals, int64_t *nnz_out, std::vector<int> *len_internal)
{
    TMVB_CHECK_ARG(s->corpus_set, "set_corpus must precede the reader lists");
    TMVB_CHECK_ARG(cumsum != nullptr && cumsum[0] == 0, "R_cumsum[0] must be 0");
    const int64_t M = s->M, nnz = cumsum[M];
    TMVB_CHECK_ARG(nnz >= 0 && (nnz == 0 || (ids && vals)), "reader arrays are NULL");
    std::vector<long long> src_off(std::max<int64_t>(M, 1)), dst_off(M + 1);
    len_internal->assign(M, 0);
    dst_off[0] = 0;
    for (int64_t p = 0; p < M; p++) {
        const int d = s->h_perm[p];
        const int64_t l = cumsum[d + 1] - cumsum[d];
        if (l < 0 || l > (1 << 30)) return fail(-1, "invalid argument: R_cumsum must be nondecreasing (document %lld)", (long long)d);
        src_off[p] = cumsum[d];
        (*len_internal)[p] = (int)l;
        dst_off[p + 1] = dst_off[p] + l;
    }
    cudaFree(*d_off);
    cudaFree(*d_ids);
    cudaFree(*d_vals);
    *d_off = nullptr;
    *d_ids = nullptr;
    *d_vals = nullptr;
    const size_t nz = (size_t)std::max<int64_t>(nnz, 1);
    long long *d_src = nullptr;
    TMVB_CUDA(cudaMalloc((void **)d_off, (M + 1) * 8));
    TMVB_CUDA(cudaMalloc((void **)d_ids, nz * 4));
    TMVB_CUDA(cudaMalloc((void **)d_vals, nz * 4));
    TMVB_CUDA(cudaMalloc((void **)&d_src, std::max<int64_t>(M, 1) * 8));
    TMVB_CUDA(cudaMemcpyAsync(*d_off, dst_off.data(), (M + 1) * 8, cudaMemcpyHostToDevice, s->stream));
    if (M > 0) TMVB_CUDA(cudaMemcpyAsync(d_src, src_off.data(), M * 8, cudaMemcpyHostToDevice, s->stream));
    s->st.h2d_bytes += (2 * M + 1) * 8;
    int err = 0;
    if (nnz > 0) {
        TMVB_TRY(shard_scratch(s, (size_t)nnz * 16));
        long long *t64 = (long long *)s->d_scratch, *c64 = t64 + nnz;
        TMVB_CUDA(cudaMemcpyAsync(t64, ids, nnz * 8, cudaMemcpyHostToDevice, s->stream));
        TMVB_CUDA(cudaMemcpyAsync(c64, vals, nnz * 8, cudaMemcpyHostToDevice, s->stream));
        s->st.h2d_bytes += nnz * 16;
        TMVB_CUDA(cudaMemsetAsync(s->d_counters + 63, 0, 4, s->stream));
        pack_corpus_kernel<long long><<<grid_for(M * 32, 256, s->n_sm), 256, 0, s->stream>>>(t64, c64, d_src, *d_off, M, (int)id_limit, *d_ids, *d_vals,
                                                                                 s->d_counters + 63, nullptr);
        s->st.kernel_launches++;
        TMVB_CUDA(cudaGetLastError());
        TMVB_CUDA(cudaMemcpyAsync(&err, s->d_counters + 63, 4, cudaMemcpyDeviceToHost, s->stream));
    }
    TMVB_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(d_src);
    if (err & 1) return fail(-1, "invalid argument: readers must lie in [0, U)");
    if (err & 2) return fail(-1, "invalid argument: all ratings must be positive integers");  // Corpus.jl:46
    *nnz_out = nnz;
    return 0;
}

int shard_get_stats(Shard *s, const double *d_sweeps, tmvb_stats *out)
{
    TMVB_CUDA(cudaSetDevice(s->device));
    TMVB_CUDA(cudaStreamSynchronize(s->stream));
    float ms = 0.f;
    if (s->estep_timed) {
        TMVB_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
        s->st.estep_ms = ms;
        if (d_sweeps) {
            TMVB_CUDA(cudaMemcpy(s->h_pinned, d_sweeps, 8, cudaMemcpyDeviceToHost));
            s->st.sweeps = (int64_t)s->h_pinned[0];
        }
    }
    if (s->mstep_timed) {
        TMVB_CUDA(cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]));
        s->st.mstep_ms = ms;
    }
    *out = s->st;
    return 0;
}

}  // namespace tmvb
