// tmvb_shard.cuh -- host-side state and plumbing common to the three model handles: one shard of documents
// on one device (CSR re-layout, length buckets, topic-word table + statistics, staging, streams, counters).
#pragma once

#include <string>
#include <vector>

#include "tmvb_estep.cuh"

namespace tmvb {

struct Bucket {
    int doc_begin, doc_end, cap, grid;
    int cap2;  // second tile capacity (CTPF reader lists); 0 otherwise
    int warps; // warps cooperating on one document in this launch (CTA = 32 * warps threads)
    int nr;    // > 0: register rounds per warp (register capacity warps * nr * S tokens); 0: shared-memory tile kernel
    int hyb;   // 1: hybrid kernel (nr register rounds per warp + a tile of `cap` tokens); 0: pure register (nr > 0) / pure tile kernel
    size_t smem;
};

struct Shard {
    int device = 0, n_sm = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    static constexpr int kAux = 11;  // bucket launches are spread over stream + aux streams so their tails overlap
    cudaStream_t aux[kAux] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kAux] = {};
    int n_streams = 1;

    int64_t K = 0, M = 0, V = 0, nnz = 0;
    size_t nnz_cap = 0;
    int K_ld = 0, RS = 0, lpt = 0, cpl = 0, layout = -1;
    bool corpus_set = false;

    // corpus, internal order = documents sorted by length (descending)
    long long *d_doc_off = nullptr, *d_src_off = nullptr;
    int *d_terms = nullptr, *d_perm = nullptr;
    float *d_counts = nullptr;
    float *d_doc_c = nullptr;   // [M] sum of a document's counts (C_d of the reference's structs, gpuLDA.jl:52), internal order
    std::vector<int> h_perm, len_sorted;
    std::vector<Bucket> buckets;
    size_t per_tok_extra = 0;   // shared-memory bytes per staged token beyond the row, its count and its id (fLDA: tau, tau_old, kappa weight)
    int tile_cap_max = 0;   // > 0: largest shared-memory tile (tokens) a launch bucket may stage; longer documents read the rest from L2
    // captured launch sequences of one E-step (see shard_launch), keyed by kernel set + by-value parameter block
    struct LaunchGraph {
        std::string key;
        cudaGraphExec_t exec = nullptr;
    };
    std::vector<LaunchGraph> graphs;
    bool use_graphs = true;

    // topic-word table (double buffered: [cur] current, [cur^1] previous) and its sufficient statistics
    float *d_beta[2] = {nullptr, nullptr}, *d_stats = nullptr;
    int cur = 0;

    int *d_counters = nullptr;  // [0..47] bucket work counters, [61] tau range flag, [62] validation flags, [63] corpus error flags
    double *d_rowchk = nullptr; // [K_ld] row sums of an uploaded topic-word table (shard_check_stochastic)
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    void *d_sort_ws = nullptr;
    size_t sort_ws_bytes = 0;
    double *h_pinned = nullptr;  // small pinned read-back buffer
    size_t pinned_doubles = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool estep_timed = false, mstep_timed = false;
    tmvb_stats st{};
};

constexpr int kMaxBuckets = 48;

int env_int(const char *name, int dflt);
int grid_for(long long work, int block, int n_sm);

int shard_create(Shard *s, int64_t K, int64_t M, int64_t V, int device, void *stream, size_t pinned_doubles);
void shard_free(Shard *s);
int shard_scratch(Shard *s, size_t bytes);
// CSR upload + device re-layout (modelutils.jl:371-388); buckets are planned with smem(cap) = cap*(RS*4+8) + fixed_bytes
// terms / counts: Int64 (elem_bytes = 8, what update_buffer! builds) or Int32 (elem_bytes = 4, a host-side packed cache)
int shard_set_corpus(Shard *s, const int64_t *N_cumsum, const void *terms, const void *counts, size_t fixed_bytes, int elem_bytes = 8);
// host [rows][K] (caller order) -> device [rows][K_ld] (internal order when perm); validate: -1 none, 0 x>=0, 1 x<=0, 2 x>0
int shard_upload_rows(Shard *s, const float *host, float *d_dst, int64_t rows, const int *d_perm, int validate);
int shard_download_rows(Shard *s, const float *d_src, float *host, int64_t rows, const int *d_perm);
// isstochastic(table, dims=2) of check_model (modelutils.jl:268,293,...) on the device copy: every row sum within sqrt(eps(Float32)) of one;
// asynchronous, a violation sets bit 14 of the validation mask
int shard_check_stochastic(Shard *s, const float *d_table);
// returns and clears the validation bit mask accumulated by shard_upload_rows (2 bits per `validate` code)
int shard_validation(Shard *s, int *mask);
// launch an E-step kernel `fn(Dev, doc_begin, doc_end, cap, cap2, counter)` over every bucket; pick(bucket, ctx) returns the
// instantiation for the bucket's (warps, nr)
typedef const void *(*BucketKernelFn)(const Bucket &b, const void *ctx);
int shard_launch(Shard *s, BucketKernelFn pick, const void *ctx, void *dev_struct, size_t dev_struct_bytes);
// the pieces of shard_launch for callers that capture a longer sequence themselves (tmvb_lda_iterate): grids + cache key of the
// launch set, and the bare enqueue of the bucket launches on s->stream and its auxiliary streams (fork / join through events)
int shard_launch_key(Shard *s, BucketKernelFn pick, const void *ctx, const void *dev_struct, size_t dev_struct_bytes, std::string *key_out);
int shard_enqueue_buckets(Shard *s, BucketKernelFn pick, const void *ctx, void *dev_struct);
void shard_drop_graphs(Shard *s);
// update_alpha! (LDA.jl:97-118 == fLDA.jl:122-146) on the device: small[0..K) = sum_d Elogtheta_d; alpha64 / alpha32 are updated in place (tmvb_lda.cu)
int lda_launch_alpha(double *alpha64, float *alpha32, const double *small, int K, int K_ld, double Md, int niter, double ntol, cudaStream_t stream);
// pick for kernels that only come in "warps per document" flavours: ctx = const void *const fn_by_warps[]
const void *pick_by_warps(const Bucket &b, const void *ctx);
// beta_new = stats ./ rowsum ; stats <- 0 ; [elbo_w = sum stats ln(beta_new + eps)].  d_acc: double[2*K_ld] (rowsum | elbo_w)
// entropy_term: elbo_w additionally receives - sum stats ln(beta_old + eps) (see normalize_kernel)
int shard_normalize(Shard *s, double *d_acc, bool want_elbo, bool entropy_term);
int shard_topics(Shard *s, const float *d_mat, const float *d_scale, int32_t *out);
int shard_get_stats(Shard *s, const double *d_sweeps, tmvb_stats *out);
// a second per-document list (CTPF reader lists, modelutils.jl:443-472) re-laid-out in the shard's internal document order
int shard_pack_aux(Shard *s, const int64_t *cumsum, const int64_t *ids, const int64_t *vals, int64_t id_limit, long long **d_off, int **d_ids,
                   float **d_vals, int64_t *nnz_out, std::vector<int> *len_internal);

}  // namespace tmvb
