/*
 * tmvb.h -- C ABI of libtmvb.so, the sm_100a CAVI engine that stands in for the OpenCL layer of
 * ericproffitt/TopicModelsVB.jl (src/gpuLDA.jl, src/gpuCTM.jl, src/gpuCTPF.jl and the device
 * plumbing update_buffer!/update_host!/@buffer/@host in src/modelutils.jl:369-570, src/macros.jl:60-104).
 *
 * The reference has no FFI: its seam is the set of Julia methods that touch `cl.*`.  Every entry
 * point below names the reference call site(s) it replaces (file:line under the reference root).
 * INTEGRATION.md shows the Julia `ccall` binding; topicmodelsvb.jl_b200/_lib.py is the ctypes one.
 *
 * Conventions
 *  - plain pointers and sizes only; the caller owns every host array for the duration of the call.
 *  - dense host matrices are Julia column-major Float32: beta[K*j + i] = topic i of term j (K x V),
 *    Elogtheta[K*d + i] (K x M), ... exactly what update_buffer! uploads (modelutils.jl:390-392).
 *  - corpus indices are Int64, 0-based, flattened (modelutils.jl:371-380).
 *  - every function returns 0 on success, <0 for an invalid argument, >0 for a CUDA error code;
 *    tmvb_last_error() returns a thread-local message.  No exception crosses the ABI.
 *  - calls on one handle must be serialised by the caller.  Work is enqueued on the handle's
 *    stream; functions that return host data synchronise that stream first.
 *  - one handle owns one shard of documents on one device.  Multi-GPU runs use one handle per
 *    GPU and sum the buffers exposed by tmvb_*_reduce_buffers() between estep and mstep.
 *  - there is NO CPU fallback: without a CUDA device every create() fails.
 */
#ifndef TMVB_H
#define TMVB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TMVB_VERSION 100

typedef struct tmvb_lda_s *tmvb_lda_t;
typedef struct tmvb_ctm_s *tmvb_ctm_t;
typedef struct tmvb_ctpf_s *tmvb_ctpf_t;

/* Timings (ms, CUDA events on the handle's stream) and counters of the most recent calls. */
typedef struct tmvb_stats {
    double estep_ms;        /* last tmvb_*_estep                                          */
    double mstep_ms;        /* last tmvb_*_mstep                                          */
    int64_t sweeps;         /* total inner sweeps of the last estep (sum over documents)  */
    int64_t kernel_launches;/* kernels launched by this handle since creation             */
    int64_t h2d_bytes;      /* host->device bytes copied by this handle since creation    */
    int64_t d2h_bytes;      /* device->host bytes copied by this handle since creation    */
} tmvb_stats;

int tmvb_version(void);
const char *tmvb_last_error(void);
/* number of visible CUDA devices (0 and an error code when there is none) */
int tmvb_device_count(int *count);
/* page-locked host memory for the caller's arrays (update_buffer!/update_host! then run at PCIe speed) */
int tmvb_alloc_pinned(void **ptr, int64_t bytes);
int tmvb_free_pinned(void *ptr);

/* ------------------------------------------------------------------ LDA ------------------ */

/* gpuLDA(corp, K) device side (gpuLDA.jl:64-80: cl.create_compute_context + 7 cl.Program builds).
 * M, V may be 0 (the `gpuLDA(Corpus(), 1)` construction of macros.jl:114).  `device` < 0 picks the
 * current device; `stream` may be NULL (a private stream is created) or a caller-owned cudaStream_t. */
int tmvb_lda_create(tmvb_lda_t *h, int64_t K, int64_t M, int64_t V, int device, void *stream);
int tmvb_lda_destroy(tmvb_lda_t h); /* idempotent on NULL */

/* update_buffer!(model::gpuLDA), corpus half (modelutils.jl:371-388).  N_cumsum[M+1], terms[sumN]
 * (0-based), counts[sumN].  The inverted index (J_cumsum / terms_sortperm) is not needed.
 * Pointers may be pinned or pageable host memory. */
int tmvb_lda_set_corpus(tmvb_lda_t h, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts);
/* The same with Int32 token arrays: for hosts that keep a packed copy of the flattened corpus between train! calls (the
 * reference rebuilds `terms`/`counts` with vcat on every update_buffer!, modelutils.jl:371-373); halves the corpus upload. */
int tmvb_lda_set_corpus32(tmvb_lda_t h, const int64_t *N_cumsum, const int32_t *terms, const int32_t *counts);

/* update_buffer!(model::gpuLDA), parameter half (modelutils.jl:390-392) and `@buffer model.alpha`
 * (macros.jl:64).  Any pointer may be NULL (left unchanged).  gamma is optional (the reference
 * allocates it uninitialised, modelutils.jl:395).  Also resets beta_old/Elogtheta_old := beta/Elogtheta
 * as the constructors do (LDA.jl:36,39). */
int tmvb_lda_upload(tmvb_lda_t h, const float *alpha, const float *beta, const float *Elogtheta, const float *gamma);
int tmvb_lda_set_alpha(tmvb_lda_t h, const float *alpha);

/* The whole inner loop `for v in 1:viter` (gpuLDA.jl:356-364: update_phi!/update_gamma!/
 * update_Elogtheta!) with the CPU model's per-document stopping rule (LDA.jl:170-178), followed by
 * the scatter half of update_beta! (gpuLDA.jl:156-177 / LDA.jl:129-132) into the K x V statistics.
 * want_elbo != 0 also accumulates the per-document ELBO terms.  Asynchronous. */
int tmvb_lda_estep(tmvb_lda_t h, int viter, float vtol, int want_elbo);

/* The inner loop of predict(corp, train_model::gpuLDA) (modelutils.jl:846-855): update_phi!/update_gamma!/update_Elogtheta! per
 * document with the uploaded alpha / beta frozen; unlike tmvb_lda_estep no statistics are scattered and no ELBO partials are kept. */
int tmvb_lda_predict(tmvb_lda_t h, int viter, float vtol);

/* Device buffers that a multi-GPU driver must sum over ranks between estep and mstep:
 * stats = float[n_stats] (K_ld x V padded statistics), small = double[n_small]
 * (sum_d Elogtheta_d, ELBO partials, sweep counter). */
int tmvb_lda_reduce_buffers(tmvb_lda_t h, void **stats, int64_t *n_stats, void **small, int64_t *n_small);

/* normalize_beta (gpuLDA.jl:179-204 / LDA.jl:121-125): beta_old <- beta; beta = stats ./ rowsum;
 * stats <- 0.  M_total is the corpus-wide document count (== M unless sharded).  Asynchronous. */
int tmvb_lda_mstep(tmvb_lda_t h);

/* Multi-GPU exchange over peer memory (no reference counterpart: the reference is single-device).
 * One process per GPU of one NVSwitch box.  Each rank exports CUDA IPC handles of its statistics / table / small buffers
 * (comm_export -> blob), the host driver all-gathers the blobs (any transport), every rank maps its peers (comm_connect), and
 * from then on tmvb_lda_exchange_mstep() replaces { sum reduce_buffers over ranks; tmvb_lda_mstep }: ONE kernel that
 * reduce-scatters the statistics with loads from the peers' memory, normalises its slice of beta and all-gathers it with
 * stores into the peers' memory (gpuLDA.jl:179-204 semantics on the summed statistics).  Every rank must call it once per
 * outer iteration; a rank that does not arrive within TMVB_COMM_TIMEOUT_MS (environment, default 4000) makes the others give
 * up instead of hanging: the condition is raised (return code 901/902) by the next tmvb_lda_elbo(mode 0) / tmvb_lda_iterate
 * read-back, by tmvb_lda_download and by tmvb_lda_comm_status, and cleared by the report.  world <= 8. */
#define TMVB_COMM_BLOB_BYTES 512
int tmvb_lda_comm_export(tmvb_lda_t h, void *blob, int64_t blob_bytes);
int tmvb_lda_comm_connect(tmvb_lda_t h, int rank, int world, const void *blobs /* [world][blob_bytes] */, int64_t blob_bytes);
int tmvb_lda_exchange_mstep(tmvb_lda_t h);
int tmvb_lda_comm_status(tmvb_lda_t h, int *status);

/* `@host model.Elogtheta_sum_buffer` (macros.jl:79, gpuLDA.jl:133); fp64. */
int tmvb_lda_get_elogtheta_sum(tmvb_lda_t h, double *out);

/* update_alpha! (LDA.jl:97-118 / gpuLDA.jl:132-154): the interior-point Newton iteration in fp64 in a one-warp device kernel
 * on the reduced Elogtheta_sum, fused with the assembly of the ELBO from the partials of the last estep/mstep.  M_total as
 * above.  Asynchronous when alpha_out is NULL; otherwise alpha_out[K] receives the new alpha (synchronises). */
int tmvb_lda_update_alpha(tmvb_lda_t h, int64_t M_total, int niter, double ntol, float *alpha_out);

/* One outer iteration of train! -- the body of `for k in 1:iter` (gpuLDA.jl:355-371): the folded inner loop + scatter
 * (tmvb_lda_estep), update_beta! (tmvb_lda_mstep, or the fused peer exchange when tmvb_lda_comm_connect has been called),
 * update_alpha! (tmvb_lda_update_alpha) and, with want_elbo, the ELBO that check_elbo! reads (tmvb_lda_elbo mode 0, written to
 * *elbo) -- enqueued as ONE CUDA graph launch.  With several GPUs the exchange kernel also runs update_alpha! beside the
 * reduce-scatter, so the iteration is E-step launches + one kernel.  The host synchronises only when want_elbo is set
 * (checkelbo = Inf iterations are fully asynchronous).  Same result as the four separate calls.  Needs one GPU or connected
 * peers (a driver that sums tmvb_lda_reduce_buffers itself keeps the separate calls). */
int tmvb_lda_iterate(tmvb_lda_t h, int viter, float vtol, int want_elbo, int64_t M_total, int niter, double ntol, double *elbo);

/* update_elbo! (gpuLDA.jl:121-128 via check_elbo!, modelutils.jl:574-585) without the K x sumN phi
 * transfer.  mode 0: the value assembled on the device by estep(want_elbo=1) -> mstep -> update_alpha (one 8-byte
 * read-back).  mode 1 (fp32 elements) / mode 2 (fp64): full recomputation from device state with
 * the CPU model's lagged-phi semantics (LDA.jl:83-93); used for the initial ELBO.  For sharded runs
 * mode 1 returns this shard's document terms plus the global terms on every rank, so the caller
 * sums `*elbo_docs` over ranks and adds `*elbo_global` once. */
int tmvb_lda_elbo(tmvb_lda_t h, int mode, int64_t M_total, double *elbo_docs, double *elbo_global);

/* update_host!(model::gpuLDA) (modelutils.jl:501-514): any pointer may be NULL. */
int tmvb_lda_download(tmvb_lda_t h, float *alpha, float *beta, float *Elogtheta, float *gamma);
/* The Elogtheta / gamma half of update_host! (modelutils.jl:507-512) folded into the last E-step of a train! call: with the
 * mirror armed, the NEXT tmvb_lda_estep / tmvb_lda_iterate also writes every document's final rows straight into these
 * page-locked, device-mapped K x M arrays (tmvb_alloc_pinned, cudaHostAlloc or cudaHostRegister memory) while the other
 * documents are still being swept, and a following tmvb_lda_download with the same two pointers only waits for the stream
 * instead of copying them.  One-shot; (NULL, NULL) disarms; any later E-step or upload makes download copy again. */
int tmvb_lda_arm_host_mirror(tmvb_lda_t h, float *Elogtheta, float *gamma);
/* the lagged copies the CPU struct keeps (LDA.jl:17,20): beta_old, Elogtheta_old */
int tmvb_lda_download_old(tmvb_lda_t h, float *beta_old, float *Elogtheta_old);
/* phi of every document, K x sumN column-major, original token order (modelutils.jl:515-516) */
int tmvb_lda_materialize_phi(tmvb_lda_t h, float *phi);

/* model.topics = [reverse(sortperm(vec(beta[i,:]))) for i in 1:K] (gpuLDA.jl:374), ranked on the device.
 * topics[i*V + r] = 1-based term id of rank r in topic i (int32, K*V). */
int tmvb_lda_topics(tmvb_lda_t h, int32_t *topics);

int tmvb_lda_sync(tmvb_lda_t h);
int tmvb_lda_get_stats(tmvb_lda_t h, tmvb_stats *out);
/* padded leading dimension of the device K-vectors (rows of beta/stats are K_ld floats) */
int tmvb_lda_kld(tmvb_lda_t h, int64_t *K_ld);


/* ------------------------------------------------------------------ corpus ingest (host only) ---- */

/* readcorp(docfile=..., delim, counts, readers, ratings) (Corpus.jl:277-296) fused with the flattening of update_buffer!
 * (modelutils.jl:371-380,443-472): text -> packed CSR with 0-based Int32 keys, ready for tmvb_*_set_corpus32.  Multi-threaded
 * (nthreads <= 0: all host threads).  Arrays are malloc'ed by the library and released by tmvb_free_csr.  A document that does
 * not parse or fails check_doc (Corpus.jl:41-49) returns rc = -6 with the reference's message
 * "document d beginning on line l failed to load." (Corpus.jl:293). */
typedef struct {
    int64_t M, nnz, nr;          /* documents, sum of N_d, sum of R_d */
    int64_t max_term, max_reader; /* largest 1-based key seen (0 if none) */
    int64_t *N_cumsum;           /* [M+1] */
    int32_t *terms, *counts;     /* [nnz] terms 0-based; counts default 1 */
    int64_t *R_cumsum;           /* [M+1] */
    int32_t *readers, *ratings;  /* [nr] readers 0-based; ratings default 1 */
} tmvb_csr;
int tmvb_read_docfile(const char *path, char delim, int counts, int readers, int ratings, int nthreads, tmvb_csr *out);
int tmvb_free_csr(tmvb_csr *c);


/* ------------------------------------------------------------------ CTM ------------------ */

/* gpuCTM(corp, K) device side (gpuCTM.jl:74-92: context + 9 cl.Program builds).  K <= 64 in this build.
 * Starts from mu = 0, sigma = invsigma = I (gpuCTM.jl:63-65). */
int tmvb_ctm_create(tmvb_ctm_t *h, int64_t K, int64_t M, int64_t V, int device, void *stream);
int tmvb_ctm_destroy(tmvb_ctm_t h);

/* update_buffer!(model::gpuCTM), corpus half (modelutils.jl:401-417). */
int tmvb_ctm_set_corpus(tmvb_ctm_t h, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts);

/* update_buffer!(model::gpuCTM), parameter half (modelutils.jl:419-428) and `@buffer model.invsigma` (macros.jl:67):
 * mu[K], sigma[K*K] (invsigma is recomputed in fp64, as gpuCTM.jl:203-205 does on the host), beta[K*V],
 * lambda[K*M], vsq[K*M], logzeta[M].  Any pointer may be NULL. */
int tmvb_ctm_upload(tmvb_ctm_t h, const float *mu, const float *sigma, const float *beta, const float *lambda, const float *vsq,
                    const float *logzeta);

/* The inner loop `for v in 1:viter` (gpuCTM.jl:497-507: update_phi!/update_logzeta!/update_vsq!/update_lambda!) with the
 * CPU model's order and per-document stopping rule (CTM.jl:194-203), the scatter half of update_beta! (gpuCTM.jl:208-229)
 * and the second moments update_sigma!/update_mu! need (gpuCTM.jl:144-196).  Asynchronous. */
int tmvb_ctm_estep(tmvb_ctm_t h, int niter, float ntol, int viter, float vtol, int want_elbo);

/* The inner loop of predict(corp, train_model::gpuCTM) (modelutils.jl:903-913): phi / logzeta / vsq / lambda per document with
 * mu, sigma, beta frozen; no statistics are scattered. */
int tmvb_ctm_predict(tmvb_ctm_t h, int niter, float ntol, int viter, float vtol);

/* buffers a multi-GPU driver sums over ranks between estep and mstep (stats float, small double) */
int tmvb_ctm_reduce_buffers(tmvb_ctm_t h, void **stats, int64_t *n_stats, void **small, int64_t *n_small);

/* update_beta!() + update_sigma!() + update_mu!() (CTM.jl:102-118,207-208 / gpuCTM.jl:166-256): beta normalised on the
 * device; sigma, invsigma = inv(sigma), mu in fp64 on the host from the reduced moments, then pushed to the device. */
int tmvb_ctm_mstep(tmvb_ctm_t h, int64_t M_total);

/* update_elbo! (gpuCTM.jl:135-142 via check_elbo!).  mode 0: from the partials of the last estep(want_elbo=1)+mstep;
 * mode 1: full recomputation with the CPU model's lagged-phi semantics (CTM.jl:89-98). */
int tmvb_ctm_elbo(tmvb_ctm_t h, int mode, int64_t M_total, double *elbo_docs, double *elbo_global);

/* update_host!(model::gpuCTM) (modelutils.jl:518-537); any pointer may be NULL. */
int tmvb_ctm_download(tmvb_ctm_t h, float *mu, float *sigma, float *invsigma, float *beta, float *lambda, float *vsq, float *logzeta);
int tmvb_ctm_download_old(tmvb_ctm_t h, float *beta_old, float *lambda_old);
int tmvb_ctm_materialize_phi(tmvb_ctm_t h, float *phi);
int tmvb_ctm_topics(tmvb_ctm_t h, int32_t *topics); /* gpuCTM.jl:517 */
int tmvb_ctm_get_stats(tmvb_ctm_t h, tmvb_stats *out);


/* ------------------------------------------------------------------ CTPF ----------------- */

/* gpuCTPF(corp, K) device side (gpuCTPF.jl:121-147: context + 12 cl.Program builds).  Starts from a..h = 0.1 and all
 * rates = 1 (gpuCTPF.jl:107-119). */
int tmvb_ctpf_create(tmvb_ctpf_t *h, int64_t K, int64_t M, int64_t V, int64_t U, int device, void *stream);
int tmvb_ctpf_destroy(tmvb_ctpf_t h);

/* update_buffer!(model::gpuCTPF), corpus half (modelutils.jl:439-472): terms/counts and readers/ratings CSR (0-based Int64).
 * The inverted indices (J_cumsum, Y_cumsum, *_sortperm) are not needed. */
int tmvb_ctpf_set_corpus(tmvb_ctpf_t h, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts, const int64_t *R_cumsum,
                         const int64_t *readers, const int64_t *ratings);

/* update_buffer!(model::gpuCTPF), parameter half (modelutils.jl:474-493): hyp[8] = {a,b,c,d,e,f,g,h} (the struct fields
 * gpuCTPF.jl:26-33), alef[K*V], he[K*U], bet[K], vav[K], gimel[K*M], zayin[K*M], dalet[K], het[K].  Any pointer may be NULL. */
int tmvb_ctpf_upload(tmvb_ctpf_t h, const double *hyp, const float *alef, const float *he, const float *bet, const float *vav,
                     const float *gimel, const float *zayin, const float *dalet, const float *het);

/* The inner loop (gpuCTPF.jl:687-697: update_xi!/update_phi!/update_zayin!/update_gimel!) with the CPU model's formulas and
 * per-document stopping rule (CTPF.jl:353-362), then the scatter halves of update_he!/update_alef! (CTPF.jl:259-277). */
int tmvb_ctpf_estep(tmvb_ctpf_t h, int viter, float vtol, int want_elbo);

int tmvb_ctpf_reduce_buffers(tmvb_ctpf_t h, void **stats_alef, int64_t *n_alef, void **stats_he, int64_t *n_he, void **small, int64_t *n_small);

/* update_he!(), update_alef!(), update_dalet!(), update_het!(), update_bet!(), update_vav!() in the reference's order
 * (CTPF.jl:366-371 / gpuCTPF.jl:699-704): shapes on the device, the four rate vectors in fp64 on the host. */
int tmvb_ctpf_mstep(tmvb_ctpf_t h, int64_t M_total);

/* update_elbo! (gpuCTPF.jl:280-286).  mode 0: from the partials of the last estep(want_elbo=1)+mstep; mode 1: recomputed
 * per document with the CPU model's lagged phi / xi (CTPF.jl:232-247). */
int tmvb_ctpf_elbo(tmvb_ctpf_t h, int mode, int64_t M_total, double *elbo_docs, double *elbo_global);

/* update_host!(model::gpuCTPF) (modelutils.jl:540-570) and the *_old copies of the CPU struct (CTPF.jl:27-45) */
int tmvb_ctpf_download(tmvb_ctpf_t h, float *alef, float *he, float *bet, float *vav, float *gimel, float *zayin, float *dalet, float *het);
int tmvb_ctpf_download_old(tmvb_ctpf_t h, float *alef_old, float *he_old, float *bet_old, float *vav_old, float *gimel_old, float *zayin_old,
                           float *dalet_old, float *het_old);
int tmvb_ctpf_topics(tmvb_ctpf_t h, int32_t *topics); /* gpuCTPF.jl:706-707 */

/* The recommendation step that ends train!(::gpuCTPF) (gpuCTPF.jl:709-731; CTPF.jl:378-400), from the device-resident state:
 *   scores [M x U] Float32, column-major as Julia's model.scores: scores[d, u] = sum_i Eeta[i, u] (Etheta[i, d] + Eepsilon[i, d])
 *   urecs  concatenated: urecs[uoff[u] .. uoff[u+1]) = 1-based documents NOT in user u's library, by descending score
 *          (findall(.)[reverse(sortperm(.))] of gpuCTPF.jl:716-722, ties ordered like the reference's stable sort reversed)
 *   drecs  concatenated: drecs[doff[d] .. doff[d+1]) = 1-based users who have NOT read document d (gpuCTPF.jl:724-729)
 *   uoff [U + 1], doff [M + 1]; both rankings hold M * U - sum(R) entries in total.
 * Every output may be NULL (a ranking and its offsets together).  The contraction runs on the tensor cores (tcgen05 kind::tf32
 * with a hi/lo split of both operands: fp32-level accuracy); mode bit 0 selects the plain fp32 CUDA-core contraction instead
 * (the checker the tests compare the tensor-core kernel with).  Single-GPU handles only (M = all documents). */
int tmvb_ctpf_recs(tmvb_ctpf_t h, float *scores, int32_t *urecs, int64_t *uoff, int32_t *drecs, int64_t *doff, int mode);
int tmvb_ctpf_get_stats(tmvb_ctpf_t h, tmvb_stats *out);

/* ------------------------------------------------------------------ fLDA ----------------- */

/* Filtered LDA (src/fLDA.jl) on the device.  The reference has no gpufLDA: `@gpu` leaves fLDA / fCTM untouched (macros.jl:274-278)
 * and its todo list asks for them; these entry points are what a `gpufLDA` mirror of gpuLDA would bind.
 * Layouts as for LDA; tau / tau_old are flat over the CSR tokens in the caller's order (tau[d][n] at N_cumsum[d] + n). */
typedef struct tmvb_flda_s *tmvb_flda_t;
int tmvb_flda_create(tmvb_flda_t *h, int64_t K, int64_t M, int64_t V, int device, void *stream);
int tmvb_flda_destroy(tmvb_flda_t h);
int tmvb_flda_set_corpus(tmvb_flda_t h, const int64_t *N_cumsum, const int64_t *terms, const int64_t *counts);
/* the struct fields of fLDA.jl:14-26: eta, alpha[K], kappa[V], beta[K*V], Elogtheta[K*M], gamma[K*M], tau[sum N]; the *_old copies are
 * set to the uploaded values (fLDA.jl:42,45,48,51).  Any pointer may be NULL. */
int tmvb_flda_upload(tmvb_flda_t h, const double *eta, const float *alpha, const float *kappa, const float *beta, const float *Elogtheta,
                     const float *gamma, const float *tau);
/* the per-document inner loop of train!(::fLDA) (fLDA.jl:223-233: update_phi!/update_tau!/update_gamma!/update_Elogtheta! with the
 * per-document stopping rule), then update_beta!(model, d) and update_kappa!(model, d) (fLDA.jl:156-171) */
int tmvb_flda_estep(tmvb_flda_t h, int viter, float vtol);
/* the same without the scatter: predict(corp, train_model::fLDA) (modelutils.jl:857-884) */
int tmvb_flda_predict(tmvb_flda_t h, int viter, float vtol);
int tmvb_flda_reduce_buffers(tmvb_flda_t h, void **stats, int64_t *n_stats, void **kstats, int64_t *n_kstats, void **small, int64_t *n_small);
/* update_beta!(model), update_kappa!(model), update_alpha!(model, niter, ntol), update_eta!(model) (fLDA.jl:236-239);
 * M_total / C_total = number of documents / sum of all counts over every rank */
int tmvb_flda_mstep(tmvb_flda_t h, int64_t M_total, double C_total, int niter, double ntol);
/* update_elbo! (fLDA.jl:105-117) summed over the shard's documents */
int tmvb_flda_elbo(tmvb_flda_t h, double *elbo_docs);
int tmvb_flda_download(tmvb_flda_t h, double *eta, float *alpha, float *kappa, float *beta, float *Elogtheta, float *gamma, float *tau);
int tmvb_flda_download_old(tmvb_flda_t h, float *kappa_old, float *beta_old, float *Elogtheta_old, float *tau_old);
int tmvb_flda_topics(tmvb_flda_t h, int32_t *topics); /* fLDA.jl:246 */
int tmvb_flda_get_stats(tmvb_flda_t h, tmvb_stats *out);

/* ------------------------------------------------------------------ fCTM ----------------- */

/* Filtered CTM (src/fCTM.jl) on the device: a gpuCTM handle that additionally holds eta, kappa and the per-token tau of fCTM.jl:10-28.
 * After tmvb_fctm_create every tmvb_ctm_* entry point applies to it: tmvb_ctm_estep runs the inner loop of train!(::fCTM)
 * (fCTM.jl:258-268: update_phi!, update_tau!, update_logzeta!, update_lambda!, update_vsq!) and the scatters of fCTM.jl:162-178,
 * tmvb_ctm_mstep adds update_kappa!(model) (fCTM.jl:154-158), tmvb_ctm_elbo evaluates update_elbo! of fCTM.jl:67-130 (either mode).
 * Like fLDA the reference has no GPU version of it (macros.jl:277-278). */
int tmvb_fctm_create(tmvb_ctm_t *h, int64_t K, int64_t M, int64_t V, int device, void *stream);
int tmvb_fctm_upload(tmvb_ctm_t h, const double *eta, const float *kappa, const float *tau);
int tmvb_fctm_download(tmvb_ctm_t h, float *kappa, float *kappa_old, float *tau, float *tau_old);
int tmvb_fctm_reduce_buffers(tmvb_ctm_t h, void **kstats, int64_t *n_kstats);

/* ------------------------------------------------------------------ multi-GPU (CTM, CTPF, fLDA, fCTM) ----------------- */

/* The statistics of these models only need summing over the ranks (their M-steps take the sums as they are): ONE kernel over
 * CUDA-IPC peer memory per outer iteration (csrc/tmvb_peer.cu: rank r reduces slice r of every buffer with loads from the peers
 * and stores the sum into every rank's copy) instead of one NCCL all-reduce per buffer.  Handshake as tmvb_lda_comm_export /
 * _connect; then tmvb_*_peer_reduce(h) between estep and mstep on every rank (a filtered CTM handle includes its kappa
 * statistics).  Fallback without peer access: sum the buffers of tmvb_*_reduce_buffers with any all-reduce. */
int tmvb_ctm_comm_export(tmvb_ctm_t h, void *blob, int64_t blob_bytes);
int tmvb_ctm_comm_connect(tmvb_ctm_t h, int rank, int world, const void *blobs, int64_t blob_bytes);
int tmvb_ctm_peer_reduce(tmvb_ctm_t h);
int tmvb_ctpf_comm_export(tmvb_ctpf_t h, void *blob, int64_t blob_bytes);
int tmvb_ctpf_comm_connect(tmvb_ctpf_t h, int rank, int world, const void *blobs, int64_t blob_bytes);
int tmvb_ctpf_peer_reduce(tmvb_ctpf_t h);
int tmvb_flda_comm_export(tmvb_flda_t h, void *blob, int64_t blob_bytes);
int tmvb_flda_comm_connect(tmvb_flda_t h, int rank, int world, const void *blobs, int64_t blob_bytes);
int tmvb_flda_peer_reduce(tmvb_flda_t h);

#ifdef __cplusplus
}
#endif
#endif /* TMVB_H */
