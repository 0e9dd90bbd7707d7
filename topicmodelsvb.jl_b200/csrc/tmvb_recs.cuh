// tmvb_recs.cuh -- the recommendation post-processing of train!(::gpuCTPF) (gpuCTPF.jl:709-731, CTPF.jl:378-400) on the device:
//   scores = (Etheta + Eepsilon)' Eeta            the one dense contraction of the package: M x K times K x U (16 980 x 30 x 5 551)
//   urecs[u] / drecs[d]                           complete rankings of the documents not in a user's library / of the users who
//                                                 have not read a document, by descending score
// The contraction runs on the 5th-generation tensor cores: tcgen05.mma kind::tf32, both operands K-major in shared memory in
// the canonical no-swizzle "core matrix" layout, the 128 x 256 fp32 accumulator in tensor memory, read back with tcgen05.ld.
// fp32 fidelity (the rankings must not depend on tf32 rounding) comes from the three-product split x = hi + lo with
// hi = x truncated to tf32: D = Xlo Yhi' + Xhi Ylo' + Xhi Yhi' (|error| ~ 2^-21 relative, against 2^-11 for one tf32 product).
// The rankings are CUB stable segmented sorts over the score matrix in both orientations (tmvb_sort.cu).
#pragma once

#include "tmvb_common.cuh"

namespace tmvb {

constexpr int kRecBM = 128, kRecBN = 256, kRecBK = 32;   // CTA tile: 128 rows of X, 256 rows of Y, 32 of K per pass

// rank documents / users: keys ascending stable, then reversed (== reverse(sortperm(.)) of the reference, ties included);
// masked entries (key = -inf) sort first and are cut off after the reversal.  See tmvb_sort.cu.
int segmented_rank(float *d_keys, int nseg, int len, int ld, const int *d_nmask, const long long *d_out_off, int *d_out, void **ws, size_t *ws_bytes,
                   cudaStream_t stream, int n_sm);

#if defined(__CUDACC__) && defined(TMVB_RECS_KERNELS)   // the kernels live in one translation unit: tmvb_ctpf.cu

// X[perm[p]][0..K) = gimel[p] / dalet + zayin[p] / het  (Etheta + Eepsilon, gpuCTPF.jl:711-713), zero padded to KP columns;
// rows in the CALLER's document order
__global__ void recs_theta_kernel(const float *__restrict__ gimel, const float *__restrict__ zayin, const float *__restrict__ inv_dalet,
                                  const float *__restrict__ inv_het, const int *__restrict__ perm, long long M, int K, int K_ld, int KP,
                                  float *__restrict__ X)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < M * KP; q += (long long)gridDim.x * blockDim.x) {
        const long long p = q / KP;
        const int i = (int)(q - p * KP);
        X[(size_t)perm[p] * KP + i] = (i < K) ? gimel[p * K_ld + i] * inv_dalet[i] + zayin[p * K_ld + i] * inv_het[i] : 0.0f;
    }
}
// Y[u][0..K) = he[u] / vav  (Eeta, gpuCTPF.jl:709)
__global__ void recs_eta_kernel(const float *__restrict__ he, const float *__restrict__ inv_vav, long long U, int K, int K_ld, int KP, float *__restrict__ Y)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < U * KP; q += (long long)gridDim.x * blockDim.x) {
        const long long u = q / KP;
        const int i = (int)(q - u * KP);
        Y[q] = (i < K) ? he[u * K_ld + i] * inv_vav[i] : 0.0f;
    }
}
// the library / reader pairs leave both rankings: keys_d[d][u] = keys_u[u][d] = -inf for every reader u of document d
__global__ void recs_mask_kernel(const long long *__restrict__ r_off, const int *__restrict__ readers, const int *__restrict__ perm, long long M,
                                 float *__restrict__ keys_d, long long ld_d, float *__restrict__ keys_u, long long ld_u, int *__restrict__ nmask_u)
{
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const float ninf = __int_as_float(0xff800000);
    for (long long p = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); p < M; p += (long long)gridDim.x * wpb) {
        const long long o = r_off[p];
        const int Rd = (int)(r_off[p + 1] - o), d = perm[p];
        for (int n = lane; n < Rd; n += 32) {
            const int u = readers[o + n];
            if (keys_d) keys_d[(size_t)d * ld_d + u] = ninf;
            if (keys_u) {
                keys_u[(size_t)u * ld_u + d] = ninf;
                atomicAdd(nmask_u + u, 1);
            }
        }
    }
}

// ---- tcgen05 plumbing (PTX as in the CUTLASS sm100 headers; one CTA group of one CTA) ------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_alloc(unsigned *slot_smem, unsigned ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc]'   (kind::tf32, fp32 accumulate; issued by ONE thread)
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long desc_a, unsigned long long desc_b, unsigned idesc, unsigned accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the MMAs issued so far arrive on the mbarrier when they have completed (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(unsigned long long *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane (warp w of the CTA reads lanes 32 w .. 32 w + 31)
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32])
{
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
}

// Shared-memory matrix descriptor of a K-major operand tile in the canonical no-swizzle layout (cute::UMMA::SmemDescriptor,
// mma_sm100_desc.hpp): a "core matrix" is 8 rows x 16 bytes stored contiguously (128 B); the two core matrices an MMA of
// K = 8 tf32 values spans lie `lbo` bytes apart, consecutive groups of 8 rows `sbo` bytes apart.
__device__ __forceinline__ unsigned long long umma_smem_desc(unsigned saddr, unsigned lbo, unsigned sbo)
{
    return (unsigned long long)((saddr >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);   // bits 46-47: descriptor version 1 (Blackwell); no swizzle
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A and B tf32, both K-major, N = 256, M = 128
constexpr unsigned kRecIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(kRecBN >> 3) << 17) | ((unsigned)(kRecBM >> 4) << 24);

// one operand tile [rows][32 floats] from global (row-major, leading dimension ldx floats, zero beyond `valid` rows) into the
// canonical layout: 16-byte chunk c of row r at (r / 8) * 1024 + c * 128 + (r % 8) * 16, split into hi (tf32-truncated) and lo
template <int ROWS>
__device__ __forceinline__ void stage_operand(const float *__restrict__ g, long long ldx, int row0, int valid, int k0, unsigned char *hi,
                                              unsigned char *lo, int tid)
{
    for (int q = tid; q < ROWS * 8; q += 128) {
        const int r = q >> 3, c = q & 7;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < valid) v = __ldg(reinterpret_cast<const float4 *>(g + (size_t)(row0 + r) * ldx + k0 + 4 * c));
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
        l.x = v.x - h.x;
        l.y = v.y - h.y;
        l.z = v.z - h.z;
        l.w = v.w - h.w;
        const int off = (r >> 3) * 1024 + c * 128 + (r & 7) * 16;
        *reinterpret_cast<float4 *>(hi + off) = h;
        *reinterpret_cast<float4 *>(lo + off) = l;
    }
}

// C[p][q] = sum_k X[p][k] Y[q][k]   (X: [P][KP], Y: [Q][KP], KP a multiple of 32, zero padded; C row-major, leading dimension ldc)
// grid (ceil(P / 128), NY): CTA (bx, by) owns rows 128 bx .. of X and the column tiles by, by + NY, ...
// swap: the two small products are issued in the other order, so that the transposed launch (X and Y exchanged) accumulates
// the same three terms in the same order and C'[q][p] == C[p][q] bit for bit (both rankings then see the same scores)
__global__ void __launch_bounds__(128) recs_scores_umma_kernel(const float *__restrict__ X, const float *__restrict__ Y, float *__restrict__ C, int P, int Q,
                                                               int KP, long long ldc, int swap)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *xhi = smem, *xlo = xhi + kRecBM * 128, *yhi = xlo + kRecBM * 128, *ylo = yhi + kRecBN * 128;
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row0 = blockIdx.x * kRecBM;
    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) tmem_alloc(&tmem_slot, kRecBN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = tmem_slot;
    unsigned phase = 0;
    const int nkb = KP / kRecBK;

    for (int col0 = blockIdx.y * kRecBN; col0 < Q; col0 += gridDim.y * kRecBN) {
        for (int kb = 0; kb < nkb; kb++) {
            // the tensor core has finished reading the previous tiles (mbarrier wait below) before they are overwritten
            stage_operand<kRecBM>(X, KP, row0, P, kb * kRecBK, xhi, xlo, tid);
            stage_operand<kRecBN>(Y, KP, col0, Q, kb * kRecBK, yhi, ylo, tid);
            fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const unsigned axh = smem_u32(xhi), axl = smem_u32(xlo), ayh = smem_u32(yhi), ayl = smem_u32(ylo);
#pragma unroll
                for (int prod = 0; prod < 3; prod++) {   // small terms first: Xlo Yhi', Xhi Ylo', Xhi Yhi' (swap: the first two exchanged)
                    const int lo_x = swap ? 1 : 0, lo_y = swap ? 0 : 1;
                    const unsigned ax = prod == lo_x ? axl : axh, ay = prod == lo_y ? ayl : ayh;
#pragma unroll
                    for (int k = 0; k < kRecBK / 8; k++)   // an MMA spans K = 8 tf32 values = two 16-byte chunks = 256 bytes of the layout
                        umma_tf32(tmem, umma_smem_desc(ax + k * 256, 128, 1024), umma_smem_desc(ay + k * 256, 128, 1024), kRecIdesc,
                                  (kb | prod | k) != 0 ? 1u : 0u);
                }
                umma_commit(&bar);
            }
            mbar_wait(&bar, phase);
            phase ^= 1u;
            tc_fence_after();
        }
        // epilogue: thread t holds row row0 + t of the tile (TMEM lane t); 8 x 32 columns
        const int row = row0 + tid;
#pragma unroll 1
        for (int cb = 0; cb < kRecBN / 32; cb++) {
            float v[32];
            tmem_ld32(tmem + ((unsigned)(warp * 32) << 16) + (unsigned)(cb * 32), v);
            if (row < P) {
                float *dst = C + (size_t)row * ldc + col0 + cb * 32;
                if (col0 + cb * 32 + 32 <= Q && (ldc & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++)
                        if (col0 + cb * 32 + j < Q) dst[j] = v[j];
                }
            }
        }
        // the accumulator has been read by every warp before the next tile's first MMA overwrites it
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 0) tmem_dealloc(tmem, kRecBN);
}

// the same contraction on the CUDA cores in fp32 -- the checker of the tensor-core kernel (tests) and nothing else
__global__ void recs_scores_ref_kernel(const float *__restrict__ X, const float *__restrict__ Y, float *__restrict__ C, int P, int Q, int KP, long long ldc)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < (long long)P * Q; q += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(q / Q), u = (int)(q - (long long)p * Q);
        float a = 0.0f;
        for (int k = 0; k < KP; k++) a = fmaf(X[(size_t)p * KP + k], Y[(size_t)u * KP + k], a);
        C[(size_t)p * ldc + u] = a;
    }
}

#endif  // __CUDACC__ && TMVB_RECS_KERNELS

}  // namespace tmvb
