// tmvb_lda_hyb.cuh -- dispatch tables of lda_estep_hyb_kernel (tmvb_lda_estep.cuh).  The kernel takes K_ld as a template
// parameter, so it is instantiated per supported K_ld in its own translation unit (tmvb_lda_hyb_<K_ld>.cu, all generated
// from tmvb_lda_hyb_inst.cuh) -- the units compile in parallel and each stays within seconds.
#pragma once

#include "tmvb_lda_estep.cuh"

namespace tmvb {

typedef void (*LdaEstepFn)(const LdaDev, int, int, int, int, int *);

// variants: (warps per document, register rounds per warp, with / without a shared-memory tile).  K_ld <= 64: all;
// K_ld > 64: four warps only (the K phase keeps one topic pair per thread: 64 W >= K_ld).
constexpr int kNumHybVariants = 20;
static const int kHybVariant[kNumHybVariants][3] = {{1, 2, 0}, {1, 3, 0}, {1, 4, 0}, {2, 2, 0}, {2, 3, 0}, {2, 4, 0}, {4, 3, 0}, {4, 4, 0},
                                                    {1, 2, 1}, {2, 2, 1}, {2, 3, 1}, {4, 3, 1}, {4, 4, 1}, {1, 5, 0}, {1, 6, 0}, {2, 5, 0}, {2, 6, 0},
                                                    {4, 5, 0}, {4, 6, 0}, {4, 6, 1}};
struct LdaHybLayout {
    int lpt, cpl, K_ld;
    LdaEstepFn fn[2][kNumHybVariants];  // [want_elbo][variant]
};

// K_ld with a hybrid instantiation (the lane layout is the one pick_layout() selects for that K_ld)
extern const LdaHybLayout kLdaHyb_56, kLdaHyb_200;
#define TMVB_LDA_HYB_TABLES &kLdaHyb_56, &kLdaHyb_200

}  // namespace tmvb
