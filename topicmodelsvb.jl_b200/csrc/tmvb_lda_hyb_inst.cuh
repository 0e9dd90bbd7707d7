// tmvb_lda_hyb_inst.cuh -- body of tmvb_lda_hyb_<K_ld>.cu: define TMVB_HYB_LPT, TMVB_HYB_CPL, TMVB_HYB_KLD, TMVB_HYB_NAME, include.
#include "tmvb_lda_hyb.cuh"

namespace tmvb {

#define TMVB_HYB(W, N, T, E) (LdaEstepFn) lda_estep_hyb_kernel<TMVB_HYB_LPT, TMVB_HYB_CPL, TMVB_HYB_KLD, W, N, T, E>
#if TMVB_HYB_KLD <= 64
#define TMVB_HYB_ROW(E)                                                                                                        \
    {TMVB_HYB(1, 2, false, E), TMVB_HYB(1, 3, false, E), TMVB_HYB(1, 4, false, E), TMVB_HYB(2, 2, false, E), TMVB_HYB(2, 3, false, E), \
     TMVB_HYB(2, 4, false, E), TMVB_HYB(4, 3, false, E), TMVB_HYB(4, 4, false, E), TMVB_HYB(1, 2, true, E),  TMVB_HYB(2, 2, true, E),   \
     TMVB_HYB(2, 3, true, E),  TMVB_HYB(4, 3, true, E),  TMVB_HYB(4, 4, true, E),  TMVB_HYB(1, 5, false, E), TMVB_HYB(1, 6, false, E), \
     TMVB_HYB(2, 5, false, E), TMVB_HYB(2, 6, false, E), TMVB_HYB(4, 5, false, E), TMVB_HYB(4, 6, false, E), TMVB_HYB(4, 6, true, E)}
#else
#define TMVB_HYB_ROW(E)                                                                                                                 \
    {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, TMVB_HYB(4, 3, false, E), TMVB_HYB(4, 4, false, E), nullptr, nullptr, nullptr, \
     TMVB_HYB(4, 3, true, E), TMVB_HYB(4, 4, true, E), nullptr, nullptr, nullptr, nullptr, TMVB_HYB(4, 5, false, E),                     \
     TMVB_HYB(4, 6, false, E), TMVB_HYB(4, 6, true, E)}
#endif

extern const LdaHybLayout TMVB_HYB_NAME = {TMVB_HYB_LPT, TMVB_HYB_CPL, TMVB_HYB_KLD, {TMVB_HYB_ROW(false), TMVB_HYB_ROW(true)}};

}  // namespace tmvb
