// lda_estep_hyb_kernel for K_ld = 56 (lane layout 2 lanes per token x 7 chunks per lane)
#define TMVB_HYB_LPT 2
#define TMVB_HYB_CPL 7
#define TMVB_HYB_KLD 56
#define TMVB_HYB_NAME kLdaHyb_56
#include "tmvb_lda_hyb_inst.cuh"
