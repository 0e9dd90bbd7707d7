// lda_estep_hyb_kernel for K_ld = 200 (lane layout 8 lanes per token x 7 chunks per lane)
#define TMVB_HYB_LPT 8
#define TMVB_HYB_CPL 7
#define TMVB_HYB_KLD 200
#define TMVB_HYB_NAME kLdaHyb_200
#include "tmvb_lda_hyb_inst.cuh"
