// MSM kernels for MNT6_G2
#define PCD_CURVE CurveMnt6G2
#define PCD_OPS_NAME MSM_OPS_MNT6_G2
#include "msm_inst.cuh"
