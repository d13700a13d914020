// MSM kernels for MNT4_G2
#define PCD_CURVE CurveMnt4G2
#define PCD_OPS_NAME MSM_OPS_MNT4_G2
#include "msm_inst.cuh"
