// MSM kernels for MNT6_G1
#define PCD_CURVE CurveMnt6G1
#define PCD_OPS_NAME MSM_OPS_MNT6_G1
#include "msm_inst.cuh"
