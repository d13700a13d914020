// MSM kernels for MNT4_G1
#define PCD_CURVE CurveMnt4G1
#define PCD_OPS_NAME MSM_OPS_MNT4_G1
#include "msm_inst.cuh"
