// msm_ops.cuh -- per-curve entry table; each msm_c*.cu instantiates the templates of msm.cuh for
// one curve (separate translation units so they compile in parallel).
#pragma once
#include "common.cuh"

struct MsmPlanC {
  int c, nwin, shared;
  size_t stride, offset;  // precomputed tables: row pitch in points, first point used
};

struct MsmOps {
  size_t affine_bytes, xyzz_bytes;
  int scalar_field;
  // sum_i scalars[i] * bases[i] -> one xyzz point at d_out.  The first n scalars come from d_scalars
  // (Montgomery form if mont), n_extra more from d_extra (plain integers); bases holds n + n_extra points.
  int (*run)(pcdgpu_ctx*, const void* d_bases, const void* d_scalars, int mont, size_t n, const void* d_extra,
             size_t n_extra, MsmPlanC plan, void* d_out);
  int (*to_affine)(pcdgpu_ctx*, const void* d_in, size_t n, void* d_out);
  int (*xyzz_sum)(pcdgpu_ctx*, const void* d_in, size_t n, void* d_out_affine);
  // d_table: 75 * 15 affine points (built from d_base by fixed_table); out[i] = scalars[i] * base
  int (*fixed_table)(pcdgpu_ctx*, const void* d_base, void* d_table);
  int (*fixed_mul)(pcdgpu_ctx*, const void* d_table, const void* d_scalars, size_t n, void* d_out);
  int (*precompute)(pcdgpu_ctx*, const void* d_bases, size_t n, int c, int nwin, void* d_pre);
};

extern const MsmOps MSM_OPS_MNT4_G1, MSM_OPS_MNT4_G2, MSM_OPS_MNT6_G1, MSM_OPS_MNT6_G2;
const MsmOps* msm_ops(int curve);
int msm_auto_window_c(size_t n, int shared);
int msm_num_windows_c(int c);
