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
  // the proof / MSM tails, lane-cooperative (wec.cuh), on the context's CURRENT lane (ctx->cur()):
  //   to_affine_at  d_out_affine <- xyzz point idx of d_in
  //   sum_points    sum of the n points d_in[first + i * stride] -> xyzz at d_out_xyzz[out_idx] and / or affine (null: skip)
  //   multi_mul     d_out[out_idx] <- sum_{j < npairs} [k_j] d_pts[idx_j]; k_j ten plain words at d_k + 10 j (G1 curves
  //                 only; npairs <= 2)
  int (*to_affine_at)(pcdgpu_ctx*, const void* d_in, size_t idx, void* d_out_affine);
  int (*sum_points)(pcdgpu_ctx*, const void* d_in, size_t first, size_t stride, int n, void* d_out_xyzz, size_t out_idx,
                    void* d_out_affine);
  int (*multi_mul)(pcdgpu_ctx*, const void* d_pts, size_t idx0, size_t idx1, const void* d_k, int npairs, void* d_out,
                   size_t out_idx);
  // d_table: 75 * 15 affine points (built from d_base by fixed_table); out[i] = scalars[i] * base
  int (*fixed_table)(pcdgpu_ctx*, const void* d_base, void* d_table);
  int (*fixed_mul)(pcdgpu_ctx*, const void* d_table, const void* d_scalars, size_t n, void* d_out);
  int (*precompute)(pcdgpu_ctx*, const void* d_bases, size_t n, int c, int nwin, void* d_pre);
};

extern const MsmOps MSM_OPS_MNT4_G1, MSM_OPS_MNT4_G2, MSM_OPS_MNT6_G1, MSM_OPS_MNT6_G2;
const MsmOps* msm_ops(int curve);
int msm_auto_window_c(size_t n, int shared);
int msm_num_windows_c(int c);
