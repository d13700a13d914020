// msm_inst.cuh -- instantiates msm.cuh for the curve named by PCD_CURVE / PCD_OPS_NAME.
#include "msm.cuh"
#include "msm_ops.cuh"

namespace {
typedef PCD_CURVE CV;

int run_(pcdgpu_ctx* ctx, const void* b, const void* s, int mont, size_t n, const void* extra, size_t n_extra,
         MsmPlanC p, void* out) {
  MsmPlan plan{p.c, p.nwin, p.shared, p.stride, p.offset};
  return msm_run<CV>(ctx, b, s, mont, n, extra, n_extra, plan, out);
}
int to_affine_(pcdgpu_ctx* ctx, const void* in, size_t n, void* out) {
  if (n == 0) return 0;
  xyzz_to_affine_kernel<CV><<<(unsigned)((n + 63) / 64), 64, 0, ctx->stream>>>(in, n, out);
  PCD_CUDA(ctx, cudaGetLastError());
  return 0;
}
int to_affine_at_(pcdgpu_ctx* ctx, const void* in, size_t idx, void* out) {
  wec_to_affine_kernel<CV><<<1, 32, 0, ctx->cur()>>>(in, idx, out);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  return 0;
}
int sum_points_(pcdgpu_ctx* ctx, const void* in, size_t first, size_t stride, int n, void* out_xyzz, size_t out_idx,
                void* out_affine) {
  wec_sum_affine_kernel<CV><<<1, 32, 0, ctx->cur()>>>(in, first, stride, n, out_xyzz, out_idx, out_affine);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  return 0;
}
int multi_mul_(pcdgpu_ctx* ctx, const void* pts, size_t idx0, size_t idx1, const void* k, int npairs, void* out,
               size_t out_idx) {
  if constexpr (Wec<CV>::G <= 16 && Wec<CV>::K == 1) {
    wec_multi_mul_kernel<CV><<<1, 32, 0, ctx->cur()>>>(pts, idx0, idx1, (const u32*)k, npairs, out, out_idx);
    PCD_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return 0;
  } else {
    ctx->set_error("multi_mul is a G1 operation");
    return PCDGPU_E_ARG;
  }
}
int fixed_table_(pcdgpu_ctx* ctx, const void* base, void* table) {
  fixed_table_kernel<CV><<<3, 32, 0, ctx->stream>>>(base, table);
  PCD_CUDA(ctx, cudaGetLastError());
  return 0;
}
int fixed_mul_(pcdgpu_ctx* ctx, const void* table, const void* scalars, size_t n, void* out) {
  if (n == 0) return 0;
  fixed_mul_kernel<CV><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(table, (const u32*)scalars, n, out);
  PCD_CUDA(ctx, cudaGetLastError());
  return 0;
}
int precompute_(pcdgpu_ctx* ctx, const void* bases, size_t n, int c, int nwin, void* pre) {
  if (n == 0) return 0;
  precompute_kernel<CV><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(bases, n, c, nwin, pre);
  PCD_CUDA(ctx, cudaGetLastError());
  return 0;
}
}  // namespace

const MsmOps PCD_OPS_NAME = {sizeof(AffinePoint<CV::F>), sizeof(XYZZ<CV>), CV::ScalarParams::ID, run_,
                             to_affine_,                 to_affine_at_,    sum_points_,          multi_mul_,
                             fixed_table_,               fixed_mul_,       precompute_};
