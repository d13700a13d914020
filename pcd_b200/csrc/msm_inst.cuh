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
int xyzz_sum_(pcdgpu_ctx* ctx, const void* in, size_t n, void* out) {
  xyzz_sum_kernel<CV><<<1, 32, 0, ctx->stream>>>(in, n, out);
  PCD_CUDA(ctx, cudaGetLastError());
  return 0;
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
                             to_affine_,                 xyzz_sum_,        fixed_table_,         fixed_mul_,
                             precompute_};
