// ntt.cuh -- internal interface of ntt.cu
#pragma once
#include "common.cuh"

struct NttTablesDev {
  const u32* tw;    // omega^k, k < n/2
  const u32* cpow;  // g^i
  const u32* cinv;  // g^-i / n
  const u32* ninv;  // 1/n
  const u32* zinv;  // 1/(g^n - 1): inverse of the vanishing polynomial on the coset
};

int ntt_tables(pcdgpu_ctx* ctx, int field, int log_n, NttTablesDev* out);
// in-place transform of 2^log_n elements at d_data (device), async on ctx->stream
int ntt_run(pcdgpu_ctx* ctx, int field, void* d_data, int log_n, int inverse, int coset);
