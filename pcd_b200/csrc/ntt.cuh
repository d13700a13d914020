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
int ntt_run_batch(pcdgpu_ctx* ctx, int field, void* d_data, int log_n, int inverse, int coset, int batch);
// transform on the general domain 7^a 2^b (a > 0 only on q4); in place, async on ctx->stream
int ntt_run_general(pcdgpu_ctx* ctx, int field, void* d_data, int a, int b, int inverse, int coset);
// GeneralEvaluationDomain::new(min_size) -> n = 7^a 2^b, or PCDGPU_E_DOMAIN
int ntt_domain_shape(int field, size_t min_size, size_t* n, int* a, int* b);
// 1 / (g^n - 1) for a general domain (device pointer, one element)
int ntt_zinv_general(pcdgpu_ctx* ctx, int field, size_t n, const u32** d_zinv);
