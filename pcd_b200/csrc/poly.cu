// poly.cu -- dense-polynomial kernels and KZG10 commit / open on the GPU: the hot operations of the Marlin
// prover the reference binds as MainSNARK / HelpSNARK in its Marlin configuration
// (/root/reference/tests/mnt4_marlin.rs:68-94: MarlinSNARK over MarlinKZG10<E, DensePolynomial<Fr>>), reached
// through IC::MainSNARK::prove / IC::HelpSNARK::prove (/root/reference/src/ec_cycle_pcd/mod.rs:171,179).
//
// Replaces, under ark-marlin's AHP prover and ark-poly-commit:
//   kzg10::KZG10::commit   commitment = MSM(powers_of_g, coeffs) [+ MSM(powers_of_gamma_g, blinding coeffs)]
//   kzg10::KZG10::open     witness polynomial (p(X) - p(z)) / (X - z), committed the same way; random_v = r(z)
//   DensePolynomial mul    product of two coefficient vectors through the evaluation domain (FFT, pointwise, iFFT)
// The AHP round logic itself (sumcheck polynomials, Fiat-Shamir sponge) is host-side orchestration above these
// calls and is not part of this library.
#include "groth16.cuh"
#include "msm_ops.cuh"
#include "ntt.cuh"

#define POLY_CHECK_ARG(ctx, cond, msg)      \
  do {                                      \
    if (!(cond)) {                          \
      if (ctx) (ctx)->set_error("%s", msg); \
      return PCDGPU_E_ARG;                  \
    }                                       \
  } while (0)

static constexpr int DIV_K = 256;  // coefficients per chunk, chunks per super-chunk

// ---- division by (X - z): q_j = sum_{i > j} p_i z^(i - j - 1), p(z) = p_0 + z q_0 -----------------------------
// The Horner recurrence q_{j-1} = p_j + z q_j is cut into chunks of K coefficients and super-chunks of K chunks:
//   up1   v_c = sum_{i in chunk c} p_i z^(i - cK)                              one thread per chunk
//   up2   V_s = sum_{c in super s} v_c (z^K)^(c - sK)                          one thread per super-chunk
//   top   HS_s = V_{s+1} + (z^K)^K HS_{s+1}                                    one thread (n / K^2 steps)
//   down2 H_c = v_{c+1} + z^K H_{c+1} inside each super-chunk, seeded by HS_s  one thread per super-chunk
//   down1 q_j = p_{j+1} + z q_{j+1} inside each chunk, seeded by H_c           one thread per chunk
// (H_c = q at the last index of chunk c).  Sequential depth 4K + n / K^2 products instead of n.
template <class F>
__global__ void __launch_bounds__(128) div_up_kernel(const u32* __restrict__ in, size_t count, const u32* __restrict__ zp,
                                                     int zsel, u32* __restrict__ out) {
  // out[c] = sum_{k < K} in[cK + k] x^k with x = zp[zsel]; in has `count` elements (zero beyond)
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t nch = (count + DIV_K - 1) / DIV_K;
  if (c >= nch) return;
  F x = ld10<F>(zp, zsel);
  size_t lo = c * DIV_K, hi = lo + DIV_K < count ? lo + DIV_K : count;
  F acc = F::zero();
  for (size_t i = hi; i-- > lo;) acc = acc * x + ld10<F>(in, i);
  st10<F>(out, c, acc);
}
// single thread: hs[s] = V[s + 1] + x hs[s + 1], hs[last] = 0;  x = zp[zsel]
template <class F>
__global__ void div_top_kernel(const u32* __restrict__ V, size_t count, const u32* __restrict__ zp, int zsel,
                               u32* __restrict__ hs) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  F x = ld10<F>(zp, zsel);
  F acc = F::zero();
  for (size_t s = count; s-- > 0;) {
    st10<F>(hs, s, acc);
    acc = acc * x + ld10<F>(V, s);
  }
}
// per group g of K entries: h[e] = v[e + 1] + x h[e + 1] walking down from the top entry, h[top] = seed[g]
template <class F>
__global__ void __launch_bounds__(128) div_down_kernel(const u32* __restrict__ v, size_t count,
                                                       const u32* __restrict__ seed, const u32* __restrict__ zp, int zsel,
                                                       u32* __restrict__ h) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t ng = (count + DIV_K - 1) / DIV_K;
  if (g >= ng) return;
  F x = ld10<F>(zp, zsel);
  size_t lo = g * DIV_K, hi = lo + DIV_K < count ? lo + DIV_K : count;
  F acc = ld10<F>(seed, g);
  for (size_t e = hi; e-- > lo;) {
    st10<F>(h, e, acc);
    acc = acc * x + ld10<F>(v, e);
  }
}
// chunk c: q[j] for j in [cK, (c+1)K), seeded by H[c]; the thread of chunk 0 also writes p(z) = p_0 + z q_0
template <class F>
__global__ void __launch_bounds__(128) div_final_kernel(const u32* __restrict__ p, size_t n, const u32* __restrict__ H,
                                                        const u32* __restrict__ zp, u32* __restrict__ q,
                                                        u32* __restrict__ eval) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t nch = (n + DIV_K - 1) / DIV_K;
  if (c >= nch) return;
  F z = ld10<F>(zp, 0);
  size_t lo = c * DIV_K, hi = lo + DIV_K < n ? lo + DIV_K : n;
  F acc = ld10<F>(H, c);  // = q[hi - 1]
  for (size_t j = hi; j-- > lo;) {
    if (j + 1 < n) st10<F>(q, j, acc);  // q has n - 1 coefficients
    acc = acc * z + ld10<F>(p, j);      // -> q[j - 1]; after j = 0 this is p(z)
  }
  if (c == 0) st10<F>(eval, 0, acc);
}
// zp = {z, z^K, z^(K^2)}
template <class F>
__global__ void div_powers_kernel(u32* zp) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  F z = ld10<F>(zp, 0);
  F zk = z.pow64(DIV_K);
  st10<F>(zp, 1, zk);
  st10<F>(zp, 2, zk.pow64(DIV_K));
}

// d_p: n coefficients, d_z: one element (device); d_q: n - 1 coefficients, d_eval: one element.  n >= 1.
template <class F>
static int poly_divide_linear_t(pcdgpu_ctx* ctx, const void* d_p, size_t n, const void* d_z, void* d_q, void* d_eval) {
  cudaStream_t st = ctx->cur();
  const size_t nch = (n + DIV_K - 1) / DIV_K, nsu = (nch + DIV_K - 1) / DIV_K;
  void* w;
  PCD_TRY(ctx->scratch(SLOT_NTT_MIXED, (4 + 2 * nch + 2 * nsu) * 40, &w));
  u32* zp = (u32*)w;
  u32* v = zp + 40;
  u32* H = v + nch * 10;
  u32* V = H + nch * 10;
  u32* HS = V + nsu * 10;
  PCD_CUDA(ctx, cudaMemcpyAsync(zp, d_z, 40, cudaMemcpyDeviceToDevice, st));
  div_powers_kernel<F><<<1, 32, 0, st>>>(zp);
  div_up_kernel<F><<<(unsigned)((nch + 127) / 128), 128, 0, st>>>((const u32*)d_p, n, zp, 0, v);
  div_up_kernel<F><<<(unsigned)((nsu + 127) / 128), 128, 0, st>>>(v, nch, zp, 1, V);
  div_top_kernel<F><<<1, 32, 0, st>>>(V, nsu, zp, 2, HS);
  div_down_kernel<F><<<(unsigned)((nsu + 127) / 128), 128, 0, st>>>(v, nch, HS, zp, 1, H);
  div_final_kernel<F><<<(unsigned)((nch + 127) / 128), 128, 0, st>>>((const u32*)d_p, n, H, zp, (u32*)d_q, (u32*)d_eval);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->launches += 6;
  return 0;
}
static int poly_divide_linear_dev(pcdgpu_ctx* ctx, int field, const void* d_p, size_t n, const void* d_z, void* d_q,
                                  void* d_eval) {
  if (field == PCDGPU_FIELD_R4) return poly_divide_linear_t<FpR4>(ctx, d_p, n, d_z, d_q, d_eval);
  return poly_divide_linear_t<FpQ4>(ctx, d_p, n, d_z, d_q, d_eval);
}

// ---- product of two polynomials through the evaluation domain -------------------------------------------------
template <class F>
__global__ void __launch_bounds__(256) pointwise_mul_kernel(u32* a, const u32* __restrict__ b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st10<F>(a, i, ld10<F>(a, i) * ld10<F>(b, i));
}

static int scalar_field_of_curve(int curve) {
  return (curve == PCDGPU_MNT4_G1 || curve == PCDGPU_MNT4_G2) ? PCDGPU_FIELD_R4 : PCDGPU_FIELD_Q4;
}

// commitment (xyzz) of the coefficient vector d_c (n, Montgomery) [+ blinding d_r (n_r)] into d_out_xyzz[0..1]
static int kzg_commit_xyzz(pcdgpu_ctx* ctx, const pcdgpu_bases* pg, const void* d_c, size_t n, const pcdgpu_bases* pgg,
                           const void* d_r, size_t n_r, void* d_xyzz2, int* count) {
  const MsmOps* ops = msm_ops(pg->curve);
  PCD_TRY(bases_msm(ctx, pg, 0, d_c, 1, n, nullptr, 0, d_xyzz2));
  *count = 1;
  if (pgg && n_r) {
    PCD_TRY(bases_msm(ctx, pgg, 0, d_r, 1, n_r, nullptr, 0, (char*)d_xyzz2 + ops->xyzz_bytes));
    *count = 2;
  }
  return 0;
}

// ---- device vectors of field elements: the arithmetic ark-marlin's AHP prover does on DensePolynomial /
// EvaluationsOnDomain between its FFTs and its commitments (ark-marlin src/ahp/prover.rs; ark-poly) -------------------
enum { VEC_ADD = 0, VEC_SUB = 1, VEC_MUL = 2, VEC_RSUB = 3 };
template <class F>
__device__ __forceinline__ F vec_apply(int op, const F& x, const F& y) {
  return op == VEC_ADD ? x + y : (op == VEC_SUB ? x - y : (op == VEC_MUL ? x * y : y - x));
}
// out[i] = a[i] op b[i]
template <class F>
__global__ void __launch_bounds__(256) vec_binary_kernel(int op, u32* out, const u32* a, const u32* b, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st10<F>(out, i, vec_apply<F>(op, ld10<F>(a, i), ld10<F>(b, i)));
}
// out[i] = a[i] op s (s travels as a kernel argument: no staging buffer, nothing to race with)
template <class F>
__global__ void __launch_bounds__(256) vec_scalar_kernel(int op, u32* out, const u32* a, F s, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st10<F>(out, i, vec_apply<F>(op, ld10<F>(a, i), s));
}
// y[i] += s x[i]
template <class F>
__global__ void __launch_bounds__(256) vec_axpy_kernel(u32* y, F s, const u32* x, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st10<F>(y, i, ld10<F>(y, i) + s * ld10<F>(x, i));
}
// ark-ff batch_inversion: zeros stay zero; Montgomery's trick over VEC_INV_CHUNK elements per thread
static constexpr int VEC_INV_CHUNK = 8;
template <class F>
__global__ void __launch_bounds__(128) vec_inverse_kernel(u32* data, size_t n) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t lo = t * VEC_INV_CHUNK;
  if (lo >= n) return;
  size_t hi = lo + VEC_INV_CHUNK < n ? lo + VEC_INV_CHUNK : n;
  F pre[VEC_INV_CHUNK];
  F acc = F::one();
  for (size_t i = lo; i < hi; i++) {
    pre[i - lo] = acc;
    F v = ld10<F>(data, i);
    if (!v.is_zero()) acc = acc * v;
  }
  acc = acc.inverse();
  for (size_t i = hi; i-- > lo;) {
    F v = ld10<F>(data, i);
    if (v.is_zero()) continue;
    st10<F>(data, i, acc * pre[i - lo]);
    acc = acc * v;
  }
}
// out[i] = scale * base^i
template <class F>
__global__ void __launch_bounds__(256) vec_powers_kernel(u32* out, F base, F scale, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st10<F>(out, i, scale * base.pow64((u64)i));
}
// out[i] = src[index[i]] (0xffffffff: zero)
template <class F>
__global__ void __launch_bounds__(256) vec_gather_kernel(u32* out, const u32* __restrict__ src,
                                                         const u32* __restrict__ index, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 j = index[i];
  st10<F>(out, i, j == 0xffffffffu ? F::zero() : ld10<F>(src, j));
}
// DensePolynomial::divide_by_vanishing_poly for X^N - 1: q[i] = sum_{k >= 1} p[i + kN], r[i] = sum_{k >= 0} p[i + kN]
template <class F>
__global__ void __launch_bounds__(256) div_vanishing_kernel(const u32* __restrict__ p, size_t n, size_t N, u32* q, u32* r) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (i >= n) {
    st10<F>(r, i, F::zero());
    return;
  }
  // walk the residue class from the top so that every quotient coefficient is a running sum
  F acc = F::zero();
  size_t last = i + ((n - 1 - i) / N) * N;  // largest index = i mod N below n (caller guarantees i < n)
  for (size_t j = last; j >= N + i; j -= N) {
    acc = acc + ld10<F>(p, j);
    st10<F>(q, j - N, acc);
  }
  st10<F>(r, i, acc + ld10<F>(p, i));
}
// single thread: out = sum_s V[s] x^s, x = zp[zsel]
template <class F>
__global__ void eval_top_kernel(const u32* __restrict__ V, size_t count, const u32* __restrict__ zp, int zsel, u32* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  F x = ld10<F>(zp, zsel);
  F acc = F::zero();
  for (size_t s = count; s-- > 0;) acc = acc * x + ld10<F>(V, s);
  st10<F>(out, 0, acc);
}
// out[row] = sum_k val[k] x[col[k]] over the row's entries (one thread per row)
template <class F>
__global__ void __launch_bounds__(128) csr_matvec_kernel(CsrDev M, size_t m, const u32* __restrict__ x, u32* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const F one = F::one();
  F acc = F::zero();
  for (u32 k = M.row_ptr[i], hi = M.row_ptr[i + 1]; k < hi; k++) {
    F co = ld10<F>(M.val, k);
    F v = ld10<F>(x, M.col[k]);
    acc = acc + (co == one ? v : co * v);
  }
  st10<F>(out, i, acc);
}

struct pcdgpu_csr {
  pcdgpu_ctx* ctx;
  int field;
  size_t m, ncols, nnz;
  CsrDev M;
  void* storage;
};

template <class F>
static F host_elem(const void* p) {
  F r;
  memcpy(r.l, p, 40);
  return r;
}
#define FIELD_DISPATCH(field, CALL)            \
  do {                                         \
    if ((field) == PCDGPU_FIELD_R4) {          \
      typedef FpR4 F;                          \
      CALL;                                    \
    } else {                                   \
      typedef FpQ4 F;                          \
      CALL;                                    \
    }                                          \
  } while (0)
static inline unsigned nblk(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

template <class F>
static int poly_eval_t(pcdgpu_ctx* ctx, const void* d_p, size_t n, const void* zhost, void* d_out) {
  cudaStream_t st = ctx->cur();
  const size_t nch = (n + DIV_K - 1) / DIV_K, nsu = (nch + DIV_K - 1) / DIV_K;
  void* w;
  PCD_TRY(ctx->scratch(SLOT_NTT_MIXED, (4 + nch + nsu) * 40, &w));
  u32* zp = (u32*)w;
  u32* v = zp + 40;
  u32* V = v + nch * 10;
  vec_powers_kernel<F><<<1, 32, 0, st>>>(zp, F::one(), host_elem<F>(zhost), 1);  // zp[0] = z
  div_powers_kernel<F><<<1, 32, 0, st>>>(zp);
  div_up_kernel<F><<<nblk(nch, 128), 128, 0, st>>>((const u32*)d_p, n, zp, 0, v);
  div_up_kernel<F><<<nblk(nsu, 128), 128, 0, st>>>(v, nch, zp, 1, V);
  eval_top_kernel<F><<<1, 32, 0, st>>>(V, nsu, zp, 2, (u32*)d_out);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->launches += 5;
  return 0;
}

extern "C" {

int pcdgpu_poly_divide_linear(pcdgpu_ctx* ctx, int field, const void* coeffs, size_t n, const void* z, void* quotient,
                              void* eval) {
  if (!ctx) return PCDGPU_E_ARG;
  POLY_CHECK_ARG(ctx, field == PCDGPU_FIELD_R4 || field == PCDGPU_FIELD_Q4, "unknown field id");
  POLY_CHECK_ARG(ctx, coeffs && z && eval && n >= 1 && (n == 1 || quotient), "bad argument");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  void *dp, *dq;
  PCD_TRY(ctx->scratch(SLOT_IO, (n + 2) * 40, &dp));
  PCD_TRY(ctx->scratch(SLOT_IO2, (n + 2) * 40, &dq));
  void* dz = (char*)dp + n * 40;
  void* de = (char*)dq + n * 40;
  PCD_CUDA(ctx, cudaMemcpyAsync(dp, coeffs, n * 40, cudaMemcpyHostToDevice, ctx->stream));
  memcpy(ctx->pinned, z, 40);
  PCD_CUDA(ctx, cudaMemcpyAsync(dz, ctx->pinned, 40, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(poly_divide_linear_dev(ctx, field, dp, n, dz, dq, de));
  if (n > 1) PCD_CUDA(ctx, cudaMemcpyAsync(quotient, dq, (n - 1) * 40, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaMemcpyAsync(eval, de, 40, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int pcdgpu_poly_mul(pcdgpu_ctx* ctx, int field, const void* a, size_t na, const void* b, size_t nb, void* out) {
  if (!ctx) return PCDGPU_E_ARG;
  POLY_CHECK_ARG(ctx, field == PCDGPU_FIELD_R4 || field == PCDGPU_FIELD_Q4, "unknown field id");
  POLY_CHECK_ARG(ctx, a && b && out && na >= 1 && nb >= 1, "bad argument");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t n;
  int da, db;
  if (ntt_domain_shape(field, na + nb - 1, &n, &da, &db) != 0) {
    ctx->set_error("polynomial product of %zu coefficients needs a domain the field does not have", na + nb - 1);
    return PCDGPU_E_DOMAIN;
  }
  void *x, *y;
  PCD_TRY(ctx->scratch(SLOT_WM_A, n * 40, &x));
  PCD_TRY(ctx->scratch(SLOT_WM_B, n * 40, &y));
  cudaStream_t st = ctx->stream;
  PCD_CUDA(ctx, cudaMemsetAsync((char*)x + na * 40, 0, (n - na) * 40, st));
  PCD_CUDA(ctx, cudaMemsetAsync((char*)y + nb * 40, 0, (n - nb) * 40, st));
  PCD_CUDA(ctx, cudaMemcpyAsync(x, a, na * 40, cudaMemcpyHostToDevice, st));
  PCD_CUDA(ctx, cudaMemcpyAsync(y, b, nb * 40, cudaMemcpyHostToDevice, st));
  PCD_TRY(ntt_run_general(ctx, field, x, da, db, 0, 0));
  PCD_TRY(ntt_run_general(ctx, field, y, da, db, 0, 0));
  if (field == PCDGPU_FIELD_R4) pointwise_mul_kernel<FpR4><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((u32*)x, (const u32*)y, n);
  else pointwise_mul_kernel<FpQ4><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((u32*)x, (const u32*)y, n);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  PCD_TRY(ntt_run_general(ctx, field, x, da, db, 1, 0));
  PCD_CUDA(ctx, cudaMemcpyAsync(out, x, (na + nb - 1) * 40, cudaMemcpyDeviceToHost, st));
  PCD_CUDA(ctx, cudaStreamSynchronize(st));
  return 0;
}

int pcdgpu_kzg_commit(pcdgpu_ctx* ctx, const pcdgpu_bases* powers_of_g, const void* coeffs, size_t n,
                      const pcdgpu_bases* powers_of_gamma_g, const void* rand_coeffs, size_t n_rand, void* out_affine) {
  if (!ctx) return PCDGPU_E_ARG;
  POLY_CHECK_ARG(ctx, powers_of_g && out_affine && (n == 0 || coeffs), "null pointer");
  POLY_CHECK_ARG(ctx, n <= powers_of_g->n, "polynomial degree exceeds the committer key (powers_of_g)");
  POLY_CHECK_ARG(ctx, n_rand == 0 || (powers_of_gamma_g && rand_coeffs && n_rand <= powers_of_gamma_g->n &&
                                      powers_of_gamma_g->curve == powers_of_g->curve),
                 "bad blinding polynomial / powers_of_gamma_g");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  const MsmOps* ops = msm_ops(powers_of_g->curve);
  void *dc, *dr, *misc;
  PCD_TRY(ctx->scratch(SLOT_IO, (n + 1) * 40, &dc));
  PCD_TRY(ctx->scratch(SLOT_IO2, (n_rand + 1) * 40, &dr));
  PCD_TRY(ctx->scratch(SLOT_MISC, 8192, &misc));
  if (n) PCD_CUDA(ctx, cudaMemcpyAsync(dc, coeffs, n * 40, cudaMemcpyHostToDevice, ctx->stream));
  if (n_rand) PCD_CUDA(ctx, cudaMemcpyAsync(dr, rand_coeffs, n_rand * 40, cudaMemcpyHostToDevice, ctx->stream));
  int count = 0;
  PCD_TRY(kzg_commit_xyzz(ctx, powers_of_g, dc, n, powers_of_gamma_g, dr, n_rand, misc, &count));
  void* d_aff = (char*)misc + 2 * ops->xyzz_bytes;
  PCD_TRY(ops->sum_points(ctx, misc, 0, 1, count, nullptr, 0, d_aff));
  PCD_CUDA(ctx, cudaMemcpyAsync(out_affine, d_aff, ops->affine_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int pcdgpu_kzg_open(pcdgpu_ctx* ctx, const pcdgpu_bases* powers_of_g, const void* coeffs, size_t n,
                    const pcdgpu_bases* powers_of_gamma_g, const void* rand_coeffs, size_t n_rand, const void* z,
                    void* out_w_affine, void* out_value, void* out_random_v) {
  if (!ctx) return PCDGPU_E_ARG;
  POLY_CHECK_ARG(ctx, powers_of_g && out_w_affine && z && coeffs && n >= 1, "bad argument");
  POLY_CHECK_ARG(ctx, n <= powers_of_g->n, "polynomial degree exceeds the committer key (powers_of_g)");
  POLY_CHECK_ARG(ctx, n_rand == 0 || (powers_of_gamma_g && rand_coeffs && out_random_v &&
                                      n_rand <= powers_of_gamma_g->n && powers_of_gamma_g->curve == powers_of_g->curve),
                 "bad blinding polynomial / powers_of_gamma_g");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  const MsmOps* ops = msm_ops(powers_of_g->curve);
  const int field = scalar_field_of_curve(powers_of_g->curve);
  // IO: p | z | eval    IO2: q (witness polynomial)     WM_C: r | r's witness | r(z)
  void *dp, *dq, *drr, *misc;
  PCD_TRY(ctx->scratch(SLOT_IO, (n + 2) * 40, &dp));
  PCD_TRY(ctx->scratch(SLOT_IO2, (n + 1) * 40, &dq));
  PCD_TRY(ctx->scratch(SLOT_WM_C, (2 * n_rand + 2) * 40, &drr));
  PCD_TRY(ctx->scratch(SLOT_MISC, 8192, &misc));
  void* dz = (char*)dp + n * 40;
  void* de = (char*)dp + (n + 1) * 40;
  void* drq = (char*)drr + n_rand * 40;
  void* dre = (char*)drr + 2 * n_rand * 40;
  cudaStream_t st = ctx->stream;
  PCD_CUDA(ctx, cudaMemcpyAsync(dp, coeffs, n * 40, cudaMemcpyHostToDevice, st));
  memcpy(ctx->pinned, z, 40);
  PCD_CUDA(ctx, cudaMemcpyAsync(dz, ctx->pinned, 40, cudaMemcpyHostToDevice, st));
  PCD_TRY(poly_divide_linear_dev(ctx, field, dp, n, dz, dq, de));
  const bool hiding = n_rand > 0;
  if (hiding) {
    PCD_CUDA(ctx, cudaMemcpyAsync(drr, rand_coeffs, n_rand * 40, cudaMemcpyHostToDevice, st));
    PCD_TRY(poly_divide_linear_dev(ctx, field, drr, n_rand, dz, drq, dre));
  }
  int count = 0;
  PCD_TRY(kzg_commit_xyzz(ctx, powers_of_g, dq, n - 1, hiding ? powers_of_gamma_g : nullptr, drq,
                          hiding ? n_rand - 1 : 0, misc, &count));
  void* d_aff = (char*)misc + 2 * ops->xyzz_bytes;
  PCD_TRY(ops->sum_points(ctx, misc, 0, 1, count, nullptr, 0, d_aff));
  PCD_CUDA(ctx, cudaMemcpyAsync(out_w_affine, d_aff, ops->affine_bytes, cudaMemcpyDeviceToHost, st));
  if (out_value) PCD_CUDA(ctx, cudaMemcpyAsync(out_value, de, 40, cudaMemcpyDeviceToHost, st));
  if (hiding) PCD_CUDA(ctx, cudaMemcpyAsync(out_random_v, dre, 40, cudaMemcpyDeviceToHost, st));
  PCD_CUDA(ctx, cudaStreamSynchronize(st));
  return 0;
}


// ---- device memory and device vectors (stream-ordered: everything below is asynchronous on the context's stream
// unless it hands a value back to the host) -----------------------------------------------------------------------
int pcdgpu_dev_alloc(pcdgpu_ctx* ctx, size_t bytes, void** d_out) {
  if (!ctx || !d_out) return PCDGPU_E_ARG;
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaError_t e = cudaMallocAsync(d_out, bytes ? bytes : 8, ctx->stream);
  if (e != cudaSuccess) {
    ctx->set_error("cudaMallocAsync(%zu): %s", bytes, cudaGetErrorString(e));
    return PCDGPU_E_NOMEM;
  }
  return 0;
}
int pcdgpu_dev_free(pcdgpu_ctx* ctx, void* d) {
  if (!ctx) return PCDGPU_E_ARG;
  if (d) PCD_CUDA(ctx, cudaFreeAsync(d, ctx->stream));
  return 0;
}
int pcdgpu_dev_upload(pcdgpu_ctx* ctx, void* d_dst, const void* src, size_t bytes) {
  if (!ctx || (bytes && (!d_dst || !src))) return PCDGPU_E_ARG;
  PCD_CUDA(ctx, cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the caller's buffer is pageable and may be reused at once
  return 0;
}
int pcdgpu_dev_download(pcdgpu_ctx* ctx, void* dst, const void* d_src, size_t bytes) {
  if (!ctx || (bytes && (!dst || !d_src))) return PCDGPU_E_ARG;
  PCD_CUDA(ctx, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
int pcdgpu_dev_copy(pcdgpu_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
  if (!ctx || (bytes && (!d_dst || !d_src))) return PCDGPU_E_ARG;
  PCD_CUDA(ctx, cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}
int pcdgpu_dev_zero(pcdgpu_ctx* ctx, void* d, size_t bytes) {
  if (!ctx || (bytes && !d)) return PCDGPU_E_ARG;
  PCD_CUDA(ctx, cudaMemsetAsync(d, 0, bytes, ctx->stream));
  return 0;
}

#define VEC_PROLOGUE(ctx, field)                                                                       \
  if (!ctx) return PCDGPU_E_ARG;                                                                       \
  POLY_CHECK_ARG(ctx, field == PCDGPU_FIELD_R4 || field == PCDGPU_FIELD_Q4, "unknown field id");       \
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));                                                           \
  cudaStream_t st = ctx->stream;                                                                       \
  (void)st
#define VEC_EPILOGUE(ctx, k)            \
  PCD_CUDA(ctx, cudaGetLastError());    \
  ctx->launches += (k);                 \
  return 0

int pcdgpu_vec_binary_dev(pcdgpu_ctx* ctx, int field, int op, void* d_out, const void* d_a, const void* d_b, size_t n) {
  VEC_PROLOGUE(ctx, field);
  POLY_CHECK_ARG(ctx, op >= VEC_ADD && op <= VEC_RSUB && (n == 0 || (d_out && d_a && d_b)), "bad argument");
  if (!n) return 0;
  FIELD_DISPATCH(field, (vec_binary_kernel<F><<<nblk(n, 256), 256, 0, st>>>(op, (u32*)d_out, (const u32*)d_a, (const u32*)d_b, n)));
  VEC_EPILOGUE(ctx, 1);
}
int pcdgpu_vec_scalar_dev(pcdgpu_ctx* ctx, int field, int op, void* d_out, const void* d_a, const void* scalar, size_t n) {
  VEC_PROLOGUE(ctx, field);
  POLY_CHECK_ARG(ctx, op >= VEC_ADD && op <= VEC_RSUB && scalar && (n == 0 || (d_out && d_a)), "bad argument");
  if (!n) return 0;
  FIELD_DISPATCH(field, (vec_scalar_kernel<F><<<nblk(n, 256), 256, 0, st>>>(op, (u32*)d_out, (const u32*)d_a, host_elem<F>(scalar), n)));
  VEC_EPILOGUE(ctx, 1);
}
int pcdgpu_vec_axpy_dev(pcdgpu_ctx* ctx, int field, void* d_y, const void* scalar, const void* d_x, size_t n) {
  VEC_PROLOGUE(ctx, field);
  POLY_CHECK_ARG(ctx, scalar && (n == 0 || (d_y && d_x)), "bad argument");
  if (!n) return 0;
  FIELD_DISPATCH(field, (vec_axpy_kernel<F><<<nblk(n, 256), 256, 0, st>>>((u32*)d_y, host_elem<F>(scalar), (const u32*)d_x, n)));
  VEC_EPILOGUE(ctx, 1);
}
int pcdgpu_vec_inverse_dev(pcdgpu_ctx* ctx, int field, void* d_data, size_t n) {
  VEC_PROLOGUE(ctx, field);
  POLY_CHECK_ARG(ctx, n == 0 || d_data, "bad argument");
  if (!n) return 0;
  size_t threads = (n + VEC_INV_CHUNK - 1) / VEC_INV_CHUNK;
  FIELD_DISPATCH(field, (vec_inverse_kernel<F><<<nblk(threads, 128), 128, 0, st>>>((u32*)d_data, n)));
  VEC_EPILOGUE(ctx, 1);
}
int pcdgpu_vec_powers_dev(pcdgpu_ctx* ctx, int field, void* d_out, const void* base, const void* scale, size_t n) {
  VEC_PROLOGUE(ctx, field);
  POLY_CHECK_ARG(ctx, base && scale && (n == 0 || d_out), "bad argument");
  if (!n) return 0;
  FIELD_DISPATCH(field, (vec_powers_kernel<F><<<nblk(n, 256), 256, 0, st>>>((u32*)d_out, host_elem<F>(base), host_elem<F>(scale), n)));
  VEC_EPILOGUE(ctx, 1);
}
int pcdgpu_vec_gather_dev(pcdgpu_ctx* ctx, int field, void* d_out, const void* d_src, const uint32_t* d_index, size_t n) {
  VEC_PROLOGUE(ctx, field);
  POLY_CHECK_ARG(ctx, n == 0 || (d_out && d_src && d_index), "bad argument");
  if (!n) return 0;
  vec_gather_kernel<FpR4><<<nblk(n, 256), 256, 0, st>>>((u32*)d_out, (const u32*)d_src, d_index, n);  // a copy: field-blind
  VEC_EPILOGUE(ctx, 1);
}
int pcdgpu_poly_eval_dev(pcdgpu_ctx* ctx, int field, const void* d_coeffs, size_t n, const void* z, void* out) {
  VEC_PROLOGUE(ctx, field);
  POLY_CHECK_ARG(ctx, z && out && (n == 0 || d_coeffs), "bad argument");
  if (!n) {
    memset(out, 0, 40);
    return 0;
  }
  void* misc;
  PCD_TRY(ctx->scratch(SLOT_MISC, 8192, &misc));
  FIELD_DISPATCH(field, PCD_TRY(poly_eval_t<F>(ctx, d_coeffs, n, z, misc)));
  PCD_CUDA(ctx, cudaMemcpyAsync(out, misc, 40, cudaMemcpyDeviceToHost, st));
  PCD_CUDA(ctx, cudaStreamSynchronize(st));
  return 0;
}
int pcdgpu_poly_divide_vanishing_dev(pcdgpu_ctx* ctx, int field, const void* d_p, size_t n, size_t domain_n, void* d_q,
                                     void* d_r) {
  VEC_PROLOGUE(ctx, field);
  POLY_CHECK_ARG(ctx, d_p && d_r && n >= 1 && domain_n >= 1 && (n <= domain_n || d_q), "bad argument");
  FIELD_DISPATCH(field, (div_vanishing_kernel<F><<<nblk(domain_n, 256), 256, 0, st>>>((const u32*)d_p, n, domain_n, (u32*)d_q, (u32*)d_r)));
  VEC_EPILOGUE(ctx, 1);
}
int pcdgpu_poly_divide_linear_dev(pcdgpu_ctx* ctx, int field, const void* d_p, size_t n, const void* z, void* d_q,
                                  void* out_eval) {
  VEC_PROLOGUE(ctx, field);
  POLY_CHECK_ARG(ctx, d_p && z && n >= 1 && (n == 1 || d_q), "bad argument");
  void* misc;
  PCD_TRY(ctx->scratch(SLOT_MISC, 8192, &misc));
  FIELD_DISPATCH(field, (vec_powers_kernel<F><<<1, 32, 0, st>>>((u32*)misc, F::one(), host_elem<F>(z), 1)));
  PCD_TRY(poly_divide_linear_dev(ctx, field, d_p, n, misc, d_q, (char*)misc + 40));
  ctx->launches += 1;
  if (out_eval) {
    PCD_CUDA(ctx, cudaMemcpyAsync(out_eval, (char*)misc + 40, 40, cudaMemcpyDeviceToHost, st));
    PCD_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return 0;
}
int pcdgpu_ntt_general_dev(pcdgpu_ctx* ctx, int field, void* d_data, int pow7, int pow2, int inverse, int coset) {
  VEC_PROLOGUE(ctx, field);
  POLY_CHECK_ARG(ctx, d_data && pow7 >= 0 && pow2 >= 0, "bad argument");
  return ntt_run_general(ctx, field, d_data, pow7, pow2, inverse, coset);
}

int pcdgpu_csr_upload(pcdgpu_ctx* ctx, int field, size_t m, size_t ncols, const uint32_t* row_ptr, const uint32_t* col,
                      const void* val, pcdgpu_csr** out) {
  VEC_PROLOGUE(ctx, field);
  POLY_CHECK_ARG(ctx, out && row_ptr && m >= 1 && ncols >= 1, "bad argument");
  POLY_CHECK_ARG(ctx, row_ptr[0] == 0, "row_ptr[0] must be 0");
  for (size_t i = 0; i < m; i++) POLY_CHECK_ARG(ctx, row_ptr[i] <= row_ptr[i + 1], "row_ptr is not monotone");
  const size_t nnz = row_ptr[m];
  POLY_CHECK_ARG(ctx, nnz == 0 || (col && val), "null col / val");
  for (size_t k = 0; k < nnz; k++) POLY_CHECK_ARG(ctx, col[k] < ncols, "column index out of range");
  pcdgpu_csr* c = new pcdgpu_csr();
  c->ctx = ctx;
  c->field = field;
  c->m = m;
  c->ncols = ncols;
  c->nnz = nnz;
  const size_t off_col = ((m + 1) * 4 + 15) / 16 * 16, off_val = off_col + (nnz * 4 + 15) / 16 * 16;
  if (cudaMalloc(&c->storage, off_val + nnz * 40 + 16) != cudaSuccess) {
    delete c;
    ctx->set_error("cudaMalloc for a CSR matrix of %zu entries failed", nnz);
    return PCDGPU_E_NOMEM;
  }
  char* base = (char*)c->storage;
  cudaMemcpyAsync(base, row_ptr, (m + 1) * 4, cudaMemcpyHostToDevice, st);
  if (nnz) {
    cudaMemcpyAsync(base + off_col, col, nnz * 4, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(base + off_val, val, nnz * 40, cudaMemcpyHostToDevice, st);
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) {
    cudaFree(c->storage);
    delete c;
    ctx->set_error("upload of a CSR matrix failed");
    return PCDGPU_E_CUDA;
  }
  c->M.row_ptr = (const u32*)base;
  c->M.col = (const u32*)(base + off_col);
  c->M.val = (const u32*)(base + off_val);
  *out = c;
  return 0;
}
void pcdgpu_csr_free(pcdgpu_csr* c) {
  if (!c) return;
  cudaFree(c->storage);
  delete c;
}
int pcdgpu_csr_matvec_dev(pcdgpu_ctx* ctx, const pcdgpu_csr* c, const void* d_x, void* d_out) {
  if (!ctx || !c) return PCDGPU_E_ARG;
  const int field = c->field;
  VEC_PROLOGUE(ctx, field);
  POLY_CHECK_ARG(ctx, d_x && d_out, "bad argument");
  FIELD_DISPATCH(field, (csr_matvec_kernel<F><<<nblk(c->m, 128), 128, 0, st>>>(c->M, c->m, (const u32*)d_x, (u32*)d_out)));
  VEC_EPILOGUE(ctx, 1);
}

// KZG10::commit over device-resident coefficients.  `shift` = first power used (MarlinKZG10's shifted commitment of
// a degree-bounded polynomial: powers_of_g[max_degree - bound ..]); the blinding polynomial always uses
// powers_of_gamma_g from 0.  out_affine: host.
int pcdgpu_kzg_commit_dev(pcdgpu_ctx* ctx, const pcdgpu_bases* powers_of_g, size_t shift, const void* d_coeffs, size_t n,
                          const pcdgpu_bases* powers_of_gamma_g, const void* d_rand, size_t n_rand, void* out_affine) {
  if (!ctx) return PCDGPU_E_ARG;
  POLY_CHECK_ARG(ctx, powers_of_g && out_affine && (n == 0 || d_coeffs), "null pointer");
  POLY_CHECK_ARG(ctx, shift + n <= powers_of_g->n, "polynomial degree (+ shift) exceeds the committer key (powers_of_g)");
  POLY_CHECK_ARG(ctx, n_rand == 0 || (powers_of_gamma_g && d_rand && n_rand <= powers_of_gamma_g->n &&
                                      powers_of_gamma_g->curve == powers_of_g->curve),
                 "bad blinding polynomial / powers_of_gamma_g");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  const MsmOps* ops = msm_ops(powers_of_g->curve);
  void* misc;
  PCD_TRY(ctx->scratch(SLOT_MISC, 8192, &misc));
  PCD_TRY(bases_msm(ctx, powers_of_g, shift, d_coeffs, 1, n, nullptr, 0, misc));
  int count = 1;
  if (n_rand) {
    PCD_TRY(bases_msm(ctx, powers_of_gamma_g, 0, d_rand, 1, n_rand, nullptr, 0, (char*)misc + ops->xyzz_bytes));
    count = 2;
  }
  void* d_aff = (char*)misc + 2 * ops->xyzz_bytes;
  PCD_TRY(ops->sum_points(ctx, misc, 0, 1, count, nullptr, 0, d_aff));
  PCD_CUDA(ctx, cudaMemcpyAsync(out_affine, d_aff, ops->affine_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

}  // extern "C"
