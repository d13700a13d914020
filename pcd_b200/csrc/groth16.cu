// groth16.cu -- R1CS -> QAP witness map and the assembly of a Groth16 proof on the GPU.
//
// Replaces ark-groth16 R1CStoQAP::witness_map (r1cs_to_qap.rs) and create_proof_with_reduction
// (prover.rs) -- SURVEY.md B.1 / B.2 -- reached from IC::MainSNARK::prove / IC::HelpSNARK::prove
// (/root/reference/src/ec_cycle_pcd/mod.rs:171,179).  Field results are canonical, so h and the
// affine proof points are bit-identical to any correct CPU evaluation of the same formulas.
#include "groth16.cuh"

#include "msm_ops.cuh"
#include "ntt.cuh"

// ---- CSR sparse matrix x assignment (ark-groth16 evaluate_constraint) ---------------------------
// blockIdx.y selects the matrix.  out[i] = <M_i, z> for i < m; a[m + j] = z[j] for the instance
// variables (the rows that make the QAP's A polynomials linearly independent); 0 elsewhere.
template <class F>
__global__ void __launch_bounds__(128) spmv_kernel(CsrDev A, CsrDev B, CsrDev C, const u32* __restrict__ z, size_t m,
                                                   size_t num_inputs, size_t n, u32* a, u32* b, u32* c, int mat_base) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int mat = blockIdx.y + mat_base;
  const CsrDev& M = mat == 0 ? A : (mat == 1 ? B : C);
  u32* out = mat == 0 ? a : (mat == 1 ? b : c);
  F acc = F::zero();
  if (i < m) {
    const F one = F::one();
    u32 lo = M.row_ptr[i], hi = M.row_ptr[i + 1];
    for (u32 k = lo; k < hi; k++) {
      F co = ld10<F>(M.val, k);
      F v = ld10<F>(z, M.col[k]);
      acc = acc + (co == one ? v : co * v);
    }
  } else if (mat == 0 && i < m + num_inputs) {
    acc = ld10<F>(z, i - m);
  }
  st10<F>(out, i, acc);
}

// h[i] = (a[i] * b[i] - c[i]) / Z(g w^i),  Z constant on the coset: g^n - 1
template <class F>
__global__ void __launch_bounds__(256) qap_combine_kernel(u32* a, const u32* __restrict__ b, const u32* __restrict__ c,
                                                          const u32* __restrict__ zinv, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F zi = ld10<F>(zinv, 0);
  F r = (ld10<F>(a, i) * ld10<F>(b, i) - ld10<F>(c, i)) * zi;
  st10<F>(a, i, r);
}

template <class F>
static int witness_map_t(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, const void* d_z, void** d_h) {
  int field = r->pairing == PCDGPU_MNT4_298 ? PCDGPU_FIELD_R4 : PCDGPU_FIELD_Q4;
  size_t n = r->n;
  if (r->dom_a < 0) {
    ctx->set_error("witness map needs a domain of %zu elements; the field has none that large", r->m + r->num_inputs);
    return PCDGPU_E_DOMAIN;
  }
  // a, b, c back to back: on a radix-2 domain the three chains run as ONE batched transform per step (the kernel's
  // batch axis): two launch sequences instead of six, and three times the CTAs per launch (at 2^16 a single transform
  // is 128 CTAs on 148 SMs)
  void *a, *b, *c;
  PCD_TRY(ctx->scratch(SLOT_WM_A, 3 * n * 40, &a));
  b = (char*)a + n * 40;
  c = (char*)b + n * 40;
  dim3 grid((unsigned)((n + 127) / 128), 3);
  int ps = ctx->prof_begin(PROF_SPMV, (double)r->m * 3);
  ctx->launches += 2;
  spmv_kernel<F><<<grid, 128, 0, ctx->stream>>>(r->A, r->B, r->C, (const u32*)d_z, r->m, r->num_inputs, n, (u32*)a,
                                                (u32*)b, (u32*)c, 0);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->prof_end(ps);
  if (r->dom_a == 0) {
    PCD_TRY(ntt_run_batch(ctx, field, a, r->dom_b, 1, 0, 3));
    PCD_TRY(ntt_run_batch(ctx, field, a, r->dom_b, 0, 1, 3));
  } else {
    void* v[3] = {a, b, c};
    for (int i = 0; i < 3; i++) {
      PCD_TRY(ntt_run_general(ctx, field, v[i], r->dom_a, r->dom_b, 1, 0));
      PCD_TRY(ntt_run_general(ctx, field, v[i], r->dom_a, r->dom_b, 0, 1));
    }
  }
  const u32* zinv;
  PCD_TRY(ntt_zinv_general(ctx, field, n, &zinv));
  qap_combine_kernel<F><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((u32*)a, (const u32*)b, (const u32*)c,
                                                                              zinv, n);
  PCD_CUDA(ctx, cudaGetLastError());
  PCD_TRY(ntt_run_general(ctx, field, a, r->dom_a, r->dom_b, 1, 1));
  *d_h = a;
  return 0;
}

// ---- the witness map in two stages, for distributing the three vectors over GPUs -------------
// stage 1: d_out = coset_fft(ifft(M z)) for ONE of the matrices (which = 0: A with the instance rows, 1: B, 2: C)
template <class F>
static int qap_vector_t(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, int which, const void* d_z, void* d_out) {
  int field = r->pairing == PCDGPU_MNT4_298 ? PCDGPU_FIELD_R4 : PCDGPU_FIELD_Q4;
  if (r->dom_a < 0) {
    ctx->set_error("witness map needs a domain of %zu elements; the field has none that large", r->m + r->num_inputs);
    return PCDGPU_E_DOMAIN;
  }
  size_t n = r->n;
  dim3 grid((unsigned)((n + 127) / 128), 1);
  int ps = ctx->prof_begin(PROF_SPMV, (double)r->m);
  ctx->launches += 1;
  spmv_kernel<F><<<grid, 128, 0, ctx->stream>>>(r->A, r->B, r->C, (const u32*)d_z, r->m, r->num_inputs, n, (u32*)d_out,
                                                (u32*)d_out, (u32*)d_out, which);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->prof_end(ps);
  PCD_TRY(ntt_run_general(ctx, field, d_out, r->dom_a, r->dom_b, 1, 0));
  PCD_TRY(ntt_run_general(ctx, field, d_out, r->dom_a, r->dom_b, 0, 1));
  return 0;
}
int qap_vector_dev(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, int which, const void* d_z, void* d_out) {
  if (r->pairing == PCDGPU_MNT4_298) return qap_vector_t<FpR4>(ctx, r, which, d_z, d_out);
  return qap_vector_t<FpQ4>(ctx, r, which, d_z, d_out);
}
// stage 2: d_a <- coset_ifft((a * b - c) / Z)  (a, b, c: the three stage-1 vectors; the result replaces a)
int qap_combine_dev(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, void* d_a, const void* d_b, const void* d_c) {
  int field = r->pairing == PCDGPU_MNT4_298 ? PCDGPU_FIELD_R4 : PCDGPU_FIELD_Q4;
  size_t n = r->n;
  const u32* zinv;
  PCD_TRY(ntt_zinv_general(ctx, field, n, &zinv));
  ctx->launches += 1;
  if (field == PCDGPU_FIELD_R4)
    qap_combine_kernel<FpR4><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((u32*)d_a, (const u32*)d_b,
                                                                                   (const u32*)d_c, zinv, n);
  else
    qap_combine_kernel<FpQ4><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((u32*)d_a, (const u32*)d_b,
                                                                                   (const u32*)d_c, zinv, n);
  PCD_CUDA(ctx, cudaGetLastError());
  return ntt_run_general(ctx, field, d_a, r->dom_a, r->dom_b, 1, 1);
}

int witness_map_dev(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, const void* d_z, void** d_h) {
  if (r->pairing == PCDGPU_MNT4_298) return witness_map_t<FpR4>(ctx, r, d_z, d_h);
  return witness_map_t<FpQ4>(ctx, r, d_z, d_h);
}

// ---- proof assembly ------------------------------------------------------------------------------
// ark-groth16 prover.rs computes
//   g_a  = r delta1 + a_query[0] + MSM(a_query[1..], z[1..]) + alpha
//   g1_b = s delta1 + b_g1_query[0] + MSM(b_g1_query[1..], z[1..]) + beta1      (g2_b likewise in G2)
//   g_c  = s g_a + r g1_b - (r s) delta1 + MSM(l_query, aux) + MSM(h_query, h)
// Every term with a key point as its base is folded into the MSMs as an extra (point, scalar) pair
// (pcdgpu_pk_upload appends the points; groth16_prepare writes the scalars), so that g_a, g1_b, g2_b
// and l' = l_acc - (r s) delta1 come straight out of the bucket method.  What is left is the one
// double-scalar multiplication s g_a + r g1_b on fresh points, done jointly (Straus).
template <class SP>
__global__ void groth16_prepare_kernel(const u32* __restrict__ rs, u32* __restrict__ extras) {
  typedef Fp<SP> Fr;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Fr r, s;
#pragma unroll
  for (int i = 0; i < 10; i++) {
    r.l[i] = rs[i];
    s.l[i] = rs[10 + i];
  }
  Fr rsm = r.to_mont() * s.to_mont();
  Fr rs_ = rsm.from_mont(), nrs = rsm.neg().from_mont();
#pragma unroll
  for (int i = 0; i < 10; i++) {
    u32 one = i == 0 ? 1u : 0u;
    extras[0 * 10 + i] = r.l[i];
    extras[1 * 10 + i] = one;
    extras[2 * 10 + i] = one;
    extras[3 * 10 + i] = s.l[i];
    extras[4 * 10 + i] = one;
    extras[5 * 10 + i] = one;
    extras[6 * 10 + i] = nrs.l[i];
    // small proofs: s g_a and r g1_b as MSMs of their own (scalars s z, r z): constant pairs (delta, a[0] / b[0], alpha / beta)
    extras[7 * 10 + i] = rs_.l[i];
    extras[8 * 10 + i] = s.l[i];
    extras[9 * 10 + i] = s.l[i];
    extras[10 * 10 + i] = rs_.l[i];
    extras[11 * 10 + i] = r.l[i];
    extras[12 * 10 + i] = r.l[i];
  }
}
// sz[i] = s z[i], rz[i] = r z[i] (Montgomery in, Montgomery out); rs = r | s plain
template <class F>
__global__ void __launch_bounds__(128) groth16_scale_kernel(const u32* __restrict__ z, size_t n, const u32* __restrict__ rs,
                                                            u32* __restrict__ sz, u32* __restrict__ rz) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F r = ld10<F>(rs, 0).to_mont(), s = ld10<F>(rs, 1).to_mont();
  F v = ld10<F>(z, i);
  st10<F>(sz, i, s * v);
  st10<F>(rz, i, r * v);
}
int groth16_scale(pcdgpu_ctx* ctx, int pairing, const void* d_z, size_t n, const u32* d_rs, void* d_sz, void* d_rz) {
  if (n == 0) return 0;
  ctx->launches += 1;
  if (pairing == PCDGPU_MNT4_298)
    groth16_scale_kernel<FpR4><<<(unsigned)((n + 127) / 128), 128, 0, ctx->cur()>>>((const u32*)d_z, n, d_rs, (u32*)d_sz, (u32*)d_rz);
  else
    groth16_scale_kernel<FpQ4><<<(unsigned)((n + 127) / 128), 128, 0, ctx->cur()>>>((const u32*)d_z, n, d_rs, (u32*)d_sz, (u32*)d_rz);
  PCD_CUDA(ctx, cudaGetLastError());
  return 0;
}

int groth16_prepare(pcdgpu_ctx* ctx, int pairing, const u32* d_rs, u32* d_extras) {
  ctx->launches += 1;
  if (pairing == PCDGPU_MNT4_298) groth16_prepare_kernel<ParamsR4><<<1, 32, 0, ctx->stream>>>(d_rs, d_extras);
  else groth16_prepare_kernel<ParamsQ4><<<1, 32, 0, ctx->stream>>>(d_rs, d_extras);
  PCD_CUDA(ctx, cudaGetLastError());
  return 0;
}

// The tail of a proof is split so that only two additions and one normalisation are left after the last MSM:
//   groth16_straus   T = s g_a + r g1_b, as soon as the a and b_g1 MSMs are done (overlaps the witness map / h MSM)
//   point_to_affine  A = g_a and B = g2_b, each right after its MSM, on that MSM's lane
//   groth16_finish   C = T + l' + h
// All of it runs on the lane-cooperative group law (wec.cuh, through the per-curve MsmOps entries): round 1's
// single-thread versions cost 3.5 - 4 ms (Straus), 0.33 - 0.41 ms (each normalisation, Fermat inversion) per proof.
// sums1 = {h_acc, l', T (or s g_a), r g1_b (small proofs only), g_a, g1_b} (G1 xyzz).
// multi-GPU prover: out1[k] = sum over ranks of partials1[rank * n1 + k], k < n1 (G1); out2 = sum of partials2 (G2, n2 = 0 / 1)
int groth16_sum_partials(pcdgpu_ctx* ctx, int pairing, const void* p1, const void* p2, int world, int n1, int n2,
                         void* out1, void* out2) {
  const MsmOps *o1 = msm_ops(pcd_g1_of(pairing)), *o2 = msm_ops(pcd_g2_of(pairing));
  for (int k = 0; k < n1; k++) PCD_TRY(o1->sum_points(ctx, p1, (size_t)k, (size_t)n1, world, out1, (size_t)k, nullptr));
  if (n2) PCD_TRY(o2->sum_points(ctx, p2, 0, 1, world, out2, 0, nullptr));
  return 0;
}

int groth16_straus(pcdgpu_ctx* ctx, int pairing, const u32* d_rs, void* sums1) {
  int ps = ctx->prof_begin(PROF_ASSEMBLE, 1.0);
  // d_rs = r | s (plain): T = [r] g1_b + [s] g_a -> pairs (sums1[5], r), (sums1[4], s); T -> sums1[2]
  int rc = msm_ops(pcd_g1_of(pairing))->multi_mul(ctx, sums1, 5, 4, d_rs, 2, sums1, 2);
  ctx->prof_end(ps);
  return rc;
}
int point_to_affine(pcdgpu_ctx* ctx, int curve, const void* src, size_t idx, void* dst) {
  int ps = ctx->prof_begin(PROF_ASSEMBLE, 1.0);
  int rc = msm_ops(curve)->to_affine_at(ctx, src, idx, dst);
  ctx->prof_end(ps);
  return rc;
}
int groth16_finish(pcdgpu_ctx* ctx, int pairing, const void* sums1, void* d_out_c, int nterms) {
  int ps = ctx->prof_begin(PROF_ASSEMBLE, 1.0);
  // C = h + l' + T (nterms = 3), or h + l' + s g_a + r g1_b with the last two from MSMs of their own (nterms = 4)
  int rc = msm_ops(pcd_g1_of(pairing))->sum_points(ctx, sums1, 0, 1, nterms, nullptr, 0, d_out_c);
  ctx->prof_end(ps);
  return rc;
}

// ---- ark-serialize compressed proof bytes -------------------------------------------------------
template <class F>
__device__ void ser_fp(const F& a, unsigned char* out, unsigned char flags) {
  F v = a.from_mont();
  for (int i = 0; i < 38; i++) out[i] = (unsigned char)(v.l[i >> 2] >> ((i & 3) * 8));
  out[37] |= flags;
}
template <class P>
__device__ unsigned char* ser_x(const Fp<P>& x, unsigned char* out, unsigned char flags) {
  ser_fp(x, out, flags);
  return out + 38;
}
template <class B, u32 NR>
__device__ unsigned char* ser_x(const Fp2T<B, NR>& x, unsigned char* out, unsigned char flags) {
  ser_fp(x.c0, out, 0);
  ser_fp(x.c1, out + 38, flags);
  return out + 76;
}
template <class B, u32 NR>
__device__ unsigned char* ser_x(const Fp3T<B, NR>& x, unsigned char* out, unsigned char flags) {
  ser_fp(x.c0, out, 0);
  ser_fp(x.c1, out + 38, 0);
  ser_fp(x.c2, out + 76, flags);
  return out + 114;
}
template <class F>
__device__ unsigned char* ser_point(const AffinePoint<F>& p, unsigned char* out) {
  if (p.is_inf()) return ser_x(F::zero(), out, 0x40);
  return ser_x(p.x, out, p.y.lexicographically_largest() ? 0x80 : 0);
}
template <class G1, class G2>
__global__ void groth16_serialize_kernel(const void* __restrict__ proof, unsigned char* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  typedef typename G1::F F1;
  typedef typename G2::F F2;
  const char* p = reinterpret_cast<const char*>(proof);
  unsigned char* o = out;
  o = ser_point<F1>(ld_aff<G1>(p, 0), o);
  o = ser_point<F2>(ld_aff<G2>(p + sizeof(AffinePoint<F1>), 0), o);
  o = ser_point<F1>(ld_aff<G1>(p + sizeof(AffinePoint<F1>) + sizeof(AffinePoint<F2>), 0), o);
}
int groth16_serialize(pcdgpu_ctx* ctx, int pairing, const void* d_proof, unsigned char* d_out) {
  if (pairing == PCDGPU_MNT4_298) groth16_serialize_kernel<CurveMnt4G1, CurveMnt4G2><<<1, 32, 0, ctx->stream>>>(d_proof, d_out);
  else groth16_serialize_kernel<CurveMnt6G1, CurveMnt6G2><<<1, 32, 0, ctx->stream>>>(d_proof, d_out);
  PCD_CUDA(ctx, cudaGetLastError());
  return 0;
}

// ---- integer-pipe microbenchmarks -----------------------------------------------------------------
// 8 independent 32x32+64 multiply-add chains per thread (IMAD.WIDE.U32 R, a, b, R): the IMAD roof of
// SURVEY.md 8d.  The multiplier changes with the data every round: with loop-invariant operands ptxas hoists
// the products out of the loop and leaves 64-bit ADDS (the first version of this kernel did exactly that --
// ncu showed the alu pipe 96 % busy and the fma pipe 3 % -- and reported a "17.9 T IMAD/s" roof that does not
// exist).  IMAD.WIDE issues on the fmaheavy pipe only: 32 lanes/clk/SM, with or without a carry predicate.
__global__ void __launch_bounds__(256) bench_imad_kernel(unsigned long long* out, int iters, u32 seed) {
  u32 lo[8], hi[8], a[8];
  u32 b = seed * 3 + blockIdx.x;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    lo[j] = j + threadIdx.x;
    hi[j] = j * 5 + blockIdx.x;
    a[j] = (seed + threadIdx.x) * (2 * j + 1);
  }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int rep = 0; rep < 8; rep++) {
#pragma unroll
      for (int j = 0; j < 8; j++) prims::mac(lo[j], hi[j], a[j], b);
      b = hi[rep];
    }
  }
  u32 s = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) s ^= lo[j] ^ hi[j];
  if (s == 0x1234567u) out[0] = s;
}
// the same multiply-adds written as carry chains (mad.lo.cc / madc.hi.cc -> IMAD.WIDE.U32.X): this is
// the form a multi-limb product needs, and what fp.cuh's operator* is made of.
__global__ void __launch_bounds__(256) bench_imadx_kernel(u32* out, int iters, u32 seed) {
  u32 acc[18];
  u32 a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
#pragma unroll
  for (int j = 0; j < 18; j++) acc[j] = j + threadIdx.x;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int rep = 0; rep < 8; rep++) {
      acc[0] = prims::mad_lo_cc(a, b, acc[0]);
      acc[1] = prims::madc_hi_cc(a, b, acc[1]);
#pragma unroll
      for (int j = 2; j < 16; j += 2) {
        acc[j] = prims::madc_lo_cc(a + j, b, acc[j]);
        acc[j + 1] = prims::madc_hi_cc(a + j, b, acc[j + 1]);
      }
      acc[16] = prims::addc(acc[16], 0);
    }
  }
  u32 s = 0;
#pragma unroll
  for (int j = 0; j < 18; j++) s ^= acc[j];
  if (s == 0x1234567u) out[0] = s;
}
// dependent Montgomery products (the prover's inner loop): 2 independent chains per thread
template <class F>
__global__ void __launch_bounds__(256) bench_modmul_kernel(u32* out, int iters, u32 seed) {
  F x, y;
  const u32* in = out + 64 + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 20;  // per-thread operands
#pragma unroll
  for (int i = 0; i < 10; i++) {
    x.l[i] = in[i];
    y.l[i] = in[10 + i] ^ seed;
  }
  x.l[9] &= 0xff;
  y.l[9] &= 0xff;
  for (int i = 0; i < iters; i++) {
    x = x * y;
    y = y * x;
  }
  if (x.l[0] == 0x12345u && y.l[3] == 7u) out[0] = x.l[1];
}

int bench_imad(pcdgpu_ctx* ctx, int modmul, int iters, double* ops_per_s, double* ms_out) {
  void* d;
  size_t bytes = 4096 + (size_t)ctx->sm_count * 8 * 256 * 80;
  PCD_TRY(ctx->scratch(SLOT_IO, bytes, &d));
  PCD_CUDA(ctx, cudaMemsetAsync(d, 0x5a, bytes, ctx->stream));
  cudaEvent_t e0, e1;
  PCD_CUDA(ctx, cudaEventCreate(&e0));
  PCD_CUDA(ctx, cudaEventCreate(&e1));
  int blocks = ctx->sm_count * 8;
  double per_thread = modmul == 1 || modmul == 2 || modmul == 9 ? 2.0 * iters : 64.0 * iters;
  if (modmul == 9) blocks = ctx->sm_count;  // low occupancy: 8 warps per SM
  for (int rep = 0; rep < 2; rep++) {  // first round warms up
    PCD_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (modmul == 1) bench_modmul_kernel<FpR4><<<blocks, 256, 0, ctx->stream>>>((u32*)d, iters, 7u);
    else if (modmul == 2) bench_modmul_kernel<FpQ4><<<blocks, 256, 0, ctx->stream>>>((u32*)d, iters, 7u);
    else if (modmul == 3) bench_imadx_kernel<<<blocks, 256, 0, ctx->stream>>>((u32*)d, iters, 7u);
    else if (modmul == 9) bench_modmul_kernel<FpQ4><<<blocks, 256, 0, ctx->stream>>>((u32*)d, iters, 7u);
    else bench_imad_kernel<<<blocks, 256, 0, ctx->stream>>>((unsigned long long*)d, iters, 7u);
    PCD_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    PCD_CUDA(ctx, cudaEventSynchronize(e1));
  }
  float ms = 0;
  PCD_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = ms;
  *ops_per_s = per_thread * 256.0 * blocks / (ms * 1e-3);
  return 0;
}
