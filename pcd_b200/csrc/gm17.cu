// gm17.cu -- the GM17 prover on the GPU: R1CS -> SAP witness map and proof assembly.
//
// Replaces ark-gm17 R1CStoSAP::witness_map (r1cs_to_sap.rs) and create_proof (prover.rs) -- SURVEY.md a8 /
// B.7 -- which the reference plugs into ECCyclePCD as MainSNARK / HelpSNARK
// (/root/reference/tests/mnt4_gm17.rs:27-28, tests/mnt4_mix_groth16gm17.rs, tests/mnt4_mix_gm17groth16.rs) and
// reaches through IC::MainSNARK::prove / IC::HelpSNARK::prove (/root/reference/src/ec_cycle_pcd/mod.rs:171,179).
// Same kernels underneath as Groth16: CSR row evaluation, (coset) NTTs on the general domain, the bucket MSM.
//
// SAP of an R1CS with m constraints and ni inputs (the constant included), rows of the square system u^2 = c:
//   2i     : (A_i + B_i)^2 = 4 C_i + e_i          e_i  = (<A_i,z> - <B_i,z>)^2   (one extra variable per constraint)
//   2i + 1 : (A_i - B_i)^2 = e_i
//   2m     : 1^2 = 1
//   2m + 2j - 1 : (z_j + 1)^2 = 4 z_j + e'_j      e'_j = (z_j - 1)^2             (one per public input j >= 1)
//   2m + 2j     : (z_j - 1)^2 = e'_j
// on the domain GeneralEvaluationDomain::new(2m + 2(ni - 1) + 1).  With u, c the interpolants,
//   H = ((u + d1 Z)^2 - (c + d2 Z)) / Z = (u^2 - c)/Z + 2 d1 u + d1^2 Z - d2           (n + 1 coefficients).
#include "groth16.cuh"
#include "msm_ops.cuh"
#include "ntt.cuh"

#define GM17_CHECK_ARG(ctx, cond, msg)      \
  do {                                      \
    if (!(cond)) {                          \
      if (ctx) (ctx)->set_error("%s", msg); \
      return PCDGPU_E_ARG;                  \
    }                                       \
  } while (0)

template <class F>
__device__ __forceinline__ F csr_row_dot(const CsrDev& M, size_t i, const u32* __restrict__ z) {
  const F one = F::one();
  F acc = F::zero();
  u32 lo = M.row_ptr[i], hi = M.row_ptr[i + 1];
  for (u32 k = lo; k < hi; k++) {
    F co = ld10<F>(M.val, k);
    F v = ld10<F>(z, M.col[k]);
    acc = acc + (co == one ? v : co * v);
  }
  return acc;
}

// One thread per SAP row PAIR t: t < m the constraint rows, t = m the row of the constant, m < t < m + ni the
// rows of public input t - m.  Writes the evaluations a, c (natural order) and the extra variables of `full`
// (full[0 .. nv) = z is copied by the caller); a, c are zero beyond row 2m + 2(ni - 1).
template <class F>
__global__ void __launch_bounds__(128) sap_eval_kernel(CsrDev A, CsrDev B, CsrDev C, const u32* __restrict__ z, size_t m,
                                                       size_t ni, size_t nv, u32* __restrict__ a, u32* __restrict__ c,
                                                       u32* __restrict__ full) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m + ni) return;
  const F one = F::one();
  if (t < m) {
    F az = csr_row_dot<F>(A, t, z), bz = csr_row_dot<F>(B, t, z), cz = csr_row_dot<F>(C, t, z);
    F d = az - bz;
    F e = d.sqr();
    st10<F>(full, nv + t, e);
    st10<F>(a, 2 * t, az + bz);
    st10<F>(a, 2 * t + 1, d);
    st10<F>(c, 2 * t, cz.dbl().dbl() + e);
    st10<F>(c, 2 * t + 1, e);
  } else if (t == m) {
    st10<F>(a, 2 * m, one);
    st10<F>(c, 2 * m, one);
  } else {
    size_t j = t - m;  // 1 .. ni - 1
    F x = ld10<F>(z, j);
    F d = x - one;
    F e = d.sqr();
    st10<F>(full, nv + m - 1 + j, e);
    st10<F>(a, 2 * m + 2 * j - 1, x + one);
    st10<F>(a, 2 * m + 2 * j, d);
    st10<F>(c, 2 * m + 2 * j - 1, x.dbl().dbl() + e);
    st10<F>(c, 2 * m + 2 * j, e);
  }
}

// h[i] = 2 d1 a[i] (a = coefficients of u), h[0] -= d2 + d1^2, h[n] = d1^2.  dm = {d1, d2} in Montgomery form.
template <class F>
__global__ void __launch_bounds__(256) sap_h_init_kernel(const u32* __restrict__ a, const u32* __restrict__ dm, size_t n,
                                                         u32* __restrict__ h) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  F d1 = ld10<F>(dm, 0);
  F d1sq = d1.sqr();
  if (i == n) {
    st10<F>(h, n, d1sq);
    return;
  }
  F v = d1.dbl() * ld10<F>(a, i);
  if (i == 0) v = v - ld10<F>(dm, 1) - d1sq;
  st10<F>(h, i, v);
}
// a[i] = (a[i]^2 - c[i]) / Z on the coset
template <class F>
__global__ void __launch_bounds__(256) sap_combine_kernel(u32* a, const u32* __restrict__ c, const u32* __restrict__ zinv,
                                                          size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  F x = ld10<F>(a, i);
  st10<F>(a, i, (x.sqr() - ld10<F>(c, i)) * ld10<F>(zinv, 0));
}
// h[i] += q[i], i < n - 1 (the quotient has degree <= n - 2 when the SAP is satisfied)
template <class F>
__global__ void __launch_bounds__(256) sap_h_add_kernel(u32* h, const u32* __restrict__ q, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 >= n) return;
  st10<F>(h, i, ld10<F>(h, i) + ld10<F>(q, i));
}

// plain d1, d2, r (as they cross the ABI) -> Montgomery d1, d2 and the plain scalars of the constant pairs:
//   extras = { r + d1, 1,   r^2 + 2 r d1, r + d1,   1,   d2 }      dm = { d1 R, d2 R }
template <class SP>
__global__ void gm17_prepare_kernel(const u32* __restrict__ ddr, u32* __restrict__ extras, u32* __restrict__ dm) {
  typedef Fp<SP> Fr;
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Fr d1, d2, r;
#pragma unroll
  for (int i = 0; i < 10; i++) {
    d1.l[i] = ddr[i];
    d2.l[i] = ddr[10 + i];
    r.l[i] = ddr[20 + i];
  }
  Fr d1m = d1.to_mont(), d2m = d2.to_mont(), rm = r.to_mont();
  Fr rd1 = (rm + d1m).from_mont();
  Fr k = (rm.sqr() + rm.dbl() * d1m).from_mont();
#pragma unroll
  for (int i = 0; i < 10; i++) {
    u32 one = i == 0 ? 1u : 0u;
    extras[0 * 10 + i] = rd1.l[i];
    extras[1 * 10 + i] = one;
    extras[2 * 10 + i] = k.l[i];
    extras[3 * 10 + i] = rd1.l[i];
    extras[4 * 10 + i] = one;
    extras[5 * 10 + i] = d2.l[i];
    dm[i] = d1m.l[i];
    dm[10 + i] = d2m.l[i];
  }
}

// The proof's tail, split like Groth16's (groth16.cu): [r] C2' right after the c_query_2 MSM on its lane, A and B
// normalised on their lanes, and after the join only C = C1' + G' + [r] C2'.  sums1 = {G', C1', C2', A} (G1 xyzz).
static int sap_domain(int pairing, size_t m, size_t ni, size_t* n, int* da, int* db) {
  int field = pairing == PCDGPU_MNT4_298 ? PCDGPU_FIELD_R4 : PCDGPU_FIELD_Q4;
  return ntt_domain_shape(field, 2 * m + 2 * (ni - 1) + 1, n, da, db);
}

// d_full: nsap elements, d_h: n + 1 coefficients (both Montgomery, in the context's scratch).  d_dm: {d1, d2}.
// `fork_after_eval`, when set, is called once the SAP assignment is complete (before the NTTs).
template <class F, class Fn>
static int sap_witness_map_t(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, const void* d_z, const u32* d_dm, void** d_full,
                             void** d_h, size_t* n_out, Fn fork_after_eval) {
  int field = r->pairing == PCDGPU_MNT4_298 ? PCDGPU_FIELD_R4 : PCDGPU_FIELD_Q4;
  size_t n;
  int da, db;
  if (sap_domain(r->pairing, r->m, r->num_inputs, &n, &da, &db) != 0) {
    ctx->set_error("SAP witness map needs a domain of %zu elements; the field has none that large",
                   2 * r->m + 2 * (r->num_inputs - 1) + 1);
    return PCDGPU_E_DOMAIN;
  }
  const size_t m = r->m, ni = r->num_inputs, nv = ni + r->num_witness, nsap = nv + m + ni - 1;
  void *a, *c, *h, *full;
  PCD_TRY(ctx->scratch(SLOT_WM_A, n * 40, &a));
  PCD_TRY(ctx->scratch(SLOT_WM_B, n * 40, &c));
  PCD_TRY(ctx->scratch(SLOT_WM_C, (n + 1) * 40, &h));
  PCD_TRY(ctx->scratch(SLOT_SAP_FULL, nsap * 40, &full));
  cudaStream_t st = ctx->stream;
  int ps = ctx->prof_begin(PROF_SPMV, (double)m * 3);
  const size_t rows = 2 * m + 2 * (ni - 1) + 1;
  PCD_CUDA(ctx, cudaMemsetAsync((char*)a + rows * 40, 0, (n - rows) * 40, st));
  PCD_CUDA(ctx, cudaMemsetAsync((char*)c + rows * 40, 0, (n - rows) * 40, st));
  PCD_CUDA(ctx, cudaMemcpyAsync(full, d_z, nv * 40, cudaMemcpyDeviceToDevice, st));
  sap_eval_kernel<F><<<(unsigned)((m + ni + 127) / 128), 128, 0, st>>>(r->A, r->B, r->C, (const u32*)d_z, m, ni, nv,
                                                                       (u32*)a, (u32*)c, (u32*)full);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->prof_end(ps);
  ctx->launches += 5;
  PCD_TRY(fork_after_eval(full));
  PCD_TRY(ntt_run_general(ctx, field, a, da, db, 1, 0));
  sap_h_init_kernel<F><<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>((const u32*)a, d_dm, n, (u32*)h);
  PCD_CUDA(ctx, cudaGetLastError());
  PCD_TRY(ntt_run_general(ctx, field, a, da, db, 0, 1));
  PCD_TRY(ntt_run_general(ctx, field, c, da, db, 1, 0));
  PCD_TRY(ntt_run_general(ctx, field, c, da, db, 0, 1));
  const u32* zinv;
  PCD_TRY(ntt_zinv_general(ctx, field, n, &zinv));
  sap_combine_kernel<F><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((u32*)a, (const u32*)c, zinv, n);
  PCD_CUDA(ctx, cudaGetLastError());
  PCD_TRY(ntt_run_general(ctx, field, a, da, db, 1, 1));
  sap_h_add_kernel<F><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((u32*)h, (const u32*)a, n);
  PCD_CUDA(ctx, cudaGetLastError());
  *d_full = full;
  *d_h = h;
  *n_out = n;
  return 0;
}

template <class Fn>
static int sap_witness_map_dev(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, const void* d_z, const u32* d_dm, void** d_full,
                               void** d_h, size_t* n_out, Fn fork_after_eval) {
  if (r->pairing == PCDGPU_MNT4_298) return sap_witness_map_t<FpR4>(ctx, r, d_z, d_dm, d_full, d_h, n_out, fork_after_eval);
  return sap_witness_map_t<FpQ4>(ctx, r, d_z, d_dm, d_full, d_h, n_out, fork_after_eval);
}

static int gm17_prepare(pcdgpu_ctx* ctx, int pairing, const u32* d_ddr, u32* d_extras, u32* d_dm) {
  ctx->launches += 1;
  if (pairing == PCDGPU_MNT4_298) gm17_prepare_kernel<ParamsR4><<<1, 32, 0, ctx->stream>>>(d_ddr, d_extras, d_dm);
  else gm17_prepare_kernel<ParamsQ4><<<1, 32, 0, ctx->stream>>>(d_ddr, d_extras, d_dm);
  PCD_CUDA(ctx, cudaGetLastError());
  return 0;
}

extern "C" {

size_t pcdgpu_sap_domain_size(int pairing, size_t m, size_t num_inputs) {
  size_t n;
  int da, db;
  if (num_inputs < 1 || (pairing != PCDGPU_MNT4_298 && pairing != PCDGPU_MNT6_298)) return 0;
  return sap_domain(pairing, m, num_inputs, &n, &da, &db) == 0 ? n : 0;
}

int pcdgpu_sap_witness_map(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, const void* z, const void* d1, const void* d2,
                           void* full, void* h) {
  if (!ctx) return PCDGPU_E_ARG;
  GM17_CHECK_ARG(ctx, r && z && d1 && d2 && full && h, "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t nv = r->num_inputs + r->num_witness, nsap = nv + r->m + r->num_inputs - 1;
  void *dz, *misc;
  PCD_TRY(ctx->scratch(SLOT_Z, nv * 40, &dz));
  PCD_TRY(ctx->scratch(SLOT_MISC, 8192, &misc));
  u32* d_ddr = (u32*)misc;
  u32* extras = d_ddr + 32;
  u32* d_dm = extras + 64;
  memcpy(ctx->pinned, d1, 40);
  memcpy((char*)ctx->pinned + 40, d2, 40);
  memset((char*)ctx->pinned + 80, 0, 40);
  PCD_CUDA(ctx, cudaMemcpyAsync(d_ddr, ctx->pinned, 120, cudaMemcpyHostToDevice, ctx->stream));
  PCD_CUDA(ctx, cudaMemcpyAsync(dz, z, nv * 40, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(gm17_prepare(ctx, r->pairing, d_ddr, extras, d_dm));
  void *dfull, *dh;
  size_t n;
  PCD_TRY(sap_witness_map_dev(ctx, r, dz, d_dm, &dfull, &dh, &n, [](void*) { return 0; }));
  PCD_CUDA(ctx, cudaMemcpyAsync(full, dfull, nsap * 40, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaMemcpyAsync(h, dh, (n + 1) * 40, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int pcdgpu_gm17_pk_upload(pcdgpu_ctx* ctx, int pairing, size_t num_sap_vars, size_t num_inputs, size_t h_len,
                          const void* a_query, const void* b_query, const void* c_query_1, const void* c_query_2,
                          const void* g_gamma2_z_t, const void* g_gamma_z, const void* h_gamma_z,
                          const void* g_ab_gamma_z, const void* g_gamma2_z2, int precompute, pcdgpu_gm17_pk** out) {
  if (!ctx) return PCDGPU_E_ARG;
  GM17_CHECK_ARG(ctx, pairing == PCDGPU_MNT4_298 || pairing == PCDGPU_MNT6_298, "unknown pairing id");
  GM17_CHECK_ARG(ctx, out && a_query && b_query && c_query_2 && g_gamma2_z_t && g_gamma_z && h_gamma_z && g_ab_gamma_z &&
                          g_gamma2_z2, "null pointer");
  GM17_CHECK_ARG(ctx, num_inputs >= 1 && num_sap_vars >= num_inputs && h_len >= 1, "bad counts");
  GM17_CHECK_ARG(ctx, num_sap_vars == num_inputs || c_query_1, "null query");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  int g1 = pcd_g1_of(pairing), g2 = pcd_g2_of(pairing);
  pcdgpu_gm17_pk* pk = new pcdgpu_gm17_pk();
  memset(pk, 0, sizeof(*pk));
  pk->ctx = ctx;
  pk->pairing = pairing;
  pk->num_sap_vars = num_sap_vars;
  pk->num_inputs = num_inputs;
  pk->h_len = h_len;
  // ark-gm17 adds the constant-variable elements and the r / d1 / d2 multiples of single key points outside its
  // MSMs (prover.rs); here they ride inside as extra (point, scalar) pairs, exactly like pcdgpu_pk_upload:
  //   a_query      : a[1..], g_gamma_z, a[0]                    scalars full[1..], r + d1, 1
  //   b_query      : b[1..], h_gamma_z, b[0]                    scalars full[1..], r + d1, 1
  //   c_query_1    : c1[..], g_gamma2_z2, g_ab_gamma_z          scalars full[ni..], r^2 + 2 r d1, r + d1
  //   c_query_2    : c2[1..], c2[0]                             scalars full[1..], 1          (then times r)
  //   g_gamma2_z_t : t[..], t[0]                                scalars h[..], d2
  auto upload_ext = [&](int curve, const void* q, size_t skip, size_t n, const void* const* extra, int n_extra,
                        pcdgpu_bases** o) -> int {
    size_t pb = msm_ops(curve)->affine_bytes;
    std::vector<char> buf((n + n_extra) * pb);
    if (n) memcpy(buf.data(), (const char*)q + skip * pb, n * pb);
    for (int i = 0; i < n_extra; i++) memcpy(buf.data() + (n + i) * pb, extra[i], pb);
    return pcdgpu_bases_upload(ctx, curve, buf.data(), n + n_extra, precompute, o);
  };
  const void* ea[2] = {g_gamma_z, a_query};
  const void* eb[2] = {h_gamma_z, b_query};
  const void* ec1[2] = {g_gamma2_z2, g_ab_gamma_z};
  const void* ec2[1] = {c_query_2};
  const void* eg[1] = {g_gamma2_z_t};
  int rc = 0;
  ctx->key_upload = true;  // window rule of concurrent MSMs (pcdgpu_bases_upload)
  rc = rc ? rc : upload_ext(g1, a_query, 1, num_sap_vars - 1, ea, 2, &pk->a_query);
  rc = rc ? rc : upload_ext(g2, b_query, 1, num_sap_vars - 1, eb, 2, &pk->b_query);
  rc = rc ? rc : upload_ext(g1, c_query_1, 0, num_sap_vars - num_inputs, ec1, 2, &pk->c_query_1);
  rc = rc ? rc : upload_ext(g1, c_query_2, 1, num_sap_vars - 1, ec2, 1, &pk->c_query_2);
  rc = rc ? rc : upload_ext(g1, g_gamma2_z_t, 0, h_len, eg, 1, &pk->g_gamma2_z_t);
  ctx->key_upload = false;
  if (rc) {
    pcdgpu_gm17_pk_free(pk);
    return rc;
  }
  *out = pk;
  return 0;
}

void pcdgpu_gm17_pk_free(pcdgpu_gm17_pk* pk) {
  if (!pk) return;
  pcdgpu_bases_free(pk->a_query);
  pcdgpu_bases_free(pk->b_query);
  pcdgpu_bases_free(pk->c_query_1);
  pcdgpu_bases_free(pk->c_query_2);
  pcdgpu_bases_free(pk->g_gamma2_z_t);
  delete pk;
}

int pcdgpu_gm17_prove_dev(pcdgpu_ctx* ctx, const pcdgpu_gm17_pk* pk, const pcdgpu_r1cs* r1cs, const void* d_z,
                          const void* d1, const void* d2, const void* r, void* out_proof) {
  if (!ctx) return PCDGPU_E_ARG;
  GM17_CHECK_ARG(ctx, pk && r1cs && d_z && d1 && d2 && r && out_proof, "null pointer");
  GM17_CHECK_ARG(ctx, pk->pairing == r1cs->pairing, "key and constraint system are over different pairings");
  const size_t ni = r1cs->num_inputs, nv = ni + r1cs->num_witness, nsap = nv + r1cs->m + ni - 1;
  GM17_CHECK_ARG(ctx, pk->num_sap_vars == nsap && pk->num_inputs == ni,
                 "key and constraint system disagree on the variable counts");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  InProofGuard in_proof(ctx);
  GM17_CHECK_ARG(ctx, in_proof.ok, "the context is running a prover call on another host thread (one context per thread)");
  int g1 = pcd_g1_of(pk->pairing), g2 = pcd_g2_of(pk->pairing);
  const MsmOps *o1 = msm_ops(g1), *o2 = msm_ops(g2);
  size_t x1 = o1->xyzz_bytes, x2 = o2->xyzz_bytes;
  // misc layout: d1 d2 r (120 B) | extras: 6 scalars | dm: 2 elements | sums1: G', C1', C2', A | sum2: B | proof
  void* misc;
  PCD_TRY(ctx->scratch(SLOT_MISC, 8192, &misc));
  char* mb = (char*)misc;
  u32* d_ddr = (u32*)mb;
  char* extras = mb + 128;
  u32* d_dm = (u32*)(mb + 384);
  void* sums1 = mb + 512;
  void* sum2 = (char*)sums1 + 4 * x1;
  char* d_proof = (char*)sum2 + x2;
  size_t proof_bytes = 2 * o1->affine_bytes + o2->affine_bytes;
  char* d_A = d_proof;
  char* d_B = d_proof + o1->affine_bytes;
  char* d_C = d_proof + o1->affine_bytes + o2->affine_bytes;
  memcpy(ctx->pinned, d1, 40);
  memcpy((char*)ctx->pinned + 40, d2, 40);
  memcpy((char*)ctx->pinned + 80, r, 40);
  PCD_CUDA(ctx, cudaMemcpyAsync(d_ddr, ctx->pinned, 120, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(gm17_prepare(ctx, pk->pairing, d_ddr, (u32*)extras, d_dm));
  const bool fork = ctx->concurrent;
  // The four MSMs over the SAP assignment start (lanes 1-4) as soon as the extra variables exist; lane 0 goes
  // on with the five NTTs and the MSM over H.
  auto start_msms = [&](void* d_full) -> int {
    const char* f = (const char*)d_full;
    if (fork) {
      PCD_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
      for (int l = 1; l < pcdgpu_ctx::NLANE_PROOF; l++) PCD_CUDA(ctx, cudaStreamWaitEvent(ctx->lane_stream[l], ctx->ev_fork, 0));
    }
    struct Job { const pcdgpu_bases* b; const char* sc; size_t n; const char* ex; size_t nex; void* out; };
    Job jobs[4] = {{pk->b_query, f + 40, nsap - 1, extras, 2, sum2},
                   {pk->a_query, f + 40, nsap - 1, extras, 2, (char*)sums1 + 3 * x1},
                   {pk->c_query_2, f + 40, nsap - 1, extras + 4 * 40, 1, (char*)sums1 + 2 * x1},
                   {pk->c_query_1, f + 40 * ni, nsap - ni, extras + 2 * 40, 2, (char*)sums1 + 1 * x1}};
    int rc = 0;
    for (int j = 0; j < 4 && rc == 0; j++) {
      ctx->lane = fork ? j + 1 : 0;
      rc = bases_msm(ctx, jobs[j].b, 0, jobs[j].sc, 1, jobs[j].n, jobs[j].ex, jobs[j].nex, jobs[j].out);
      if (rc == 0 && j == 0) rc = point_to_affine(ctx, g2, sum2, 0, d_B);
      if (rc == 0 && j == 1) rc = point_to_affine(ctx, g1, sums1, 3, d_A);
      if (rc == 0 && j == 2) {  // [r] C2' on the lane that produced C2'
        int ps = ctx->prof_begin(PROF_ASSEMBLE, 1.0);
        rc = o1->multi_mul(ctx, sums1, 2, 2, d_ddr + 20, 1, sums1, 2);  // lane-cooperative double-and-add (wec.cuh)
        ctx->prof_end(ps);
      }
      if (fork && rc == 0 && cudaEventRecord(ctx->ev_join[j + 1], ctx->lane_stream[j + 1]) != cudaSuccess) rc = PCDGPU_E_CUDA;
    }
    ctx->lane = 0;
    return rc;
  };
  void *d_full, *d_h;
  size_t n = 0;
  int rc = sap_witness_map_dev(ctx, r1cs, d_z, d_dm, &d_full, &d_h, &n, start_msms);
  if (rc == 0 && pk->h_len != n + 1) {
    ctx->set_error("key's g_gamma2_z_t length is not the SAP domain size + 1");
    rc = PCDGPU_E_ARG;
  }
  if (rc == 0) rc = bases_msm(ctx, pk->g_gamma2_z_t, 0, d_h, 1, n + 1, extras + 5 * 40, 1, (char*)sums1 + 0 * x1);
  if (rc) {
    ctx->drain_lanes();
    return rc;
  }
  if (fork)
    for (int l = 1; l < pcdgpu_ctx::NLANE_PROOF; l++) PCD_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join[l], 0));
  int ps = ctx->prof_begin(PROF_ASSEMBLE, 1.0);
  PCD_TRY(o1->sum_points(ctx, sums1, 0, 1, 3, nullptr, 0, d_C));  // C = G' + C1' + [r] C2'
  ctx->prof_end(ps);
  PCD_CUDA(ctx, cudaMemcpyAsync(out_proof, d_proof, proof_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int pcdgpu_gm17_prove(pcdgpu_ctx* ctx, const pcdgpu_gm17_pk* pk, const pcdgpu_r1cs* r1cs, const void* z, const void* d1,
                      const void* d2, const void* r, void* out_proof) {
  if (!ctx) return PCDGPU_E_ARG;
  GM17_CHECK_ARG(ctx, pk && r1cs && z && d1 && d2 && r && out_proof, "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t nv = r1cs->num_inputs + r1cs->num_witness;
  void* dz;
  PCD_TRY(ctx->scratch(SLOT_Z, nv * 40, &dz));
  PCD_CUDA(ctx, cudaMemcpyAsync(dz, z, nv * 40, cudaMemcpyHostToDevice, ctx->stream));
  return pcdgpu_gm17_prove_dev(ctx, pk, r1cs, dz, d1, d2, r, out_proof);
}

}  // extern "C"
