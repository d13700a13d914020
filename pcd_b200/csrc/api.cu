// api.cu -- the extern "C" surface of libpcdgpu.so (include/pcdgpu.h).  Host-pointer entry points
// stage through device scratch; _dev entry points are asynchronous on the context's stream.
#include <atomic>
#include <cstdlib>

#include "groth16.cuh"
#include "msm_ops.cuh"
#include "ntt.cuh"

// ---- small helpers -----------------------------------------------------------------------------
const MsmOps* msm_ops(int curve) {
  switch (curve) {
    case PCDGPU_MNT4_G1: return &MSM_OPS_MNT4_G1;
    case PCDGPU_MNT4_G2: return &MSM_OPS_MNT4_G2;
    case PCDGPU_MNT6_G1: return &MSM_OPS_MNT6_G1;
    case PCDGPU_MNT6_G2: return &MSM_OPS_MNT6_G2;
  }
  return nullptr;
}
int pcd_g1_of(int pairing) { return pairing == PCDGPU_MNT4_298 ? PCDGPU_MNT4_G1 : PCDGPU_MNT6_G1; }
int pcd_g2_of(int pairing) { return pairing == PCDGPU_MNT4_298 ? PCDGPU_MNT4_G2 : PCDGPU_MNT6_G2; }
static int g1_of(int pairing) { return pcd_g1_of(pairing); }
static int g2_of(int pairing) { return pcd_g2_of(pairing); }

static const int MSM_BITS = 298;
int msm_num_windows_c(int c) { return (MSM_BITS + 1 + c - 1) / c; }
int msm_auto_window_c(size_t n, int shared) {
  // lg = round(log2 n): a key query of 2^20 + 3 points must not be treated as 2^21
  int lg = 0;
  while (((size_t)3 << lg) < 2 * (n ? n : 1)) lg++;  // smallest lg with 1.5 * 2^lg >= n
  int c;
  if (shared) {
    // One bucket set for all windows (balanced widths, msm_win_start).  Accumulation costs ~ nwin * n mixed
    // additions and needs >= ~2 * 10^5 buckets to give every SM its threads; the reduction costs ~ 2^(c-1)
    // buckets at ~14 x a mixed addition.  Measured optimum (tools/probe_msm.py): c = lg up to 2^18, 19 at 2^20.
    c = lg <= 18 ? lg : lg - 1;
    // ~10^3 points (the default-circuit proofs of a PCD step): the walk of one thread per bucket has nothing to
    // parallelise over, so use FEW buckets and let every bucket go down the chunked (CTA tree sum) path: measured on
    // the whole 2^10 proof c = 10: 1.78 / 2.10 ms (MNT4 / MNT6), 7: 1.37 / 1.70, 6: 1.32 / 1.79, 4: 1.31 / 2.11; measured
    // again with lane-cooperative heavy-bucket trees and two buckets per reduction group (inside the PCD step): c = 9:
    // 1.32 / 1.47, 8: 1.36 / 1.53, 7: 0.92 / 1.10, 6: 0.88 / 1.03, 5: 0.89 / 1.04
    if (lg <= 12) c = 6;
    if (c < 6) c = 6;
    if (c > 21) c = 21;
  } else {
    c = lg - 4;  // ~32 entries per bucket and window
    if (c < 4) c = 4;
    if (c > 16) c = 16;
  }
  return c;
}

static MsmPlanC plan_plain(pcdgpu_ctx* ctx, size_t n) {
  int c = ctx->msm_window > 0 ? ctx->msm_window : msm_auto_window_c(n, 0);
  return MsmPlanC{c, msm_num_windows_c(c), 0, 0, 0};
}

#define CHECK_ARG(ctx, cond, msg)         \
  do {                                    \
    if (!(cond)) {                        \
      if (ctx) (ctx)->set_error("%s", msg); \
      return PCDGPU_E_ARG;                \
    }                                     \
  } while (0)

// identities of uploaded keys and constraint systems (pcdgpu_ctx::ProofGraph: an address can be reused, a uid cannot)
static unsigned long long next_uid() {
  static std::atomic<unsigned long long> n{0};
  return ++n;
}

extern "C" {

const char* pcdgpu_strerror(int code) {
  switch (code) {
    case PCDGPU_OK: return "ok";
    case PCDGPU_E_ARG: return "bad argument";
    case PCDGPU_E_NODEVICE: return "no usable CUDA device (libpcdgpu has no CPU fallback)";
    case PCDGPU_E_CUDA: return "CUDA call failed";
    case PCDGPU_E_DOMAIN: return "evaluation domain exceeds the field's 2-adicity";
    case PCDGPU_E_NOMEM: return "out of device memory";
  }
  return "unknown error";
}
const char* pcdgpu_last_error(const pcdgpu_ctx* ctx) { return ctx ? ctx->err : ""; }
size_t pcdgpu_affine_bytes(int curve) {
  const MsmOps* o = msm_ops(curve);
  return o ? o->affine_bytes : 0;
}

int pcdgpu_ctx_create(int device, pcdgpu_ctx** out) {
  if (!out) return PCDGPU_E_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return PCDGPU_E_NODEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PCDGPU_E_NODEVICE;
  if (prop.major != 10) return PCDGPU_E_NODEVICE;  // the only code in the library is sm_100a SASS
  if (cudaSetDevice(device) != cudaSuccess) return PCDGPU_E_NODEVICE;
  pcdgpu_ctx* ctx = new pcdgpu_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  // Stream priorities: lane 0 carries the witness map and the MSM over h -- a chain of short kernels (CSR product,
  // NTT passes) that the other lanes' accumulation kernels would otherwise starve (measured at 2^18: seven NTTs that
  // take 0.1 ms each alone stretched over 6 ms, delaying the h MSM to the very end of the proof); lanes 2 and 3 (a,
  // b_g1) feed the double-scalar multiplication and come next; the G2 and l MSMs fill the rest.
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  static const bool no_prio = getenv("PCDGPU_NO_PRIORITIES") != nullptr;  // development aid (A/B runs)
  if (no_prio) prio_greatest = prio_least;
  // (numerically smaller = more urgent) lane 0 > lanes 2, 3, 5, 6 (a, b_g1 -- the double-scalar multiplication waits for
  // them -- and the extra MSMs of small proofs) > lane 1 (the G2 MSM) > lane 4 (l)
  auto prio_at = [&](int k) { return prio_greatest + k < prio_least ? prio_greatest + k : prio_least; };
  // every partial failure goes through pcdgpu_ctx_destroy, which releases whatever was created so far
  int rc = PCDGPU_OK;
  if (cudaStreamCreateWithPriority(&ctx->own_stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) rc = PCDGPU_E_CUDA;
  ctx->stream = ctx->own_stream;
  if (rc == 0 && cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess) rc = PCDGPU_E_CUDA;
  for (int l = 1; l < pcdgpu_ctx::NLANE && rc == 0; l++)
    if (cudaStreamCreateWithPriority(&ctx->lane_stream[l], cudaStreamNonBlocking,
                                     (l == 2 || l == 3 || l >= 5) ? prio_at(1) : (l == 1 ? prio_at(2) : prio_least)) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join[l], cudaEventDisableTiming) != cudaSuccess)
      rc = PCDGPU_E_CUDA;
  for (int i = 0; i < 3 && rc == 0; i++)
    if (cudaEventCreateWithFlags(&ctx->ev_acc[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_sorted[i], cudaEventDisableTiming) != cudaSuccess)
      rc = PCDGPU_E_CUDA;
  ctx->pinned_bytes = 1 << 16;
  if (rc == 0 && cudaMallocHost(&ctx->pinned, ctx->pinned_bytes) != cudaSuccess) rc = PCDGPU_E_NOMEM;
  if (rc) {
    pcdgpu_ctx_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return PCDGPU_OK;
}

void pcdgpu_ctx_destroy(pcdgpu_ctx* ctx) {
  if (!ctx) return;
  pcdgpu_comm_destroy(ctx);
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < pcdgpu_ctx::NSLOT; i++)
    if (ctx->slot[i]) cudaFree(ctx->slot[i]);
  for (int l = 1; l < pcdgpu_ctx::NLANE; l++) {
    if (ctx->lane_stream[l]) cudaStreamDestroy(ctx->lane_stream[l]);
    if (ctx->ev_join[l]) cudaEventDestroy(ctx->ev_join[l]);
  }
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  for (auto& g : ctx->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  for (int i = 0; i < 3; i++) {
    if (ctx->ev_acc[i]) cudaEventDestroy(ctx->ev_acc[i]);
    if (ctx->ev_sorted[i]) cudaEventDestroy(ctx->ev_sorted[i]);
  }
  for (auto& kv : ctx->ntt_tables) cudaFree(kv.second.twiddles);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->prof_pinned) cudaFreeHost(ctx->prof_pinned);
  for (auto& sp : ctx->spans) {
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
}

int pcdgpu_sync(pcdgpu_ctx* ctx) {
  if (!ctx) return PCDGPU_E_ARG;
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}
int pcdgpu_set_stream(pcdgpu_ctx* ctx, void* stream) {
  if (!ctx) return PCDGPU_E_ARG;
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->stream = stream ? (cudaStream_t)stream : ctx->own_stream;
  return 0;
}
int pcdgpu_set_concurrency(pcdgpu_ctx* ctx, int on) {
  if (!ctx) return PCDGPU_E_ARG;
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->concurrent = on != 0;
  return 0;
}
int pcdgpu_set_proof_graphs(pcdgpu_ctx* ctx, int on) {
  if (!ctx) return PCDGPU_E_ARG;
  ctx->use_graphs = on != 0;
  return 0;
}
int pcdgpu_proof_graph_stats(pcdgpu_ctx* ctx, uint64_t* captured, uint64_t* replayed) {
  if (!ctx) return PCDGPU_E_ARG;
  if (captured) *captured = ctx->graphs_captured;
  if (replayed) *replayed = ctx->graphs_replayed;
  return 0;
}
int pcdgpu_set_msm_window(pcdgpu_ctx* ctx, int c) {
  if (!ctx || c < 0 || c > 21 || c == 1) return PCDGPU_E_ARG;
  ctx->msm_window = c;
  return 0;
}

// ---- NTT ---------------------------------------------------------------------------------------
int pcdgpu_ntt_dev(pcdgpu_ctx* ctx, int field, void* d_data, uint32_t log_n, int inverse, int coset) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, d_data, "null data pointer");
  CHECK_ARG(ctx, field == PCDGPU_FIELD_R4 || field == PCDGPU_FIELD_Q4, "unknown field id");
  CHECK_ARG(ctx, log_n < 40, "log_n out of range");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  return ntt_run(ctx, field, d_data, (int)log_n, inverse != 0, coset != 0);
}

size_t pcdgpu_domain_size(int field, size_t min_size, int* pow7, int* pow2) {
  size_t n;
  int a, b;
  if ((field != PCDGPU_FIELD_R4 && field != PCDGPU_FIELD_Q4) || ntt_domain_shape(field, min_size, &n, &a, &b) != 0) return 0;
  if (pow7) *pow7 = a;
  if (pow2) *pow2 = b;
  return n;
}

int pcdgpu_ntt_general(pcdgpu_ctx* ctx, int field, void* data, int pow7, int pow2, int inverse, int coset) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, data, "null data pointer");
  CHECK_ARG(ctx, field == PCDGPU_FIELD_R4 || field == PCDGPU_FIELD_Q4, "unknown field id");
  CHECK_ARG(ctx, pow7 >= 0 && pow7 <= 2 && pow2 >= 0 && pow2 < 40, "domain exponents out of range");
  if (pow7 == 0) return pcdgpu_ntt(ctx, field, data, (uint32_t)pow2, inverse, coset);
  if (field != PCDGPU_FIELD_Q4 || pow2 > 17) {
    ctx->set_error("no evaluation domain of size 7^%d 2^%d on this field", pow7, pow2);
    return PCDGPU_E_DOMAIN;
  }
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t bytes = ((size_t)(pow7 == 1 ? 7 : 49) << pow2) * 40;
  void* d;
  PCD_TRY(ctx->scratch(SLOT_IO, bytes, &d));
  PCD_CUDA(ctx, cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(ntt_run_general(ctx, field, d, pow7, pow2, inverse != 0, coset != 0));
  PCD_CUDA(ctx, cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int pcdgpu_ntt(pcdgpu_ctx* ctx, int field, void* data, uint32_t log_n, int inverse, int coset) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, data, "null data pointer");
  CHECK_ARG(ctx, field == PCDGPU_FIELD_R4 || field == PCDGPU_FIELD_Q4, "unknown field id");
  CHECK_ARG(ctx, log_n < 40, "log_n out of range");
  int two_adicity = field == PCDGPU_FIELD_R4 ? 34 : 17;
  if ((int)log_n > two_adicity) {
    ctx->set_error("radix-2 domain 2^%u exceeds the field's 2-adicity %d", log_n, two_adicity);
    return PCDGPU_E_DOMAIN;
  }
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t bytes = ((size_t)40) << log_n;
  void* d;
  PCD_TRY(ctx->scratch(SLOT_IO, bytes, &d));
  PCD_CUDA(ctx, cudaMemcpyAsync(d, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(ntt_run(ctx, field, d, (int)log_n, inverse != 0, coset != 0));
  PCD_CUDA(ctx, cudaMemcpyAsync(data, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---- MSM ---------------------------------------------------------------------------------------
int pcdgpu_msm_dev(pcdgpu_ctx* ctx, int curve, const void* d_bases, const void* d_scalars, int scalars_mont, size_t n,
                   void* d_out_xyzz) {
  if (!ctx) return PCDGPU_E_ARG;
  const MsmOps* ops = msm_ops(curve);
  CHECK_ARG(ctx, ops, "unknown curve id");
  CHECK_ARG(ctx, d_out_xyzz && (n == 0 || (d_bases && d_scalars)), "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  return ops->run(ctx, d_bases, d_scalars, scalars_mont, n, nullptr, 0, plan_plain(ctx, n), d_out_xyzz);
}

static int finish_to_host(pcdgpu_ctx* ctx, const MsmOps* ops, void* d_xyzz, void* out_affine) {
  void* d_aff = (char*)d_xyzz + ops->xyzz_bytes;
  PCD_TRY(ops->to_affine_at(ctx, d_xyzz, 0, d_aff));
  PCD_CUDA(ctx, cudaMemcpyAsync(out_affine, d_aff, ops->affine_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int pcdgpu_msm(pcdgpu_ctx* ctx, int curve, const void* bases, const void* scalars, size_t n, void* out_affine) {
  if (!ctx) return PCDGPU_E_ARG;
  const MsmOps* ops = msm_ops(curve);
  CHECK_ARG(ctx, ops, "unknown curve id");
  CHECK_ARG(ctx, out_affine && (n == 0 || (bases && scalars)), "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  void *db, *ds, *dres;
  PCD_TRY(ctx->scratch(SLOT_IO, n * ops->affine_bytes + 16, &db));
  PCD_TRY(ctx->scratch(SLOT_IO2, n * 40 + 16, &ds));
  PCD_TRY(ctx->scratch(SLOT_MISC, 4096, &dres));
  PCD_CUDA(ctx, cudaMemcpyAsync(db, bases, n * ops->affine_bytes, cudaMemcpyHostToDevice, ctx->stream));
  PCD_CUDA(ctx, cudaMemcpyAsync(ds, scalars, n * 40, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(ops->run(ctx, db, ds, 0, n, nullptr, 0, plan_plain(ctx, n), dres));
  return finish_to_host(ctx, ops, dres, out_affine);
}

int pcdgpu_bases_upload(pcdgpu_ctx* ctx, int curve, const void* bases, size_t n, int precompute, pcdgpu_bases** out) {
  if (!ctx) return PCDGPU_E_ARG;
  const MsmOps* ops = msm_ops(curve);
  CHECK_ARG(ctx, ops, "unknown curve id");
  CHECK_ARG(ctx, out && (n == 0 || bases), "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  pcdgpu_bases* b = new pcdgpu_bases();
  b->ctx = ctx;
  b->curve = curve;
  b->n = n;
  b->c = 0;
  b->nwin = 0;
  b->points = nullptr;
  b->table = nullptr;
  size_t rows = 1;
  if (precompute && n > 0) {
    // key queries (ctx->key_upload): the five MSMs of a proof run side by side, so the latency-bound bucket reduction
    // (cost ~ 2^(c-1) buckets) competes with the other lanes' accumulation (cost ~ 299 / c additions per point) and a
    // smaller window than a lone MSM's wins.  Measured with the whole proof (tools/probe_pcd.py, round 2, lane-cooperative
    // reduction): 2^18 main proof c = 15: 8.41 ms, 16: 8.34, 17: 8.90, 18: 10.1; 2^16 helper proof (G2 over Fq3) c = 14:
    // 9.4 ms, 15: 6.27, 16: 7.50, 17: 8.4 (below c = 15 the accumulation runs out of buckets to spread over the SMs).
    b->c = ctx->msm_window > 0 ? ctx->msm_window : msm_auto_window_c(n, 1);
    // Measured again at the end of round 2 inside the PCD step (unit buckets, lane-cooperative heavy-bucket trees;
    // tools/probe_step.py with PCD_WINDOW_*): main 2^18 c = 13: 14.7 ms, 14: 9.9, 15: 6.25 - 6.45, 16: 6.57 - 6.68, 17: 6.99;
    // helper 2^16 c = 13: 7.2, 14: 3.87, 15: 3.88 - 3.96, 16: 4.90.  Hence 15 for both (below it the cliff: too few buckets).
    if (ctx->msm_window <= 0 && ctx->key_upload) {
      if (b->c == 18 && n < ((size_t)3 << 17)) b->c = 15;  // ~2^18 points
      else if (b->c >= 18) b->c -= 2;
      else if (b->c >= 16) b->c -= 1;
    }
    b->nwin = msm_num_windows_c(b->c);
    rows = b->nwin;
  }
  cudaError_t e = cudaMalloc(&b->points, rows * (n ? n : 1) * ops->affine_bytes);
  if (e != cudaSuccess) {
    ctx->set_error("cudaMalloc for %zu x %zu base points: %s", rows, n, cudaGetErrorString(e));
    delete b;
    return PCDGPU_E_NOMEM;
  }
  if (n) {
    e = cudaMemcpyAsync(b->points, bases, n * ops->affine_bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && b->nwin) {
      b->table = b->points;  // row 0 of the table is the points themselves
      int rc = ops->precompute(ctx, b->points, n, b->c, b->nwin, b->table);
      if (rc) {
        cudaFree(b->points);
        delete b;
        return rc;
      }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      ctx->set_error("uploading base points: %s", cudaGetErrorString(e));
      cudaFree(b->points);
      delete b;
      return PCDGPU_E_CUDA;
    }
  }
  *out = b;
  return 0;
}

void pcdgpu_bases_free(pcdgpu_bases* b) {
  if (!b) return;
  cudaSetDevice(b->ctx->device);
  cudaStreamSynchronize(b->ctx->stream);
  if (b->points) cudaFree(b->points);
  delete b;
}

// MSM over points [offset, offset + n) of a resident vector followed by its last n_extra points
// (the per-proof constant pairs appended at key upload) with plain scalars d_extra.
}  // extern "C"
int bases_msm(pcdgpu_ctx* ctx, const pcdgpu_bases* b, size_t offset, const void* d_scalars, int scalars_mont,
                     size_t n, const void* d_extra, size_t n_extra, void* d_out_xyzz) {
  const MsmOps* ops = msm_ops(b->curve);
  size_t avail = b->n - n_extra;  // ordinary points
  if (offset > avail) offset = avail;
  if (n > avail - offset) n = avail - offset;  // truncate to the shorter input, as arkworks does
  if (n_extra && offset + n != avail) {
    ctx->set_error("internal: extra pairs need the ordinary points to end where the extras begin");
    return PCDGPU_E_ARG;
  }
  if (b->table) {
    MsmPlanC plan{b->c, b->nwin, 1, b->n, offset};
    return ops->run(ctx, b->table, d_scalars, scalars_mont, n, d_extra, n_extra, plan, d_out_xyzz);
  }
  return ops->run(ctx, (const char*)b->points + offset * ops->affine_bytes, d_scalars, scalars_mont, n, d_extra,
                  n_extra, plan_plain(ctx, n + n_extra), d_out_xyzz);
}
extern "C" {

int pcdgpu_msm_bases_dev(pcdgpu_ctx* ctx, const pcdgpu_bases* b, size_t offset, const void* d_scalars, int scalars_mont,
                         size_t n, void* d_out_xyzz) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, b && d_out_xyzz && (n == 0 || d_scalars), "null pointer");
  CHECK_ARG(ctx, offset <= b->n, "offset beyond the base vector");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  return bases_msm(ctx, b, offset, d_scalars, scalars_mont, n, nullptr, 0, d_out_xyzz);
}

int pcdgpu_msm_bases(pcdgpu_ctx* ctx, const pcdgpu_bases* b, size_t offset, const void* scalars, size_t n,
                     void* out_affine) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, b && out_affine && (n == 0 || scalars), "null pointer");
  const MsmOps* ops = msm_ops(b->curve);
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  void *ds, *dres;
  PCD_TRY(ctx->scratch(SLOT_IO2, n * 40 + 16, &ds));
  PCD_TRY(ctx->scratch(SLOT_MISC, 4096, &dres));
  PCD_CUDA(ctx, cudaMemcpyAsync(ds, scalars, n * 40, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(pcdgpu_msm_bases_dev(ctx, b, offset, ds, 0, n, dres));
  return finish_to_host(ctx, ops, dres, out_affine);
}

int pcdgpu_xyzz_sum(pcdgpu_ctx* ctx, int curve, const void* xyzz, size_t n, void* out_affine) {
  if (!ctx) return PCDGPU_E_ARG;
  const MsmOps* ops = msm_ops(curve);
  CHECK_ARG(ctx, ops, "unknown curve id");
  CHECK_ARG(ctx, out_affine && (n == 0 || xyzz), "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  void* d;
  PCD_TRY(ctx->scratch(SLOT_IO, (n + 1) * ops->xyzz_bytes, &d));
  void* d_aff = (char*)d + n * ops->xyzz_bytes;
  PCD_CUDA(ctx, cudaMemcpyAsync(d, xyzz, n * ops->xyzz_bytes, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(ops->sum_points(ctx, d, 0, 1, (int)n, nullptr, 0, d_aff));
  PCD_CUDA(ctx, cudaMemcpyAsync(out_affine, d_aff, ops->affine_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int pcdgpu_xyzz_download(pcdgpu_ctx* ctx, int curve, const void* d_xyzz, void* out_xyzz) {
  if (!ctx) return PCDGPU_E_ARG;
  const MsmOps* ops = msm_ops(curve);
  CHECK_ARG(ctx, ops && d_xyzz && out_xyzz, "bad argument");
  PCD_CUDA(ctx, cudaMemcpyAsync(out_xyzz, d_xyzz, ops->xyzz_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---- fixed-base multiplication -------------------------------------------------------------------
int pcdgpu_fixed_base_mul_dev(pcdgpu_ctx* ctx, int curve, const void* base_host, const void* d_scalars, size_t n,
                              void* d_out) {
  if (!ctx) return PCDGPU_E_ARG;
  const MsmOps* ops = msm_ops(curve);
  CHECK_ARG(ctx, ops, "unknown curve id");
  CHECK_ARG(ctx, base_host && (n == 0 || (d_scalars && d_out)), "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  void* t;
  PCD_TRY(ctx->scratch(SLOT_MSM_SEG, (75 * 15 + 1) * ops->affine_bytes, &t));
  void* d_base = (char*)t + 75 * 15 * ops->affine_bytes;
  memcpy(ctx->pinned, base_host, ops->affine_bytes);
  PCD_CUDA(ctx, cudaMemcpyAsync(d_base, ctx->pinned, ops->affine_bytes, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(ops->fixed_table(ctx, d_base, t));
  PCD_TRY(ops->fixed_mul(ctx, t, d_scalars, n, d_out));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the pinned staging buffer is reusable afterwards
  return 0;
}

int pcdgpu_fixed_base_mul(pcdgpu_ctx* ctx, int curve, const void* base, const void* scalars, size_t n, void* out) {
  if (!ctx) return PCDGPU_E_ARG;
  const MsmOps* ops = msm_ops(curve);
  CHECK_ARG(ctx, ops, "unknown curve id");
  CHECK_ARG(ctx, base && (n == 0 || (scalars && out)), "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  void *ds, *dout;
  PCD_TRY(ctx->scratch(SLOT_IO2, n * 40 + 16, &ds));
  PCD_TRY(ctx->scratch(SLOT_IO, n * ops->affine_bytes + 16, &dout));
  PCD_CUDA(ctx, cudaMemcpyAsync(ds, scalars, n * 40, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(pcdgpu_fixed_base_mul_dev(ctx, curve, base, ds, n, dout));
  PCD_CUDA(ctx, cudaMemcpyAsync(out, dout, n * ops->affine_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ---- R1CS / witness map ----------------------------------------------------------------------------
int pcdgpu_r1cs_upload(pcdgpu_ctx* ctx, int pairing, size_t m, size_t num_inputs, size_t num_witness,
                       const uint32_t* a_ptr, const uint32_t* a_col, const void* a_val, const uint32_t* b_ptr,
                       const uint32_t* b_col, const void* b_val, const uint32_t* c_ptr, const uint32_t* c_col,
                       const void* c_val, pcdgpu_r1cs** out) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, pairing == PCDGPU_MNT4_298 || pairing == PCDGPU_MNT6_298, "unknown pairing id");
  CHECK_ARG(ctx, out && a_ptr && b_ptr && c_ptr, "null pointer");
  CHECK_ARG(ctx, num_inputs >= 1, "num_inputs counts the constant 1 and must be >= 1");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t* ptrs[3] = {a_ptr, b_ptr, c_ptr};
  const uint32_t* cols[3] = {a_col, b_col, c_col};
  const void* vals[3] = {a_val, b_val, c_val};
  size_t nnz[3], total = 0, off[9];
  for (int i = 0; i < 3; i++) {
    // row_ptr must be a monotone prefix array ending at nnz: a bad one would send spmv_kernel out of bounds on the device
    CHECK_ARG(ctx, ptrs[i][0] == 0, "row_ptr[0] must be 0");
    for (size_t k = 0; k < m; k++) CHECK_ARG(ctx, ptrs[i][k] <= ptrs[i][k + 1], "row_ptr is not monotone");
    nnz[i] = ptrs[i][m];
    CHECK_ARG(ctx, nnz[i] == 0 || (cols[i] && vals[i]), "null column / value array");
    for (size_t k = 0; k < nnz[i]; k++) CHECK_ARG(ctx, cols[i][k] < num_inputs + num_witness, "column index out of range");
    off[3 * i] = total;
    total += ((m + 1) * 4 + 15) & ~(size_t)15;
    off[3 * i + 1] = total;
    total += (nnz[i] * 4 + 15) & ~(size_t)15;
    off[3 * i + 2] = total;
    total += (nnz[i] * 40 + 15) & ~(size_t)15;
  }
  pcdgpu_r1cs* r = new pcdgpu_r1cs();
  r->uid = next_uid();
  r->ctx = ctx;
  r->pairing = pairing;
  r->m = m;
  r->num_inputs = num_inputs;
  r->num_witness = num_witness;
  {
    int field = pairing == PCDGPU_MNT4_298 ? PCDGPU_FIELD_R4 : PCDGPU_FIELD_Q4;
    if (ntt_domain_shape(field, m + num_inputs, &r->n, &r->dom_a, &r->dom_b) != 0) {
      r->dom_a = -1;  // no domain: the witness map reports PCDGPU_E_DOMAIN (the host still needs a size for h)
      r->dom_b = ilog2_ceil(m + num_inputs);
      r->n = (size_t)1 << r->dom_b;
    }
  }
  cudaError_t e = cudaMalloc(&r->storage, total ? total : 16);
  if (e != cudaSuccess) {
    ctx->set_error("cudaMalloc(%zu) for R1CS matrices: %s", total, cudaGetErrorString(e));
    delete r;
    return PCDGPU_E_NOMEM;
  }
  char* base = (char*)r->storage;
  CsrDev* M[3] = {&r->A, &r->B, &r->C};
  for (int i = 0; i < 3 && e == cudaSuccess; i++) {
    M[i]->row_ptr = (const u32*)(base + off[3 * i]);
    M[i]->col = (const u32*)(base + off[3 * i + 1]);
    M[i]->val = (const u32*)(base + off[3 * i + 2]);
    e = cudaMemcpyAsync(base + off[3 * i], ptrs[i], (m + 1) * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz[i])
      e = cudaMemcpyAsync(base + off[3 * i + 1], cols[i], nnz[i] * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && nnz[i])
      e = cudaMemcpyAsync(base + off[3 * i + 2], vals[i], nnz[i] * 40, cudaMemcpyHostToDevice, ctx->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    ctx->set_error("uploading R1CS matrices: %s", cudaGetErrorString(e));
    cudaFree(r->storage);
    delete r;
    return PCDGPU_E_CUDA;
  }
  *out = r;
  return 0;
}

void pcdgpu_r1cs_free(pcdgpu_r1cs* r) {
  if (!r) return;
  cudaSetDevice(r->ctx->device);
  cudaStreamSynchronize(r->ctx->stream);
  cudaFree(r->storage);
  delete r;
}

size_t pcdgpu_r1cs_domain_size(const pcdgpu_r1cs* r) { return r ? r->n : 0; }

int pcdgpu_witness_map(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, const void* z, void* h) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, r && z && h, "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t nv = r->num_inputs + r->num_witness;
  void* dz;
  PCD_TRY(ctx->scratch(SLOT_Z, nv * 40, &dz));
  PCD_CUDA(ctx, cudaMemcpyAsync(dz, z, nv * 40, cudaMemcpyHostToDevice, ctx->stream));
  void* dh;
  PCD_TRY(witness_map_dev(ctx, r, dz, &dh));
  PCD_CUDA(ctx, cudaMemcpyAsync(h, dh, r->n * 40, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int pcdgpu_qap_vector_dev(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, int which, const void* d_z, void* d_out) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, r && d_z && d_out && which >= 0 && which <= 2, "bad argument");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  return qap_vector_dev(ctx, r, which, d_z, d_out);
}

int pcdgpu_qap_combine_dev(pcdgpu_ctx* ctx, const pcdgpu_r1cs* r, void* d_a, const void* d_b, const void* d_c) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, r && d_a && d_b && d_c, "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  return qap_combine_dev(ctx, r, d_a, d_b, d_c);
}

// ---- Groth16 ------------------------------------------------------------------------------------------
}  // extern "C"
static void shard_of(size_t n, int world, int rank, size_t* lo, size_t* hi) {  // contiguous, balanced (pcd_b200/sharding.py)
  size_t base = n / world, extra = n % world;
  *lo = rank * base + ((size_t)rank < extra ? (size_t)rank : extra);
  *hi = *lo + base + ((size_t)rank < extra ? 1 : 0);
}
static int pk_upload_impl(pcdgpu_ctx* ctx, int pairing, size_t num_vars, size_t num_inputs, size_t h_len,
                     const void* alpha_g1, const void* beta_g1, const void* delta_g1, const void* beta_g2,
                     const void* delta_g2, const void* a_query, const void* b_g1_query, const void* b_g2_query,
                     const void* h_query, const void* l_query, int precompute, int rank, int world, pcdgpu_pk** out) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, pairing == PCDGPU_MNT4_298 || pairing == PCDGPU_MNT6_298, "unknown pairing id");
  CHECK_ARG(ctx, out && alpha_g1 && beta_g1 && delta_g1 && beta_g2 && delta_g2 && a_query && b_g1_query && b_g2_query,
            "null pointer");
  CHECK_ARG(ctx, num_inputs >= 1 && num_vars >= num_inputs, "bad variable counts");
  CHECK_ARG(ctx, (h_len == 0 || h_query) && (num_vars == num_inputs || l_query), "null query");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  int g1 = g1_of(pairing), g2 = g2_of(pairing);
  size_t s1 = msm_ops(g1)->affine_bytes, s2 = msm_ops(g2)->affine_bytes;
  pcdgpu_pk* pk = new pcdgpu_pk();
  memset(pk, 0, sizeof(*pk));
  pk->uid = next_uid();
  pk->ctx = ctx;
  pk->pairing = pairing;
  pk->num_vars = num_vars;
  pk->num_inputs = num_inputs;
  pk->h_len = h_len;
  pk->shard_rank = rank;
  pk->shard_world = world;
  shard_of(num_vars - 1, world, rank, &pk->v_lo, &pk->v_hi);
  shard_of(h_len, world, rank, &pk->h_lo, &pk->h_hi);
  shard_of(num_vars - num_inputs, world, rank, &pk->l_lo, &pk->l_hi);
  const int nex3 = rank == 0 ? 3 : 0, nex1 = rank == 0 ? 1 : 0;  // the constant points ride with rank 0
  int rc = 0;
  // Element 0 of the a/b queries belongs to the constant 1 and ark-groth16 adds it, the vk elements
  // and r*delta / s*delta outside its MSMs (prover.rs).  Here those pairs ride inside the MSMs: every
  // query vector is uploaded with its constant points appended, and the prover supplies their
  // scalars (r or s, 1, 1) as "extra" scalars -- no serial scalar multiplication is left for them.
  //   a_query    : a[1..], delta_g1, a[0], alpha_g1          scalars z[1..], r, 1, 1
  //   b_g1_query : b[1..], delta_g1, b[0], beta_g1           scalars z[1..], s, 1, 1
  //   b_g2_query : b2[1..], delta_g2, b2[0], beta_g2         scalars z[1..], s, 1, 1
  //   l_query    : l[..], delta_g1                           scalars z[num_inputs..], -(r s)
  auto upload_ext = [&](int curve, const void* q, size_t skip, size_t n, const void* const* extra, int n_extra,
                        pcdgpu_bases** out) -> int {
    size_t pb = msm_ops(curve)->affine_bytes;
    std::vector<char> buf((n + n_extra) * pb);
    if (n) memcpy(buf.data(), (const char*)q + skip * pb, n * pb);
    for (int i = 0; i < n_extra; i++) memcpy(buf.data() + (n + i) * pb, extra[i], pb);
    return pcdgpu_bases_upload(ctx, curve, buf.data(), n + n_extra, precompute, out);
  };
  ctx->key_upload = true;
  const void* ea[3] = {delta_g1, a_query, alpha_g1};
  const void* eb1[3] = {delta_g1, b_g1_query, beta_g1};
  const void* eb2[3] = {delta_g2, b_g2_query, beta_g2};
  const void* el[1] = {delta_g1};
  rc = rc ? rc : upload_ext(g1, a_query, 1 + pk->v_lo, pk->v_hi - pk->v_lo, ea, nex3, &pk->a_query);
  rc = rc ? rc : upload_ext(g1, b_g1_query, 1 + pk->v_lo, pk->v_hi - pk->v_lo, eb1, nex3, &pk->b_g1_query);
  rc = rc ? rc : upload_ext(g2, b_g2_query, 1 + pk->v_lo, pk->v_hi - pk->v_lo, eb2, nex3, &pk->b_g2_query);
  rc = rc ? rc : pcdgpu_bases_upload(ctx, g1, h_query ? (const char*)h_query + pk->h_lo * s1 : nullptr,
                                     pk->h_hi - pk->h_lo, precompute, &pk->h_query);
  rc = rc ? rc : upload_ext(g1, l_query, pk->l_lo, pk->l_hi - pk->l_lo, el, nex1, &pk->l_query);
  ctx->key_upload = false;
  (void)s2;
  if (rc) {
    pcdgpu_pk_free(pk);
    return rc;
  }
  *out = pk;
  return 0;
}

extern "C" {
int pcdgpu_pk_upload(pcdgpu_ctx* ctx, int pairing, size_t num_vars, size_t num_inputs, size_t h_len,
                     const void* alpha_g1, const void* beta_g1, const void* delta_g1, const void* beta_g2,
                     const void* delta_g2, const void* a_query, const void* b_g1_query, const void* b_g2_query,
                     const void* h_query, const void* l_query, int precompute, pcdgpu_pk** out) {
  return pk_upload_impl(ctx, pairing, num_vars, num_inputs, h_len, alpha_g1, beta_g1, delta_g1, beta_g2, delta_g2, a_query,
                        b_g1_query, b_g2_query, h_query, l_query, precompute, 0, 1, out);
}
int pcdgpu_pk_upload_sharded(pcdgpu_ctx* ctx, int pairing, size_t num_vars, size_t num_inputs, size_t h_len,
                             const void* alpha_g1, const void* beta_g1, const void* delta_g1, const void* beta_g2,
                             const void* delta_g2, const void* a_query, const void* b_g1_query, const void* b_g2_query,
                             const void* h_query, const void* l_query, int precompute, pcdgpu_pk** out) {
  if (!ctx) return PCDGPU_E_ARG;
  ctx->in_proof = true;  // window rule of MSMs that run side by side, as in the single-GPU prover
  int rc = pk_upload_impl(ctx, pairing, num_vars, num_inputs, h_len, alpha_g1, beta_g1, delta_g1, beta_g2, delta_g2, a_query,
                          b_g1_query, b_g2_query, h_query, l_query, precompute, ctx->comm_rank, ctx->comm_world, out);
  ctx->in_proof = false;
  return rc;
}

void pcdgpu_pk_free(pcdgpu_pk* pk) {
  if (!pk) return;
  pcdgpu_bases_free(pk->a_query);
  pcdgpu_bases_free(pk->b_g1_query);
  pcdgpu_bases_free(pk->b_g2_query);
  pcdgpu_bases_free(pk->h_query);
  pcdgpu_bases_free(pk->l_query);
  delete pk;
}

}  // extern "C"
// Assembly state of a Groth16 proof in SLOT_MISC (shared by the prover and the multi-GPU assembly calls):
//   rs (r | s, plain) | extras: 13 plain scalars of the constant pairs | sums1: h, l', T | s g_a, r g1_b, g_a, g1_b
//   (G1 xyzz) | sum2: g2_b | proof (A || B || C affine)
struct G16Misc {
  u32* d_rs;
  char* extras;
  void* sums1;
  void* sum2;
  char* d_proof;
  size_t x1, x2, a1, a2;
};
static int g16_misc(pcdgpu_ctx* ctx, int pairing, G16Misc* m) {
  const MsmOps *o1 = msm_ops(pcd_g1_of(pairing)), *o2 = msm_ops(pcd_g2_of(pairing));
  void* misc;
  PCD_TRY(ctx->scratch(SLOT_MISC, 8192, &misc));
  char* mb = (char*)misc;
  m->x1 = o1->xyzz_bytes;
  m->x2 = o2->xyzz_bytes;
  m->a1 = o1->affine_bytes;
  m->a2 = o2->affine_bytes;
  m->d_rs = (u32*)mb;
  m->extras = mb + 128;  // 13 x 40 B
  m->sums1 = mb + 1024;
  m->sum2 = (char*)m->sums1 + 6 * m->x1;
  m->d_proof = (char*)m->sum2 + m->x2;
  return 0;
}
extern "C" {

// Proofs of at most this many variables compute s g_a + r g1_b as two MORE MSMs (scalars s z and r z over the a and
// b_g1 queries) instead of a double-scalar multiplication of the finished g_a and g1_b: the two MSMs start with the
// others, so nothing waits for a 298-doubling chain -- the tiny default-circuit proofs of a PCD step
// (/root/reference/src/ec_cycle_pcd/data_structures.rs:139-143,343-350) are pure latency.  Above it the extra
// accumulation work (two MSMs with DENSE scalars) costs more than the chain, which there hides under the h MSM.
// MEASURED inside the PCD step (B200, tools/probe_step.py): the helper proof (MNT6, 2^16 variables) 4.87 ms with the
// chain, 4.23 ms with the two MSMs; the main proof (MNT4, 2^18) 6.7 ms with the chain, 9.2 ms with the MSMs -- hence
// 2^17.  PCDGPU_SMALL_NV_LOG overrides it (development aid for A/B runs).
static size_t groth16_small_nv() {
  static const int lg = getenv("PCDGPU_SMALL_NV_LOG") ? atoi(getenv("PCDGPU_SMALL_NV_LOG")) : 17;
  return (size_t)1 << (lg < 0 ? 0 : (lg > 40 ? 40 : lg));
}
#define GROTH16_SMALL_NV groth16_small_nv()

// Everything a proof enqueues, from the upload of (r, s) (staged in ctx->pinned) to the copy of the proof into
// h_proof (pinned): no host synchronisation, no host read of device data -- the sequence depends only on the key, the
// constraint system and the addresses, so it can be captured into a CUDA graph (pcdgpu_groth16_prove_dev).
static int groth16_enqueue(pcdgpu_ctx* ctx, const pcdgpu_pk* pk, const pcdgpu_r1cs* r1cs, const void* d_z, const G16Misc& m,
                           void* h_proof) {
  int g1 = g1_of(pk->pairing), g2 = g2_of(pk->pairing);
  const size_t x1 = m.x1;
  u32* d_rs = m.d_rs;
  char* extras = m.extras;
  void *sums1 = m.sums1, *sum2 = m.sum2;
  size_t proof_bytes = 2 * m.a1 + m.a2;
  char* d_A = m.d_proof;
  char* d_B = m.d_proof + m.a1;
  char* d_C = m.d_proof + m.a1 + m.a2;
  PCD_CUDA(ctx, cudaMemcpyAsync(d_rs, ctx->pinned, 80, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(groth16_prepare(ctx, pk->pairing, d_rs, (u32*)extras));
  const char* z = (const char*)d_z;
  size_t nv = pk->num_vars, ni = pk->num_inputs;
  const bool fork = ctx->concurrent;
  const bool small = nv <= GROTH16_SMALL_NV;
  void *d_sz = nullptr, *d_rz = nullptr;
  if (small) {  // s z and r z (Montgomery), before the fork: both extra lanes read them
    PCD_TRY(ctx->scratch(SLOT_SAP_FULL, 2 * nv * 40 + 80, &d_sz));
    d_rz = (char*)d_sz + nv * 40;
    PCD_TRY(groth16_scale(ctx, pk->pairing, z + 40, nv - 1, d_rs, d_sz, d_rz));
  }
  // Fork: the MSMs over the assignment only need z and the extra scalars, so they start on lanes 1-4 (5, 6) while
  // lane 0 runs the witness map and then the h MSM (the G2 MSM, the longest, goes first).  Each lane finishes its
  // own piece of the proof: lane 1 normalises B, lane 2 (a) normalises A, lane 3 (b_g1) waits for lane 2 and runs the
  // double-scalar multiplication T = s g_a + r g1_b (large proofs) -- all of it under the witness map / h MSM, so that
  // after the join only C = h + l' + T is left.
  const int nlanes = small ? pcdgpu_ctx::NLANE : pcdgpu_ctx::NLANE_PROOF;
  if (fork) {
    PCD_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    for (int l = 1; l < nlanes; l++) PCD_CUDA(ctx, cudaStreamWaitEvent(ctx->lane_stream[l], ctx->ev_fork, 0));
  }
  struct Job { const pcdgpu_bases* b; const char* sc; size_t n; const char* ex; size_t nex; void* out; };
  Job jobs[6] = {{pk->b_g2_query, z + 40, nv - 1, extras + 3 * 40, 3, sum2},
                 {pk->a_query, z + 40, nv - 1, extras, 3, (char*)sums1 + 4 * x1},
                 {pk->b_g1_query, z + 40, nv - 1, extras + 3 * 40, 3, (char*)sums1 + 5 * x1},
                 {pk->l_query, z + 40 * ni, nv - ni, extras + 6 * 40, 1, (char*)sums1 + 1 * x1},
                 {pk->a_query, (const char*)d_sz, nv - 1, extras + 7 * 40, 3, (char*)sums1 + 2 * x1},
                 {pk->b_g1_query, (const char*)d_rz, nv - 1, extras + 10 * 40, 3, (char*)sums1 + 3 * x1}};
  int rc = 0;
  // the witness map is enqueued FIRST (lane 0, most urgent): its short kernels then start while the GPU is still
  // empty instead of queueing behind the lanes' accumulation grids; when the lanes are serialised (profiling) it keeps
  // its place after them so that the scratch it returns stays untouched until the h MSM
  void* d_h = nullptr;
  ctx->lane = 0;
  if (fork) rc = witness_map_dev(ctx, r1cs, d_z, &d_h);
  // Order of the accumulation grids of a large proof (the throughput-bound part; everything else hides under it): a and
  // b_g1 first -- the double-scalar multiplication behind them is the longest serial tail (2 - 3 ms on one warp) and
  // must start early --, then b_g2 (most urgent of the rest: its bucket reduction is the next longest tail), l and h.  Only the accumulate
  // kernels are ordered (events recorded right after / waited right before their launch): sorting, bucket reduction
  // and the tails of every lane still overlap freely.  Enqueue order = dependency order (an event must have been
  // recorded before it is waited for).
  // MEASURED (B200, tools/probe_pcd.py).  While b_g2's accumulation was the longest chain this order was slower (main
  // 2^18: 9.05 vs 8.15 ms); since base points at infinity are skipped and the heavy-bucket threshold follows the real
  // entry count, the a / b_g1 -> double-scalar chain is the critical one and the order wins: main 6.16 -> 6.10 ms,
  // helper (MNT6, 2^16) 4.28 -> 3.88 ms.  PCDGPU_NO_ACC_ORDER turns it off (A/B runs).
  // Also measured and NOT kept: every lane's accumulation grid waiting for the witness map (whose transforms take ~0.7 ms
  // alone and ~3.1 ms beside four accumulating lanes, and the h MSM cannot start before them): the map then ends at
  // 1.1 ms, but the h MSM's sorting kernels crawl behind the four grids that start together (0.2 -> 2.2 ms) and the
  // double-scalar chain starts a millisecond later: main 6.77 -> 7.2 - 7.3 ms.  h's grid not waiting for b_g2's: 6.8 - 6.9.
  // The h MSM enqueued right behind the witness map with every lane's grid waiting for h's SORTING (witness map + h's
  // sort on a GPU that only sorts: 0.97 ms instead of 3.3): the grids then all start at ~1 ms but the a / b_g1 ->
  // double-scalar chain starts 1.5 ms later than today: main 6.59 -> 7.41 ms.
  static const bool want_gates = getenv("PCDGPU_NO_ACC_ORDER") == nullptr;
  const bool gates = fork && !small && want_gates;
  // A lighter ordering, also measured: b_g2's accumulation grid starts only when the a, b_g1 and l lanes have
  // finished SORTING.  Its CTAs fill the register file (Fq2: 234 registers, Fq3 sliced: four CTAs per SM) and its first
  // work items are the largest buckets, so for the first millisecond nothing retires and the other lanes' sort kernels
  // (0.1 - 0.25 ms each alone) crawled for 0.9 - 1.4 ms, delaying the a / b_g1 accumulation and with it the
  // double-scalar multiplication.
  // MEASURED: helper (MNT6, 2^16) 5.4 - 5.6 -> 5.3 - 5.4 ms, main (MNT4, 2^18) 8.0 - 8.2 -> 8.6 - 8.7 ms: off by default
  // (PCDGPU_SORT_GATE enables it for experiments)
  static const bool want_sort_gate = getenv("PCDGPU_SORT_GATE") != nullptr;
  const bool sort_gate = fork && !small && !gates && want_sort_gate;
  const int order_gated[6] = {1, 2, 0, 3, 4, 5}, order_plain[6] = {0, 1, 2, 3, 4, 5}, order_sorted[6] = {1, 2, 3, 0, 4, 5};
  const int* order = gates ? order_gated : (sort_gate ? order_sorted : order_plain);
  for (int q = 0; q < nlanes - 1 && rc == 0; q++) {
    const int j = order[q];
    // job -> lane.  Lanes differ in stream priority (pcdgpu_ctx_create: lanes 2, 3, 5, 6 above lane 1 above lane 4).  A
    // large proof keeps the a / b_g1 MSMs on the urgent lanes 2, 3 (the double-scalar chain waits for them); a proof
    // without the chain (small) is bound by its G2 MSM, which then takes lane 2 and leaves lane 1 to the a MSM.
    // MEASURED inside the PCD step (tools/probe_step.py): helper (MNT6, 2^16) 4.08 -> 3.90 ms, default-circuit proofs
    // unchanged.  PCDGPU_NO_SMALL_G2_URGENT restores the plain mapping (A/B runs).
    static const bool g2_urgent = getenv("PCDGPU_NO_SMALL_G2_URGENT") == nullptr;
    const int lane_of_job = (small && g2_urgent) ? (j == 0 ? 2 : (j == 1 ? 1 : j + 1)) : j + 1;
    ctx->lane = fork ? lane_of_job : 0;
    if (sort_gate) {
      if (j >= 1 && j <= 3) ctx->sort_done = ctx->ev_sorted[j - 1];
      if (j == 0)
        for (int k = 0; k < 3; k++) ctx->gate_wait[k] = ctx->ev_sorted[k];
    }
    if (gates) {
      if (j == 1 || j == 2) ctx->gate_done = ctx->ev_acc[j - 1];
      if (j == 0 || j == 3) {
        ctx->gate_wait[0] = ctx->ev_acc[0];
        ctx->gate_wait[1] = ctx->ev_acc[1];
        if (j == 0) ctx->gate_done = ctx->ev_acc[2];
      }
    }
    rc = bases_msm(ctx, jobs[j].b, 0, jobs[j].sc, 1, jobs[j].n, jobs[j].ex, jobs[j].nex, jobs[j].out);
    ctx->gate_wait[0] = ctx->gate_wait[1] = ctx->gate_wait[2] = ctx->gate_done = ctx->sort_done = nullptr;
    if (rc == 0 && j == 0) rc = point_to_affine(ctx, g2, sum2, 0, d_B);
    if (rc == 0 && j == 1) rc = point_to_affine(ctx, g1, sums1, 4, d_A);
    if (rc == 0 && j == 2 && !small) {
      // lane 3's double-scalar multiplication needs g_a from lane 2
      if (fork && cudaStreamWaitEvent(ctx->lane_stream[3], ctx->ev_join[2], 0) != cudaSuccess) rc = PCDGPU_E_CUDA;
      if (rc == 0) rc = groth16_straus(ctx, pk->pairing, d_rs, sums1);
    }
    if (fork && rc == 0 && cudaEventRecord(ctx->ev_join[lane_of_job], ctx->lane_stream[lane_of_job]) != cudaSuccess)
      rc = PCDGPU_E_CUDA;
  }
  ctx->lane = 0;
  if (rc == 0 && !fork) rc = witness_map_dev(ctx, r1cs, d_z, &d_h);
  // h: n coefficients vs n - 1 query points: truncated to the shorter
  if (rc == 0 && gates) ctx->gate_wait[0] = ctx->ev_acc[2];  // h after b_g2's grid: lane 0 is the most urgent stream
  if (rc == 0) rc = bases_msm(ctx, pk->h_query, 0, d_h, 1, r1cs->n, nullptr, 0, (char*)sums1 + 0 * x1);
  ctx->gate_wait[0] = ctx->gate_wait[1] = ctx->gate_done = nullptr;
  if (rc == 0 && fork)
    for (int l = 1; l < nlanes && rc == 0; l++)
      if (cudaStreamWaitEvent(ctx->stream, ctx->ev_join[l], 0) != cudaSuccess) rc = PCDGPU_E_CUDA;
  if (rc == 0) rc = groth16_finish(ctx, pk->pairing, sums1, d_C, small ? 4 : 3);
  if (rc == 0 && cudaMemcpyAsync(h_proof, m.d_proof, proof_bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    rc = PCDGPU_E_CUDA;
  return rc;
}

static pcdgpu_ctx::ProofGraph* proof_graph_slot(pcdgpu_ctx* ctx, const pcdgpu_pk* pk, const pcdgpu_r1cs* r1cs, const void* d_z) {
  pcdgpu_ctx::ProofGraph* lru = &ctx->graphs[0];
  for (auto& g : ctx->graphs) {
    if (g.seen && g.pk_uid == pk->uid && g.r1cs_uid == r1cs->uid && g.d_z == d_z) {
      g.last_use = ++ctx->graph_clock;
      return &g;
    }
    if (g.last_use < lru->last_use) lru = &g;
  }
  if (lru->exec) cudaGraphExecDestroy(lru->exec);
  *lru = pcdgpu_ctx::ProofGraph();
  lru->pk_uid = pk->uid;
  lru->r1cs_uid = r1cs->uid;
  lru->d_z = d_z;
  lru->last_use = ++ctx->graph_clock;
  return lru;
}

int pcdgpu_groth16_prove_dev(pcdgpu_ctx* ctx, const pcdgpu_pk* pk, const pcdgpu_r1cs* r1cs, const void* d_z,
                             const void* r, const void* s, void* out_proof) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, pk && r1cs && d_z && r && s && out_proof, "null pointer");
  CHECK_ARG(ctx, pk->pairing == r1cs->pairing, "key and constraint system are over different pairings");
  CHECK_ARG(ctx, pk->num_vars == r1cs->num_inputs + r1cs->num_witness && pk->num_inputs == r1cs->num_inputs,
            "key and constraint system disagree on the variable counts");
  CHECK_ARG(ctx, pk->shard_world == 1, "sharded key: use pcdgpu_groth16_prove_sharded");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  InProofGuard in_proof(ctx);
  CHECK_ARG(ctx, in_proof.ok, "the context is running a prover call on another host thread (one context per thread)");
  G16Misc m;
  PCD_TRY(g16_misc(ctx, pk->pairing, &m));
  const size_t proof_bytes = 2 * m.a1 + m.a2;
  char* h_proof = (char*)ctx->pinned + 128;
  memcpy(ctx->pinned, r, 40);
  memcpy((char*)ctx->pinned + 40, s, 40);
  // Proof graphs (common.cuh: ProofGraph; off unless pcdgpu_set_proof_graphs).  First call with a (key, system,
  // assignment address): eager.  Second call in
  // the same scratch epoch: captured, instantiated and launched.  Later calls: one cudaGraphLaunch.
  static const bool env_graphs = getenv("PCDGPU_GRAPHS") != nullptr;  // development aid (A/B runs): on without the call
  pcdgpu_ctx::ProofGraph* ge =
      ((ctx->use_graphs || env_graphs) && ctx->concurrent && !ctx->profiling) ? proof_graph_slot(ctx, pk, r1cs, d_z) : nullptr;
  if (ge && ge->exec && ge->epoch != ctx->scratch_epoch) {  // scratch moved since the capture
    cudaGraphExecDestroy(ge->exec);
    ge->exec = nullptr;
    ge->seen = 1;
  }
  int rc = 0;
  bool launched = false;
  if (ge && ge->exec) {
    if (cudaGraphLaunch(ge->exec, ctx->stream) != cudaSuccess) {
      ctx->set_error("cudaGraphLaunch: %s", cudaGetErrorString(cudaGetLastError()));
      rc = PCDGPU_E_CUDA;
    }
    ctx->launches += ge->launches;
    ctx->graphs_replayed++;
    launched = true;
  } else if (ge && ge->seen == 1 && ge->epoch == ctx->scratch_epoch) {
    const unsigned long long l0 = ctx->launches;
    cudaGraph_t graph = nullptr;
    bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
    if (ok) {
      ctx->capturing = true;
      int crc = groth16_enqueue(ctx, pk, r1cs, d_z, m, h_proof);
      ctx->capturing = false;
      ctx->lane = 0;
      ok = cudaStreamEndCapture(ctx->stream, &graph) == cudaSuccess && graph != nullptr && crc == 0;
      if (ok) ok = cudaGraphInstantiate(&ge->exec, graph, 0) == cudaSuccess;
      if (graph) cudaGraphDestroy(graph);
      if (ok) {
        ge->launches = ctx->launches - l0;
        ge->epoch = ctx->scratch_epoch;
        ctx->launches = l0;
        if (cudaGraphLaunch(ge->exec, ctx->stream) != cudaSuccess) {
          ctx->set_error("cudaGraphLaunch: %s", cudaGetErrorString(cudaGetLastError()));
          rc = PCDGPU_E_CUDA;
        }
        ctx->launches += ge->launches;
        ctx->graphs_captured++;
        ctx->graphs_replayed++;
        launched = true;
      } else {
        ctx->launches = l0;
        if (crc != 0 && crc != PCD_E_RETRY) rc = crc;  // a real error: nothing was executed, report it
      }
    }
    if (!ok) {  // no graph for this entry, now or later; the call is redone eagerly below
      cudaGetLastError();
      ge->exec = nullptr;
      ge->seen = 2;
    }
  }
  if (!launched && rc == 0) {
    rc = groth16_enqueue(ctx, pk, r1cs, d_z, m, h_proof);
    if (ge && ge->seen == 0) ge->seen = 1;
    if (ge) ge->epoch = ctx->scratch_epoch;  // compared at the next call: captured only if nothing grew in between
  }
  // every exit, error or not, drains the lanes and the stream: the next call reuses the pinned staging buffer and
  // the scratch the lanes are working in
  if (rc) ctx->drain_lanes();
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (rc == 0 && e != cudaSuccess) {
    ctx->set_error("cudaStreamSynchronize: %s", cudaGetErrorString(e));
    rc = PCDGPU_E_CUDA;
  }
  if (rc == 0) memcpy(out_proof, h_proof, proof_bytes);
  return rc;
}

// One proof over the GPUs of a communicator (pcdgpu_comm_init): every rank holds a slice of each query
// (pcdgpu_pk_upload_sharded), computes the partial sums of the five MSMs over its slice -- the witness map is
// recomputed on every rank (0.2 - 1 ms; cheaper than moving h) -- and two all-gathers of xyzz points move them:
// (a, b_g1 | b_g2) as soon as those three MSMs are done, on the b_g1 lane, so that A, B and s g_a + r g1_b are ready
// before the h MSM ends, then (h, l').  Every rank assembles and returns the proof (bit-identical on all ranks and to
// the single-GPU proof: affine points are canonical).
int pcdgpu_groth16_prove_sharded_dev(pcdgpu_ctx* ctx, const pcdgpu_pk* pk, const pcdgpu_r1cs* r1cs, const void* d_z,
                                     const void* r, const void* s, void* out_proof) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, pk && r1cs && d_z && r && s && out_proof, "null pointer");
  CHECK_ARG(ctx, pk->pairing == r1cs->pairing, "key and constraint system are over different pairings");
  CHECK_ARG(ctx, pk->num_vars == r1cs->num_inputs + r1cs->num_witness && pk->num_inputs == r1cs->num_inputs,
            "key and constraint system disagree on the variable counts");
  CHECK_ARG(ctx, pk->shard_world == ctx->comm_world && pk->shard_rank == ctx->comm_rank,
            "the key was sharded for another communicator");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  InProofGuard in_proof(ctx);
  CHECK_ARG(ctx, in_proof.ok, "the context is running a prover call on another host thread (one context per thread)");
  const int world = ctx->comm_world;
  int g1 = g1_of(pk->pairing), g2 = g2_of(pk->pairing);
  G16Misc m;
  PCD_TRY(g16_misc(ctx, pk->pairing, &m));
  const size_t x1 = m.x1, x2 = m.x2;
  // comm layout: mine_ab (2 x1) | mine_g2 (x2) | mine_hl (2 x1) | all_ab (world 2 x1) | all_g2 (world x2) | all_hl (world 2 x1)
  void* comm;
  PCD_TRY(ctx->scratch(SLOT_COMM, (size_t)(world + 1) * (4 * x1 + x2) + 256, &comm));
  char* mine_ab = (char*)comm;
  char* mine_g2 = mine_ab + 2 * x1;
  char* mine_hl = mine_g2 + x2;
  char* all_ab = mine_hl + 2 * x1;
  char* all_g2 = all_ab + (size_t)world * 2 * x1;
  char* all_hl = all_g2 + (size_t)world * x2;
  memcpy(ctx->pinned, r, 40);
  memcpy((char*)ctx->pinned + 40, s, 40);
  PCD_CUDA(ctx, cudaMemcpyAsync(m.d_rs, ctx->pinned, 80, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(groth16_prepare(ctx, pk->pairing, m.d_rs, (u32*)m.extras));
  const char* z = (const char*)d_z;
  const size_t ni = pk->num_inputs;
  const bool fork = ctx->concurrent;
  const bool r0 = pk->shard_rank == 0;
  if (fork) {
    PCD_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    for (int l = 1; l < pcdgpu_ctx::NLANE_PROOF; l++) PCD_CUDA(ctx, cudaStreamWaitEvent(ctx->lane_stream[l], ctx->ev_fork, 0));
  }
  struct Job { const pcdgpu_bases* b; const char* sc; size_t n; const char* ex; size_t nex; void* out; };
  const size_t nvs = pk->v_hi - pk->v_lo;
  Job jobs[4] = {{pk->b_g2_query, z + 40 * (1 + pk->v_lo), nvs, m.extras + 3 * 40, r0 ? 3u : 0u, mine_g2},
                 {pk->a_query, z + 40 * (1 + pk->v_lo), nvs, m.extras, r0 ? 3u : 0u, mine_ab},
                 {pk->b_g1_query, z + 40 * (1 + pk->v_lo), nvs, m.extras + 3 * 40, r0 ? 3u : 0u, mine_ab + x1},
                 {pk->l_query, z + 40 * (ni + pk->l_lo), pk->l_hi - pk->l_lo, m.extras + 6 * 40, r0 ? 1u : 0u, mine_hl + x1}};
  int rc = 0;
  for (int j = 0; j < 4 && rc == 0; j++) {
    ctx->lane = fork ? j + 1 : 0;
    rc = bases_msm(ctx, jobs[j].b, 0, jobs[j].sc, 1, jobs[j].n, jobs[j].ex, jobs[j].nex, jobs[j].out);
    if (rc == 0 && j == 2) {
      // first exchange, on this lane: needs the b_g2 and a partial sums of lanes 1 and 2 as well
      if (fork && (cudaStreamWaitEvent(ctx->lane_stream[3], ctx->ev_join[1], 0) != cudaSuccess ||
                   cudaStreamWaitEvent(ctx->lane_stream[3], ctx->ev_join[2], 0) != cudaSuccess))
        rc = PCDGPU_E_CUDA;
      if (rc == 0) rc = comm_group(ctx, true);
      if (rc == 0) rc = comm_allgather(ctx, mine_ab, all_ab, 2 * x1, ctx->cur());
      if (rc == 0) rc = comm_allgather(ctx, mine_g2, all_g2, x2, ctx->cur());
      if (rc == 0) rc = comm_group(ctx, false);
      // g_a, g1_b -> sums1[4], sums1[5]; g2_b -> sum2; A, B; T = s g_a + r g1_b
      if (rc == 0) rc = groth16_sum_partials(ctx, pk->pairing, all_ab, all_g2, world, 2, 1, (char*)m.sums1 + 4 * x1, m.sum2);
      if (rc == 0) rc = point_to_affine(ctx, g1, m.sums1, 4, m.d_proof);
      if (rc == 0) rc = point_to_affine(ctx, g2, m.sum2, 0, m.d_proof + m.a1);
      if (rc == 0) rc = groth16_straus(ctx, pk->pairing, m.d_rs, m.sums1);
    }
    if (fork && rc == 0 && cudaEventRecord(ctx->ev_join[j + 1], ctx->lane_stream[j + 1]) != cudaSuccess) rc = PCDGPU_E_CUDA;
  }
  ctx->lane = 0;
  void* d_h = nullptr;
  if (rc == 0) rc = witness_map_dev(ctx, r1cs, d_z, &d_h);
  if (rc == 0) {
    size_t hn = pk->h_hi - pk->h_lo;
    if (pk->h_lo >= r1cs->n) hn = 0;
    else if (pk->h_lo + hn > r1cs->n) hn = r1cs->n - pk->h_lo;
    rc = bases_msm(ctx, pk->h_query, 0, (const char*)d_h + 40 * pk->h_lo, 1, hn, nullptr, 0, mine_hl);
  }
  if (rc == 0 && fork)
    for (int l = 1; l < pcdgpu_ctx::NLANE_PROOF && rc == 0; l++)
      if (cudaStreamWaitEvent(ctx->stream, ctx->ev_join[l], 0) != cudaSuccess) rc = PCDGPU_E_CUDA;
  if (rc == 0) rc = comm_allgather(ctx, mine_hl, all_hl, 2 * x1, ctx->stream);
  if (rc == 0) rc = groth16_sum_partials(ctx, pk->pairing, all_hl, nullptr, world, 2, 0, m.sums1, nullptr);
  if (rc == 0) rc = groth16_finish(ctx, pk->pairing, m.sums1, m.d_proof + m.a1 + m.a2, 3);
  if (rc == 0 && cudaMemcpyAsync(out_proof, m.d_proof, 2 * m.a1 + m.a2, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
    rc = PCDGPU_E_CUDA;
  if (rc) ctx->drain_lanes();
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (rc == 0 && e != cudaSuccess) {
    ctx->set_error("cudaStreamSynchronize: %s", cudaGetErrorString(e));
    rc = PCDGPU_E_CUDA;
  }
  return rc;
}

int pcdgpu_groth16_prove_sharded(pcdgpu_ctx* ctx, const pcdgpu_pk* pk, const pcdgpu_r1cs* r1cs, const void* z, const void* r,
                                 const void* s, void* out_proof) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, pk && r1cs && z && r && s && out_proof, "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t nv = r1cs->num_inputs + r1cs->num_witness;
  void* dz;
  PCD_TRY(ctx->scratch(SLOT_Z, nv * 40, &dz));
  PCD_CUDA(ctx, cudaMemcpyAsync(dz, z, nv * 40, cudaMemcpyHostToDevice, ctx->stream));
  return pcdgpu_groth16_prove_sharded_dev(ctx, pk, r1cs, dz, r, s, out_proof);
}

// One MSM sharded by point range: every rank passes ITS slice of the bases (resident) and of the scalars; the xyzz
// partial sums are all-gathered (one exchange of 160 - 480 B per rank) and every rank adds them and normalises.
int pcdgpu_msm_bases_sharded_dev(pcdgpu_ctx* ctx, const pcdgpu_bases* slice, const void* d_scalars, int scalars_mont,
                                 size_t n, void* d_out_affine) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, slice && d_out_affine && (n == 0 || d_scalars), "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  const MsmOps* ops = msm_ops(slice->curve);
  const int world = ctx->comm_world;
  void* comm;
  PCD_TRY(ctx->scratch(SLOT_COMM, (size_t)(world + 1) * ops->xyzz_bytes + 256, &comm));
  char* mine = (char*)comm;
  char* all = mine + ops->xyzz_bytes;
  PCD_TRY(bases_msm(ctx, slice, 0, d_scalars, scalars_mont, n, nullptr, 0, mine));
  PCD_TRY(comm_allgather(ctx, mine, all, ops->xyzz_bytes, ctx->stream));
  return ops->sum_points(ctx, all, 0, 1, world, nullptr, 0, d_out_affine);
}
int pcdgpu_msm_bases_sharded(pcdgpu_ctx* ctx, const pcdgpu_bases* slice, const void* scalars, size_t n, void* out_affine) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, slice && out_affine && (n == 0 || scalars), "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  const MsmOps* ops = msm_ops(slice->curve);
  void *ds, *dres;
  PCD_TRY(ctx->scratch(SLOT_IO2, n * 40 + 16, &ds));
  PCD_TRY(ctx->scratch(SLOT_MISC, 8192, &dres));
  PCD_CUDA(ctx, cudaMemcpyAsync(ds, scalars, n * 40, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(pcdgpu_msm_bases_sharded_dev(ctx, slice, ds, 0, n, dres));
  PCD_CUDA(ctx, cudaMemcpyAsync(out_affine, dres, ops->affine_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int pcdgpu_groth16_assemble_begin_dev(pcdgpu_ctx* ctx, int pairing, const void* r, const void* s, int world,
                                      const void* d_partials_ab, const void* d_partials_g2) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, pairing == PCDGPU_MNT4_298 || pairing == PCDGPU_MNT6_298, "unknown pairing id");
  CHECK_ARG(ctx, r && s && d_partials_ab && d_partials_g2 && world >= 1, "bad argument");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  int g1 = g1_of(pairing), g2 = g2_of(pairing);
  G16Misc m;
  PCD_TRY(g16_misc(ctx, pairing, &m));
  memcpy(ctx->pinned, r, 40);
  memcpy((char*)ctx->pinned + 40, s, 40);
  PCD_CUDA(ctx, cudaMemcpyAsync(m.d_rs, ctx->pinned, 80, cudaMemcpyHostToDevice, ctx->stream));
  // g_a, g1_b -> sums1[4], sums1[5]; g2_b -> sum2
  PCD_TRY(groth16_sum_partials(ctx, pairing, d_partials_ab, d_partials_g2, world, 2, 1, (char*)m.sums1 + 4 * m.x1, m.sum2));
  PCD_TRY(point_to_affine(ctx, g1, m.sums1, 4, m.d_proof));
  PCD_TRY(point_to_affine(ctx, g2, m.sum2, 0, m.d_proof + m.a1));
  return groth16_straus(ctx, pairing, m.d_rs, m.sums1);
}

int pcdgpu_groth16_assemble_finish_dev(pcdgpu_ctx* ctx, int pairing, int world, const void* d_partials_hl,
                                       void* out_proof) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, pairing == PCDGPU_MNT4_298 || pairing == PCDGPU_MNT6_298, "unknown pairing id");
  CHECK_ARG(ctx, d_partials_hl && out_proof && world >= 1, "bad argument");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  G16Misc m;
  PCD_TRY(g16_misc(ctx, pairing, &m));
  // h, l' -> sums1[0], sums1[1]
  PCD_TRY(groth16_sum_partials(ctx, pairing, d_partials_hl, nullptr, world, 2, 0, m.sums1, nullptr));
  PCD_TRY(groth16_finish(ctx, pairing, m.sums1, m.d_proof + m.a1 + m.a2));
  PCD_CUDA(ctx, cudaMemcpyAsync(out_proof, m.d_proof, 2 * m.a1 + m.a2, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int pcdgpu_groth16_prove(pcdgpu_ctx* ctx, const pcdgpu_pk* pk, const pcdgpu_r1cs* r1cs, const void* z, const void* r,
                         const void* s, void* out_proof) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, pk && r1cs && z && r && s && out_proof, "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t nv = r1cs->num_inputs + r1cs->num_witness;
  void* dz;
  PCD_TRY(ctx->scratch(SLOT_Z, nv * 40, &dz));
  PCD_CUDA(ctx, cudaMemcpyAsync(dz, z, nv * 40, cudaMemcpyHostToDevice, ctx->stream));
  return pcdgpu_groth16_prove_dev(ctx, pk, r1cs, dz, r, s, out_proof);
}

int pcdgpu_serialize_proof(pcdgpu_ctx* ctx, int pairing, const void* proof_affine, uint8_t* out, size_t* out_len) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, pairing == PCDGPU_MNT4_298 || pairing == PCDGPU_MNT6_298, "unknown pairing id");
  CHECK_ARG(ctx, proof_affine && out && out_len, "null pointer");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  size_t in_bytes = 2 * msm_ops(g1_of(pairing))->affine_bytes + msm_ops(g2_of(pairing))->affine_bytes;
  size_t out_bytes = pairing == PCDGPU_MNT4_298 ? 152 : 190;
  void* d;
  PCD_TRY(ctx->scratch(SLOT_MISC, 8192, &d));
  unsigned char* d_out = (unsigned char*)d + 1024;
  PCD_CUDA(ctx, cudaMemcpyAsync(d, proof_affine, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  PCD_TRY(groth16_serialize(ctx, pairing, d, d_out));
  PCD_CUDA(ctx, cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *out_len = out_bytes;
  return 0;
}

int pcdgpu_profile_enable(pcdgpu_ctx* ctx, int on) {
  if (!ctx) return PCDGPU_E_ARG;
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (on && !ctx->prof_pinned) PCD_CUDA(ctx, cudaMallocHost((void**)&ctx->prof_pinned, 4096 * sizeof(unsigned)));
  ctx->profiling = on != 0;
  ctx->spans_used = 0;
  ctx->launches = 0;
  return 0;
}

int pcdgpu_profile_read(pcdgpu_ctx* ctx, double* ms, double* units, uint64_t* spans, uint64_t* launches) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, ms && units && spans && launches, "null pointer");
  PCD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < PROF_NSLOT; i++) {
    ms[i] = 0;
    units[i] = 0;
    spans[i] = 0;
  }
  for (size_t i = 0; i < ctx->spans_used; i++) {
    ProfSpan& sp = ctx->spans[i];
    float t = 0;
    if (cudaEventElapsedTime(&t, sp.a, sp.b) != cudaSuccess) continue;
    ms[sp.slot] += t;
    units[sp.slot] += sp.units_pinned >= 0 ? (double)ctx->prof_pinned[sp.units_pinned] : sp.units;
    spans[sp.slot] += 1;
  }
  *launches = ctx->launches;
  ctx->spans_used = 0;
  ctx->launches = 0;
  return 0;
}

int pcdgpu_profile_timeline(pcdgpu_ctx* ctx, double* t0_ms, double* t1_ms, int* cls, size_t cap, size_t* count) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, t0_ms && t1_ms && cls && count, "null pointer");
  PCD_CUDA(ctx, cudaDeviceSynchronize());
  size_t n = ctx->spans_used < cap ? ctx->spans_used : cap;
  for (size_t i = 0; i < n; i++) {
    float a = 0, b = 0;
    cudaEventElapsedTime(&a, ctx->spans[0].a, ctx->spans[i].a);
    cudaEventElapsedTime(&b, ctx->spans[0].a, ctx->spans[i].b);
    t0_ms[i] = a;
    t1_ms[i] = b;
    cls[i] = ctx->spans[i].slot;
  }
  *count = n;
  return 0;
}

int pcdgpu_set_msm_side_by_side(pcdgpu_ctx* ctx, int on) {
  if (!ctx) return PCDGPU_E_ARG;
  ctx->in_proof = on != 0;
  ctx->key_upload = on != 0;
  return 0;
}

int pcdgpu_bench_imad(pcdgpu_ctx* ctx, int modmul, int iters, double* out_ops_per_s, double* out_ms) {
  if (!ctx) return PCDGPU_E_ARG;
  CHECK_ARG(ctx, out_ops_per_s && out_ms && iters > 0, "bad argument");
  PCD_CUDA(ctx, cudaSetDevice(ctx->device));
  return bench_imad(ctx, modmul, iters, out_ops_per_s, out_ms);
}

}  // extern "C"
