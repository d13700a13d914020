// msm.cuh -- variable-base multi-scalar multiplication (Pippenger bucket method), templated on the
// curve.  Replaces ark-ec VariableBaseMSM::multi_scalar_mul under
// ark-groth16's create_proof_with_reduction; reached from
// /root/reference/src/ec_cycle_pcd/mod.rs:171,179.  The result is compared with the reference
// after into_affine(), which is representation independent, so the GPU is free to choose its own
// decomposition:
//
//   1. msm_digits     one thread per scalar: leave Montgomery form if asked, signed base-2^c
//                     recoding (digits in [-2^(c-1), 2^(c-1)]: half the buckets of arkworks'
//                     unsigned windows), histogram of bucket sizes with global atomics;
//   2. exclusive scan of the histogram (cub::DeviceScan, ~10^5..10^6 counters);
//   3. msm_scatter    counting-sort the (point, sign) entries by bucket;
//   4. msm_accumulate one thread per bucket walks its entries with XYZZ mixed additions
//                     (8M + 2S); buckets beyond HEAVY entries (e.g. the "scalar == 1" bucket of a
//                     real witness) are left to msm_accumulate_heavy, one CTA each;
//   5. msm_reduce     per window sum_b (b+1) B_b: every thread takes L consecutive buckets with the
//                     running-sum trick and adds [t L] * (its plain sum);
//   6. msm_window_sum / msm_horner  tree-sum per window, then the 2^c Horner chain.
//
// With precomputed bases (pcdgpu_bases_upload(..., precompute = 1)) the table holds 2^(c j) P for
// every window j, all windows share ONE bucket set and step 6's doubling chain disappears.
#pragma once
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "vecio.cuh"

static constexpr int MSM_SCALAR_BITS = 298;
static constexpr int MSM_HEAVY = 1024;      // entries per bucket handled by one thread
static constexpr int MSM_HEAVY_THREADS = 128;
static constexpr int MSM_MAX_HEAVY = 4096;  // size of the heavy-bucket list

// ---- 1. digits ------------------------------------------------------------------------------
// dig[w * n + i] = signed digit of scalar i in window w; counts[bucket]++ for non-zero digits.
// shared != 0: all windows share one bucket set (precomputed bases).
template <class SP>
__global__ void msm_digits_kernel(const u32* __restrict__ scalars, int mont, size_t n, int c, int nwin, int shared,
                                  int* __restrict__ dig, u32* __restrict__ counts) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<SP> k;
  const uint2* p = reinterpret_cast<const uint2*>(scalars + i * 10);
#pragma unroll
  for (int j = 0; j < 5; j++) {
    uint2 v = __ldg(p + j);
    k.l[2 * j] = v.x;
    k.l[2 * j + 1] = v.y;
  }
  if (mont) k = k.from_mont();
  u32 w32[12];
#pragma unroll
  for (int j = 0; j < 10; j++) w32[j] = k.l[j];
  w32[10] = 0;
  w32[11] = 0;
  const u32 B = 1u << (c - 1);
  const u32 mask = (1u << c) - 1;
  u32 carry = 0;
  for (int w = 0; w < nwin; w++) {
    int bit = w * c;
    int limb = bit >> 5, off = bit & 31;
    u32 d = 0;
    if (limb < 10) {
      unsigned long long v = w32[limb] | ((unsigned long long)w32[limb + 1] << 32);
      d = (u32)(v >> off) & mask;
    }
    d += carry;
    int sd;
    if (d > B) {
      sd = (int)d - (int)(1u << c);
      carry = 1;
    } else {
      sd = (int)d;
      carry = 0;
    }
    dig[(size_t)w * n + i] = sd;
    if (sd != 0) {
      u32 b = (u32)(sd < 0 ? -sd : sd) - 1;
      atomicAdd(&counts[(shared ? 0 : (size_t)w * B) + b], 1u);
    }
  }
}

// ---- 3. scatter -----------------------------------------------------------------------------
// entry = (shared ? w * stride + offset + i : i) | sign << 31  (index into the base table)
static __global__ void msm_scatter_kernel(const int* __restrict__ dig, size_t n, int c, int nwin, int shared, size_t stride,
                                   size_t offset, u32* __restrict__ cursor, u32* __restrict__ entries) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * (size_t)nwin) return;
  int sd = dig[t];
  if (sd == 0) return;
  size_t w = t / n, i = t - w * n;
  const u32 B = 1u << (c - 1);
  u32 b = (u32)(sd < 0 ? -sd : sd) - 1;
  u32 pos = atomicAdd(&cursor[(shared ? 0 : w * B) + b], 1u);
  entries[pos] = (u32)(shared ? w * stride + offset + i : i) | (sd < 0 ? 0x80000000u : 0u);
}

// ---- 4. bucket accumulation -------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(128) msm_accumulate_kernel(const void* __restrict__ bases,
                                                             const u32* __restrict__ offsets,
                                                             const u32* __restrict__ entries, size_t nbuckets,
                                                             void* __restrict__ buckets, u32* __restrict__ heavy) {
  typedef typename C::F F;
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nbuckets) return;
  u32 lo = offsets[g], hi = offsets[g + 1];
  XYZZ<C> acc = XYZZ<C>::inf();
  if (hi - lo > (u32)MSM_HEAVY) {
    u32 slot = atomicAdd(&heavy[0], 1u);
    if (slot < (u32)MSM_MAX_HEAVY) {
      heavy[1 + slot] = (u32)g;
      return;  // msm_accumulate_heavy writes the bucket
    }
    // list full: fall through and do it serially (correct, slow)
  }
  for (u32 e = lo; e < hi; e++) {
    u32 ent = entries[e];
    AffinePoint<F> p = ld_vec<AffinePoint<F>>(bases, ent & 0x7fffffffu);
    if (ent >> 31) p.y = p.y.neg();
    acc.madd(p);
  }
  st_vec(buckets, g, acc);
}

template <class C>
__global__ void __launch_bounds__(MSM_HEAVY_THREADS) msm_accumulate_heavy_kernel(
    const void* __restrict__ bases, const u32* __restrict__ offsets, const u32* __restrict__ entries,
    void* __restrict__ buckets, const u32* __restrict__ heavy) {
  typedef typename C::F F;
  extern __shared__ uint4 sm4[];
  XYZZ<C>* sm = reinterpret_cast<XYZZ<C>*>(sm4);
  u32 nheavy = heavy[0] < (u32)MSM_MAX_HEAVY ? heavy[0] : (u32)MSM_MAX_HEAVY;
  for (u32 h = blockIdx.x; h < nheavy; h += gridDim.x) {
    u32 g = heavy[1 + h];
    u32 lo = offsets[g], hi = offsets[g + 1];
    XYZZ<C> acc = XYZZ<C>::inf();
    for (u32 e = lo + threadIdx.x; e < hi; e += MSM_HEAVY_THREADS) {
      u32 ent = entries[e];
      AffinePoint<F> p = ld_vec<AffinePoint<F>>(bases, ent & 0x7fffffffu);
      if (ent >> 31) p.y = p.y.neg();
      acc.madd(p);
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = MSM_HEAVY_THREADS / 2; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) {
        XYZZ<C> a = sm[threadIdx.x];
        a.add(sm[threadIdx.x + s]);
        sm[threadIdx.x] = a;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) st_vec(buckets, g, sm[0]);
    __syncthreads();
  }
}

// ---- 5. bucket reduction ----------------------------------------------------------------------
// thread t of window w: seg[w * T + t] = sum_{b in [tL, tL+L)} (b + 1) * bucket[w * B + b]
template <class C>
__global__ void __launch_bounds__(128) msm_reduce_kernel(const void* __restrict__ buckets, int c, int logL, int nwin,
                                                         void* __restrict__ seg) {
  const u32 B = 1u << (c - 1);
  const u32 L = 1u << logL;
  const u32 T = B >> logL;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)T * nwin) return;
  u32 w = (u32)(gid / T), t = (u32)(gid - (size_t)w * T);
  size_t base = (size_t)w * B + (size_t)t * L;
  XYZZ<C> run = XYZZ<C>::inf(), acc = XYZZ<C>::inf();
  for (int b = (int)L - 1; b >= 0; b--) {
    XYZZ<C> p = ld_vec_rw<XYZZ<C>>(buckets, base + b);
    run.add(p);
    acc.add(run);
  }
  // acc = sum (b - tL + 1) B_b ; add [tL] * run
  u32 k = t * L;
  if (k != 0 && !run.is_inf()) {
    XYZZ<C> m = XYZZ<C>::mul(run, &k, 1);
    acc.add(m);
  }
  st_vec(seg, gid, acc);
}

// ---- 6. per-window tree sum and Horner ----------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(MSM_HEAVY_THREADS) msm_window_sum_kernel(const void* __restrict__ seg, u32 T,
                                                                           void* __restrict__ wsum) {
  extern __shared__ uint4 sm4[];
  XYZZ<C>* sm = reinterpret_cast<XYZZ<C>*>(sm4);
  u32 w = blockIdx.x;
  XYZZ<C> acc = XYZZ<C>::inf();
  for (u32 t = threadIdx.x; t < T; t += MSM_HEAVY_THREADS) {
    XYZZ<C> p = ld_vec_rw<XYZZ<C>>(seg, (size_t)w * T + t);
    acc.add(p);
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  for (int s = MSM_HEAVY_THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      XYZZ<C> a = sm[threadIdx.x];
      a.add(sm[threadIdx.x + s]);
      sm[threadIdx.x] = a;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_vec(wsum, w, sm[0]);
}

// out = sum_w 2^(c w) wsum[w]  (+ *addend if given)
template <class C>
__global__ void msm_horner_kernel(const void* __restrict__ wsum, int c, int nwin, void* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  XYZZ<C> total = ld_vec_rw<XYZZ<C>>(wsum, nwin - 1);
  for (int w = nwin - 2; w >= 0; w--) {
    for (int i = 0; i < c; i++) total = total.dbl();
    XYZZ<C> p = ld_vec_rw<XYZZ<C>>(wsum, w);
    total.add(p);
  }
  st_vec(out, 0, total);
}

// ---- helpers: normalisation, fixed-base multiplication, table precomputation ---------------------
template <class C>
__global__ void xyzz_to_affine_kernel(const void* __restrict__ in, size_t n, void* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<C> p = ld_vec_rw<XYZZ<C>>(in, i);
  st_vec(out, i, p.to_affine());
}

// sum of n xyzz points -> affine (single thread; n is a handful of per-GPU partials)
template <class C>
__global__ void xyzz_sum_kernel(const void* __restrict__ in, size_t n, void* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  XYZZ<C> acc = XYZZ<C>::inf();
  for (size_t i = 0; i < n; i++) {
    XYZZ<C> p = ld_vec_rw<XYZZ<C>>(in, i);
    acc.add(p);
  }
  st_vec(out, 0, acc.to_affine());
}

// table[j * 15 + (d - 1)] = d * 16^j * base, j < 75, d in 1..15 (4-bit fixed windows), affine
template <class C>
__global__ void fixed_table_kernel(const void* __restrict__ base, void* __restrict__ table) {
  typedef typename C::F F;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= 75) return;
  AffinePoint<F> b = ld_vec<AffinePoint<F>>(base, 0);
  XYZZ<C> p = XYZZ<C>::from_affine(b);
  for (int i = 0; i < 4 * j; i++) p = p.dbl();
  AffinePoint<F> pj = p.to_affine();
  XYZZ<C> acc = XYZZ<C>::inf();
  for (int d = 1; d <= 15; d++) {
    acc.madd(pj);
    st_vec(table, (size_t)j * 15 + (d - 1), acc.to_affine());
  }
}
template <class C>
__global__ void __launch_bounds__(128) fixed_mul_kernel(const void* __restrict__ table,
                                                        const u32* __restrict__ scalars, size_t n,
                                                        void* __restrict__ out) {
  typedef typename C::F F;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  u32 k[10];
  const uint2* p = reinterpret_cast<const uint2*>(scalars + i * 10);
#pragma unroll
  for (int j = 0; j < 5; j++) {
    uint2 v = __ldg(p + j);
    k[2 * j] = v.x;
    k[2 * j + 1] = v.y;
  }
  XYZZ<C> acc = XYZZ<C>::inf();
  for (int j = 0; j < 75; j++) {
    u32 d = (k[j >> 3] >> ((j & 7) * 4)) & 15u;
    if (d) acc.madd(ld_vec<AffinePoint<F>>(table, (size_t)j * 15 + (d - 1)));
  }
  st_vec(out, i, acc.to_affine());
}

// pre[j * n + i] = 2^(c j) * bases[i]  (affine), j < nwin
template <class C>
__global__ void __launch_bounds__(128) precompute_kernel(const void* __restrict__ bases, size_t n, int c, int nwin,
                                                         void* __restrict__ pre) {
  typedef typename C::F F;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  AffinePoint<F> a = ld_vec<AffinePoint<F>>(bases, i);
  st_vec(pre, i, a);
  XYZZ<C> p = XYZZ<C>::from_affine(a);
  for (int j = 1; j < nwin; j++) {
    for (int b = 0; b < c; b++) p = p.dbl();
    a = p.to_affine();
    st_vec(pre, (size_t)j * n + i, a);
    p = XYZZ<C>::from_affine(a);
  }
}

// ---- host driver --------------------------------------------------------------------------------
struct MsmPlan {
  int c, nwin, shared;
  size_t stride, offset;  // shared (precomputed) tables: row pitch in points and first point used
};

// bases: n affine points (or the precomputed table nwin x n when plan.shared); result: one xyzz
// point at d_out.
template <class C>
static int msm_run(pcdgpu_ctx* ctx, const void* d_bases, const void* d_scalars, int scalars_mont, size_t n,
                   MsmPlan plan, void* d_out) {
  typedef typename C::ScalarParams SP;
  cudaStream_t st = ctx->stream;
  if (n == 0) {
    PCD_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(XYZZ<C>), st));
    return 0;
  }
  const int c = plan.c, nwin = plan.nwin, shared = plan.shared;
  if ((size_t)nwin * (shared ? plan.stride : n) >= ((size_t)1 << 31)) {
    ctx->set_error("MSM of %zu points x %d windows exceeds the 31-bit entry index", n, nwin);
    return PCDGPU_E_ARG;
  }
  const size_t B = (size_t)1 << (c - 1);
  const size_t nbuckets = shared ? B : B * nwin;
  const int rwin = shared ? 1 : nwin;  // windows seen by the reduction
  void *dig, *ent, *cnt, *bkt, *seg, *cub_tmp;
  PCD_TRY(ctx->scratch(SLOT_MSM_DIG, (size_t)nwin * n * 4, &dig));
  PCD_TRY(ctx->scratch(SLOT_MSM_ENT, (size_t)nwin * n * 4, &ent));
  // counts | offsets | cursor, each nbuckets + 1, then the heavy list
  size_t cstride = (nbuckets + 1 + 3) & ~(size_t)3;
  PCD_TRY(ctx->scratch(SLOT_MSM_CNT, (3 * cstride + MSM_MAX_HEAVY + 4) * 4, &cnt));
  PCD_TRY(ctx->scratch(SLOT_MSM_BKT, nbuckets * sizeof(XYZZ<C>), &bkt));
  u32* counts = (u32*)cnt;
  u32* offsets = counts + cstride;
  u32* cursor = offsets + cstride;
  u32* heavy = cursor + cstride;
  PCD_CUDA(ctx, cudaMemsetAsync(counts, 0, cstride * 4, st));
  PCD_CUDA(ctx, cudaMemsetAsync(heavy, 0, 4, st));
  const int acc_slot = sizeof(typename C::F) > 40 ? PROF_MSM_ACC_G2 : PROF_MSM_ACC_G1;
  int ps = ctx->prof_begin(PROF_MSM_SORT, (double)n * nwin);
  ctx->launches += 7 + 3;  // seven kernels of this file + cub's scan (init, scan) + one d2d copy
  msm_digits_kernel<SP><<<(unsigned)((n + 127) / 128), 128, 0, st>>>((const u32*)d_scalars, scalars_mont, n, c, nwin,
                                                                    shared, (int*)dig, counts);
  PCD_CUDA(ctx, cudaGetLastError());
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, counts, offsets, (int)(nbuckets + 1), st);
  PCD_TRY(ctx->scratch(SLOT_CUB, cub_bytes + 16, &cub_tmp));
  PCD_CUDA(ctx, cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, counts, offsets, (int)(nbuckets + 1), st));
  PCD_CUDA(ctx, cudaMemcpyAsync(cursor, offsets, (nbuckets + 1) * 4, cudaMemcpyDeviceToDevice, st));
  size_t total = (size_t)nwin * n;
  msm_scatter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const int*)dig, n, c, nwin, shared,
                                                                      plan.stride, plan.offset, cursor, (u32*)ent);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->prof_end(ps);
  ps = ctx->prof_begin(acc_slot, (double)n * nwin);
  if (ps >= 0 && ctx->prof_pinned) {  // exact number of bucket entries (non-zero digits) for the roofline
    ctx->spans[ps].units_pinned = ps;
    cudaMemcpyAsync(ctx->prof_pinned + ps, offsets + nbuckets, 4, cudaMemcpyDeviceToHost, st);
  }
  msm_accumulate_kernel<C><<<(unsigned)((nbuckets + 127) / 128), 128, 0, st>>>(d_bases, offsets, (const u32*)ent,
                                                                               nbuckets, bkt, heavy);
  PCD_CUDA(ctx, cudaGetLastError());
  size_t heavy_smem = MSM_HEAVY_THREADS * sizeof(XYZZ<C>);
  PCD_CUDA(ctx, cudaFuncSetAttribute(msm_accumulate_heavy_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)heavy_smem));
  msm_accumulate_heavy_kernel<C><<<ctx->sm_count, MSM_HEAVY_THREADS, heavy_smem, st>>>(d_bases, offsets,
                                                                                      (const u32*)ent, bkt, heavy);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->prof_end(ps);
  ps = ctx->prof_begin(PROF_MSM_REDUCE, (double)nbuckets);
  // reduction: L buckets per thread
  int logL = c - 1 >= 12 ? 4 : (c - 1 >= 6 ? 2 : 0);
  size_t T = B >> logL;
  PCD_TRY(ctx->scratch(SLOT_MSM_SEG, (T * rwin + rwin + 2) * sizeof(XYZZ<C>), &seg));
  void* wsum = (char*)seg + T * rwin * sizeof(XYZZ<C>);
  msm_reduce_kernel<C><<<(unsigned)((T * rwin + 127) / 128), 128, 0, st>>>(bkt, c, logL, rwin, seg);
  PCD_CUDA(ctx, cudaGetLastError());
  PCD_CUDA(ctx, cudaFuncSetAttribute(msm_window_sum_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)heavy_smem));
  msm_window_sum_kernel<C><<<rwin, MSM_HEAVY_THREADS, heavy_smem, st>>>(seg, (u32)T, wsum);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->prof_end(ps);
  ps = ctx->prof_begin(PROF_MSM_TAIL, (double)rwin);
  msm_horner_kernel<C><<<1, 32, 0, st>>>(wsum, c, rwin, d_out);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->prof_end(ps);
  return 0;
}
