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
//   5. wec_reduce     per window sum_b (b+1) B_b: every GROUP of lanes (wec.cuh) takes L consecutive buckets
//                     with the running-sum trick and adds [t L] * (its plain sum); the groups of a CTA and then
//                     wec_sum fold the contributions;
//   6. wec_combine    the 2^c Horner chain over the window sums (one window with precomputed tables).
//
// With precomputed bases (pcdgpu_bases_upload(..., precompute = 1)) the table holds 2^(c j) P for
// every window j, all windows share ONE bucket set and step 6's doubling chain disappears.
#pragma once
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"
#include "vecio.cuh"
#include "wec.cuh"

static constexpr int MSM_SCALAR_BITS = 298;
static constexpr int MSM_HEAVY = 1024;      // upper limit of the entries one thread may walk (see heavy_thr)
static constexpr int MSM_HEAVY_THREADS = 128;
static constexpr int MSM_MAX_HEAVY = 16384;  // size of the heavy-bucket list
static constexpr int MSM_HEAVY_CHUNK = 1024;  // entries of a heavy bucket summed by one CTA at a time
static constexpr int MSM_REDUCE_LOGL = 3;     // bucket reduction: 2^3 buckets per group of lanes
static constexpr int MSM_SUM_PER_GROUP = 4;   // wec_sum_kernel: points added serially by a group before the CTA tree
static constexpr int MSM_MAX_SPLIT = 4;       // parts a bucket's entry list may be cut into (msm_accumulate_kernel)

// Window layout.  Plain MSMs use windows of c bits from bit 0 up.  With precomputed tables all windows
// share one bucket set, and a short top window (299 mod c real bits) would pile every scalar's top digit
// into the first few buckets (c = 17 at 2^18: 2^10 buckets receive 2^18 / 2^10 extra entries each, which
// tripled the accumulation time); there the 299 bits are cut into nwin windows of equal width +- 1.
__host__ __device__ __forceinline__ int msm_win_start(int j, int c, int nwin, int balanced) {
  if (!balanced) return j * c;
  int base = (MSM_SCALAR_BITS + 1) / nwin, rem = (MSM_SCALAR_BITS + 1) % nwin;
  return j * base + (j < rem ? j : rem);
}

// ---- 1. digits ------------------------------------------------------------------------------
// dig[w * n + i] = signed digit of scalar i in window w; counts[bucket]++ for non-zero digits.
// shared != 0: all windows share one bucket set (precomputed bases).
// Scalars i < n_main come from `scalars` (Montgomery form if mont), the n - n_main trailing ones from
// `extra` (always plain integers): the per-proof pairs (delta, r), (a_query[0], 1), ... that the
// Groth16 prover folds into its MSMs.
// A base point at infinity (x = y = 0: the query point of a variable that never occurs in the matrix -- common in
// b_query) gets all-zero digits: left in, its bucket entries cost a full mixed addition of WARP time each (the lane
// that holds it sits out while its neighbours add), measured as 10 - 16 active lanes of 32 in the a / b_g1 MSMs of
// the synthetic key.  bases / point_u4 / base_first: row 0 of the (table of) base points, 16-byte words per point,
// index of scalar 0's point.
template <class SP>
__global__ void msm_digits_kernel(const u32* __restrict__ scalars, int mont, size_t n_main,
                                  const u32* __restrict__ extra, size_t n, int c, int nwin, int shared,
                                  int* __restrict__ dig, u32* __restrict__ counts, const uint4* __restrict__ bases,
                                  int point_u4, size_t base_first, u32 unit_k) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  {
    const uint4* bp = bases + (base_first + i) * (size_t)point_u4;
    u32 any = 0;
    for (int q = 0; q < point_u4; q++) {
      uint4 v = __ldg(bp + q);
      any |= v.x | v.y | v.z | v.w;
    }
    if (any == 0) {
      for (int w = 0; w < nwin; w++) dig[(size_t)w * n + i] = 0;
      return;
    }
  }
  Fp<SP> k;
  const bool is_extra = i >= n_main;
  const uint2* p = reinterpret_cast<const uint2*>(is_extra ? extra + (i - n_main) * 10 : scalars + i * 10);
  if (is_extra) mont = 0;
#pragma unroll
  for (int j = 0; j < 5; j++) {
    uint2 v = __ldg(p + j);
    k.l[2 * j] = v.x;
    k.l[2 * j + 1] = v.y;
  }
  if (mont) k = k.from_mont();
  const u32 B = 1u << (c - 1);  // buckets per window (bucket-set pitch)
  // Unit buckets (unit_k > 0, shared bucket set only).  A witness is full of bits: every scalar equal to 1 would land in
  // bucket 0 of window 0 -- a fifth of all points of a real assignment in ONE bucket, which then has to go through the
  // latency-bound heavy-bucket kernels (measured: 2.1 ms of serialised kernel time per main proof for 3 % of the entries).
  // Instead scalar 1 of point i goes to one of unit_k extra buckets behind the B weighted ones (digit B + 1 + i mod
  // unit_k), which the accumulate kernel walks like any other bucket and the reduction adds with weight 1.  ark-ec's
  // Pippenger has the same special case ("we only process unit scalars once in the first window").
  if (unit_k) {
    u32 rest = 0;
#pragma unroll
    for (int j = 1; j < 10; j++) rest |= k.l[j];
    if (rest == 0 && k.l[0] == 1u) {
      const u32 b = B + (u32)(i & (size_t)(unit_k - 1));
      dig[i] = (int)(b + 1);
      for (int w = 1; w < nwin; w++) dig[(size_t)w * n + i] = 0;
      atomicAdd(&counts[b], 1u);
      return;
    }
  }
  u32 w32[12];
#pragma unroll
  for (int j = 0; j < 10; j++) w32[j] = k.l[j];
  w32[10] = 0;
  w32[11] = 0;
  u32 carry = 0;
  for (int w = 0; w < nwin; w++) {
    int bit = msm_win_start(w, c, nwin, shared);
    int cw = (w + 1 < nwin ? msm_win_start(w + 1, c, nwin, shared) : (shared ? MSM_SCALAR_BITS + 1 : bit + c)) - bit;
    const u32 mask = (1u << cw) - 1;
    const u32 half = 1u << (cw - 1);
    int limb = bit >> 5, off = bit & 31;
    u32 d = 0;
    if (limb < 10) {
      unsigned long long v = w32[limb] | ((unsigned long long)w32[limb + 1] << 32);
      d = (u32)(v >> off) & mask;
    }
    d += carry;
    int sd;
    if (d > half) {
      sd = (int)d - (int)(1u << cw);
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
// Buckets are visited in order of decreasing size (perm, from a radix sort of the clamped sizes): the 32
// buckets of a warp then hold (almost) the same number of entries, so no lane idles while its
// neighbours finish -- with buckets in natural order only 70 % of the lanes were active (ncu, round 1).
static __global__ void msm_sizekey_kernel(const u32* __restrict__ counts, size_t nbuckets, u32* __restrict__ key,
                                          u32* __restrict__ id) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nbuckets) return;
  u32 c = counts[g];
  key[g] = c < 2047u ? c : 2047u;
  id[g] = (u32)g;
}

// profiling only: out[0] = entries the accumulate kernel walked, out[1] = entries left to the heavy-bucket kernels
static __global__ void msm_prof_entries_kernel(const u32* __restrict__ heavy, const u32* __restrict__ offsets, size_t nbuckets,
                                               u32* __restrict__ out) {
  u32 nheavy = heavy[0] < (u32)MSM_MAX_HEAVY ? heavy[0] : (u32)MSM_MAX_HEAVY;
  u32 h = 0;
  for (u32 i = threadIdx.x; i < nheavy; i += 32) h += offsets[heavy[1 + i] + 1] - offsets[heavy[1 + i]];
  for (int d = 16; d > 0; d >>= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
  if (threadIdx.x == 0) {
    out[0] = offsets[nbuckets] - h;
    out[1] = h;
  }
}

// The heavy-bucket threshold the host passes is derived from n x windows entries; the REAL count (non-zero digits of
// points that are not at infinity) is only known on the device and can be several times smaller for a witness (bits,
// zero products, absent variables).  A bucket of a few hundred copies of one repeated value (-1, 2, ...: one such bucket
// per window) then sits just under the threshold and ONE thread walks it while the rest of the grid has long finished:
// measured on the helper proof's G2 MSM (Fq3, 0.36 M real entries of 1.3 M nominal): accumulate 9.2 ms with whole
// buckets, 5.6 / 3.6 ms with 2 / 4 parts -- inversely proportional, i.e. the time of the longest item.  So the kernel
// lowers the threshold to `split` x half the entries a resident work slot gets when the real work is spread evenly.
// ... but never below twice the real average bucket (with few buckets for the resident threads the average bucket is
// larger than a slot's share and `split` is what spreads it: without this floor every bucket of the dense h MSM at
// 2^16 went to the heavy path, 18 ms).
__device__ __forceinline__ u32 msm_dynamic_heavy_thr(u32 host_thr, u32 real_total, u32 resident_items, size_t nbuckets,
                                                     u32 split) {
  u32 share = real_total / (resident_items ? resident_items : 1u);
  u32 item = share / 2 > 8u ? share / 2 : 8u;
  u32 dyn = split * item;
  u32 floor_thr = 2u * (u32)(real_total / nbuckets) + 8u;
  if (dyn < floor_thr) dyn = floor_thr;
  if (dyn < 24u) dyn = 24u;
  return dyn < host_thr ? dyn : host_thr;
}

// Small bucket sets (<= 1024 buckets: the default-circuit proofs' MSMs, ~10^3 points at c = 7) get their offsets, cursors
// and visiting order from ONE single-CTA kernel instead of cub's scan + radix sort (seven launches): these MSMs are
// bound by the host's enqueue rate (seven MSMs per proof, ~75 us of launches each before this).  The order is the same
// rule -- decreasing clamped size, ties by bucket index -- computed as a rank by comparison (n^2 / 1024 per thread).
static __global__ void __launch_bounds__(1024) msm_plan_small_kernel(const u32* __restrict__ counts, u32 nbuckets,
                                                                      u32* __restrict__ offsets, u32* __restrict__ cursor,
                                                                      u32* __restrict__ perm, u32* __restrict__ heavy,
                                                                      u32* __restrict__ queue) {
  __shared__ u32 key[1024];
  __shared__ u32 warp_sum[32];
  const u32 t = threadIdx.x;
  const u32 c = t < nbuckets ? counts[t] : 0u;
  key[t] = c < 2047u ? c : 2047u;
  // exclusive scan of the counts (nbuckets + 1 outputs)
  u32 v = c;
  for (int d = 1; d < 32; d <<= 1) {
    u32 o = __shfl_up_sync(0xffffffffu, v, d);
    if ((t & 31u) >= (u32)d) v += o;
  }
  if ((t & 31u) == 31u) warp_sum[t >> 5] = v;
  __syncthreads();
  if (t < 32) {
    u32 w = warp_sum[t];
    for (int d = 1; d < 32; d <<= 1) {
      u32 o = __shfl_up_sync(0xffffffffu, w, d);
      if (t >= (u32)d) w += o;
    }
    warp_sum[t] = w;
  }
  __syncthreads();
  const u32 incl = v + ((t >> 5) ? warp_sum[(t >> 5) - 1] : 0u);
  if (t < nbuckets) {
    offsets[t] = incl - c;
    cursor[t] = incl - c;
    if (t == nbuckets - 1) {
      offsets[nbuckets] = incl;
      cursor[nbuckets] = incl;
    }
    u32 rank = 0;
    const u32 mine = key[t];
    for (u32 j = 0; j < nbuckets; j++) {
      const u32 kj = key[j];
      rank += (kj > mine || (kj == mine && j < t)) ? 1u : 0u;
    }
    perm[rank] = t;
  }
  if (t == 0) {
    heavy[0] = 0;
    queue[0] = 0;
  }
}

// Persistent warps pull 32 buckets at a time from a global queue (dynamic scheduling): with one thread per
// bucket and a plain grid the kernel ran in a few "waves" of equally long threads, and its time was
// the wave count rounded up (c = 18 at 2^20: 2.3 waves cost 3).
// PRE: the bases are a precomputed window table (same point encoding; kept as a template parameter because the two
// kernels are tuned and profiled separately)
template <class C, bool PRE>
__device__ __forceinline__ AffinePoint<typename C::F> ld_base(const void* bases, size_t idx) {
  return ld_vec<AffinePoint<typename C::F>>(bases, idx);
}

// `split` > 1: every bucket's entry list is cut into `split` equal parts walked by different threads (their sums land in
// partials[g * split + part] and msm_fold_parts adds them up).  With one thread per bucket and equally long buckets the
// kernel's time is the number of bucket "waves" rounded UP (2^17 buckets on 296 x 128 threads: 3.46 -> 4); parts make
// the waves short enough for the rounding not to matter.
template <class C, bool PRE>
__global__ void __launch_bounds__(128) msm_accumulate_kernel(const void* __restrict__ bases,
                                                             const u32* __restrict__ offsets,
                                                             const u32* __restrict__ entries,
                                                             const u32* __restrict__ perm, size_t nbuckets,
                                                             void* __restrict__ buckets, u32* __restrict__ heavy,
                                                             u32* __restrict__ queue, u32 heavy_thr, u32 split,
                                                             u32* __restrict__ hflag, u32 quantum, u32 persistent_from,
                                                             u32 resident_items, u32 nweighted) {
  typedef typename C::F F;
  typedef C CF;
  typedef typename CF::F FF;
  const unsigned lane = threadIdx.x & 31;
  const size_t nitems = nbuckets * split;
  // CTAs below `persistent_from` retire once their warps have walked `quantum` entries per lane: every retirement is a
  // point where the block scheduler can start a higher-priority lane's kernel (witness map, sorts, the G2 chain); the
  // last CTAs of the grid stay until the queue is empty, so all the work is always done
  const bool may_retire = blockIdx.x < persistent_from;
  heavy_thr = msm_dynamic_heavy_thr(heavy_thr, offsets[nbuckets], resident_items, nbuckets, split);
  u32 walked = 0;
  for (;;) {
    if (may_retire && __shfl_sync(0xffffffffu, walked, 0) >= quantum) break;
    u32 first = 0;
    if (lane == 0) first = atomicAdd(queue, 32u);
    first = __shfl_sync(0xffffffffu, first, 0);
    if (first >= nitems) break;
    size_t t = (size_t)first + lane;
    if (t >= nitems) continue;
    const u32 part = (u32)(t % split);
    size_t g = perm[t / split];
    u32 lo = offsets[g], hi = offsets[g + 1];
    XYZZ<CF> acc = XYZZ<CF>::inf();
    if (hi - lo > heavy_thr && g < nweighted) {  // unit buckets (msm_digits_kernel) are sized by the host: never heavy
      if (part != 0) continue;  // part 0 speaks for the whole bucket
      u32 slot = atomicAdd(&heavy[0], 1u);
      if (slot < (u32)MSM_MAX_HEAVY) {
        heavy[1 + slot] = (u32)g;
        hflag[g] = 0xffffffffu;  // msm_accumulate_heavy writes the bucket; msm_fold_parts leaves it alone
        continue;
      }
      // list full: fall through and do it serially (correct, slow); the other parts stay infinity
      if (split > 1) {
        for (u32 q = 1; q < split; q++) st_vec(buckets, g * split + q, acc);
      }
    } else if (split > 1) {
      const u32 len = hi - lo;
      const u32 a = lo + (u32)(((unsigned long long)len * part) / split);
      hi = lo + (u32)(((unsigned long long)len * (part + 1)) / split);
      lo = a;
    }
    walked += hi - lo;
    for (u32 e = lo; e < hi; e++) {
      u32 ent = entries[e];
      AffinePoint<FF> p = ld_base<C, PRE>(bases, ent & 0x7fffffffu);
      if (ent >> 31) p.y = p.y.neg();
      acc.madd_impl(p);  // inline even for the G2 curves: one call level less in the hottest loop
    }
    st_vec(buckets, g * split + part, acc);
  }
}
// ---- the same walk with the point sliced over the three lanes of a group (Fq3: fp3s.cuh) -----------------------------
// Lane l of a group holds coefficient l of every coordinate: a group walks one work item, a warp ten (lanes 30, 31
// idle).  Registers per lane drop to a third of the one-thread form, so nothing spills and every product is inline.
template <class CS>
__device__ __forceinline__ AffinePoint<typename CS::F> ld_base_sliced(const void* bases, size_t idx, int l) {
  AffinePoint<typename CS::F> p;
  const char* src = reinterpret_cast<const char*>(bases) + idx * (size_t)240;
  const uint2* px = reinterpret_cast<const uint2*>(src + 40 * l);
  const uint2* py = reinterpret_cast<const uint2*>(src + 120 + 40 * l);
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint2 a = __ldg(px + i), b = __ldg(py + i);
    p.x.c.l[2 * i] = a.x;
    p.x.c.l[2 * i + 1] = a.y;
    p.y.c.l[2 * i] = b.x;
    p.y.c.l[2 * i + 1] = b.y;
  }
  return p;
}
template <class CS>
__device__ __forceinline__ void st_xyzz_sliced(void* buckets, size_t idx, const XYZZ<CS>& a, int l) {
  char* dst = reinterpret_cast<char*>(buckets) + idx * (size_t)480 + 40 * l;
  const FpR4* co[4] = {&a.x.c, &a.y.c, &a.zz.c, &a.zzz.c};
#pragma unroll
  for (int q = 0; q < 4; q++) {
    uint2* d = reinterpret_cast<uint2*>(dst + 120 * q);
#pragma unroll
    for (int i = 0; i < 5; i++) d[i] = make_uint2(co[q]->l[2 * i], co[q]->l[2 * i + 1]);
  }
}
template <class CS>
__device__ __forceinline__ XYZZ<CS> ld_xyzz_sliced(const void* buckets, size_t idx, int l) {
  XYZZ<CS> a;
  const char* src = reinterpret_cast<const char*>(buckets) + idx * (size_t)480 + 40 * l;
  FpR4* co[4] = {&a.x.c, &a.y.c, &a.zz.c, &a.zzz.c};
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint2* d = reinterpret_cast<const uint2*>(src + 120 * q);
#pragma unroll
    for (int i = 0; i < 5; i++) {
      uint2 v = d[i];
      co[q]->l[2 * i] = v.x;
      co[q]->l[2 * i + 1] = v.y;
    }
  }
  return a;
}

#ifndef PCD_SLICED_MIN_CTAS
#define PCD_SLICED_MIN_CTAS 3
#endif
template <class CS, bool PRE>
__global__ void __launch_bounds__(128, PCD_SLICED_MIN_CTAS) msm_accumulate_sliced_kernel(const void* __restrict__ bases,
                                                                    const u32* __restrict__ offsets,
                                                                    const u32* __restrict__ entries,
                                                                    const u32* __restrict__ perm, size_t nbuckets,
                                                                    void* __restrict__ buckets, u32* __restrict__ heavy,
                                                                    u32* __restrict__ queue, u32 heavy_thr, u32 split,
                                                                    u32* __restrict__ hflag, u32 quantum, u32 persistent_from,
                                                             u32 resident_items, u32 nweighted) {
  typedef typename CS::F FF;
  const unsigned lane = threadIdx.x & 31;
  const int l = (int)(lane % 3u);
  const unsigned grp = lane / 3u;  // 0..9 (10: the two idle lanes)
  const size_t nitems = nbuckets * split;
  const bool may_retire = blockIdx.x < persistent_from;  // see msm_accumulate_kernel
  heavy_thr = msm_dynamic_heavy_thr(heavy_thr, offsets[nbuckets], resident_items, nbuckets, split);
  u32 walked = 0;
  for (;;) {
    if (may_retire && __shfl_sync(0xffffffffu, walked, 0) >= quantum) break;
    u32 first = 0;
    if (lane == 0) first = atomicAdd(queue, 10u);
    first = __shfl_sync(0xffffffffu, first, 0);
    if (first >= nitems) break;
    if (grp >= 10) continue;
    size_t t = (size_t)first + grp;
    if (t >= nitems) continue;
    const u32 part = (u32)(t % split);
    size_t g = perm[t / split];
    u32 lo = offsets[g], hi = offsets[g + 1];
    XYZZ<CS> acc = XYZZ<CS>::inf();
    if (hi - lo > heavy_thr && g < nweighted) {  // unit buckets (msm_digits_kernel) are sized by the host: never heavy
      if (part != 0) continue;  // part 0 speaks for the whole bucket
      u32 slot = 0;
      if (l == 0) slot = atomicAdd(&heavy[0], 1u);
      slot = __shfl_sync(FF::gmask(), slot, (int)(lane - l));
      if (slot < (u32)MSM_MAX_HEAVY) {
        if (l == 0) {
          heavy[1 + slot] = (u32)g;
          hflag[g] = 0xffffffffu;
        }
        continue;
      }
      if (split > 1) {  // list full: walk it here (correct, slow); the other parts stay infinity
        for (u32 q = 1; q < split; q++) st_xyzz_sliced<CS>(buckets, g * split + q, acc, l);
      }
    } else if (split > 1) {
      const u32 len = hi - lo;
      const u32 a = lo + (u32)(((unsigned long long)len * part) / split);
      hi = lo + (u32)(((unsigned long long)len * (part + 1)) / split);
      lo = a;
    }
    walked += hi - lo;
    for (u32 e = lo; e < hi; e++) {
      u32 ent = entries[e];
      AffinePoint<FF> p = ld_base_sliced<CS>(bases, ent & 0x7fffffffu, l);
      if (ent >> 31) p.y = p.y.neg();
      acc.madd_impl(p);
    }
    st_xyzz_sliced<CS>(buckets, g * split + part, acc, l);
  }
}
// fold of the bucket parts, sliced (three lanes per bucket)
template <class CS>
__global__ void __launch_bounds__(96) msm_fold_parts_sliced_kernel(const void* __restrict__ partials, size_t nbuckets, u32 split,
                                                                   const u32* __restrict__ hflag, void* __restrict__ buckets) {
  // 96 threads = 3 warps of 10 groups: group index from the warp, so that a group never straddles two warps
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane >= 30) return;
  const int l = (int)(lane % 3u);
  size_t g = ((size_t)blockIdx.x * 3 + warp) * 10 + lane / 3u;
  if (g >= nbuckets || hflag[g] == 0xffffffffu) return;
  XYZZ<CS> acc = ld_xyzz_sliced<CS>(partials, g * split, l);
  for (u32 q = 1; q < split; q++) acc.add_impl(ld_xyzz_sliced<CS>(partials, g * split + q, l));
  st_xyzz_sliced<CS>(buckets, g, acc, l);
}

// buckets[g] = sum of the `split` partial sums of bucket g (unless the heavy path owns the bucket)
template <class C>
__global__ void __launch_bounds__(128) msm_fold_parts_kernel(const void* __restrict__ partials, size_t nbuckets, u32 split,
                                                             const u32* __restrict__ hflag, void* __restrict__ buckets) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nbuckets || hflag[g] == 0xffffffffu) return;
  XYZZ<C> acc = ld_vec_rw<XYZZ<C>>(partials, g * split);
  for (u32 q = 1; q < split; q++) acc.add(ld_vec_rw<XYZZ<C>>(partials, g * split + q));
  st_vec(buckets, g, acc);
}

// CTA-wide tree sum of one xyzz point per thread (shared memory: MSM_HEAVY_THREADS points, then the temporaries of
// the CTA's lane groups); the sum ends in sm[0].  The additions are lane-cooperative (wec.cuh: G lanes per addition,
// its cost the multiplicative depth of the formula): with one thread per addition the seven levels of this tree were
// seven full one-thread additions deep (G1 ~9 us, G2 over Fq2 ~27 us each) and made the heavy-bucket kernels -- a few
// thousand entries -- as long as an accumulate launch over millions (profile class msm_accumulate_tail).
template <class C>
__device__ __forceinline__ void cta_tree_sum(XYZZ<C>* sm, const XYZZ<C>& mine) {
  typedef Wec<C> WG;
  constexpr int NG = MSM_HEAVY_THREADS / WG::G;
  sm[threadIdx.x] = mine;
  u32* pts = reinterpret_cast<u32*>(sm);
  WG wg(pts + (size_t)MSM_HEAVY_THREADS * WG::PW + (size_t)(threadIdx.x / WG::G) * WG::NTW);
  const int g = (int)threadIdx.x / WG::G;
  for (int s = MSM_HEAVY_THREADS / 2; s > 0; s >>= 1) {
    __syncthreads();
    for (int i = g; i < s; i += NG) wg.add(pts + (size_t)i * WG::PW, pts + (size_t)(i + s) * WG::PW);
  }
  __syncthreads();
}
// dynamic shared memory of a kernel that calls cta_tree_sum
template <class C>
constexpr size_t cta_tree_smem() {
  return (size_t)MSM_HEAVY_THREADS * sizeof(XYZZ<C>) + (size_t)(MSM_HEAVY_THREADS / Wec<C>::G) * Wec<C>::NTW * 4;
}

// Heavy buckets (more than MSM_HEAVY entries: the "scalar == 1" bucket of a real witness holds a
// fifth of all points) are cut into chunks of MSM_HEAVY_CHUNK entries; every CTA of the grid takes
// chunks round-robin, sums one with a tree and appends (bucket, partial sum) to a list ...
template <class C, bool PRE>
__global__ void __launch_bounds__(MSM_HEAVY_THREADS) msm_accumulate_heavy_kernel(
    const void* __restrict__ bases, const u32* __restrict__ offsets, const u32* __restrict__ entries,
    const u32* __restrict__ heavy, u32* __restrict__ hp_count, u32* __restrict__ hp_id, void* __restrict__ hp_sum) {
  typedef typename C::F F;
  extern __shared__ uint4 sm4[];
  XYZZ<C>* sm = reinterpret_cast<XYZZ<C>*>(sm4);
  u32 nheavy = heavy[0] < (u32)MSM_MAX_HEAVY ? heavy[0] : (u32)MSM_MAX_HEAVY;
  u32 job = 0;  // running (bucket, chunk) index, identical in every CTA
  for (u32 h = 0; h < nheavy; h++) {
    u32 g = heavy[1 + h];
    u32 lo = offsets[g], hi = offsets[g + 1];
    u32 nchunks = (hi - lo + MSM_HEAVY_CHUNK - 1) / MSM_HEAVY_CHUNK;
    for (u32 ch = 0; ch < nchunks; ch++, job++) {
      if (job % gridDim.x != blockIdx.x) continue;
      u32 e0 = lo + ch * MSM_HEAVY_CHUNK;
      u32 e1 = e0 + MSM_HEAVY_CHUNK < hi ? e0 + MSM_HEAVY_CHUNK : hi;
      XYZZ<C> acc = XYZZ<C>::inf();
      for (u32 e = e0 + threadIdx.x; e < e1; e += MSM_HEAVY_THREADS) {
        u32 ent = entries[e];
        AffinePoint<typename C::F> p = ld_base<C, PRE>(bases, ent & 0x7fffffffu);
        if (ent >> 31) p.y = p.y.neg();
        acc.madd(p);  // out of line for G2 (inlining it here as well takes msm_c3.cu from 80 s to 5 min of ptxas)
      }
      cta_tree_sum<C>(sm, acc);
      if (threadIdx.x == 0) {
        u32 slot = atomicAdd(hp_count, 1u);
        hp_id[slot] = h;
        st_vec(hp_sum, slot, sm[0]);
      }
      __syncthreads();
    }
  }
}
// ... and one CTA per heavy bucket adds up that bucket's partial sums.
template <class C>
__global__ void __launch_bounds__(MSM_HEAVY_THREADS) msm_heavy_finish_kernel(
    const u32* __restrict__ heavy, const u32* __restrict__ hp_count, const u32* __restrict__ hp_id,
    const void* __restrict__ hp_sum, void* __restrict__ buckets) {
  extern __shared__ uint4 sm4[];
  XYZZ<C>* sm = reinterpret_cast<XYZZ<C>*>(sm4);
  u32 nheavy = heavy[0] < (u32)MSM_MAX_HEAVY ? heavy[0] : (u32)MSM_MAX_HEAVY;
  u32 np = *hp_count;
  for (u32 h = blockIdx.x; h < nheavy; h += gridDim.x) {
    XYZZ<C> acc = XYZZ<C>::inf();
    for (u32 s = threadIdx.x; s < np; s += MSM_HEAVY_THREADS)
      if (hp_id[s] == h) acc.add(ld_vec_rw<XYZZ<C>>(hp_sum, s));
    cta_tree_sum<C>(sm, acc);
    if (threadIdx.x == 0) st_vec(buckets, heavy[1 + h], sm[0]);
    __syncthreads();
  }
}

// ---- the heavy path with the point sliced over three lanes (Fq3) --------------------------------------------------------
// One thread per Fq3 point pays 58 - 82 dependent base-field products per group operation (40 - 60 us): the 128-thread
// form above took 1.2 ms + 0.55 ms for the ONE heavy bucket of the helper proof's witness (the "scalar == 1" bucket) and
// 1.0 ms of a 1.65 ms default-circuit proof.  Sliced, a group operation is a third as deep, a CTA holds 40 groups, and a
// chunk is 320 entries (8 per group) followed by a 6-level tree.
static constexpr int MSM_HEAVY_CHUNK_SLICED = 320;
static constexpr int MSM_HEAVY_GROUPS = 40;  // 4 warps x 10 groups
// one out-of-line copy of each group operation for the heavy path (by value: see ec.cuh on nvcc's stack colouring):
// inlined into the three kernels below they took ptxas from 80 s to 9 min on this file
template <class CS>
__device__ __noinline__ XYZZ<CS> sliced_add_ni(XYZZ<CS> a, XYZZ<CS> b) {
  a.add_impl(b);
  return a;
}
template <class CS>
__device__ __noinline__ XYZZ<CS> sliced_madd_ni(XYZZ<CS> a, AffinePoint<typename CS::F> p) {
  a.madd_impl(p);
  return a;
}
template <class CS>
__device__ __forceinline__ void cta_tree_sum_sliced(void* sm, XYZZ<CS>& acc, int G, int l, bool live) {
  if (live) st_xyzz_sliced<CS>(sm, (size_t)G, acc, l);
  __syncthreads();
  for (int s = 32; s > 0; s >>= 1) {
    if (live && G < s && G + s < MSM_HEAVY_GROUPS) {
      acc = sliced_add_ni<CS>(acc, ld_xyzz_sliced<CS>(sm, (size_t)(G + s), l));
      st_xyzz_sliced<CS>(sm, (size_t)G, acc, l);
    }
    __syncthreads();
  }
}
template <class CS, bool PRE>
__global__ void __launch_bounds__(128) msm_accumulate_heavy_sliced_kernel(
    const void* __restrict__ bases, const u32* __restrict__ offsets, const u32* __restrict__ entries,
    const u32* __restrict__ heavy, u32* __restrict__ hp_count, u32* __restrict__ hp_id, void* __restrict__ hp_sum) {
  typedef typename CS::F FF;
  extern __shared__ uint4 sm4[];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool live = lane < 30;
  const int l = (int)(lane % 3u);
  const int G = (int)(warp * 10 + lane / 3u);
  u32 nheavy = heavy[0] < (u32)MSM_MAX_HEAVY ? heavy[0] : (u32)MSM_MAX_HEAVY;
  u32 job = 0;
  for (u32 h = 0; h < nheavy; h++) {
    u32 g = heavy[1 + h];
    u32 lo = offsets[g], hi = offsets[g + 1];
    u32 nchunks = (hi - lo + MSM_HEAVY_CHUNK_SLICED - 1) / MSM_HEAVY_CHUNK_SLICED;
    for (u32 ch = 0; ch < nchunks; ch++, job++) {
      if (job % gridDim.x != blockIdx.x) continue;
      u32 e0 = lo + ch * MSM_HEAVY_CHUNK_SLICED;
      u32 e1 = e0 + MSM_HEAVY_CHUNK_SLICED < hi ? e0 + MSM_HEAVY_CHUNK_SLICED : hi;
      XYZZ<CS> acc = XYZZ<CS>::inf();
      if (live) {
        for (u32 e = e0 + (u32)G; e < e1; e += MSM_HEAVY_GROUPS) {
          u32 ent = entries[e];
          AffinePoint<FF> p = ld_base_sliced<CS>(bases, ent & 0x7fffffffu, l);
          if (ent >> 31) p.y = p.y.neg();
          acc = sliced_madd_ni<CS>(acc, p);
        }
      }
      cta_tree_sum_sliced<CS>(sm4, acc, G, l, live);
      if (warp == 0 && lane < 3) {
        u32 slot = 0;
        if (lane == 0) {
          slot = atomicAdd(hp_count, 1u);
          hp_id[slot] = h;
        }
        slot = __shfl_sync(7u, slot, 0);
        st_xyzz_sliced<CS>(hp_sum, slot, acc, l);
      }
      __syncthreads();
    }
  }
}
template <class CS>
__global__ void __launch_bounds__(128) msm_heavy_finish_sliced_kernel(
    const u32* __restrict__ heavy, const u32* __restrict__ hp_count, const u32* __restrict__ hp_id,
    const void* __restrict__ hp_sum, void* __restrict__ buckets) {
  extern __shared__ uint4 sm4[];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool live = lane < 30;
  const int l = (int)(lane % 3u);
  const int G = (int)(warp * 10 + lane / 3u);
  u32 nheavy = heavy[0] < (u32)MSM_MAX_HEAVY ? heavy[0] : (u32)MSM_MAX_HEAVY;
  u32 np = *hp_count;
  for (u32 h = blockIdx.x; h < nheavy; h += gridDim.x) {
    XYZZ<CS> acc = XYZZ<CS>::inf();
    if (live)
      for (u32 s = (u32)G; s < np; s += MSM_HEAVY_GROUPS)
        if (hp_id[s] == h) acc = sliced_add_ni<CS>(acc, ld_xyzz_sliced<CS>(hp_sum, s, l));
    cta_tree_sum_sliced<CS>(sm4, acc, G, l, live);
    if (warp == 0 && lane < 3) st_xyzz_sliced<CS>(buckets, heavy[1 + h], acc, l);
    __syncthreads();
  }
}

// ---- 5. / 6. bucket reduction and window combination: wec.cuh (lane-cooperative group law) -------------------------

// ---- helpers: normalisation, fixed-base multiplication, table precomputation ---------------------
template <class C>
__global__ void xyzz_to_affine_kernel(const void* __restrict__ in, size_t n, void* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<C> p = ld_vec_rw<XYZZ<C>>(in, i);
  st_vec(out, i, p.to_affine());
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

// pre[j * n + i] = 2^(start_j) * bases[i]  (affine), j < nwin, start_j = msm_win_start(j, c, nwin, balanced).
template <class C>
__global__ void __launch_bounds__(128) precompute_kernel(const void* __restrict__ bases, size_t n, int c, int nwin,
                                                         void* __restrict__ pre) {
  typedef typename C::F F;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  AffinePoint<F> a = ld_vec_rw<AffinePoint<F>>(bases, i);  // pre may alias bases (row 0 in place)
  st_vec(pre, i, a);
  XYZZ<C> p = XYZZ<C>::from_affine(a);
  for (int j = 1; j < nwin; j++) {
    int nd = msm_win_start(j, c, nwin, 1) - msm_win_start(j - 1, c, nwin, 1);
    for (int b = 0; b < nd; b++) p = p.dbl();
    a = p.to_affine();
    st_vec(pre, (size_t)j * n + i, a);
    p = XYZZ<C>::from_affine(a);
  }
}

// ---- host driver --------------------------------------------------------------------------------
// curves whose accumulation runs sliced over three lanes (fp3s.cuh)
template <class C> struct MsmSliced { static constexpr bool value = false; typedef C type; };
#if defined(__CUDACC__)
template <> struct MsmSliced<CurveMnt6G2> { static constexpr bool value = true; typedef CurveMnt6G2S type; };
#endif
struct MsmPlan {
  int c, nwin, shared;
  size_t stride, offset;  // shared (precomputed) tables: row pitch in points and first point used
};

// bases: n affine points (or the precomputed table nwin x n when plan.shared); result: one xyzz
// point at d_out.
template <class C>
static int msm_run(pcdgpu_ctx* ctx, const void* d_bases, const void* d_scalars, int scalars_mont, size_t n_main,
                   const void* d_extra, size_t n_extra, MsmPlan plan, void* d_out) {
  const size_t n = n_main + n_extra;
  typedef typename C::ScalarParams SP;
  cudaStream_t st = ctx->cur();
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
  const int rwin = shared ? 1 : nwin;  // windows seen by the reduction
  // unit buckets (msm_digits_kernel): ~256 unit scalars per bucket if every scalar were 1, a power of two, a multiple of
  // the reduction's group size; only with a shared bucket set and from 2^14 points up (smaller MSMs are launch-bound)
  static const bool no_units = getenv("PCDGPU_NO_UNIT_BUCKETS") != nullptr;  // development aid (A/B runs)
  size_t unit_k = 0;
  if (shared && n >= ((size_t)1 << 14) && !no_units) {
    unit_k = 64;
    while (unit_k < (n >> 8) && unit_k < 16384) unit_k <<= 1;
  }
  const size_t nbuckets = (shared ? B : B * nwin) + unit_k;
  void *dig, *ent, *cnt, *bkt, *seg, *cub_tmp;
  PCD_TRY(ctx->scratch(SLOT_MSM_DIG, (size_t)nwin * n * 4, &dig));
  PCD_TRY(ctx->scratch(SLOT_MSM_ENT, (size_t)nwin * n * 4, &ent));
  // counts | offsets | cursor | size keys (in, out) | bucket ids (in, out), each nbuckets + 1, then the heavy list
  size_t cstride = (nbuckets + 1 + 3) & ~(size_t)3;
  PCD_TRY(ctx->scratch(SLOT_MSM_CNT, (7 * cstride + MSM_MAX_HEAVY + 8) * 4, &cnt));
  u32* counts = (u32*)cnt;
  u32* offsets = counts + cstride;
  u32* cursor = offsets + cstride;
  u32* key_in = cursor + cstride;
  u32* key_out = key_in + cstride;
  u32* id_in = key_out + cstride;
  u32* perm = id_in + cstride;
  u32* heavy = perm + cstride;
  const bool small_plan = nbuckets <= 1024;  // msm_plan_small_kernel: it also clears the heavy list and the queue
  PCD_CUDA(ctx, cudaMemsetAsync(counts, 0, cstride * 4, st));
  u32* queue = heavy + MSM_MAX_HEAVY + 2;
  if (!small_plan) {
    PCD_CUDA(ctx, cudaMemsetAsync(heavy, 0, 4, st));
    PCD_CUDA(ctx, cudaMemsetAsync(queue, 0, 4, st));
  }
  const int acc_slot = n < ((size_t)1 << 14) ? PROF_MSM_ACC_SMALL
                       : (sizeof(typename C::F) > 80 ? PROF_MSM_ACC_G2Q3
                                                     : (sizeof(typename C::F) > 40 ? PROF_MSM_ACC_G2 : PROF_MSM_ACC_G1));
  int ps = ctx->prof_begin(PROF_MSM_SORT, (double)n * nwin);
  ctx->launches += 5 + 2;  // digits, scatter, accumulate, heavy x2 + cub's scan (init, scan); the reduction and
                           // combination count themselves
  msm_digits_kernel<SP><<<(unsigned)((n + 127) / 128), 128, 0, st>>>((const u32*)d_scalars, scalars_mont, n_main,
                                                                    (const u32*)d_extra, n, c, nwin,
                                                                    shared, (int*)dig, counts, (const uint4*)d_bases,
                                                                    (int)(sizeof(AffinePoint<typename C::F>) / 16),
                                                                    shared ? plan.offset : (size_t)0, (u32)unit_k);
  PCD_CUDA(ctx, cudaGetLastError());
  if (small_plan) {
    msm_plan_small_kernel<<<1, 1024, 0, st>>>(counts, (u32)nbuckets, offsets, cursor, perm, heavy, queue);
    PCD_CUDA(ctx, cudaGetLastError());
  } else {
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, counts, offsets, (int)(nbuckets + 1), st);
  PCD_TRY(ctx->scratch(SLOT_CUB, cub_bytes + 16, &cub_tmp));
  PCD_CUDA(ctx, cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, counts, offsets, (int)(nbuckets + 1), st));
  PCD_CUDA(ctx, cudaMemcpyAsync(cursor, offsets, (nbuckets + 1) * 4, cudaMemcpyDeviceToDevice, st));
  // bucket visiting order: decreasing size
  msm_sizekey_kernel<<<(unsigned)((nbuckets + 255) / 256), 256, 0, st>>>(counts, nbuckets, key_in, id_in);
  PCD_CUDA(ctx, cudaGetLastError());
  size_t sort_bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, sort_bytes, key_in, key_out, id_in, perm, (int)nbuckets, 0, 11, st);
  void* sort_tmp;
  PCD_TRY(ctx->scratch(SLOT_CUB2, sort_bytes + 16, &sort_tmp));
  PCD_CUDA(ctx, cub::DeviceRadixSort::SortPairsDescending(sort_tmp, sort_bytes, key_in, key_out, id_in, perm,
                                                          (int)nbuckets, 0, 11, st));
  ctx->launches += 4;  // size keys + cub's radix sort passes
  }
  size_t total = (size_t)nwin * n;
  msm_scatter_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const int*)dig, n, c, nwin, shared,
                                                                      plan.stride, plan.offset, cursor, (u32*)ent);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->prof_end(ps);
  if (ctx->sort_done) {
    PCD_CUDA(ctx, cudaEventRecord(ctx->sort_done, st));
    ctx->sort_done = nullptr;
  }
  for (int gi = 0; gi < 3; gi++)  // the prover's order of the accumulation grids (common.cuh: gate_wait / gate_done)
    if (ctx->gate_wait[gi]) {
      PCD_CUDA(ctx, cudaStreamWaitEvent(st, ctx->gate_wait[gi], 0));
      ctx->gate_wait[gi] = nullptr;
    }
  ps = ctx->prof_begin(acc_slot, (double)n * nwin);
  constexpr bool SLICED = MsmSliced<C>::value;      // three lanes per work item (Fq3)
  constexpr size_t ITEMS_PER_CTA = SLICED ? 40 : 128;  // work items a CTA of 128 threads holds at a time
  // occupancy and function attributes are per device and never change: asked once (every query / set is a few
  // microseconds of host time, and the default-circuit proofs are bound by the host's enqueue rate)
  static int occ_cache[2][16] = {{0}};
  static bool attr_done[16] = {false};
  const int dev_slot = ctx->device & 15;
  int acc_ctas = occ_cache[shared ? 1 : 0][dev_slot];
  if (acc_ctas == 0) {
  if constexpr (SLICED) {
    if (shared) PCD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&acc_ctas, msm_accumulate_sliced_kernel<typename MsmSliced<C>::type, true>, 128, 0));
    else PCD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&acc_ctas, msm_accumulate_sliced_kernel<typename MsmSliced<C>::type, false>, 128, 0));
  } else {
    if (shared) PCD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&acc_ctas, msm_accumulate_kernel<C, true>, 128, 0));
    else PCD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&acc_ctas, msm_accumulate_kernel<C, false>, 128, 0));
  }
  if (acc_ctas < 1) acc_ctas = 1;
  occ_cache[shared ? 1 : 0][dev_slot] = acc_ctas;
  }
  // CTAs per SM (measured, bench.py): side by side with the other lanes of a proof two CTAs (8 warps) are best -- the
  // registers left free let the latency-bound kernels (reduction, sorting, tails) run beside this one; a lone MSM
  // gains 4 % from a third CTA (5.99 -> 5.77 ms at 2^20); a fourth, forced to 128 registers, is slower.
  static const int acc_cap_env = getenv("PCDGPU_ACC_CTAS") ? atoi(getenv("PCDGPU_ACC_CTAS")) : 0;  // development aid
  const int acc_cap = acc_cap_env > 0 ? acc_cap_env : (SLICED ? 4 : (ctx->concurrent && ctx->in_proof ? 2 : 3));
  if (acc_ctas > acc_cap) acc_ctas = acc_cap;
  size_t acc_grid = (size_t)acc_ctas * ctx->sm_count;
  // One thread walks a bucket only up to 4 x the average size.  Real witnesses repeat values (0, 1, 2, -1,
  // ...): every copy of a value lands in the same bucket of each window, and a 500-entry bucket walked by
  // one thread (7 ms) would set the kernel's duration; such buckets go to the CTA-parallel heavy path.
  size_t avg_entries = total / nbuckets;
  u32 heavy_thr = (u32)(4 * avg_entries < 64 ? 64 : (4 * avg_entries > (size_t)MSM_HEAVY ? (size_t)MSM_HEAVY : 4 * avg_entries));
  // ... and never longer than a few times the entries a resident thread would get if the work were spread evenly: with
  // few buckets (small MSMs: 2^3 .. 2^9 buckets for ~10^3 points) "4 x the average" is thousands of entries, and a
  // bucket just below it was walked by ONE thread for milliseconds (measured: 2^9-point MSM at c = 4, 6.3 ms) while
  // the chunked path does the same bucket in 0.2 ms
  {
    size_t resident = acc_grid * 128;
    size_t even = 4 * (total / (resident ? resident : 1)) + 32;
    if (nbuckets * 8 < resident && even < heavy_thr) heavy_thr = (u32)even;
  }
  // bucket parts: at least ~6 waves of work items, at least 8 entries per part
  u32 split = 1;
  static const int split_env = getenv("PCDGPU_ACC_SPLIT") ? atoi(getenv("PCDGPU_ACC_SPLIT")) : 0;  // development aid
  const u32 max_split = split_env > 0 ? (u32)split_env : (u32)MSM_MAX_SPLIT;
  while (split < max_split && nbuckets * split < 6 * acc_grid * ITEMS_PER_CTA && avg_entries / (2 * split) >= 8) split *= 2;
  // ... and eight parts when four leave fewer than two work items per resident thread (2^18-point key tables with
  // c = 15: 17 K buckets x 4 on 38 - 57 K threads ran as 1.2 - 1.8 "waves" of ~80-entry items, the last of them alone)
  if (split_env <= 0 && split == (u32)MSM_MAX_SPLIT && nbuckets * split < 2 * acc_grid * ITEMS_PER_CTA &&
      avg_entries / (4 * split) >= 8)
    split *= 2;
  if (acc_grid > (nbuckets * split + ITEMS_PER_CTA - 1) / ITEMS_PER_CTA) acc_grid = (nbuckets * split + ITEMS_PER_CTA - 1) / ITEMS_PER_CTA;
  // inside a proof the CTAs retire after ~a quarter of a millisecond of work (quantum entries per lane) so that the
  // other lanes' kernels are not starved by a persistent grid; a lone MSM keeps the persistent form
  static const int quantum_env = getenv("PCDGPU_ACC_QUANTUM") ? atoi(getenv("PCDGPU_ACC_QUANTUM")) : -1;  // development aid
  u32 quantum = 0, persistent_from = 0;
  const u32 resident_items = (u32)(acc_grid * ITEMS_PER_CTA);
  {
    const int prods = SLICED ? 20 : (sizeof(typename C::F) / 40 == 1 ? 10 : 28);
    int q = quantum_env >= 0 ? quantum_env : (ctx->concurrent && ctx->in_proof ? 320 / prods : 0);
    if (q > 0) {
      const size_t resident = acc_grid;
      const size_t per_cta = (size_t)q * (SLICED ? 40 : 128);
      size_t extra = (total + per_cta - 1) / per_cta;
      if (extra > 65535) extra = 65535;
      quantum = (u32)q;
      persistent_from = (u32)extra;
      acc_grid = resident + extra;
    }
  }
  PCD_TRY(ctx->scratch(SLOT_MSM_BKT, nbuckets * sizeof(XYZZ<C>) * (split > 1 ? 1 + split : 1), &bkt));  // buckets | parts
  void* acc_out = split > 1 ? (void*)((char*)bkt + nbuckets * sizeof(XYZZ<C>)) : bkt;
  if constexpr (SLICED) {
    if (shared)
      msm_accumulate_sliced_kernel<typename MsmSliced<C>::type, true><<<(unsigned)acc_grid, 128, 0, st>>>(
          d_bases, offsets, (const u32*)ent, perm, nbuckets, acc_out, heavy, queue, heavy_thr, split, cursor, quantum, persistent_from, resident_items, (u32)(nbuckets - unit_k));
    else
      msm_accumulate_sliced_kernel<typename MsmSliced<C>::type, false><<<(unsigned)acc_grid, 128, 0, st>>>(
          d_bases, offsets, (const u32*)ent, perm, nbuckets, acc_out, heavy, queue, heavy_thr, split, cursor, quantum, persistent_from, resident_items, (u32)(nbuckets - unit_k));
  } else {
    if (shared)
      msm_accumulate_kernel<C, true><<<(unsigned)acc_grid, 128, 0, st>>>(d_bases, offsets, (const u32*)ent, perm, nbuckets,
                                                                       acc_out, heavy, queue, heavy_thr, split, cursor, quantum, persistent_from, resident_items, (u32)(nbuckets - unit_k));
    else
      msm_accumulate_kernel<C, false><<<(unsigned)acc_grid, 128, 0, st>>>(d_bases, offsets, (const u32*)ent, perm, nbuckets,
                                                                        acc_out, heavy, queue, heavy_thr, split, cursor, quantum, persistent_from, resident_items, (u32)(nbuckets - unit_k));
  }
  PCD_CUDA(ctx, cudaGetLastError());
  if (ctx->gate_done) {
    PCD_CUDA(ctx, cudaEventRecord(ctx->gate_done, st));
    ctx->gate_done = nullptr;
  }
  // Large MSMs: the accumulate kernel is a span (and a roofline) of its own; the part fold and the heavy-bucket kernels
  // that follow are the ACC_TAIL class.  The exact entry counts (non-zero digits of points that are not at infinity,
  // minus the heavy buckets the accumulate kernel handed over) are read back from the device.  Small MSMs keep one span.
  if (ps >= 0 && acc_slot != PROF_MSM_ACC_SMALL) {
    ctx->prof_end(ps);
    const int ps_acc = ps;
    ps = ctx->prof_begin(PROF_MSM_ACC_TAIL, 0.0);
    if (ctx->prof_pinned) {
      u32* pe = queue + 2;  // two spare words behind the queue counter
      msm_prof_entries_kernel<<<1, 32, 0, st>>>(heavy, offsets, nbuckets, pe);
      ctx->spans[ps_acc].units_pinned = ps_acc;
      cudaMemcpyAsync(ctx->prof_pinned + ps_acc, pe, 4, cudaMemcpyDeviceToHost, st);
      if (ps >= 0) {
        ctx->spans[ps].units_pinned = ps;
        cudaMemcpyAsync(ctx->prof_pinned + ps, pe + 1, 4, cudaMemcpyDeviceToHost, st);
      }
    }
  } else if (ps >= 0 && ctx->prof_pinned) {
    ctx->spans[ps].units_pinned = ps;
    cudaMemcpyAsync(ctx->prof_pinned + ps, offsets + nbuckets, 4, cudaMemcpyDeviceToHost, st);
  }
  if (split > 1) {
    if constexpr (SLICED)
      msm_fold_parts_sliced_kernel<typename MsmSliced<C>::type><<<(unsigned)((nbuckets + 29) / 30), 96, 0, st>>>(
          acc_out, nbuckets, split, cursor, bkt);
    else
      msm_fold_parts_kernel<C><<<(unsigned)((nbuckets + 127) / 128), 128, 0, st>>>(acc_out, nbuckets, split, cursor, bkt);
    PCD_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
  }
  size_t heavy_smem = cta_tree_smem<C>();
  // heavy-bucket partial list: at most one partial per MSM_HEAVY_CHUNK entries plus one per bucket
  size_t hp_cap = total / (SLICED ? MSM_HEAVY_CHUNK_SLICED : MSM_HEAVY_CHUNK) + MSM_MAX_HEAVY + 8;
  void* hp;
  PCD_TRY(ctx->scratch(SLOT_MSM_HP, hp_cap * (sizeof(XYZZ<C>) + 4) + 64, &hp));
  u32* hp_count = (u32*)hp;
  u32* hp_id = hp_count + 4;
  void* hp_sum = (char*)hp + 64 + ((hp_cap * 4 + 63) & ~(size_t)63);
  PCD_CUDA(ctx, cudaMemsetAsync(hp_count, 0, 4, st));
  if constexpr (SLICED) {
    typedef typename MsmSliced<C>::type CS;
    const size_t sl_smem = MSM_HEAVY_GROUPS * sizeof(XYZZ<C>);
    if (shared)
      msm_accumulate_heavy_sliced_kernel<CS, true><<<ctx->sm_count * 4, 128, sl_smem, st>>>(
          d_bases, offsets, (const u32*)ent, heavy, hp_count, hp_id, hp_sum);
    else
      msm_accumulate_heavy_sliced_kernel<CS, false><<<ctx->sm_count * 4, 128, sl_smem, st>>>(
          d_bases, offsets, (const u32*)ent, heavy, hp_count, hp_id, hp_sum);
    PCD_CUDA(ctx, cudaGetLastError());
    msm_heavy_finish_sliced_kernel<CS><<<64, 128, sl_smem, st>>>(heavy, hp_count, hp_id, hp_sum, bkt);
  } else {
    if (!attr_done[dev_slot]) {
      PCD_CUDA(ctx, cudaFuncSetAttribute(msm_accumulate_heavy_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)heavy_smem));
      PCD_CUDA(ctx, cudaFuncSetAttribute(msm_accumulate_heavy_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)heavy_smem));
      PCD_CUDA(ctx, cudaFuncSetAttribute(msm_heavy_finish_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)heavy_smem));
    }
    if (shared)
      msm_accumulate_heavy_kernel<C, true><<<ctx->sm_count * 4, MSM_HEAVY_THREADS, heavy_smem, st>>>(
          d_bases, offsets, (const u32*)ent, heavy, hp_count, hp_id, hp_sum);
    else
      msm_accumulate_heavy_kernel<C, false><<<ctx->sm_count * 4, MSM_HEAVY_THREADS, heavy_smem, st>>>(
          d_bases, offsets, (const u32*)ent, heavy, hp_count, hp_id, hp_sum);
    PCD_CUDA(ctx, cudaGetLastError());
    msm_heavy_finish_kernel<C><<<64, MSM_HEAVY_THREADS, heavy_smem, st>>>(heavy, hp_count, hp_id, hp_sum, bkt);
  }
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->prof_end(ps);
  ps = ctx->prof_begin(PROF_MSM_REDUCE, (double)nbuckets);
  // seg layout, in points: CTA partial sums [rwin][ctas] | sum levels 2 x [rwin][pitch] | wsum [rwin]
  typedef Wec<C> WG;
  static const int logL_env = getenv("PCDGPU_REDUCE_LOGL") ? atoi(getenv("PCDGPU_REDUCE_LOGL")) : 0;  // development aid
  // small bucket sets (the default-circuit proofs' MSMs: 64 buckets) are pure latency: fewer buckets per group, a shorter
  // running-sum chain.  MEASURED inside the PCD step: default-circuit proofs 1.20 / 1.04 ms with 8 buckets per group,
  // 1.12 / 0.94 ms with 4, 1.10 / 0.92 ms with 2; the large proofs lose with 4 (main 6.58 -> 6.80 ms: twice the groups,
  // twice the instructions)
  static const int logL_small_env = getenv("PCDGPU_REDUCE_LOGL_SMALL") ? atoi(getenv("PCDGPU_REDUCE_LOGL_SMALL")) : 0;  // development aid
  const int logL = logL_env > 0 ? logL_env : (B <= 1024 ? (logL_small_env > 0 ? logL_small_env : 1) : MSM_REDUCE_LOGL);
  const size_t L = (size_t)1 << logL;
  if (unit_k % L != 0) {
    ctx->set_error("unit buckets (%zu) are not a multiple of the reduction's group size (%zu)", unit_k, L);
    return PCDGPU_E_ARG;
  }
  const size_t T = ((B + L - 1) >> logL) + (unit_k >> logL);  // groups: weighted buckets, then unit buckets
  const size_t GPC = WEC_THREADS / WG::G;
  const size_t ctas = (T + GPC - 1) / GPC;
  const size_t per_cta = GPC * MSM_SUM_PER_GROUP;
  const size_t pitch = (ctas + per_cta - 1) / per_cta + 1;
  const size_t PB = sizeof(XYZZ<C>);
  PCD_TRY(ctx->scratch(SLOT_MSM_SEG, ((size_t)rwin * ctas + 2 * rwin * pitch + rwin + 8) * PB, &seg));
  char* sp = (char*)seg;
  void* lvl[2] = {sp + (size_t)rwin * ctas * PB, sp + ((size_t)rwin * ctas + rwin * pitch) * PB};
  void* wsum = (char*)lvl[1] + (size_t)rwin * pitch * PB;
  void* fin = rwin == 1 ? d_out : wsum;  // one window (precomputed tables): its sum IS the result
  const size_t red_smem = wec_smem_bytes<C>(WEC_THREADS, 4), sum_smem = wec_smem_bytes<C>(WEC_THREADS, 2);
  if (!attr_done[dev_slot]) {
    PCD_CUDA(ctx, cudaFuncSetAttribute(wec_reduce_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)red_smem));
    PCD_CUDA(ctx, cudaFuncSetAttribute(wec_sum_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sum_smem));
    attr_done[dev_slot] = true;
  }
  wec_reduce_kernel<C><<<dim3((unsigned)ctas, (unsigned)rwin), WEC_THREADS, red_smem, st>>>(bkt, B, logL,
                                                                                         ctas == 1 ? fin : seg, unit_k);
  PCD_CUDA(ctx, cudaGetLastError());
  ctx->launches += 1;
  const void* in = seg;
  size_t in_pitch = ctas, count = ctas;
  int pp = 0;
  while (count > 1) {
    size_t nct = (count + per_cta - 1) / per_cta;
    void* out = nct == 1 ? fin : lvl[pp];
    size_t out_pitch = nct == 1 ? 1 : pitch;
    wec_sum_kernel<C><<<dim3((unsigned)nct, (unsigned)rwin), WEC_THREADS, sum_smem, st>>>(in, in_pitch, count, out, out_pitch,
                                                                                       MSM_SUM_PER_GROUP);
    PCD_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    in = out;
    in_pitch = out_pitch;
    count = nct;
    pp ^= 1;
  }
  ctx->prof_end(ps);
  if (rwin > 1) {
    ps = ctx->prof_begin(PROF_MSM_TAIL, (double)rwin);
    wec_combine_kernel<C><<<1, 32, 0, st>>>(wsum, c, rwin, d_out);
    PCD_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    ctx->prof_end(ps);
  }
  return 0;
}
