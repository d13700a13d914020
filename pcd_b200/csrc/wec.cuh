// wec.cuh -- warp-cooperative group law: G lanes of a warp compute ONE group operation.
//
// The tails of an MSM and of a proof -- bucket reduction, window combination, s g_a + r g1_b, normalisation to
// affine (ark-ec GroupProjective::{add_assign, double_in_place, into_affine} under VariableBaseMSM::multi_scalar_mul
// and ark-groth16's create_proof_with_reduction; reached from /root/reference/src/ec_cycle_pcd/mod.rs:171,179) -- are
// chains of DEPENDENT group operations on few points.  One thread per chain pays the product COUNT of every operation
// (14 products in Fq, 42 in Fq2, 82 in Fq3 for a general addition, 0.7 us per product on B200): round 1 measured
// 3.4 - 5.4 ms for the G2 bucket reduction and 3.5 - 4 ms for the double-scalar multiplication of ONE proof.  The
// products inside an operation are mostly independent, so here the G lanes of a group (4 for Fq, 16 for Fq2, 32 for
// Fq3) each run one base-field instruction per step on operands in shared memory, following a schedule generated
// from the same formulas (tools/gen_wec.py -> wec_programs.cuh): an operation costs its multiplicative DEPTH (3 - 4
// products) plus its linear steps.  A warp holds 32 / G groups, each with its own points, running the same schedule.
//
// Memory: a point is 4 K slots (x, y, zz, zzz; K = extension degree), a slot is ten u32 words -- the layout of an
// XYZZ point in global memory, so points are copied in and out verbatim.
#pragma once
#include "ec.cuh"
#include "wec_programs.cuh"

enum { WEC_NOP = 0, WEC_MUL, WEC_LIN, WEC_RED, WEC_INV, WEC_CPY };

template <class C>
struct WecTraits;
#define WEC_DEFINE_TRAITS(CURVE, NAME, BASE, KDEG)                                   \
  template <>                                                                        \
  struct WecTraits<CURVE> {                                                          \
    typedef BASE B;                                                                  \
    static constexpr int K = KDEG;                                                   \
    static constexpr int G = WEC_##NAME##_G;                                         \
    static constexpr int NT = WEC_##NAME##_NTEMPS;                                   \
    __device__ static const u32* add1() { return WEC_##NAME##_ADD1; }                \
    __device__ static const u32* add2() { return WEC_##NAME##_ADD2; }                \
    __device__ static const u32* madd1() { return WEC_##NAME##_MADD1; }              \
    __device__ static const u32* madd2() { return WEC_##NAME##_MADD2; }              \
    __device__ static const u32* dbl() { return WEC_##NAME##_DBL; }                  \
    __device__ static const u32* toaff() { return WEC_##NAME##_TOAFF; }              \
    static constexpr int add1_rows = WEC_##NAME##_ADD1_ROWS, add2_rows = WEC_##NAME##_ADD2_ROWS,       \
                         madd1_rows = WEC_##NAME##_MADD1_ROWS, madd2_rows = WEC_##NAME##_MADD2_ROWS,   \
                         dbl_rows = WEC_##NAME##_DBL_ROWS, toaff_rows = WEC_##NAME##_TOAFF_ROWS;       \
  };
WEC_DEFINE_TRAITS(CurveMnt4G1, MNT4_G1, FpQ4, 1)
WEC_DEFINE_TRAITS(CurveMnt4G2, MNT4_G2, FpQ4, 2)
WEC_DEFINE_TRAITS(CurveMnt6G1, MNT6_G1, FpR4, 1)
WEC_DEFINE_TRAITS(CurveMnt6G2, MNT6_G2, FpR4, 3)

template <class B>
__device__ __forceinline__ B wec_ld(const u32* p) {
  B r;
  const uint2* q = reinterpret_cast<const uint2*>(p);
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint2 v = q[i];
    r.l[2 * i] = v.x;
    r.l[2 * i + 1] = v.y;
  }
  return r;
}
template <class B>
__device__ __forceinline__ void wec_st(u32* p, const B& a) {
  uint2* q = reinterpret_cast<uint2*>(p);
#pragma unroll
  for (int i = 0; i < 5; i++) q[i] = make_uint2(a.l[2 * i], a.l[2 * i + 1]);
}
__device__ __forceinline__ u32* wec_slot(u32 s, u32* xb, u32* yb, u32* tb) {
  return s < 16u ? xb + s * 10u : (s < 32u ? yb + (s - 16u) * 10u : tb + (s - 32u) * 10u);
}

// ---- unreduced ("lazy") linear arithmetic on ten-limb integers ---------------------------------------------------
// The generator (tools/gen_wec.py) tracks for every value a bound m with value < m p < 2^320 and never lets a
// subtraction underflow, so LIN needs no comparison and no conditional subtraction at all.
template <class B>
__device__ __forceinline__ void wec_raw_add(u32* r, const u32* a, const u32* b) {  // r = a + b (no carry out by the bound)
  r[0] = prims::add_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < FP_LIMBS - 1; i++) r[i] = prims::addc_cc(a[i], b[i]);
  r[FP_LIMBS - 1] = prims::addc(a[FP_LIMBS - 1], b[FP_LIMBS - 1]);
}
template <class B>
__device__ __forceinline__ void wec_raw_sub(u32* r, const u32* a, const u32* b) {  // r = a - b, a >= b
  r[0] = prims::sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < FP_LIMBS - 1; i++) r[i] = prims::subc_cc(a[i], b[i]);
  r[FP_LIMBS - 1] = prims::subc(a[FP_LIMBS - 1], b[FP_LIMBS - 1]);
}
template <class B>
__device__ __forceinline__ void wec_p_shl(u32* r, u32 j) {  // r = p << j, j < 22
  typedef typename B::Params P;
  r[0] = P::mod(0) << j;
#pragma unroll
  for (int i = 1; i < FP_LIMBS; i++) r[i] = __funnelshift_l(P::mod(i - 1), P::mod(i), j);
}
// d = a + k b (a absent when zero_a); k != 0 small; for k < 0: a + (p << j) - |k| b
template <class B>
__device__ __forceinline__ void wec_lin(u32* d, const u32* a, const u32* b, int k, u32 j, bool zero_a) {
  u32 t[FP_LIMBS], x[FP_LIMBS];
  const uint2* bq = reinterpret_cast<const uint2*>(b);
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint2 v = bq[i];
    x[2 * i] = v.x;
    x[2 * i + 1] = v.y;
  }
  const u32 ka = (u32)(k < 0 ? -k : k);
#pragma unroll
  for (int i = 0; i < FP_LIMBS; i++) t[i] = x[i];
  for (int bit = 30 - __clz(ka); bit >= 0; bit--) {  // |k| b by double-and-add from the top bit (no iterations for |k| = 1)
    wec_raw_add<B>(t, t, t);
    if ((ka >> bit) & 1u) wec_raw_add<B>(t, t, x);
  }
  if (k < 0) {
    wec_p_shl<B>(x, j);
    wec_raw_sub<B>(t, x, t);
  }
  if (!zero_a) {
    const uint2* aq = reinterpret_cast<const uint2*>(a);
#pragma unroll
    for (int i = 0; i < 5; i++) {
      uint2 v = aq[i];
      x[2 * i] = v.x;
      x[2 * i + 1] = v.y;
    }
    wec_raw_add<B>(t, t, x);
  }
  uint2* dq = reinterpret_cast<uint2*>(d);
#pragma unroll
  for (int i = 0; i < 5; i++) dq[i] = make_uint2(t[2 * i], t[2 * i + 1]);
}
// d = a mod p for a < 2^m p: conditional subtraction of p << s for s = m - 1 .. 0
template <class B>
__device__ __forceinline__ void wec_red(u32* d, const u32* a, u32 m) {
  u32 v[FP_LIMBS], ps[FP_LIMBS], t[FP_LIMBS];
  const uint2* aq = reinterpret_cast<const uint2*>(a);
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint2 w = aq[i];
    v[2 * i] = w.x;
    v[2 * i + 1] = w.y;
  }
  for (int s = (int)m - 1; s >= 0; s--) {
    wec_p_shl<B>(ps, (u32)s);
    t[0] = prims::sub_cc(v[0], ps[0]);
#pragma unroll
    for (int i = 1; i < FP_LIMBS; i++) t[i] = prims::subc_cc(v[i], ps[i]);
    const u32 borrow = prims::subc(0, 0);  // 0xffffffff if v < p << s
#pragma unroll
    for (int i = 0; i < FP_LIMBS; i++) v[i] = borrow ? v[i] : t[i];
  }
  uint2* dq = reinterpret_cast<uint2*>(d);
#pragma unroll
  for (int i = 0; i < 5; i++) dq[i] = make_uint2(v[2 * i], v[2 * i + 1]);
}

// The interpreter: `rows` rows of G instructions (two words each); lane gl of the group executes instruction
// [row * G + gl] and the group synchronises after every row (the generator never lets a row read a slot that another
// lane of the same row writes, and gives all the instructions of a row the same opcode: no divergence inside a row).
// One copy per base field and group size in a kernel (noinline): it holds the only inlined product.
template <class B, int G>
__device__ __noinline__ void wec_exec(const u32* __restrict__ prog, int rows, int gl, unsigned gmask, u32* xb, u32* yb,
                                      u32* tb) {
  const uint2* pr = reinterpret_cast<const uint2*>(prog);
  uint2 wn = __ldg(pr + gl);
  for (int r = 0; r < rows; r++) {
    const uint2 w = wn;
    if (r + 1 < rows) wn = __ldg(pr + (r + 1) * G + gl);  // the next row's words travel while this row computes
    const u32 op = w.x >> 28;
    if (op != WEC_NOP) {
      u32* d = wec_slot((w.x >> 18) & 511u, xb, yb, tb);
      u32* a = wec_slot((w.x >> 9) & 511u, xb, yb, tb);
      u32* b = wec_slot(w.x & 511u, xb, yb, tb);
      if (op == WEC_MUL) {
        wec_st<B>(d, wec_ld<B>(a) * wec_ld<B>(b));
      } else if (op == WEC_LIN) {
        const int k = (int)(signed char)(w.y & 0xffu);
        wec_lin<B>(d, a, b, k, (w.y >> 8) & 31u, (w.y >> 18) & 1u);
      } else if (op == WEC_RED) {
        wec_red<B>(d, a, (w.y >> 13) & 31u);
      } else if (op == WEC_INV) {
        wec_st<B>(d, wec_ld<B>(a).inverse());
      } else {  // WEC_CPY
        wec_st<B>(d, wec_ld<B>(a));
      }
    }
    __syncwarp(gmask);
  }
}

template <class C>
struct Wec {
  typedef WecTraits<C> W;
  typedef typename W::B B;
  static constexpr int K = W::K, G = W::G;
  static constexpr int EW = 10 * K;                        // words of one coordinate
  static constexpr int PW = 4 * EW;                        // words of one xyzz point
  static constexpr int NTW = (W::NT * 10 + 3) & ~3;        // words of the temporaries
  static constexpr int GROUPS_PER_WARP = 32 / G;
  int gl;          // lane inside the group
  unsigned gmask;  // the group's lanes
  u32* tb;         // the group's temporaries (NTW words)

  __device__ explicit Wec(u32* temps) : tb(temps) {
    const unsigned lane = threadIdx.x & 31u;
    gl = (int)(lane % G);
    gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane - gl));
  }
  __device__ void sync() const { __syncwarp(gmask); }
  // is the extension-field element at e (K slots) zero?  Same answer on every lane of the group.
  __device__ bool ext_zero(const u32* e) const {
    u32 t = 0;
    if (gl < K) {
#pragma unroll
      for (int i = 0; i < 10; i++) t |= e[gl * 10 + i];
    }
    const bool z = __ballot_sync(gmask, t != 0) == 0;
    sync();  // the vote orders execution, __syncwarp also orders the reads above against the group's later writes
    return z;
  }
  __device__ bool is_inf(const u32* p) const { return ext_zero(p + 2 * EW); }
  __device__ void copy(u32* dst, const u32* src, int nwords = PW) const {
    for (int i = gl; i < nwords; i += G) dst[i] = src[i];
    sync();
  }
  __device__ void set_inf(u32* p) const {
    for (int i = gl; i < PW; i += G) p[i] = 0;
    sync();
  }
  // xyzz point idx of a global array -> shared memory (and back)
  __device__ void load(u32* dst, const void* src, size_t idx) const {
    const uint4* s = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(src) + idx * (size_t)(PW * 4));
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = gl; i < PW / 4; i += G) d[i] = s[i];
    sync();
  }
  __device__ void store(void* dst, size_t idx, const u32* src, int nwords = PW) const {
    uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<char*>(dst) + idx * (size_t)(nwords * 4));
    const uint4* s = reinterpret_cast<const uint4*>(src);
    for (int i = gl; i < nwords / 4; i += G) d[i] = s[i];
  }
  // affine (x, y) at a -> xyzz at p (zz = zzz = 1)
  __device__ void from_affine(u32* p, const u32* a) const {
    for (int i = gl; i < 2 * EW; i += G) {
      p[i] = a[i];
      const int j = i % EW;  // zz = zzz = 1: coefficient 0 is the Montgomery one, the others zero
      p[2 * EW + i] = j < 10 ? B::Params::one(j) : 0u;
    }
    sync();
  }
  __device__ bool affine_is_inf(const u32* a) const { return ext_zero(a) && ext_zero(a + EW); }
  __device__ void run(const u32* prog, int rows, u32* x, u32* y) const {
    wec_exec<B, G>(prog, rows, gl, gmask, x, y, tb);
  }
  // x <- 2 x
  __device__ void dbl(u32* x) const {
    if (is_inf(x)) return;
    if (ext_zero(x + EW)) {
      set_inf(x);
      return;
    }
    run(W::dbl(), W::dbl_rows, x, x);
  }
  // x <- x + y (both xyzz; y is not modified and must not alias x)
  __device__ void add(u32* x, u32* y) const {
    if (is_inf(y)) return;
    if (is_inf(x)) {
      copy(x, y);
      return;
    }
    run(W::add1(), W::add1_rows, x, y);
    const bool pz = ext_zero(tb), rz = ext_zero(tb + EW);
    if (pz) {
      if (rz) dbl(x);
      else set_inf(x);
      return;
    }
    run(W::add2(), W::add2_rows, x, y);
  }
  // x <- x + (affine point at y)
  __device__ void madd(u32* x, u32* y) const {
    if (affine_is_inf(y)) return;
    if (is_inf(x)) {
      from_affine(x, y);
      return;
    }
    run(W::madd1(), W::madd1_rows, x, y);
    const bool pz = ext_zero(tb), rz = ext_zero(tb + EW);
    if (pz) {
      if (rz) {
        from_affine(x, y);
        dbl(x);
      } else {
        set_inf(x);
      }
      return;
    }
    run(W::madd2(), W::madd2_rows, x, y);
  }
  // x[0 .. 2 EW) <- the affine coordinates (zeros for the point at infinity)
  __device__ void to_affine(u32* x) const {
    if (is_inf(x)) {
      for (int i = gl; i < 2 * EW; i += G) x[i] = 0;
      sync();
      return;
    }
    run(W::toaff(), W::toaff_rows, x, x);
  }
  // x <- [k] y, k < 2^32 (tmp: nothing; x must not alias y)
  __device__ void mul_u32(u32* x, u32* y, u32 k) const {
    set_inf(x);
    if (k == 0) return;
    for (int bit = 31 - __clz(k); bit >= 0; bit--) {
      dbl(x);
      if ((k >> bit) & 1u) add(x, y);
    }
  }
};

// shared-memory words a CTA of `threads` threads needs for `npoints` points per group
template <class C>
constexpr size_t wec_smem_bytes(int threads, int npoints) {
  return (size_t)(threads / Wec<C>::G) * (size_t)(npoints * Wec<C>::PW + Wec<C>::NTW) * 4;
}

static constexpr int WEC_THREADS = 128;

// ---- bucket reduction ----------------------------------------------------------------------------------------------
// S_w = sum_b (b + 1) B_{w,b} per window.  Group t of a window takes the L = 2^logL buckets [t L, t L + L) from the top
// down with the running-sum trick (2 L additions), adds [t L] * (their plain sum), and the groups of a CTA fold their
// contributions in a tree: part[w * gridDim.x + blockIdx.x] = the CTA's share of S_w.
// K > 0 (one window only): K unit buckets follow the B weighted ones (msm_digits_kernel: scalars equal to 1); groups
// T .. T + K / L - 1 add L of them each with weight 1.
template <class C>
__global__ void __launch_bounds__(WEC_THREADS) wec_reduce_kernel(const void* __restrict__ buckets, size_t B, int logL,
                                                                  void* __restrict__ part, size_t K) {
  typedef Wec<C> WG;
  constexpr int G = WG::G, GPC = WEC_THREADS / G, AREA = 4 * WG::PW + WG::NTW;
  extern __shared__ uint4 wec_sm4[];
  u32* sm = reinterpret_cast<u32*>(wec_sm4);
  const int g = threadIdx.x / G;
  u32* run = sm + (size_t)g * AREA;
  u32* acc = run + WG::PW;
  u32* q = acc + WG::PW;
  u32* tmp = q + WG::PW;
  WG wg(tmp + WG::PW);
  const size_t L = (size_t)1 << logL;
  const size_t T = (B + L - 1) >> logL;
  const size_t w = blockIdx.y, t = (size_t)blockIdx.x * GPC + g;
  wg.set_inf(run);
  wg.set_inf(acc);
  if (t < T) {
    const size_t base = w * B + t * L;
    const size_t lim = B - t * L < L ? B - t * L : L;
    for (size_t j = lim; j-- > 0;) {
      wg.load(q, buckets, base + j);
      wg.add(run, q);
      wg.add(acc, run);
    }
    const u32 k = (u32)(t * L);
    if (k != 0 && !wg.is_inf(run)) {
      wg.mul_u32(tmp, run, k);
      wg.add(acc, tmp);
    }
  } else if (t < T + (K >> logL)) {
    const size_t base = B + (t - T) * L;
    for (size_t j = 0; j < L; j++) {
      wg.load(q, buckets, base + j);
      wg.add(acc, q);
    }
  }
  for (int s = GPC / 2; s > 0; s >>= 1) {
    __syncthreads();
    if (g < s) wg.add(acc, acc + (size_t)s * AREA);
  }
  if (g == 0) wg.store(part, w * gridDim.x + blockIdx.x, acc);
}

// out[w * out_pitch + blockIdx.x] = sum of the CTA's share of in[w * in_pitch + 0 .. count): every group adds
// `per_group` consecutive points, the groups of the CTA fold in a tree
template <class C>
__global__ void __launch_bounds__(WEC_THREADS) wec_sum_kernel(const void* __restrict__ in, size_t in_pitch, size_t count,
                                                               void* __restrict__ out, size_t out_pitch, int per_group) {
  typedef Wec<C> WG;
  constexpr int G = WG::G, GPC = WEC_THREADS / G, AREA = 2 * WG::PW + WG::NTW;
  extern __shared__ uint4 wec_sm4[];
  u32* sm = reinterpret_cast<u32*>(wec_sm4);
  const int g = threadIdx.x / G;
  u32* acc = sm + (size_t)g * AREA;
  u32* q = acc + WG::PW;
  WG wg(q + WG::PW);
  const size_t w = blockIdx.y;
  const size_t lo = ((size_t)blockIdx.x * GPC + g) * per_group;
  wg.set_inf(acc);
  for (size_t i = lo; i < lo + per_group && i < count; i++) {
    wg.load(q, in, w * in_pitch + i);
    wg.add(acc, q);
  }
  for (int s = GPC / 2; s > 0; s >>= 1) {
    __syncthreads();
    if (g < s) wg.add(acc, acc + (size_t)s * AREA);
  }
  if (g == 0) wg.store(out, w * out_pitch + blockIdx.x, acc);
}

// result = sum_w 2^(c w) wsum[w] by Horner: c doublings and one addition per window, on one group (launch with 32 threads)
template <class C>
__global__ void __launch_bounds__(32) wec_combine_kernel(const void* __restrict__ wsum, int c, int nwin, void* __restrict__ out) {
  typedef Wec<C> WG;
  __shared__ uint4 sm4[(2 * WG::PW + WG::NTW) / 4];
  if ((int)threadIdx.x >= WG::G) return;
  u32* acc = reinterpret_cast<u32*>(sm4);
  u32* q = acc + WG::PW;
  WG wg(q + WG::PW);
  wg.load(acc, wsum, nwin - 1);
  for (int ww = nwin - 2; ww >= 0; ww--) {
    for (int i = 0; i < c; i++) wg.dbl(acc);
    wg.load(q, wsum, ww);
    wg.add(acc, q);
  }
  wg.store(out, 0, acc);
}

// dst (affine) <- xyzz point idx of src
template <class C>
__global__ void __launch_bounds__(32) wec_to_affine_kernel(const void* __restrict__ src, size_t idx, void* __restrict__ dst) {
  typedef Wec<C> WG;
  __shared__ uint4 sm4[(WG::PW + WG::NTW) / 4];
  if ((int)threadIdx.x >= WG::G) return;
  u32* p = reinterpret_cast<u32*>(sm4);
  WG wg(p + WG::PW);
  wg.load(p, src, idx);
  wg.to_affine(p);
  wg.store(dst, 0, p, 2 * WG::EW);
}

// sum of the n xyzz points src[first + i * stride] (n small: per-GPU partial sums, the terms of C) -> xyzz at
// dst_xyzz[out_idx] and / or affine at dst_affine (null pointers are skipped)
template <class C>
__global__ void __launch_bounds__(32) wec_sum_affine_kernel(const void* __restrict__ src, size_t first, size_t stride, int n,
                                                             void* __restrict__ dst_xyzz, size_t out_idx,
                                                             void* __restrict__ dst_affine) {
  typedef Wec<C> WG;
  __shared__ uint4 sm4[(2 * WG::PW + WG::NTW) / 4];
  if ((int)threadIdx.x >= WG::G) return;
  u32* acc = reinterpret_cast<u32*>(sm4);
  u32* q = acc + WG::PW;
  WG wg(q + WG::PW);
  wg.set_inf(acc);
  for (int i = 0; i < n; i++) {
    wg.load(q, src, first + (size_t)i * stride);
    wg.add(acc, q);
  }
  if (dst_xyzz) wg.store(dst_xyzz, out_idx, acc);
  if (dst_affine) {
    wg.sync();
    wg.to_affine(acc);
    wg.store(dst_affine, 0, acc, 2 * WG::EW);
  }
}

// out[iout] = sum_{j < npairs} [k_j] pts[i_j], npairs <= 2, k_j 320-bit plain integers (ten words at k + 10 j): one
// double-and-add chain per pair with SIGNED 4-bit windows (digits -7 .. 8, table P .. 8 P: 4 doublings + 3 additions to
// build, then <= 300 doublings and ~76 additions; the 2-bit unsigned form before it took 320 + 120).  The chains are
// independent, so they run as two groups of the same warp at the cost of one (G1 curves: G = 4); a final addition
// joins them.  This is s g_a + r g1_b of a Groth16 proof -- the longest serial tail of a proof -- and [r] C2' of a
// GM17 proof.
template <class C>
__global__ void __launch_bounds__(32) wec_multi_mul_kernel(const void* __restrict__ pts, size_t i0, size_t i1,
                                                            const u32* __restrict__ k, int npairs, void* __restrict__ out,
                                                            size_t iout) {
  typedef Wec<C> WG;
  typedef typename WG::B B;
  static_assert(WG::G <= 16, "two groups in one warp");
  constexpr int NDIG = 81;                             // 320 bits / 4 + the last carry
  constexpr int DIGW = 24;                             // words holding the digits (one byte each)
  constexpr int AREA = 10 * WG::PW + WG::NTW + DIGW;  // acc | table 1 P .. 8 P | q | temporaries | digits
  __shared__ uint4 sm4[2 * AREA / 4];
  static_assert(AREA % 4 == 0, "16-byte aligned areas");
  const int g = threadIdx.x / WG::G;
  if (g >= 2) return;
  u32* acc = reinterpret_cast<u32*>(sm4) + (size_t)g * AREA;
  u32* tab = acc + WG::PW;  // tab + (m - 1) PW = m P
  u32* q = tab + 8 * WG::PW;
  WG wg(q + WG::PW);
  signed char* dig = reinterpret_cast<signed char*>(q + WG::PW + WG::NTW);
  const unsigned both = (1u << (2 * WG::G)) - 1u;
  wg.set_inf(acc);
  if (g < npairs) {
    const u32* kk = k + 10 * g;
    if (wg.gl == 0) {  // signed recoding from the low end: d in -7 .. 8
      u32 carry = 0;
      for (int j = 0; j < NDIG; j++) {
        u32 d = (j < 80 ? (kk[j >> 3] >> ((j & 7) * 4)) & 15u : 0u) + carry;
        carry = d > 8u ? 1u : 0u;
        dig[j] = (signed char)(carry ? (int)d - 16 : (int)d);
      }
    }
    wg.load(tab, pts, g == 0 ? i0 : i1);
    for (int m = 2; m <= 8; m++) {  // m P = 2 (m / 2) P for even m, (m - 1) P + P for odd m
      u32* t = tab + (m - 1) * WG::PW;
      if ((m & 1) == 0) {
        wg.copy(t, tab + (m / 2 - 1) * WG::PW);
        wg.dbl(t);
      } else {
        wg.copy(t, tab + (m - 2) * WG::PW);
        wg.add(t, tab);
      }
    }
    for (int j = NDIG - 1; j >= 0; j--) {
      wg.dbl(acc);
      wg.dbl(acc);
      wg.dbl(acc);
      wg.dbl(acc);
      // ONE addition site for both signs and both groups of the warp: with `add` called from a positive and a negative
      // branch the two chains (different digits) serialised the two calls -- 1.5 additions of warp time per window
      const int d = dig[j];
      const int m = d < 0 ? -d : d;
      if (m) {
        wg.copy(q, tab + (m - 1) * WG::PW);
        if (d < 0 && wg.gl < WG::K) {  // y <- -y, coefficient by coefficient (stored points are canonical)
          u32* yc = q + WG::EW + wg.gl * 10;
          wec_st<B>(yc, wec_ld<B>(yc).neg());
        }
        wg.sync();
      } else {
        wg.set_inf(q);
      }
      wg.add(acc, q);
    }
  }
  __syncwarp(both);
  if (g == 0) {
    if (npairs > 1) wg.add(acc, acc + AREA);
    wg.store(out, iout, acc);
  }
}
