// pcd_oracle.cpp -- CPU oracle (C++17, 5 x u64 limbs) for the Groth16 proving path of
// arkworks-rs/pcd on the MNT4-298 / MNT6-298 cycle.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load liboracle.so, and there only as the checker / the CPU arm.  The
// product (libpcdgpu.so) never links or calls it and has no CPU fallback.
//
// PARITY UNPINNED (see oracle/pcd_oracle.py header and DESIGN.md): the arithmetic of the path is
// in un-vendored, un-pinned arkworks git dependencies (/root/reference/Cargo.toml:16-42) and the
// reference's tests hold no golden vectors (/root/reference/tests/mnt4_groth16.rs:87,103,117,119
// only assert verify == true/false).  This file restates the published algorithms of those
// crates in the *shape* arkworks runs them on a CPU (so that it can also serve as the CPU
// baseline, SURVEY.md 8d / BASELINE.md 3):
//   * ark-ff Fp320: 5 x u64 Montgomery CIOS, R = 2^320               (Fp::mul below)
//   * ark-ff Fp2/Fp3 Karatsuba                                       (Fp2, Fp3)
//   * ark-ec short-Weierstrass Jacobian add_assign_mixed / double_in_place / add_assign
//   * ark-ec VariableBaseMSM::multi_scalar_mul: unsigned windows, c = 3 | floor(.69 lg N)+2,
//     zero scalars dropped, unit scalars added in window 0, one task per window (rayon shape)
//   * ark-poly Radix2EvaluationDomain in-order FFT, coset shift by GENERATOR, ifft 1/n scale
//   * ark-groth16 R1CStoQAP::witness_map and create_proof_with_reduction (r, s supplied)
// reached from /root/reference/src/ec_cycle_pcd/mod.rs:171,179 (IC::MainSNARK::prove /
// IC::HelpSNARK::prove).  It is pinned against oracle/pcd_oracle.py (big-int restatement with
// algebraic self-checks) through tests/golden/*.json and tests/test_oracle_c.py.
//
// Memory encodings are the ABI's (include/pcdgpu.h): field element = 40 B little-endian limbs in
// Montgomery form; MSM scalar = 40 B little-endian plain integer; affine point = x || y with the
// point at infinity encoded as x = y = 0.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

typedef uint64_t u64;
typedef unsigned __int128 u128;

// ------------------------------------------------------------------------------------------
// Prime fields
// ------------------------------------------------------------------------------------------
struct FieldConsts {
  u64 mod[5];
  u64 inv;     // -p^-1 mod 2^64
  u64 one[5];  // R mod p
  u64 r2[5];   // R^2 mod p
  u64 generator_small;
  int two_adicity;
};

static bool geq5(const u64* a, const u64* b) {
  for (int i = 4; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] > b[i];
  }
  return true;
}
static void sub5(u64* a, const u64* b) {
  u64 borrow = 0;
  for (int i = 0; i < 5; i++) {
    u128 t = (u128)a[i] - b[i] - borrow;
    a[i] = (u64)t;
    borrow = (u64)(t >> 64) & 1;
  }
}
static u64 add5(u64* a, const u64* b) {
  u64 carry = 0;
  for (int i = 0; i < 5; i++) {
    u128 t = (u128)a[i] + b[i] + carry;
    a[i] = (u64)t;
    carry = (u64)(t >> 64);
  }
  return carry;
}

static FieldConsts make_consts(const u64 mod[5], u64 gen, int two_adicity) {
  FieldConsts c;
  memcpy(c.mod, mod, 40);
  u64 x = 1;  // Newton: x = p^-1 mod 2^64
  for (int i = 0; i < 6; i++) x *= 2 - mod[0] * x;
  c.inv = (u64)0 - x;
  // R mod p and R^2 mod p by doubling 1, 320 / 640 times
  u64 v[5] = {1, 0, 0, 0, 0};
  for (int i = 0; i < 640; i++) {
    if (i == 320) memcpy(c.one, v, 40);
    u64 t[5];
    memcpy(t, v, 40);
    add5(v, t);  // p < 2^298 so 2v < 2^300: no carry out
    if (geq5(v, mod)) sub5(v, mod);
  }
  memcpy(c.r2, v, 40);
  c.generator_small = gen;
  c.two_adicity = two_adicity;
  return c;
}

static const u64 MOD_R4[5] = {0xbb4334a400000001ULL, 0xfb494c07925d6ad3ULL, 0xcaeec9635cf44194ULL,
                              0xa266249da7b0548eULL, 0x000003bcf7bcd473ULL};
static const u64 MOD_Q4[5] = {0xc90cd65a71660001ULL, 0x41a9e35e51200e12ULL, 0xcaeec9635d1330eaULL,
                              0xa266249da7b0548eULL, 0x000003bcf7bcd473ULL};

struct PR4 {
  static const FieldConsts& C() { static FieldConsts c = make_consts(MOD_R4, 10, 34); return c; }
};
struct PQ4 {
  static const FieldConsts& C() { static FieldConsts c = make_consts(MOD_Q4, 17, 17); return c; }
};

template <class P>
struct Fp {
  u64 l[5];
  static constexpr int DEG = 1;
  typedef Fp Base;
  static Fp zero() { Fp r; memset(r.l, 0, 40); return r; }
  static Fp one() { Fp r; memcpy(r.l, P::C().one, 40); return r; }
  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3] | l[4]) == 0; }
  bool operator==(const Fp& o) const { return memcmp(l, o.l, 40) == 0; }
  bool operator!=(const Fp& o) const { return !(*this == o); }
  Fp operator+(const Fp& o) const {
    Fp r = *this;
    add5(r.l, o.l);
    if (geq5(r.l, P::C().mod)) sub5(r.l, P::C().mod);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r = *this;
    if (!geq5(r.l, o.l)) add5(r.l, P::C().mod);
    sub5(r.l, o.l);
    return r;
  }
  Fp neg() const {
    if (is_zero()) return *this;
    Fp r; memcpy(r.l, P::C().mod, 40);
    sub5(r.l, l);
    return r;
  }
  Fp dbl() const { return *this + *this; }
  // CIOS Montgomery product (ark-ff Fp320 mul_assign shape)
  Fp operator*(const Fp& o) const {
    const FieldConsts& c = P::C();
    u64 t[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 5; i++) {
      u64 carry = 0;
      for (int j = 0; j < 5; j++) {
        u128 s = (u128)l[j] * o.l[i] + t[j] + carry;
        t[j] = (u64)s;
        carry = (u64)(s >> 64);
      }
      u128 s = (u128)t[5] + carry;
      t[5] = (u64)s;
      t[6] = (u64)(s >> 64);
      u64 m = t[0] * c.inv;
      s = (u128)m * c.mod[0] + t[0];
      carry = (u64)(s >> 64);
      for (int j = 1; j < 5; j++) {
        s = (u128)m * c.mod[j] + t[j] + carry;
        t[j - 1] = (u64)s;
        carry = (u64)(s >> 64);
      }
      s = (u128)t[5] + carry;
      t[4] = (u64)s;
      t[5] = t[6] + (u64)(s >> 64);
    }
    Fp r;
    memcpy(r.l, t, 40);
    if (geq5(r.l, c.mod)) sub5(r.l, c.mod);
    return r;
  }
  Fp sqr() const { return (*this) * (*this); }
  Fp mul_u64(u64 k) const { return (*this) * from_u64(k); }
  static Fp from_u64(u64 v) {
    Fp r = zero();
    r.l[0] = v;
    return r.to_mont();
  }
  Fp to_mont() const { Fp r2; memcpy(r2.l, P::C().r2, 40); return (*this) * r2; }
  Fp from_mont() const { Fp o = zero(); o.l[0] = 1; return (*this) * o; }
  // exponent as plain little-endian limbs
  Fp pow(const u64* e, int nl) const {
    Fp r = one();
    bool started = false;
    for (int i = nl - 1; i >= 0; i--)
      for (int b = 63; b >= 0; b--) {
        if (started) r = r.sqr();
        if ((e[i] >> b) & 1) { r = started ? r * (*this) : *this; started = true; }
      }
    return r;
  }
  Fp pow_u64(u64 e) const { return pow(&e, 1); }
  Fp inverse() const {
    u64 e[5];
    memcpy(e, P::C().mod, 40);
    e[0] -= 2;  // p is odd and p[0] >= 2
    return pow(e, 5);
  }
  // is the plain integer value larger than that of the negation (ark-serialize y flag)
  bool lex_largest() const {
    Fp a = from_mont(), b = neg().from_mont();
    for (int i = 4; i >= 0; i--)
      if (a.l[i] != b.l[i]) return a.l[i] > b.l[i];
    return false;
  }
  static Fp generator() { return from_u64(P::C().generator_small); }
};
typedef Fp<PR4> FpR4;
typedef Fp<PQ4> FpQ4;

template <class B, u64 NR>
struct Fp2 {
  B c0, c1;
  static constexpr int DEG = 2;
  typedef B Base;
  static Fp2 zero() { return {B::zero(), B::zero()}; }
  static Fp2 one() { return {B::one(), B::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fp2& o) const { return c0 == o.c0 && c1 == o.c1; }
  bool operator!=(const Fp2& o) const { return !(*this == o); }
  Fp2 operator+(const Fp2& o) const { return {c0 + o.c0, c1 + o.c1}; }
  Fp2 operator-(const Fp2& o) const { return {c0 - o.c0, c1 - o.c1}; }
  Fp2 neg() const { return {c0.neg(), c1.neg()}; }
  Fp2 dbl() const { return {c0.dbl(), c1.dbl()}; }
  Fp2 operator*(const Fp2& o) const {
    static const B nr = B::from_u64(NR);
    B v0 = c0 * o.c0, v1 = c1 * o.c1;
    return {v0 + nr * v1, (c0 + c1) * (o.c0 + o.c1) - v0 - v1};
  }
  Fp2 sqr() const { return (*this) * (*this); }
  Fp2 mul_u64(u64 k) const { B f = B::from_u64(k); return {c0 * f, c1 * f}; }
  Fp2 inverse() const {
    static const B nr = B::from_u64(NR);
    B n = (c0.sqr() - nr * c1.sqr()).inverse();
    return {c0 * n, (c1 * n).neg()};
  }
  bool lex_largest() const { return !c1.is_zero() ? c1.lex_largest() : c0.lex_largest(); }
};

template <class B, u64 NR>
struct Fp3 {
  B c0, c1, c2;
  static constexpr int DEG = 3;
  typedef B Base;
  static Fp3 zero() { return {B::zero(), B::zero(), B::zero()}; }
  static Fp3 one() { return {B::one(), B::zero(), B::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero() && c2.is_zero(); }
  bool operator==(const Fp3& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
  bool operator!=(const Fp3& o) const { return !(*this == o); }
  Fp3 operator+(const Fp3& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
  Fp3 operator-(const Fp3& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
  Fp3 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
  Fp3 dbl() const { return {c0.dbl(), c1.dbl(), c2.dbl()}; }
  Fp3 operator*(const Fp3& o) const {
    static const B nr = B::from_u64(NR);
    B v0 = c0 * o.c0, v1 = c1 * o.c1, v2 = c2 * o.c2;
    B x = (c1 + c2) * (o.c1 + o.c2) - v1 - v2;
    B y = (c0 + c1) * (o.c0 + o.c1) - v0 - v1;
    B z = (c0 + c2) * (o.c0 + o.c2) - v0 - v2 + v1;
    return {v0 + nr * x, y + nr * v2, z};
  }
  Fp3 sqr() const { return (*this) * (*this); }
  Fp3 mul_u64(u64 k) const { B f = B::from_u64(k); return {c0 * f, c1 * f, c2 * f}; }
  Fp3 inverse() const {
    static const B nr = B::from_u64(NR);
    B t0 = c0.sqr() - nr * (c1 * c2);
    B t1 = nr * c2.sqr() - c0 * c1;
    B t2 = c1.sqr() - c0 * c2;
    B n = (c0 * t0 + nr * (c2 * t1 + c1 * t2)).inverse();
    return {t0 * n, t1 * n, t2 * n};
  }
  bool lex_largest() const {
    if (!c2.is_zero()) return c2.lex_largest();
    if (!c1.is_zero()) return c1.lex_largest();
    return c0.lex_largest();
  }
};
typedef Fp2<FpQ4, 17> Fq2;
typedef Fp3<FpR4, 5> Fq3;

// ------------------------------------------------------------------------------------------
// Curves: y^2 = x^3 + a x + b (a != 0), Jacobian coordinates as ark-ec's GroupProjective.
// ------------------------------------------------------------------------------------------
struct C4G1 { typedef FpQ4 F; typedef PR4 SP; static F a() { return F::from_u64(2); } };
struct C4G2 { typedef Fq2 F; typedef PR4 SP; static F a() { return {FpQ4::from_u64(34), FpQ4::zero()}; } };
struct C6G1 { typedef FpR4 F; typedef PQ4 SP; static F a() { return F::from_u64(11); } };
struct C6G2 { typedef Fq3 F; typedef PQ4 SP; static F a() { return {FpR4::zero(), FpR4::zero(), FpR4::from_u64(11)}; } };

template <class C>
struct Aff {
  typename C::F x, y;
  bool is_inf() const { return x.is_zero() && y.is_zero(); }
};

template <class C>
struct Jac {
  typedef typename C::F F;
  F x, y, z;
  static Jac inf() { return {F::one(), F::one(), F::zero()}; }
  bool is_inf() const { return z.is_zero(); }
  static Jac from_affine(const Aff<C>& p) {
    if (p.is_inf()) return inf();
    return {p.x, p.y, F::one()};
  }
  // dbl-2007-bl (a general), ark-ec double_in_place
  void dbl() {
    if (is_inf()) return;
    if (y.is_zero()) { *this = inf(); return; }
    static const F ca = C::a();
    F xx = x.sqr(), yy = y.sqr(), yyyy = yy.sqr(), zz = z.sqr();
    F s = ((x + yy).sqr() - xx - yyyy).dbl();
    F m = xx.dbl() + xx + ca * zz.sqr();
    F t = m.sqr() - s.dbl();
    F z3 = (y + z).sqr() - yy - zz;
    x = t;
    y = m * (s - t) - yyyy.dbl().dbl().dbl();
    z = z3;
  }
  // madd-2007-bl, ark-ec add_assign_mixed
  void add_mixed(const Aff<C>& p) {
    if (p.is_inf()) return;
    if (is_inf()) { *this = from_affine(p); return; }
    F z1z1 = z.sqr();
    F u2 = p.x * z1z1;
    F s2 = (p.y * z) * z1z1;
    if (x == u2 && y == s2) { dbl(); return; }
    F h = u2 - x;
    if (h.is_zero()) { *this = inf(); return; }
    F hh = h.sqr();
    F i = hh.dbl().dbl();
    F j = h * i;
    F r = (s2 - y).dbl();
    F v = x * i;
    F x3 = r.sqr() - j - v.dbl();
    F y3 = r * (v - x3) - (y * j).dbl();
    F z3 = (z + h).sqr() - z1z1 - hh;
    x = x3; y = y3; z = z3;
  }
  // add-2007-bl, ark-ec add_assign
  void add(const Jac& o) {
    if (o.is_inf()) return;
    if (is_inf()) { *this = o; return; }
    F z1z1 = z.sqr(), z2z2 = o.z.sqr();
    F u1 = x * z2z2, u2 = o.x * z1z1;
    F s1 = y * o.z * z2z2, s2 = o.y * z * z1z1;
    if (u1 == u2) {
      if (s1 == s2) { dbl(); return; }
      *this = inf();
      return;
    }
    F h = u2 - u1;
    F i = h.dbl().sqr();
    F j = h * i;
    F r = (s2 - s1).dbl();
    F v = u1 * i;
    F x3 = r.sqr() - j - v.dbl();
    F y3 = r * (v - x3) - (s1 * j).dbl();
    F z3 = ((z + o.z).sqr() - z1z1 - z2z2) * h;
    x = x3; y = y3; z = z3;
  }
  Aff<C> to_affine() const {
    if (is_inf()) return {F::zero(), F::zero()};
    F zi = z.inverse();
    F zi2 = zi.sqr();
    return {x * zi2, y * (zi2 * zi)};
  }
  // [k]P, k plain little-endian limbs
  static Jac mul(const Jac& p, const u64* k, int nl) {
    Jac acc = inf();
    for (int i = nl - 1; i >= 0; i--)
      for (int b = 63; b >= 0; b--) {
        acc.dbl();
        if ((k[i] >> b) & 1) acc.add(p);
      }
    return acc;
  }
};

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
static void parallel_for(size_t n, int threads, const std::function<void(size_t, size_t, int)>& fn) {
  if (threads <= 1 || n < 2) { fn(0, n, 0); return; }
  size_t t = std::min<size_t>(threads, n);
  std::vector<std::thread> th;
  size_t chunk = (n + t - 1) / t;
  for (size_t i = 0; i < t; i++) {
    size_t lo = i * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back(fn, lo, hi, (int)i);
  }
  for (auto& x : th) x.join();
}

static int ark_window_size(size_t n) {
  if (n < 32) return 3;
  int lg = 0;
  while (((size_t)1 << lg) < n) lg++;
  return lg * 69 / 100 + 2;
}

static inline u64 scalar_bits(const u64* k, int start, int c) {
  int limb = start / 64, off = start % 64;
  if (limb >= 5) return 0;
  u64 v = k[limb] >> off;
  if (off + c > 64 && limb + 1 < 5) v |= k[limb + 1] << (64 - off);
  return v & (((u64)1 << c) - 1);
}

// ------------------------------------------------------------------------------------------
// VariableBaseMSM::multi_scalar_mul
// ------------------------------------------------------------------------------------------
template <class C>
static Jac<C> msm_pippenger(const Aff<C>* bases, const u64* scalars, size_t n, int threads, int c_override) {
  const int num_bits = 298;
  int c = c_override > 0 ? c_override : ark_window_size(n);
  int nwin = (num_bits + c - 1) / c;
  std::vector<Jac<C>> wsum(nwin);
  std::atomic<int> next(0);
  auto work = [&]() {
    std::vector<Jac<C>> buckets(((size_t)1 << c) - 1);
    for (;;) {
      int w = next.fetch_add(1);
      if (w >= nwin) break;
      int w_start = w * c;
      Jac<C> res = Jac<C>::inf();
      for (auto& b : buckets) b = Jac<C>::inf();
      for (size_t i = 0; i < n; i++) {
        const u64* k = scalars + 5 * i;
        if ((k[0] | k[1] | k[2] | k[3] | k[4]) == 0) continue;
        if (k[0] == 1 && (k[1] | k[2] | k[3] | k[4]) == 0) {
          if (w_start == 0) res.add_mixed(bases[i]);
          continue;
        }
        u64 d = scalar_bits(k, w_start, c);
        if (d) buckets[d - 1].add_mixed(bases[i]);
      }
      Jac<C> running = Jac<C>::inf();
      for (size_t b = buckets.size(); b-- > 0;) {
        running.add(buckets[b]);
        res.add(running);
      }
      wsum[w] = res;
    }
  };
  int t = std::max(1, std::min(threads, nwin));
  if (t == 1) work();
  else {
    std::vector<std::thread> th;
    for (int i = 0; i < t; i++) th.emplace_back(work);
    for (auto& x : th) x.join();
  }
  Jac<C> total = wsum[nwin - 1];
  for (int w = nwin - 2; w >= 0; w--) {
    for (int i = 0; i < c; i++) total.dbl();
    total.add(wsum[w]);
  }
  return total;
}

// ------------------------------------------------------------------------------------------
// Radix-2 evaluation domain
// ------------------------------------------------------------------------------------------
template <class P>
static Fp<P> omega_for(int log_n) {
  const FieldConsts& c = P::C();
  u64 e[5];
  memcpy(e, c.mod, 40);
  e[0] -= 1;
  int s = c.two_adicity;
  // e = (p - 1) >> s  (s < 64 here: 34 or 17)
  u64 out[5];
  for (int i = 0; i < 5; i++) {
    u64 v = e[i] >> s;
    if (i + 1 < 5) v |= e[i + 1] << (64 - s);
    out[i] = v;
  }
  Fp<P> w = Fp<P>::generator().pow(out, 5);
  for (int i = log_n; i < s; i++) w = w.sqr();
  return w;
}

static inline size_t bitrev(size_t i, int bits) {
  size_t r = 0;
  for (int b = 0; b < bits; b++) { r = (r << 1) | (i & 1); i >>= 1; }
  return r;
}

// in-order radix-2 DIT; per-stage work is split across threads (ark-poly `parallel` shape)
template <class P>
static void fft_in_place(Fp<P>* a, int log_n, const Fp<P>& omega, int threads) {
  typedef Fp<P> F;
  size_t n = (size_t)1 << log_n;
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev(i, log_n);
    if (i < j) std::swap(a[i], a[j]);
  }
  // twiddle table omega^k, k < n/2
  std::vector<F> tw(std::max<size_t>(n / 2, 1));
  tw[0] = F::one();
  for (size_t k = 1; k < n / 2; k++) tw[k] = tw[k - 1] * omega;
  for (int s = 0; s < log_n; s++) {
    size_t m = (size_t)1 << s;
    size_t step = n / (2 * m);
    parallel_for(n / 2, threads, [&](size_t lo, size_t hi, int) {
      for (size_t idx = lo; idx < hi; idx++) {
        size_t k = (idx / m) * 2 * m, j = idx % m;
        F t = a[k + j + m] * tw[j * step];
        F u = a[k + j];
        a[k + j] = u + t;
        a[k + j + m] = u - t;
      }
    });
  }
}

template <class P>
static void domain_transform(Fp<P>* a, int log_n, int inverse, int coset, int threads) {
  typedef Fp<P> F;
  size_t n = (size_t)1 << log_n;
  F omega = omega_for<P>(log_n);
  F g = F::generator();
  if (!inverse) {
    if (coset) {
      F pw = F::one();
      for (size_t i = 0; i < n; i++) { a[i] = a[i] * pw; pw = pw * g; }
    }
    fft_in_place(a, log_n, omega, threads);
  } else {
    fft_in_place(a, log_n, omega.inverse(), threads);
    F ninv = F::from_u64((u64)n).inverse();
    if (coset) {
      F gi = g.inverse();
      F pw = ninv;
      for (size_t i = 0; i < n; i++) { a[i] = a[i] * pw; pw = pw * gi; }
    } else {
      parallel_for(n, threads, [&](size_t lo, size_t hi, int) {
        for (size_t i = lo; i < hi; i++) a[i] = a[i] * ninv;
      });
    }
  }
}

// ------------------------------------------------------------------------------------------
// GeneralEvaluationDomain: radix-2 when the size's log fits the field's 2-adicity,
// otherwise -- on q4, whose multiplicative group also has a subgroup of order 7^2 -- the smallest
// 7^a 2^b >= m (ark-poly MixedRadixEvaluationDomain).  The group generator of a mixed domain is
// LARGE_SUBGROUP_ROOT^((2^s 7^2) / n), LARGE_SUBGROUP_ROOT = GENERATOR^((p-1) / (2^s 7^2)), whose 49th
// power is TWO_ADIC_ROOT.  Any correct DFT with that generator gives the same (canonical) values; this
// one is Cooley-Tukey: split off factors of 7 (naive 7-point DFTs + twiddles), radix-2 below.
// ------------------------------------------------------------------------------------------
struct DomainShape {
  size_t n;
  int a;  // power of 7
  int b;  // power of 2
};

template <class P>
static bool domain_shape(size_t min_size, DomainShape* out) {
  const FieldConsts& c = P::C();
  int lg = 0;
  while (((size_t)1 << lg) < (min_size ? min_size : 1)) lg++;
  if (lg <= c.two_adicity) {
    *out = {(size_t)1 << lg, 0, lg};
    return true;
  }
  if (c.two_adicity != 17) return false;  // only q4 has the small subgroup (7, adicity 2)
  bool found = false;
  for (int a = 0; a <= 2; a++)
    for (int b = 0; b <= c.two_adicity; b++) {
      size_t s = (a == 0 ? 1 : (a == 1 ? 7 : 49)) * ((size_t)1 << b);
      if (s >= min_size && (!found || s < out->n)) {
        *out = {s, a, b};
        found = true;
      }
    }
  return found;
}

template <class P>
static Fp<P> omega_general(const DomainShape& d) {
  if (d.a == 0) return omega_for<P>(d.b);
  const FieldConsts& c = P::C();
  // e = (p - 1) / (2^s * 49): exact division, done limb-wise
  u64 e[5];
  memcpy(e, c.mod, 40);
  e[0] -= 1;
  int s = c.two_adicity;
  u64 sh[5];
  for (int i = 0; i < 5; i++) {
    u64 v = e[i] >> s;
    if (i + 1 < 5) v |= e[i + 1] << (64 - s);
    sh[i] = v;
  }
  u64 q[5];
  u128 rem = 0;
  for (int i = 4; i >= 0; i--) {
    u128 cur = (rem << 64) | sh[i];
    q[i] = (u64)(cur / 49);
    rem = cur % 49;
  }
  Fp<P> w = Fp<P>::generator().pow(q, 5);  // LARGE_SUBGROUP_ROOT, order 2^s * 49
  u64 k = (((u64)1 << s) * 49) / (u64)d.n;
  return w.pow_u64(k);
}

// out-of-place mixed-radix DFT of x (length n = 7^a 2^b) with generator w; x is overwritten
template <class P>
static void mixed_fft(std::vector<Fp<P>>& x, int a, int b, const Fp<P>& w, int threads) {
  typedef Fp<P> F;
  size_t n = x.size();
  if (a == 0) {
    if (b > 0) fft_in_place(x.data(), b, w, threads);
    return;
  }
  size_t M = n / 7;
  F w7 = w.pow_u64((u64)M);  // primitive 7th root
  F w7p[7];
  w7p[0] = F::one();
  for (int i = 1; i < 7; i++) w7p[i] = w7p[i - 1] * w7;
  std::vector<std::vector<F>> y(7, std::vector<F>(M));
  parallel_for(M, threads, [&](size_t lo, size_t hi, int) {
    F tw = w.pow_u64((u64)lo);  // w^{n2}
    for (size_t n2 = lo; n2 < hi; n2++) {
      F twk = F::one();  // w^{n2 k1}
      for (int k1 = 0; k1 < 7; k1++) {
        F acc = F::zero();
        for (int n1 = 0; n1 < 7; n1++) acc = acc + x[M * n1 + n2] * w7p[(n1 * k1) % 7];
        y[k1][n2] = acc * twk;
        twk = twk * tw;
      }
      tw = tw * w;
    }
  });
  F wsub = w.pow_u64(7);
  for (int k1 = 0; k1 < 7; k1++) mixed_fft<P>(y[k1], a - 1, b, wsub, threads);
  for (int k1 = 0; k1 < 7; k1++)
    for (size_t k2 = 0; k2 < M; k2++) x[k1 + 7 * k2] = y[k1][k2];
}

template <class P>
static void domain_transform_general(Fp<P>* v, const DomainShape& d, int inverse, int coset, int threads) {
  typedef Fp<P> F;
  if (d.a == 0) {
    domain_transform<P>(v, d.b, inverse, coset, threads);
    return;
  }
  size_t n = d.n;
  F omega = omega_general<P>(d);
  F g = F::generator();
  std::vector<F> x(v, v + n);
  if (!inverse) {
    if (coset) {
      F pw = F::one();
      for (size_t i = 0; i < n; i++) { x[i] = x[i] * pw; pw = pw * g; }
    }
    mixed_fft<P>(x, d.a, d.b, omega, threads);
  } else {
    mixed_fft<P>(x, d.a, d.b, omega.inverse(), threads);
    F pw = F::from_u64((u64)n).inverse();
    F gi = coset ? g.inverse() : F::one();
    for (size_t i = 0; i < n; i++) { x[i] = x[i] * pw; pw = pw * gi; }
  }
  memcpy((void*)v, (const void*)x.data(), n * sizeof(F));
}

// ------------------------------------------------------------------------------------------
// R1CS in CSR form and the QAP witness map
// ------------------------------------------------------------------------------------------
struct Csr {
  const uint32_t* row_ptr;  // m + 1
  const uint32_t* col;      // nnz
  const u64* coeff;         // nnz x 5, Montgomery
};

template <class P>
static Fp<P> row_dot(const Csr& M, size_t i, const Fp<P>* z) {
  typedef Fp<P> F;
  F acc = F::zero();
  F one = F::one();
  for (uint32_t k = M.row_ptr[i]; k < M.row_ptr[i + 1]; k++) {
    F co;
    memcpy(co.l, M.coeff + 5 * (size_t)k, 40);
    const F& v = z[M.col[k]];
    acc = acc + (co == one ? v : co * v);  // ark-groth16 evaluate_constraint skips the unit mul
  }
  return acc;
}

template <class P>
static int witness_map(const Csr& A, const Csr& B, const Csr& Cm, size_t m, size_t num_inputs, const Fp<P>* z,
                       Fp<P>* h, int threads) {
  typedef Fp<P> F;
  DomainShape dom;
  if (!domain_shape<P>(m + num_inputs, &dom)) return -1;
  size_t n = dom.n;
  std::vector<F> a(n, F::zero()), b(n, F::zero()), c(n, F::zero());
  parallel_for(m, threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) {
      a[i] = row_dot<P>(A, i, z);
      b[i] = row_dot<P>(B, i, z);
      c[i] = row_dot<P>(Cm, i, z);
    }
  });
  for (size_t j = 0; j < num_inputs; j++) a[m + j] = z[j];
  domain_transform_general<P>(a.data(), dom, 1, 0, threads);
  domain_transform_general<P>(a.data(), dom, 0, 1, threads);
  domain_transform_general<P>(b.data(), dom, 1, 0, threads);
  domain_transform_general<P>(b.data(), dom, 0, 1, threads);
  domain_transform_general<P>(c.data(), dom, 1, 0, threads);
  domain_transform_general<P>(c.data(), dom, 0, 1, threads);
  // (g^n - 1)^-1
  F gn = F::generator().pow_u64((u64)n);
  F zinv = (gn - F::one()).inverse();
  parallel_for(n, threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) h[i] = (a[i] * b[i] - c[i]) * zinv;
  });
  domain_transform_general<P>(h, dom, 1, 1, threads);
  return (int)n;
}

// ------------------------------------------------------------------------------------------
// Groth16 prover
// ------------------------------------------------------------------------------------------
struct Groth16PkView {
  const void *alpha_g1, *beta_g1, *delta_g1;  // affine G1
  const void *beta_g2, *delta_g2;             // affine G2
  const void *a_query, *b_g1_query;           // num_vars affine G1
  const void* b_g2_query;                     // num_vars affine G2
  const void* h_query;                        // n - 1 affine G1
  const void* l_query;                        // num_witness affine G1
};

template <class G1, class G2, class P>
static int groth16_prove(const Groth16PkView& pk, const Csr& A, const Csr& B, const Csr& Cm, size_t m,
                         size_t num_inputs, size_t num_witness, const Fp<P>* z, const u64* r, const u64* s,
                         void* out, int threads) {
  typedef Fp<P> F;
  size_t nv = num_inputs + num_witness;
  DomainShape dom;
  if (!domain_shape<P>(m + num_inputs, &dom)) return -1;
  size_t n = dom.n;
  std::vector<F> h(n);
  if (witness_map<P>(A, B, Cm, m, num_inputs, z, h.data(), threads) < 0) return -1;
  // scalars leave Montgomery form (into_repr)
  std::vector<u64> hs(5 * n), zs(5 * nv);
  parallel_for(n, threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) { F v = h[i].from_mont(); memcpy(&hs[5 * i], v.l, 40); }
  });
  parallel_for(nv, threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) { F v = z[i].from_mont(); memcpy(&zs[5 * i], v.l, 40); }
  });
  const Aff<G1>* aq = (const Aff<G1>*)pk.a_query;
  const Aff<G1>* bq1 = (const Aff<G1>*)pk.b_g1_query;
  const Aff<G2>* bq2 = (const Aff<G2>*)pk.b_g2_query;
  Jac<G1> h_acc = msm_pippenger<G1>((const Aff<G1>*)pk.h_query, hs.data(), n - 1, threads, 0);
  Jac<G1> l_acc = msm_pippenger<G1>((const Aff<G1>*)pk.l_query, zs.data() + 5 * num_inputs, num_witness, threads, 0);
  Jac<G1> delta1 = Jac<G1>::from_affine(*(const Aff<G1>*)pk.delta_g1);
  Jac<G2> delta2 = Jac<G2>::from_affine(*(const Aff<G2>*)pk.delta_g2);
  // g_a = r*delta + a_query[0] + MSM(a_query[1..], z[1..]) + alpha
  Jac<G1> g_a = Jac<G1>::mul(delta1, r, 5);
  g_a.add_mixed(aq[0]);
  g_a.add(msm_pippenger<G1>(aq + 1, zs.data() + 5, nv - 1, threads, 0));
  g_a.add_mixed(*(const Aff<G1>*)pk.alpha_g1);
  Jac<G1> g1_b = Jac<G1>::mul(delta1, s, 5);
  g1_b.add_mixed(bq1[0]);
  g1_b.add(msm_pippenger<G1>(bq1 + 1, zs.data() + 5, nv - 1, threads, 0));
  g1_b.add_mixed(*(const Aff<G1>*)pk.beta_g1);
  Jac<G2> g2_b = Jac<G2>::mul(delta2, s, 5);
  g2_b.add_mixed(bq2[0]);
  g2_b.add(msm_pippenger<G2>(bq2 + 1, zs.data() + 5, nv - 1, threads, 0));
  g2_b.add_mixed(*(const Aff<G2>*)pk.beta_g2);
  // g_c = s*g_a + r*g1_b - (r s) delta + l_acc + h_acc
  F rm, sm;
  memcpy(rm.l, r, 40);
  memcpy(sm.l, s, 40);
  F rs = (rm.to_mont() * sm.to_mont()).from_mont();
  Jac<G1> g_c = Jac<G1>::mul(g_a, s, 5);
  g_c.add(Jac<G1>::mul(g1_b, r, 5));
  Jac<G1> rsd = Jac<G1>::mul(delta1, rs.l, 5);
  rsd.y = rsd.y.neg();
  g_c.add(rsd);
  g_c.add(l_acc);
  g_c.add(h_acc);
  char* o = (char*)out;
  Aff<G1> A_ = g_a.to_affine();
  Aff<G2> B_ = g2_b.to_affine();
  Aff<G1> C_ = g_c.to_affine();
  memcpy(o, &A_, sizeof(A_));
  memcpy(o + sizeof(A_), &B_, sizeof(B_));
  memcpy(o + sizeof(A_) + sizeof(B_), &C_, sizeof(C_));
  return 0;
}

// ------------------------------------------------------------------------------------------
// GM17 prover (ark-gm17 r1cs_to_sap.rs / prover.rs).  Mirrors oracle/pcd_oracle.py
// sap_witness_map / gm17_prove; the reference binds it at /root/reference/tests/mnt4_gm17.rs:27-28.
// ------------------------------------------------------------------------------------------
// full: num_inputs + num_witness + m + (num_inputs - 1) elements (the SAP assignment), h: n + 1 coefficients
template <class P>
static int sap_witness_map(const Csr& A, const Csr& B, const Csr& Cm, size_t m, size_t num_inputs, size_t num_witness,
                           const Fp<P>* z, const Fp<P>& d1, const Fp<P>& d2, Fp<P>* full, Fp<P>* h, int threads) {
  typedef Fp<P> F;
  DomainShape dom;
  if (!domain_shape<P>(2 * m + 2 * (num_inputs - 1) + 1, &dom)) return -1;
  const size_t n = dom.n, nv = num_inputs + num_witness;
  const size_t ev1 = nv, ev2 = nv + m - 1, off = 2 * m;
  const F one = F::one();
  std::vector<F> a(n, F::zero()), c(n, F::zero());
  for (size_t i = 0; i < nv; i++) full[i] = z[i];
  parallel_for(m, threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) {
      F az = row_dot<P>(A, i, z), bz = row_dot<P>(B, i, z), cz = row_dot<P>(Cm, i, z);
      F e = (az - bz) * (az - bz);
      full[ev1 + i] = e;
      a[2 * i] = az + bz;
      a[2 * i + 1] = az - bz;
      F c4 = cz + cz;
      c4 = c4 + c4;
      c[2 * i] = c4 + e;
      c[2 * i + 1] = e;
    }
  });
  a[off] = one;
  c[off] = one;
  for (size_t i = 1; i < num_inputs; i++) {
    F e = (z[i] - one) * (z[i] - one);
    full[ev2 + i] = e;
    a[off + 2 * i - 1] = z[i] + one;
    a[off + 2 * i] = z[i] - one;
    F x4 = z[i] + z[i];
    x4 = x4 + x4;
    c[off + 2 * i - 1] = x4 + e;
    c[off + 2 * i] = e;
  }
  domain_transform_general<P>(a.data(), dom, 1, 0, threads);
  F d1d = d1 + d1, d1sq = d1 * d1;
  parallel_for(n, threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) h[i] = d1d * a[i];
  });
  h[0] = h[0] - d2 - d1sq;
  h[n] = d1sq;
  domain_transform_general<P>(a.data(), dom, 0, 1, threads);
  domain_transform_general<P>(c.data(), dom, 1, 0, threads);
  domain_transform_general<P>(c.data(), dom, 0, 1, threads);
  F zinv = (F::generator().pow_u64((u64)n) - one).inverse();
  parallel_for(n, threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) a[i] = (a[i] * a[i] - c[i]) * zinv;
  });
  domain_transform_general<P>(a.data(), dom, 1, 1, threads);
  parallel_for(n - 1, threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) h[i] = h[i] + a[i];
  });
  return (int)n;
}

struct Gm17PkView {
  const void *a_query, *b_query, *c_query_1, *c_query_2, *g_gamma2_z_t;  // nsap G1, nsap G2, nsap - ni G1, nsap G1, n + 1 G1
  const void *g_gamma_z, *h_gamma_z, *g_ab_gamma_z, *g_gamma2_z2;         // G1, G2, G1, G1
};

template <class G1, class G2, class P>
static int gm17_prove(const Gm17PkView& pk, const Csr& A, const Csr& B, const Csr& Cm, size_t m, size_t num_inputs,
                      size_t num_witness, const Fp<P>* z, const u64* d1p, const u64* d2p, const u64* rp, void* out,
                      int threads) {
  typedef Fp<P> F;
  DomainShape dom;
  if (!domain_shape<P>(2 * m + 2 * (num_inputs - 1) + 1, &dom)) return -1;
  const size_t n = dom.n, nsap = num_inputs + num_witness + m + num_inputs - 1;
  F d1, d2, r;
  memcpy(d1.l, d1p, 40);
  memcpy(d2.l, d2p, 40);
  memcpy(r.l, rp, 40);
  F d1m = d1.to_mont(), d2m = d2.to_mont(), rm = r.to_mont();
  std::vector<F> full(nsap), h(n + 1);
  if (sap_witness_map<P>(A, B, Cm, m, num_inputs, num_witness, z, d1m, d2m, full.data(), h.data(), threads) < 0) return -1;
  std::vector<u64> hs(5 * (n + 1)), fs(5 * nsap);
  parallel_for(n + 1, threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) { F v = h[i].from_mont(); memcpy(&hs[5 * i], v.l, 40); }
  });
  parallel_for(nsap, threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) { F v = full[i].from_mont(); memcpy(&fs[5 * i], v.l, 40); }
  });
  const Aff<G1>* aq = (const Aff<G1>*)pk.a_query;
  const Aff<G2>* bq = (const Aff<G2>*)pk.b_query;
  const Aff<G1>* c2q = (const Aff<G1>*)pk.c_query_2;
  const Aff<G1>* gzt = (const Aff<G1>*)pk.g_gamma2_z_t;
  Jac<G1> ggz = Jac<G1>::from_affine(*(const Aff<G1>*)pk.g_gamma_z);
  Jac<G2> hgz = Jac<G2>::from_affine(*(const Aff<G2>*)pk.h_gamma_z);
  Jac<G1> gabz = Jac<G1>::from_affine(*(const Aff<G1>*)pk.g_ab_gamma_z);
  Jac<G1> gz2 = Jac<G1>::from_affine(*(const Aff<G1>*)pk.g_gamma2_z2);
  F rd1 = (rm + d1m).from_mont();
  // A = (r + d1) g_gamma_z + a_query[0] + MSM(a_query[1..], full[1..]);  B likewise in G2
  Jac<G1> g_a = Jac<G1>::mul(ggz, rd1.l, 5);
  g_a.add_mixed(aq[0]);
  g_a.add(msm_pippenger<G1>(aq + 1, fs.data() + 5, nsap - 1, threads, 0));
  Jac<G2> g_b = Jac<G2>::mul(hgz, rd1.l, 5);
  g_b.add_mixed(bq[0]);
  g_b.add(msm_pippenger<G2>(bq + 1, fs.data() + 5, nsap - 1, threads, 0));
  // C = MSM(c_query_1, aux) + (r^2 + 2 r d1) g_gamma2_z2 + (r + d1) g_ab_gamma_z + r (c_query_2[0] + MSM(c_query_2[1..]))
  //     + d2 g_gamma2_z_t[0] + MSM(g_gamma2_z_t, h)
  Jac<G1> g_c = msm_pippenger<G1>((const Aff<G1>*)pk.c_query_1, fs.data() + 5 * num_inputs, nsap - num_inputs, threads, 0);
  F k = (rm * rm + (rm + rm) * d1m).from_mont();
  g_c.add(Jac<G1>::mul(gz2, k.l, 5));
  g_c.add(Jac<G1>::mul(gabz, rd1.l, 5));
  Jac<G1> c2 = msm_pippenger<G1>(c2q + 1, fs.data() + 5, nsap - 1, threads, 0);
  c2.add_mixed(c2q[0]);
  g_c.add(Jac<G1>::mul(c2, r.l, 5));
  g_c.add(Jac<G1>::mul(Jac<G1>::from_affine(gzt[0]), d2.l, 5));
  g_c.add(msm_pippenger<G1>(gzt, hs.data(), n + 1, threads, 0));
  char* o = (char*)out;
  Aff<G1> A_ = g_a.to_affine();
  Aff<G2> B_ = g_b.to_affine();
  Aff<G1> C_ = g_c.to_affine();
  memcpy(o, &A_, sizeof(A_));
  memcpy(o + sizeof(A_), &B_, sizeof(B_));
  memcpy(o + sizeof(A_) + sizeof(B_), &C_, sizeof(C_));
  return 0;
}

// ------------------------------------------------------------------------------------------
// ark-serialize compressed encodings
// ------------------------------------------------------------------------------------------
template <class P>
static void ser_fp(const Fp<P>& a, uint8_t* out, uint8_t flags) {
  Fp<P> v = a.from_mont();
  memcpy(out, v.l, 38);
  out[37] |= flags;
}
template <class P> static uint8_t* ser_coords(const Fp<P>& x, uint8_t* out, uint8_t flags) { ser_fp(x, out, flags); return out + 38; }
template <class B, u64 NR> static uint8_t* ser_coords(const Fp2<B, NR>& x, uint8_t* out, uint8_t flags) {
  ser_fp(x.c0, out, 0); ser_fp(x.c1, out + 38, flags); return out + 76;
}
template <class B, u64 NR> static uint8_t* ser_coords(const Fp3<B, NR>& x, uint8_t* out, uint8_t flags) {
  ser_fp(x.c0, out, 0); ser_fp(x.c1, out + 38, 0); ser_fp(x.c2, out + 76, flags); return out + 114;
}
template <class C>
static uint8_t* ser_point(const Aff<C>& p, uint8_t* out) {
  if (p.is_inf()) return ser_coords(C::F::zero(), out, 0x40);
  return ser_coords(p.x, out, p.y.lex_largest() ? 0x80 : 0);
}

// ------------------------------------------------------------------------------------------
// C API
// ------------------------------------------------------------------------------------------
template <class F> static F ldf(const void* p) { F r; memcpy((void*)&r, p, sizeof(F)); return r; }
template <class F> static void stf(void* p, const F& a) { memcpy(p, (const void*)&a, sizeof(F)); }

template <class F>
static void field_op(int op, const void* a, const void* b, void* out) {
  F x = ldf<F>(a), y = ldf<F>(b), r = F::zero();
  switch (op) {
    case 0: r = x + y; break;
    case 1: r = x - y; break;
    case 2: r = x * y; break;
    case 3: r = x.sqr(); break;
    case 4: r = x.inverse(); break;
    case 5: r = x.neg(); break;
    case 10: r = x.dbl(); break;
  }
  stf(out, r);
}

template <class C>
static void msm_entry(const void* bases, const void* scalars, size_t n, void* out, int threads, int c) {
  Jac<C> r = msm_pippenger<C>((const Aff<C>*)bases, (const u64*)scalars, n, threads, c);
  stf(out, r.to_affine());
}
template <class C>
static void fixed_mul_entry(const void* base, const void* scalars, size_t n, void* out, int threads) {
  Jac<C> b = Jac<C>::from_affine(ldf<Aff<C>>(base));
  Aff<C>* o = (Aff<C>*)out;
  const u64* k = (const u64*)scalars;
  parallel_for(n, threads, [&](size_t lo, size_t hi, int) {
    for (size_t i = lo; i < hi; i++) o[i] = Jac<C>::mul(b, k + 5 * i, 5).to_affine();
  });
}
template <class C>
static void point_sum_entry(const void* pts, size_t n, void* out) {
  Jac<C> acc = Jac<C>::inf();
  const Aff<C>* p = (const Aff<C>*)pts;
  for (size_t i = 0; i < n; i++) acc.add_mixed(p[i]);
  stf(out, acc.to_affine());
}
template <class C>
static int on_curve_entry(const void* pts, size_t n, const void* b_coeff) {
  typedef typename C::F F;
  F b = ldf<F>(b_coeff);
  const Aff<C>* p = (const Aff<C>*)pts;
  for (size_t i = 0; i < n; i++) {
    if (p[i].is_inf()) continue;
    F rhs = p[i].x.sqr() * p[i].x + C::a() * p[i].x + b;
    if (p[i].y.sqr() != rhs) return 0;
  }
  return 1;
}

extern "C" {
// field: 0 = r4, 1 = q4, 2 = Fq2 (over q4), 3 = Fq3 (over r4)
void orc_field_op(int field, int op, const void* a, const void* b, void* out) {
  switch (field) {
    case 0: field_op<FpR4>(op, a, b, out); break;
    case 1: field_op<FpQ4>(op, a, b, out); break;
    case 2: field_op<Fq2>(op, a, b, out); break;
    default: field_op<Fq3>(op, a, b, out); break;
  }
}
// data: n = 2^log_n Montgomery elements, transformed in place
void orc_ntt(int field, void* data, int log_n, int inverse, int coset, int threads) {
  if (field == 0) domain_transform<PR4>((FpR4*)data, log_n, inverse, coset, threads);
  else domain_transform<PQ4>((FpQ4*)data, log_n, inverse, coset, threads);
}
// The DEFINITION of the forward transform at a few output indices (ark-poly EvaluationDomain::fft: out[i] =
// sum_j in[j] (g^c omega^i)^j, c = 1 on the coset): Horner evaluation of the input as a polynomial at the point
// g^c omega_n^i, n multiplications per index, independent of every FFT code path.  Used to check transforms too large
// to recompute in full (2^24: the benchmarked size).  out: count elements; idx: count output indices < 2^log_n.
void orc_dft_at(int field, const void* data, int log_n, int coset, const u64* idx, size_t count, void* out, int threads) {
  auto run = [&](auto tag) {
    typedef decltype(tag) P;
    typedef Fp<P> F;
    const F* a = (const F*)data;
    F* o = (F*)out;
    const size_t n = (size_t)1 << log_n;
    const F omega = omega_for<P>(log_n);
    const F g = F::generator();
    parallel_for(count, threads, [&](size_t lo, size_t hi, int) {
      for (size_t t = lo; t < hi; t++) {
        u64 e[1] = {idx[t]};
        F x = omega.pow(e, 1);
        if (coset) x = x * g;
        F acc = a[n - 1];
        for (size_t j = n - 1; j-- > 0;) acc = acc * x + a[j];
        o[t] = acc;
      }
    });
  };
  if (field == 0) run(PR4{}); else run(PQ4{});
}
// GeneralEvaluationDomain::new(min_size): size chosen (0 if none), with its 7-adic and 2-adic exponents
size_t orc_domain_size(int field, size_t min_size, int* a, int* b) {
  DomainShape d;
  bool ok = field == 0 ? domain_shape<PR4>(min_size, &d) : domain_shape<PQ4>(min_size, &d);
  if (!ok) return 0;
  *a = d.a;
  *b = d.b;
  return d.n;
}
// transform on an explicit domain 7^a 2^b (a > 0 only on q4)
void orc_ntt_general(int field, void* data, int a, int b, int inverse, int coset, int threads) {
  DomainShape d{(size_t)(a == 0 ? 1 : (a == 1 ? 7 : 49)) << b, a, b};
  if (field == 0) domain_transform_general<PR4>((FpR4*)data, d, inverse, coset, threads);
  else domain_transform_general<PQ4>((FpQ4*)data, d, inverse, coset, threads);
}
// curve: 0 = MNT4 G1, 1 = MNT4 G2, 2 = MNT6 G1, 3 = MNT6 G2.  c = 0 -> arkworks window rule
void orc_msm(int curve, const void* bases, const void* scalars, size_t n, void* out_affine, int threads, int c) {
  switch (curve) {
    case 0: msm_entry<C4G1>(bases, scalars, n, out_affine, threads, c); break;
    case 1: msm_entry<C4G2>(bases, scalars, n, out_affine, threads, c); break;
    case 2: msm_entry<C6G1>(bases, scalars, n, out_affine, threads, c); break;
    default: msm_entry<C6G2>(bases, scalars, n, out_affine, threads, c); break;
  }
}
// out[i] = [scalars[i]] base  (double-and-add; the definition, used to build test keys/points)
void orc_fixed_base_mul(int curve, const void* base, const void* scalars, size_t n, void* out, int threads) {
  switch (curve) {
    case 0: fixed_mul_entry<C4G1>(base, scalars, n, out, threads); break;
    case 1: fixed_mul_entry<C4G2>(base, scalars, n, out, threads); break;
    case 2: fixed_mul_entry<C6G1>(base, scalars, n, out, threads); break;
    default: fixed_mul_entry<C6G2>(base, scalars, n, out, threads); break;
  }
}
void orc_point_sum(int curve, const void* pts, size_t n, void* out) {
  switch (curve) {
    case 0: point_sum_entry<C4G1>(pts, n, out); break;
    case 1: point_sum_entry<C4G2>(pts, n, out); break;
    case 2: point_sum_entry<C6G1>(pts, n, out); break;
    default: point_sum_entry<C6G2>(pts, n, out); break;
  }
}
int orc_on_curve(int curve, const void* pts, size_t n, const void* b_coeff) {
  switch (curve) {
    case 0: return on_curve_entry<C4G1>(pts, n, b_coeff);
    case 1: return on_curve_entry<C4G2>(pts, n, b_coeff);
    case 2: return on_curve_entry<C6G1>(pts, n, b_coeff);
    default: return on_curve_entry<C6G2>(pts, n, b_coeff);
  }
}
// Montgomery <-> plain conversion of n elements
void orc_from_mont(int field, const void* in, void* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (field == 0) stf((char*)out + 40 * i, ldf<FpR4>((const char*)in + 40 * i).from_mont());
    else stf((char*)out + 40 * i, ldf<FpQ4>((const char*)in + 40 * i).from_mont());
  }
}
void orc_to_mont(int field, const void* in, void* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    if (field == 0) stf((char*)out + 40 * i, ldf<FpR4>((const char*)in + 40 * i).to_mont());
    else stf((char*)out + 40 * i, ldf<FpQ4>((const char*)in + 40 * i).to_mont());
  }
}
// pairing: 0 = MNT4-298 (Fr = r4), 1 = MNT6-298 (Fr = q4).  h: n = next_pow2(m + num_inputs) elements.
int orc_witness_map(int pairing, const uint32_t* a_ptr, const uint32_t* a_col, const void* a_val,
                    const uint32_t* b_ptr, const uint32_t* b_col, const void* b_val, const uint32_t* c_ptr,
                    const uint32_t* c_col, const void* c_val, size_t m, size_t num_inputs, const void* z, void* h,
                    int threads) {
  Csr A{a_ptr, a_col, (const u64*)a_val}, B{b_ptr, b_col, (const u64*)b_val}, C{c_ptr, c_col, (const u64*)c_val};
  if (pairing == 0) return witness_map<PR4>(A, B, C, m, num_inputs, (const FpR4*)z, (FpR4*)h, threads);
  return witness_map<PQ4>(A, B, C, m, num_inputs, (const FpQ4*)z, (FpQ4*)h, threads);
}
// pk: 10 pointers in Groth16PkView order.  out: A (G1 affine) || B (G2 affine) || C (G1 affine).
int orc_groth16_prove(int pairing, const void* const* pk, const uint32_t* a_ptr, const uint32_t* a_col,
                      const void* a_val, const uint32_t* b_ptr, const uint32_t* b_col, const void* b_val,
                      const uint32_t* c_ptr, const uint32_t* c_col, const void* c_val, size_t m, size_t num_inputs,
                      size_t num_witness, const void* z, const void* r, const void* s, void* out, int threads) {
  Groth16PkView v{pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7], pk[8], pk[9]};
  Csr A{a_ptr, a_col, (const u64*)a_val}, B{b_ptr, b_col, (const u64*)b_val}, C{c_ptr, c_col, (const u64*)c_val};
  if (pairing == 0)
    return groth16_prove<C4G1, C4G2, PR4>(v, A, B, C, m, num_inputs, num_witness, (const FpR4*)z, (const u64*)r,
                                          (const u64*)s, out, threads);
  return groth16_prove<C6G1, C6G2, PQ4>(v, A, B, C, m, num_inputs, num_witness, (const FpQ4*)z, (const u64*)r,
                                        (const u64*)s, out, threads);
}
// GM17.  full: the SAP assignment (num_inputs + num_witness + m + num_inputs - 1), h: n + 1 coefficients;
// d1, d2 in Montgomery form here.  Returns the SAP domain size n (or -1).
int orc_sap_witness_map(int pairing, const uint32_t* a_ptr, const uint32_t* a_col, const void* a_val,
                        const uint32_t* b_ptr, const uint32_t* b_col, const void* b_val, const uint32_t* c_ptr,
                        const uint32_t* c_col, const void* c_val, size_t m, size_t num_inputs, size_t num_witness,
                        const void* z, const void* d1, const void* d2, void* full, void* h, int threads) {
  Csr A{a_ptr, a_col, (const u64*)a_val}, B{b_ptr, b_col, (const u64*)b_val}, C{c_ptr, c_col, (const u64*)c_val};
  if (pairing == 0)
    return sap_witness_map<PR4>(A, B, C, m, num_inputs, num_witness, (const FpR4*)z, ldf<FpR4>(d1), ldf<FpR4>(d2),
                                (FpR4*)full, (FpR4*)h, threads);
  return sap_witness_map<PQ4>(A, B, C, m, num_inputs, num_witness, (const FpQ4*)z, ldf<FpQ4>(d1), ldf<FpQ4>(d2),
                              (FpQ4*)full, (FpQ4*)h, threads);
}
// pk: 9 pointers in Gm17PkView order; d1, d2, r plain integers (like r, s of orc_groth16_prove).
// out: A (G1 affine) || B (G2 affine) || C (G1 affine).
int orc_gm17_prove(int pairing, const void* const* pk, const uint32_t* a_ptr, const uint32_t* a_col, const void* a_val,
                   const uint32_t* b_ptr, const uint32_t* b_col, const void* b_val, const uint32_t* c_ptr,
                   const uint32_t* c_col, const void* c_val, size_t m, size_t num_inputs, size_t num_witness,
                   const void* z, const void* d1, const void* d2, const void* r, void* out, int threads) {
  Gm17PkView v{pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7], pk[8]};
  Csr A{a_ptr, a_col, (const u64*)a_val}, B{b_ptr, b_col, (const u64*)b_val}, C{c_ptr, c_col, (const u64*)c_val};
  if (pairing == 0)
    return gm17_prove<C4G1, C4G2, PR4>(v, A, B, C, m, num_inputs, num_witness, (const FpR4*)z, (const u64*)d1,
                                       (const u64*)d2, (const u64*)r, out, threads);
  return gm17_prove<C6G1, C6G2, PQ4>(v, A, B, C, m, num_inputs, num_witness, (const FpQ4*)z, (const u64*)d1,
                                     (const u64*)d2, (const u64*)r, out, threads);
}
// proof affine bytes (as written by orc_groth16_prove / pcdgpu_groth16_prove) -> canonical
// compressed bytes; returns the length (152 MNT4, 190 MNT6)
int orc_serialize_proof(int pairing, const void* proof_affine, uint8_t* out) {
  const char* p = (const char*)proof_affine;
  uint8_t* o = out;
  memset(out, 0, pairing == 0 ? 152 : 190);
  if (pairing == 0) {
    o = ser_point(ldf<Aff<C4G1>>(p), o);
    o = ser_point(ldf<Aff<C4G2>>(p + 80), o);
    o = ser_point(ldf<Aff<C4G1>>(p + 240), o);
  } else {
    o = ser_point(ldf<Aff<C6G1>>(p), o);
    o = ser_point(ldf<Aff<C6G2>>(p + 80), o);
    o = ser_point(ldf<Aff<C6G1>>(p + 320), o);
  }
  return (int)(o - out);
}
int orc_hw_threads() { return (int)std::thread::hardware_concurrency(); }
}  // extern "C"
