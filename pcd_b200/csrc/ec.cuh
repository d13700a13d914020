// ec.cuh -- short-Weierstrass group arithmetic y^2 = x^3 + a x + b, a != 0, in XYZZ coordinates
// (X, Y, ZZ, ZZZ) with x = X/ZZ, y = Y/ZZZ.  Replaces ark-ec's GroupProjective
// {add_assign_mixed, double_in_place, add_assign, into_affine} (Jacobian) under
// VariableBaseMSM::multi_scalar_mul; results are compared after into_affine(), which is
// representation independent.  XYZZ mixed addition is 8M + 2S against Jacobian 7M + 4S.
//
// ABI point layout: x || y, each a field element in Montgomery form; the point at infinity is
// encoded as x = y = 0 (not on any of the four curves since b != 0).
#pragma once
#include <type_traits>

#include "fp30.cuh"
#include "fpx.cuh"

template <class F>
struct AffinePoint {
  F x, y;
  PCD_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  PCD_HD static AffinePoint inf() { AffinePoint p; p.x = F::zero(); p.y = F::zero(); return p; }
  PCD_HD AffinePoint neg() const { AffinePoint p; p.x = x; p.y = y.neg(); return p; }
};

// Radix-2^30 twins of the two G1 curves: same group law over Fp30 (carry-free products, values in [0, 2p),
// Montgomery radix 2^300).  With -DPCD_FAST30 the bucket-accumulation kernels compute in them (points enter
// through fast_affine() / tables stored in that form, buckets leave through slow_xyzz()).  OFF by default:
// measured on B200 the radix-2^30 product is no faster than the carry-chain one (42.4 vs 44.1 G products/s)
// because IMAD.WIDE issues on the fmaheavy pipe at 32 lanes/clk/SM whatever its carry predicates, and both
// products keep that pipe 93-97 % busy (profiles/r01_ncu_modmul.csv); the extra shifts/masks made
// msm_accumulate 16 % slower.  Kept (and tested in tests/hostemu) as the starting point for a product that
// leaves the fmaheavy pipe (e.g. FP64 DFMA limbs), which is the only way past that roof.
struct CurveMnt4G1Fast {
  typedef Fp30Q4 F; static constexpr bool OUTLINE = false;
  PCD_HD static F mul_a(const F& v) { return v.dbl(); }
};
struct CurveMnt6G1Fast {
  typedef Fp30R4 F; static constexpr bool OUTLINE = false;
  PCD_HD static F mul_a(const F& v) { return v.template mul_small<11>(); }
};

// Curve tags: coordinate field + multiplication by the curve coefficient a.
struct CurveMnt4G1 {  // y^2 = x^3 + 2x + b over F_q4, order r4
#ifdef PCD_FAST30
  typedef CurveMnt4G1Fast Fast;
#else
  typedef CurveMnt4G1 Fast;
#endif
  typedef FpQ4 F; typedef ParamsR4 ScalarParams; typedef GenMnt4G1 Gen; static constexpr int ID = 0;
  static constexpr bool OUTLINE = false;
  PCD_HD static F mul_a(const F& v) { return v.dbl(); }
};
// Measured (B200, 2^20 proof): with the base products inlined as well the G2 walk needs 255 registers + spills and
// takes 8.3 ms against 7.3 ms with out-of-line products (and msm_c1.cu compiles in 6 min instead of 25 s) -- so
// the twin is OFF unless -DPCD_G2_INLINE_PRODUCTS; only the group law (madd_impl) is inlined in the walk.
struct CurveMnt4G2Inl {  // CurveMnt4G2 with inlined base products and group law: the accumulate kernels' twin
  typedef Fq2I F; static constexpr bool OUTLINE = false; static constexpr bool BITCAST = true;
  PCD_HD static F mul_a(const F& v) { return v.template mul_small<34>(); }
};
struct CurveMnt4G2 {  // twist over Fq2: a' = (34, 0)
#ifdef PCD_G2_INLINE_PRODUCTS
  typedef CurveMnt4G2Inl Fast;
#else
  typedef CurveMnt4G2 Fast;
#endif
  typedef Fq2 F; typedef ParamsR4 ScalarParams; typedef GenMnt4G2 Gen; static constexpr int ID = 1;
  static constexpr bool OUTLINE = true;  // group operations are real function calls (code size, compile time)
  PCD_HD static F mul_a(const F& v) { return v.template mul_small<34>(); }
};
struct CurveMnt6G1 {  // y^2 = x^3 + 11x + b over F_r4, order q4
#ifdef PCD_FAST30
  typedef CurveMnt6G1Fast Fast;
#else
  typedef CurveMnt6G1 Fast;
#endif
  typedef FpR4 F; typedef ParamsQ4 ScalarParams; typedef GenMnt6G1 Gen; static constexpr int ID = 2;
  static constexpr bool OUTLINE = false;
  PCD_HD static F mul_a(const F& v) { return v.template mul_small<11>(); }
};
struct CurveMnt6G2 {  // twist over Fq3: a' = (0, 0, 11) = 11 u^2;  u^3 = 5
  typedef CurveMnt6G2 Fast;
  typedef Fq3 F; typedef ParamsQ4 ScalarParams; typedef GenMnt6G2 Gen; static constexpr int ID = 3;
  static constexpr bool OUTLINE = true;
  PCD_HD static F mul_a(const F& v) {
    F r;
    r.c0 = v.c1.template mul_small<55>();
    r.c1 = v.c2.template mul_small<55>();
    r.c2 = v.c0.template mul_small<11>();
    return r;
  }
};

// ABI field element (ten 32-bit words of x * 2^320 mod p, canonical) <-> radix-2^30 element (x * 2^300, < 2p):
// bit repacking and one product by a constant each way.
template <class P30, class P>
PCD_HD Fp30<P30> to_fast(const Fp<P>& a) {
  static_assert(P30::ID == P::ID, "same field");
  return Fp30<P30>::from_words(a.l) * Fp30<P30>::konst(P30::abi_to_int);
}
template <class P, class P30>
PCD_HD Fp<P> from_fast(const Fp30<P30>& a) {
  static_assert(P30::ID == P::ID, "same field");
  Fp<P> r;
  (a * Fp30<P30>::konst(P30::int_to_abi)).canonical().to_words(r.l);
  return r;
}

template <class C>
struct XYZZ;
// the same for points; identity for the curves without a radix-2^30 twin (C::Fast == C)
template <class C>
PCD_HD AffinePoint<typename C::Fast::F> fast_affine(const AffinePoint<typename C::F>& p) {
  if constexpr (std::is_same<C, typename C::Fast>::value) {
    return p;
  } else if constexpr (sizeof(AffinePoint<typename C::Fast::F>) == sizeof(AffinePoint<typename C::F>) &&
                       C::Fast::F::WORDS > FP_LIMBS) {  // layout-identical twin of an extension field: same words
    AffinePoint<typename C::Fast::F> q;
    const u32* s = reinterpret_cast<const u32*>(&p);
    u32* d = reinterpret_cast<u32*>(&q);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(q) / 4); i++) d[i] = s[i];
    return q;
  } else {
    typedef typename C::Fast::F::Params P30;
    AffinePoint<typename C::Fast::F> q;
    q.x = to_fast<P30>(p.x);
    q.y = to_fast<P30>(p.y);
    return q;
  }
}
template <class C>
PCD_HD XYZZ<C> slow_xyzz(const XYZZ<typename C::Fast>& p);

template <class C>
struct XYZZ {
  typedef typename C::F F;
  F x, y, zz, zzz;

  PCD_HD static XYZZ inf() { XYZZ p; p.x = F::zero(); p.y = F::zero(); p.zz = F::zero(); p.zzz = F::zero(); return p; }
  PCD_HD bool is_inf() const { return zz.is_zero(); }
  PCD_HD static XYZZ from_affine(const AffinePoint<F>& a) {
    if (a.is_inf()) return inf();
    XYZZ p; p.x = a.x; p.y = a.y; p.zz = F::one(); p.zzz = F::one();
    return p;
  }
  PCD_HD XYZZ neg() const { XYZZ p = *this; p.y = y.neg(); return p; }

  // 2 * affine (mdbl-2008-s-1)
  PCD_HD static XYZZ dbl_affine(const AffinePoint<F>& a) {
    if (a.is_inf() || a.y.is_zero()) return inf();
    F U = a.y.dbl();
    F V = U.sqr();
    F W = U * V;
    F S = a.x * V;
    F xx = a.x.sqr();
    F M = xx.dbl() + xx + C::mul_a(F::one());
    XYZZ p;
    p.x = M.sqr() - S.dbl();
    p.y = M * (S - p.x) - W * a.y;
    p.zz = V;
    p.zzz = W;
    return p;
  }
  // Out-of-line entry points: for the G2 curves (C::OUTLINE) every group operation is a call, so a
  // kernel holds one copy of each formula instead of one per use.  Operands and results travel BY
  // VALUE (param space): passing the addresses of locals let nvcc 12.9's stack colouring overlay a
  // result temporary with the live accumulator (seen in PTX: to_affine_out(dst == src)).
  PCD_NOINLINE static XYZZ dbl_out(XYZZ src) { return src.dbl_impl(); }
  PCD_NOINLINE static XYZZ madd_out(XYZZ self, AffinePoint<F> a) { self.madd_impl(a); return self; }
  PCD_NOINLINE static XYZZ add_out(XYZZ self, XYZZ o) { self.add_impl(o); return self; }
  PCD_HD XYZZ dbl() const {
    if constexpr (C::OUTLINE) return dbl_out(*this);
    else return dbl_impl();
  }
  PCD_HD void madd(const AffinePoint<F>& a) {
    if constexpr (C::OUTLINE) *this = madd_out(*this, a);
    else madd_impl(a);
  }
  PCD_HD void add(const XYZZ& o) {
    if constexpr (C::OUTLINE) *this = add_out(*this, o);
    else add_impl(o);
  }
  // dbl-2008-s-1
  PCD_HD XYZZ dbl_impl() const {
    if (is_inf() || y.is_zero()) return inf();
    F U = y.dbl();
    F V = U.sqr();
    F W = U * V;
    F S = x * V;
    F xx = x.sqr();
    F M = xx.dbl() + xx + C::mul_a(zz.sqr());
    XYZZ p;
    p.x = M.sqr() - S.dbl();
    p.y = M * (S - p.x) - W * y;
    p.zz = V * zz;
    p.zzz = W * zzz;
    return p;
  }
  // this += affine (madd-2008-s), all exceptional cases handled
  PCD_HD void madd_impl(const AffinePoint<F>& a) {
    if (a.is_inf()) return;
    if (is_inf()) { *this = from_affine(a); return; }
    F P = a.x * zz - x;
    F R = a.y * zzz - y;
    if (P.is_zero()) {
      if (R.is_zero()) *this = dbl_affine(a);
      else *this = inf();
      return;
    }
    F PP = P.sqr();
    F PPP = P * PP;
    F Q = x * PP;
    F x3 = R.sqr() - PPP - Q.dbl();
    y = R * (Q - x3) - y * PPP;
    x = x3;
    zz = zz * PP;
    zzz = zzz * PPP;
  }
  // this += o (add-2008-s)
  PCD_HD void add_impl(const XYZZ& o) {
    if (o.is_inf()) return;
    if (is_inf()) { *this = o; return; }
    F U1 = x * o.zz;
    F U2 = o.x * zz;
    F S1 = y * o.zzz;
    F S2 = o.y * zzz;
    F P = U2 - U1;
    F R = S2 - S1;
    if (P.is_zero()) {
      if (R.is_zero()) *this = dbl();
      else *this = inf();
      return;
    }
    F PP = P.sqr();
    F PPP = P * PP;
    F Q = U1 * PP;
    F x3 = R.sqr() - PPP - Q.dbl();
    y = R * (Q - x3) - S1 * PPP;
    x = x3;
    zz = zz * o.zz * PP;
    zzz = zzz * o.zzz * PPP;
  }
  PCD_NOINLINE static AffinePoint<F> to_affine_out(XYZZ src) { return src.to_affine_impl(); }
  PCD_HD AffinePoint<F> to_affine() const { return to_affine_out(*this); }
  PCD_HD AffinePoint<F> to_affine_impl() const {
    if (is_inf()) return AffinePoint<F>::inf();
    F i = zzz.inverse();
    F t = zz * i;  // = 1/Z
    AffinePoint<F> a;
    a.x = x * t.sqr();
    a.y = y * i;
    return a;
  }
  // [k]P, k = nlimbs 32-bit little-endian limbs (plain integer), MSB-first double-and-add
  struct Scalar320 { u32 w[10]; };
  PCD_NOINLINE static XYZZ mul_out(XYZZ p, Scalar320 k, int nlimbs) { return mul_impl(p, k.w, nlimbs); }
  PCD_HD static XYZZ mul(const XYZZ& p, const u32* k, int nlimbs) {
    Scalar320 kk;
#pragma unroll
    for (int i = 0; i < 10; i++) kk.w[i] = i < nlimbs ? k[i] : 0u;
    return mul_out(p, kk, nlimbs);
  }
  PCD_HD static XYZZ mul_impl(const XYZZ& p, const u32* k, int nlimbs) {
    XYZZ acc = inf();
    bool started = false;
    for (int i = nlimbs - 1; i >= 0; i--) {
      u32 w = k[i];
      for (int b = 31; b >= 0; b--) {
        if (started) acc = acc.dbl();
        if ((w >> b) & 1) { acc.add(p); started = true; }
      }
    }
    return acc;
  }
};

template <class C>
PCD_HD XYZZ<C> slow_xyzz(const XYZZ<typename C::Fast>& p) {
  if constexpr (std::is_same<C, typename C::Fast>::value) {
    return p;
  } else if constexpr (sizeof(XYZZ<typename C::Fast>) == sizeof(XYZZ<C>) && C::Fast::F::WORDS > FP_LIMBS) {
    XYZZ<C> q;
    const u32* s = reinterpret_cast<const u32*>(&p);
    u32* d = reinterpret_cast<u32*>(&q);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(q) / 4); i++) d[i] = s[i];
    return q;
  } else {
    typedef typename C::F::Params P;
    XYZZ<C> q;
    q.x = from_fast<P>(p.x);
    q.y = from_fast<P>(p.y);
    q.zz = from_fast<P>(p.zz);
    q.zzz = from_fast<P>(p.zzz);
    return q;
  }
}
