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

#include "fp3s.cuh"
#include "fpx.cuh"

template <class F>
struct AffinePoint {
  F x, y;
  PCD_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
  PCD_HD static AffinePoint inf() { AffinePoint p; p.x = F::zero(); p.y = F::zero(); return p; }
  PCD_HD AffinePoint neg() const { AffinePoint p; p.x = x; p.y = y.neg(); return p; }
};

// Curve tags: coordinate field + multiplication by the curve coefficient a.
struct CurveMnt4G1 {  // y^2 = x^3 + 2x + b over F_q4, order r4
  typedef FpQ4 F; typedef ParamsR4 ScalarParams; typedef GenMnt4G1 Gen; static constexpr int ID = 0;
  static constexpr bool OUTLINE = false;
  PCD_HD static F mul_a(const F& v) { return v.dbl(); }
};
struct CurveMnt4G2 {  // twist over Fq2: a' = (34, 0)
  typedef Fq2 F; typedef ParamsR4 ScalarParams; typedef GenMnt4G2 Gen; static constexpr int ID = 1;
  static constexpr bool OUTLINE = true;  // group operations are real function calls (code size, compile time)
  PCD_HD static F mul_a(const F& v) { return v.template mul_small<34>(); }
};
struct CurveMnt6G1 {  // y^2 = x^3 + 11x + b over F_r4, order q4
  typedef FpR4 F; typedef ParamsQ4 ScalarParams; typedef GenMnt6G1 Gen; static constexpr int ID = 2;
  static constexpr bool OUTLINE = false;
  PCD_HD static F mul_a(const F& v) { return v.template mul_small<11>(); }
};
struct CurveMnt6G2 {  // twist over Fq3: a' = (0, 0, 11) = 11 u^2;  u^3 = 5
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

#if defined(__CUDACC__)
// CurveMnt6G2 with its coordinates sliced over three lanes (fp3s.cuh): the bucket-accumulation kernels' twin.
// Memory layout of points is CurveMnt6G2's; lane l of a group reads / writes coefficient l of every coordinate.
struct CurveMnt6G2S {
  typedef Fq3S F; typedef ParamsQ4 ScalarParams; static constexpr int ID = 3;
  static constexpr bool OUTLINE = false;
  __device__ __forceinline__ static F mul_a(const F& v) {  // a' = 11 u^2: (55 v1, 55 v2, 11 v0)
    const int l = F::li();
    const int base = (int)(threadIdx.x & 31u) - l;
    const FpR4 t = F::shfl(v.c, base + (l == 2 ? 0 : l + 1)).template mul_small<11>();
    F r;
    r.c = F::sel(l == 2, t, t.template mul_small<5>());
    return r;
  }
};
#endif

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
