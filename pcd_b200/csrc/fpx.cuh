// fpx.cuh -- quadratic and cubic extensions used by the G2 twists of the MNT cycle.
//   Fq2 = F_q4[u]/(u^2 - 17)  (MNT4-298 G2 coordinates; ark-ff Fp2, ark-mnt4-298 fq2.rs)
//   Fq3 = F_r4[u]/(u^3 - 5)   (MNT6-298 G2 coordinates; ark-ff Fp3, ark-mnt6-298 fq3.rs)
// In-memory layout = arkworks': coefficients c0, c1, (c2) contiguous, each an Fp.
#pragma once
#include "fp.cuh"

template <class B, u32 NR>
struct Fp2T {
  B c0, c1;
  typedef B Base;
  // base products are out-of-line calls: one copy of the 260-instruction product per kernel
  PCD_HD static B mulb(const B& a, const B& b) { return B::mul_ni(a, b); }
  static constexpr int WORDS = 2 * B::WORDS;
  PCD_HD static Fp2T zero() { Fp2T r; r.c0 = B::zero(); r.c1 = B::zero(); return r; }
  PCD_HD static Fp2T one() { Fp2T r; r.c0 = B::one(); r.c1 = B::zero(); return r; }
  PCD_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  PCD_HD bool operator==(const Fp2T& o) const { return c0 == o.c0 && c1 == o.c1; }
  PCD_HD bool operator!=(const Fp2T& o) const { return !(*this == o); }
  PCD_HD friend Fp2T operator+(const Fp2T& a, const Fp2T& b) { Fp2T r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
  PCD_HD friend Fp2T operator-(const Fp2T& a, const Fp2T& b) { Fp2T r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
  PCD_HD Fp2T neg() const { Fp2T r; r.c0 = c0.neg(); r.c1 = c1.neg(); return r; }
  PCD_HD Fp2T dbl() const { Fp2T r; r.c0 = c0.dbl(); r.c1 = c1.dbl(); return r; }
  // Karatsuba: 3 base products
  PCD_HD friend Fp2T operator*(const Fp2T& a, const Fp2T& b) {
    B v0 = mulb(a.c0, b.c0);
    B v1 = mulb(a.c1, b.c1);
    Fp2T r;
    r.c1 = mulb(a.c0 + a.c1, b.c0 + b.c1) - v0 - v1;
    r.c0 = v0 + v1.template mul_small<NR>();
    return r;
  }
  // complex squaring: 2 base products
  PCD_HD Fp2T sqr() const {
    B ab = mulb(c0, c1);
    B t = mulb(c0 + c1, c0 + c1.template mul_small<NR>());
    Fp2T r;
    r.c0 = t - ab - ab.template mul_small<NR>();
    r.c1 = ab.dbl();
    return r;
  }
  template <u32 K>
  PCD_HD Fp2T mul_small() const { Fp2T r; r.c0 = c0.template mul_small<K>(); r.c1 = c1.template mul_small<K>(); return r; }
  PCD_HD Fp2T inverse() const {
    B n = mulb(c0, c0) - mulb(c1, c1).template mul_small<NR>();
    B ni = n.inverse();
    Fp2T r; r.c0 = mulb(c0, ni); r.c1 = mulb(c1, ni).neg();
    return r;
  }
  // ark-ec "is y the larger of {y,-y}": compare the highest coefficient first
  PCD_HD bool lexicographically_largest() const {
    if (!c1.is_zero()) return c1.lexicographically_largest();
    return c0.lexicographically_largest();
  }
};

template <class B, u32 NR>
struct Fp3T {
  B c0, c1, c2;
  typedef B Base;
  static constexpr int WORDS = 3 * B::WORDS;
  PCD_HD static Fp3T zero() { Fp3T r; r.c0 = B::zero(); r.c1 = B::zero(); r.c2 = B::zero(); return r; }
  PCD_HD static Fp3T one() { Fp3T r; r.c0 = B::one(); r.c1 = B::zero(); r.c2 = B::zero(); return r; }
  PCD_HD bool is_zero() const { return c0.is_zero() && c1.is_zero() && c2.is_zero(); }
  PCD_HD bool operator==(const Fp3T& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
  PCD_HD bool operator!=(const Fp3T& o) const { return !(*this == o); }
  PCD_HD friend Fp3T operator+(const Fp3T& a, const Fp3T& b) { Fp3T r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; r.c2 = a.c2 + b.c2; return r; }
  PCD_HD friend Fp3T operator-(const Fp3T& a, const Fp3T& b) { Fp3T r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; r.c2 = a.c2 - b.c2; return r; }
  PCD_HD Fp3T neg() const { Fp3T r; r.c0 = c0.neg(); r.c1 = c1.neg(); r.c2 = c2.neg(); return r; }
  PCD_HD Fp3T dbl() const { Fp3T r; r.c0 = c0.dbl(); r.c1 = c1.dbl(); r.c2 = c2.dbl(); return r; }
  // Karatsuba: 6 base products
  PCD_HD friend Fp3T operator*(const Fp3T& a, const Fp3T& b) {
    B v0 = B::mul_ni(a.c0, b.c0), v1 = B::mul_ni(a.c1, b.c1), v2 = B::mul_ni(a.c2, b.c2);
    Fp3T r;
    r.c0 = v0 + (B::mul_ni(a.c1 + a.c2, b.c1 + b.c2) - v1 - v2).template mul_small<NR>();
    r.c1 = B::mul_ni(a.c0 + a.c1, b.c0 + b.c1) - v0 - v1 + v2.template mul_small<NR>();
    r.c2 = B::mul_ni(a.c0 + a.c2, b.c0 + b.c2) - v0 - v2 + v1;
    return r;
  }
  // Chung-Hasan SQR2: 5 base products
  PCD_HD Fp3T sqr() const {
    B s0 = B::mul_ni(c0, c0);
    B s1 = B::mul_ni(c0, c1).dbl();
    B t2 = c0 - c1 + c2;
    B s2 = B::mul_ni(t2, t2);
    B s3 = B::mul_ni(c1, c2).dbl();
    B s4 = B::mul_ni(c2, c2);
    Fp3T r;
    r.c0 = s0 + s3.template mul_small<NR>();
    r.c1 = s1 + s4.template mul_small<NR>();
    r.c2 = s1 + s2 + s3 - s0 - s4;
    return r;
  }
  template <u32 K>
  PCD_HD Fp3T mul_small() const { Fp3T r; r.c0 = c0.template mul_small<K>(); r.c1 = c1.template mul_small<K>(); r.c2 = c2.template mul_small<K>(); return r; }
  PCD_HD Fp3T inverse() const {
    B t0 = B::mul_ni(c0, c0) - B::mul_ni(c1, c2).template mul_small<NR>();
    B t1 = B::mul_ni(c2, c2).template mul_small<NR>() - B::mul_ni(c0, c1);
    B t2 = B::mul_ni(c1, c1) - B::mul_ni(c0, c2);
    B n = B::mul_ni(c0, t0) + (B::mul_ni(c2, t1) + B::mul_ni(c1, t2)).template mul_small<NR>();
    B ni = n.inverse();
    Fp3T r; r.c0 = B::mul_ni(t0, ni); r.c1 = B::mul_ni(t1, ni); r.c2 = B::mul_ni(t2, ni);
    return r;
  }
  PCD_HD bool lexicographically_largest() const {
    if (!c2.is_zero()) return c2.lexicographically_largest();
    if (!c1.is_zero()) return c1.lexicographically_largest();
    return c0.lexicographically_largest();
  }
};

typedef Fp2T<FpQ4, 17> Fq2;  // MNT4-298 twist field
typedef Fp3T<FpR4, 5> Fq3;   // MNT6-298 twist field
