// fp30.cuh -- 298-bit prime field arithmetic in radix 2^30: ten 30-bit limbs held in 32-bit words,
// Montgomery form with R' = 2^300, values kept in [0, 2p) ("lazy" reduction; R' > 4p makes the
// product of two such values again < 2p without a final subtraction).
//
// Why not the classic 10 x 32-bit carry-chain product (prims.cuh / fp.cuh): measured on B200, the
// carry-chained multiply-add IMAD.WIDE.U32.X issues at HALF the rate of the plain IMAD.WIDE.U32
// (9.15 vs 17.9 T/s, tools/gpu_probe.py), and the single PTX carry flag serialises the whole product
// (IPC 0.31 per scheduler in msm_accumulate, profiles/r01_ncu_msm_accumulate_g1.csv).  With 30-bit
// limbs a 64-bit column accumulator absorbs up to 16 products without overflowing, so the product
// is 210 INDEPENDENT full-rate IMAD.WIDE.U32 (t[j] += a[j] * b[i]) plus shifts/masks on the ALU pipe:
// no carry flag, no inline PTX, plain C++ that also compiles for the host (tests/hostemu).
//
// Replaces ark-ff's Fp320 arithmetic (5 x u64 CIOS) below IC::MainSNARK::prove /
// IC::HelpSNARK::prove (/root/reference/src/ec_cycle_pcd/mod.rs:171,179).  The ABI encoding
// (ark-ff's: ten 32-bit words of a * 2^320 mod p, canonical) is converted at kernel boundaries by
// from_abi_words / to_abi_words (bit repacking; the Montgomery factor is handled by the callers'
// constants, see the notes at each kernel).
#pragma once
#include <cstdint>

#include "constants30.cuh"
#include "prims.cuh"

#if defined(__CUDACC__)
#define F30_HD __host__ __device__ __forceinline__
#define F30_NOINLINE __host__ __device__ __noinline__
#else
#define F30_HD inline
#define F30_NOINLINE
#endif

typedef uint64_t u64;

static constexpr u32 MASK30 = (1u << 30) - 1;

// acc += a * b as exactly one IMAD.WIDE.U32 (the plain C++ expression makes nvcc 12.9 emit extra adds
// of zero high words).  Not volatile, no carry flag: ptxas schedules these freely.
F30_HD void mad_wide(u64& acc, u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
  asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
#else
  acc += (u64)a * b;
#endif
}

template <class P>
struct Fp30 {
  u32 l[10];  // value = sum l[i] 2^(30 i); every limb < 2^30; value < 2p
  typedef P Params;
  static constexpr int WORDS = 10;

  F30_HD static Fp30 zero() {
    Fp30 r;
#pragma unroll
    for (int i = 0; i < 10; i++) r.l[i] = 0;
    return r;
  }
  // 1 in Montgomery form: R' mod p
  F30_HD static Fp30 one() { return konst(P::one); }
  template <class FN>
  F30_HD static Fp30 konst(FN f) {
    Fp30 r;
#pragma unroll
    for (int i = 0; i < 10; i++) r.l[i] = f(i);
    return r;
  }

  // ---- comparisons ----------------------------------------------------------------------------
  F30_HD bool limbs_equal(const Fp30& o) const {
    u32 t = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) t |= l[i] ^ o.l[i];
    return t == 0;
  }
  F30_HD bool limbs_zero() const {
    u32 t = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) t |= l[i];
    return t == 0;
  }
  // value in [0, 2p): zero mod p iff 0 or p
  F30_HD bool is_zero() const { return limbs_zero() || limbs_equal(konst(P::mod)); }
  F30_HD bool operator==(const Fp30& o) const { return (*this - o).is_zero(); }
  F30_HD bool operator!=(const Fp30& o) const { return !(*this == o); }

  // ---- carry handling ---------------------------------------------------------------------------
  // t[0..9] arbitrary 64-bit columns -> limbs < 2^30 (the caller guarantees the value fits 300 bits)
  F30_HD static Fp30 normalize(u64* t) {
    Fp30 r;
#pragma unroll
    for (int j = 0; j < 9; j++) {
      t[j + 1] += t[j] >> 30;
      r.l[j] = (u32)t[j] & MASK30;
    }
    r.l[9] = (u32)t[9];
    return r;
  }
  // r = a - k if a >= k else a, for limb vectors (k given by its limb function); a, k < 2^300
  template <class FN>
  F30_HD static Fp30 cond_sub(const Fp30& a, FN k) {
    Fp30 d;
    int32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
      int32_t v = (int32_t)a.l[i] - (int32_t)k(i) + borrow;  // in (-2^31, 2^30)
      borrow = v >> 31;                                       // 0 or -1
      d.l[i] = (u32)v & MASK30;
    }
    // borrow == -1: a < k, keep a
    Fp30 r;
#pragma unroll
    for (int i = 0; i < 10; i++) r.l[i] = borrow ? a.l[i] : d.l[i];
    return r;
  }

  // ---- addition / subtraction (results in [0, 2p)) ------------------------------------------------
  F30_HD friend Fp30 operator+(const Fp30& a, const Fp30& b) {
    Fp30 s;
    u32 carry = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
      u32 v = a.l[i] + b.l[i] + carry;  // < 2^31 + 1
      carry = v >> 30;
      s.l[i] = v & MASK30;
    }
    // a + b < 4p < 2^300: no carry out of limb 9 beyond its 30 bits
    return cond_sub(s, P::mod2);
  }
  F30_HD friend Fp30 operator-(const Fp30& a, const Fp30& b) {
    Fp30 d;
    int32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
      int32_t v = (int32_t)a.l[i] - (int32_t)b.l[i] + borrow;
      borrow = v >> 31;
      d.l[i] = (u32)v & MASK30;
    }
    // negative (borrow == -1): add 2p; the discarded 2^300 wraps the two's complement back
    u32 carry = 0;
    Fp30 r;
#pragma unroll
    for (int i = 0; i < 10; i++) {
      u32 v = d.l[i] + (borrow ? P::mod2(i) : 0u) + carry;
      carry = v >> 30;
      r.l[i] = v & MASK30;
    }
    return r;
  }
  F30_HD Fp30 neg() const { return zero() - *this; }
  F30_HD Fp30 dbl() const { return *this + *this; }

  // ---- Montgomery product, R' = 2^300 ---------------------------------------------------------------
  // Operand scanning with 64-bit column accumulators held as register PAIRS (lo[j], hi[j]).  Row i adds
  // a[j] * b[i] and m * p[j] to the ten columns (20 independent multiply-adds), where m makes column 0
  // divisible by 2^30; the columns then shift down by one limb.  A column sees at most 2 * 5 products of
  // < 2^60 between the two carry sweeps (after rows 4 and 9), so it stays below 2^64.
  //
  // The multiply-add is written as mad.lo.cc / madc.hi on the 32-bit halves (prims::mac), NOT as
  // mad.wide.u32 on a 64-bit register: ptxas 12.9 rewrites chains of the latter into IMAD.WIDE + 3-input
  // IADD3 pairs (646 alu instructions per two products, 36.6 G products/s on B200), while the former
  // becomes one plain accumulate-form IMAD.WIDE.U32 each -- no carry predicate, so it issues at full rate
  // (a carry-in OR carry-out predicate on IMAD.WIDE halves its issue rate: tools/probe_modmul.py).
  F30_HD friend Fp30 operator*(const Fp30& a, const Fp30& b) {
    u32 lo[10], hi[10];
#pragma unroll
    for (int i = 0; i < 10; i++) {
      const u32 bi = b.l[i];
      if (i == 0) {
#pragma unroll
        for (int j = 0; j < 10; j++) prims::mul_wide(lo[j], hi[j], a.l[j], bi);
      } else {
#pragma unroll
        for (int j = 0; j < 10; j++) prims::mac(lo[j], hi[j], a.l[j], bi);
      }
      const u32 m = (lo[0] * P::PINV30) & MASK30;
#pragma unroll
      for (int j = 0; j < 10; j++) prims::mac(lo[j], hi[j], m, P::mod(j));
      prims::add_shr30(lo[1], hi[1], lo[0], hi[0]);
#pragma unroll
      for (int j = 0; j < 9; j++) {
        lo[j] = lo[j + 1];
        hi[j] = hi[j + 1];
      }
      lo[9] = 0;
      hi[9] = 0;
      if (i == 4) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          prims::add_shr30(lo[j + 1], hi[j + 1], lo[j], hi[j]);
          lo[j] &= MASK30;
          hi[j] = 0;
        }
      }
    }
    Fp30 r;
#pragma unroll
    for (int j = 0; j < 9; j++) {
      prims::add_shr30(lo[j + 1], hi[j + 1], lo[j], hi[j]);
      r.l[j] = lo[j] & MASK30;
    }
    r.l[9] = lo[9];
    return r;
  }
  F30_HD Fp30 sqr() const { return (*this) * (*this); }
  F30_NOINLINE static Fp30 mul_ni(Fp30 a, Fp30 b) { return a * b; }

  // multiply by a small compile-time constant with additions only
  template <u32 K>
  F30_HD Fp30 mul_small() const {
    static_assert(K >= 1 && K < 256, "small constant");
    Fp30 acc = *this;
    bool started = false;
    Fp30 r = *this;
#pragma unroll
    for (int bit = 7; bit >= 0; bit--) {
      if (started) r = r.dbl();
      if ((K >> bit) & 1) {
        if (started) r = r + acc;
        else started = true;
      }
    }
    return r;
  }

  // a^e for a 64-bit exponent
  F30_HD Fp30 pow64(u64 e) const {
    Fp30 r = one();
    Fp30 base = *this;
    while (e) {
      if (e & 1) r = r * base;
      e >>= 1;
      if (e) base = base.sqr();
    }
    return r;
  }
  // a^-1 = a^(p-2) (Fermat); zero maps to zero.  Exponent bits come from the 30-bit limbs of p - 2.
  F30_NOINLINE Fp30 inverse() const {
    Fp30 r = one();
    bool started = false;
    for (int i = 9; i >= 0; i--) {
      u32 e = 0;
#pragma unroll
      for (int k = 0; k < 10; k++)
        if (k == i) e = P::pm2(k);
      for (int b = 29; b >= 0; b--) {
        if (started) r = r.sqr();
        if ((e >> b) & 1) {
          r = started ? r * (*this) : *this;
          started = true;
        }
      }
    }
    return r;
  }

  // ---- canonical form and conversions -----------------------------------------------------------------
  // fully reduced representative in [0, p)
  F30_HD Fp30 canonical() const { return cond_sub(*this, P::mod); }
  // leave Montgomery form: plain integer value, canonical
  F30_HD Fp30 from_mont() const {
    Fp30 o = zero();
    o.l[0] = 1;
    return ((*this) * o).canonical();
  }
  F30_HD Fp30 to_mont() const { return (*this) * konst(P::r2); }
  // ark-ec "is y the larger of {y, -y}" on the plain integer value
  F30_HD bool lexicographically_largest() const {
    Fp30 v = from_mont();
    Fp30 n = neg().from_mont();
#pragma unroll
    for (int i = 9; i >= 0; i--)
      if (v.l[i] != n.l[i]) return v.l[i] > n.l[i];
    return false;
  }

  // ten 32-bit words (a 320-bit little-endian integer < 2^300) <-> ten 30-bit limbs: bit repacking only
  F30_HD static Fp30 from_words(const u32* w) {
    Fp30 r;
#pragma unroll
    for (int i = 0; i < 10; i++) {
      int bit = 30 * i, k = bit >> 5, off = bit & 31;
      u32 v = w[k] >> off;
      if (off > 2 && k + 1 < 10) v |= w[k + 1] << (32 - off);
      r.l[i] = v & MASK30;
    }
    return r;
  }
  F30_HD void to_words(u32* w) const {
#pragma unroll
    for (int k = 0; k < 10; k++) {
      // word k holds bits [32k, 32k + 32): limbs i0 = floor(32k / 30) and i0 + 1 (and i0 + 2 never)
      int bit = 32 * k, i0 = bit / 30, off = bit - 30 * i0;
      u32 v = l[i0] >> off;
      if (i0 + 1 < 10) v |= l[i0 + 1] << (30 - off);
      if (off > 28 && i0 + 2 < 10) v |= l[i0 + 2] << (60 - off);
      w[k] = v;
    }
  }
};

typedef Fp30<Params30R4> Fp30R4;
typedef Fp30<Params30Q4> Fp30Q4;
