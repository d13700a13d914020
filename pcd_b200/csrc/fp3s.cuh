// fp3s.cuh -- Fq3 = F_r4[u]/(u^3 - 5) with ONE COEFFICIENT PER LANE: lanes 3g, 3g+1, 3g+2 of a warp hold c0, c1, c2 of
// the same element and run the same code; products of coefficients from different lanes travel by warp shuffles.
//
// Why: the G2 MSM of the helper SNARK (MNT6-298, G2 over Fq3; ark-ec VariableBaseMSM::multi_scalar_mul reached from
// IC::HelpSNARK::prove, /root/reference/src/ec_cycle_pcd/mod.rs:179) spends its time in bucket accumulation, and a
// mixed addition over Fq3 held by one thread needs ~25 live field elements (255 registers, 844 B of spills, every
// product an out-of-line call: round 1 measured 25-30 % of the integer-multiply roof).  Sliced, a lane keeps a third
// of the state (~130 registers, no spills, everything inline), every extension-field product is exactly two base
// products per lane (Karatsuba's six, balanced), and 30 of 32 lanes work.  Same element values, same formulas as
// fpx.cuh's Fp3T, so results are bit-identical.
//
// All lanes of a group must call every operation together (the shuffles name the group's three lanes); groups of
// one warp may diverge from each other.
#pragma once
#include "fp.cuh"

#if defined(__CUDACC__)
template <class B, u32 NR>
struct Fp3S {
  B c;  // this lane's coefficient
  static constexpr int LANES = 3;

  __device__ __forceinline__ static int li() { return (int)((threadIdx.x & 31u) % 3u); }
  __device__ __forceinline__ static unsigned gmask() {
    const unsigned lane = threadIdx.x & 31u;
    return 7u << (lane - lane % 3u);
  }
  // the value of v on the lane `src` (absolute lane id inside the warp) of this group
  __device__ __forceinline__ static B shfl(const B& v, int src) {
    B r;
    const unsigned m = gmask();
#pragma unroll
    for (int i = 0; i < FP_LIMBS; i++) r.l[i] = __shfl_sync(m, v.l[i], src);
    return r;
  }
  __device__ __forceinline__ static B sel(bool p, const B& a, const B& b) {
    B r;
#pragma unroll
    for (int i = 0; i < FP_LIMBS; i++) r.l[i] = p ? a.l[i] : b.l[i];
    return r;
  }
  __device__ __forceinline__ static Fp3S zero() { Fp3S r; r.c = B::zero(); return r; }
  __device__ __forceinline__ static Fp3S one() { Fp3S r; r.c = li() == 0 ? B::one() : B::zero(); return r; }
  __device__ __forceinline__ bool is_zero() const { return __all_sync(gmask(), c.is_zero()); }
  __device__ __forceinline__ bool operator==(const Fp3S& o) const { return __all_sync(gmask(), c == o.c); }
  __device__ __forceinline__ bool operator!=(const Fp3S& o) const { return !(*this == o); }
  __device__ __forceinline__ friend Fp3S operator+(const Fp3S& a, const Fp3S& b) { Fp3S r; r.c = a.c + b.c; return r; }
  __device__ __forceinline__ friend Fp3S operator-(const Fp3S& a, const Fp3S& b) { Fp3S r; r.c = a.c - b.c; return r; }
  __device__ __forceinline__ Fp3S neg() const { Fp3S r; r.c = c.neg(); return r; }
  __device__ __forceinline__ Fp3S dbl() const { Fp3S r; r.c = c.dbl(); return r; }
  template <u32 K>
  __device__ __forceinline__ Fp3S mul_small() const { Fp3S r; r.c = c.template mul_small<K>(); return r; }

  // Karatsuba (fpx.cuh Fp3T::operator*), two base products per lane.  With i this lane's coefficient index,
  // j = i + 1, k = i + 2 (mod 3):  v_i = a_i b_i,  u_i = (a_j + a_k)(b_j + b_k) - v_j - v_k = a_j b_k + a_k b_j, and
  //   r_0 = v_0 + NR u_0,   r_1 = u_2 + NR v_2,   r_2 = u_1 + v_1.
  __device__ __forceinline__ friend Fp3S operator*(const Fp3S& a, const Fp3S& b) {
    const int l = li();
    const int base = (int)(threadIdx.x & 31u) - l;
    const int j = base + (l == 2 ? 0 : l + 1), k = base + (l == 0 ? 2 : l - 1);
    const B aj = shfl(a.c, j), ak = shfl(a.c, k), bj = shfl(b.c, j), bk = shfl(b.c, k);
    const B v = a.c * b.c;
    const B w = (aj + ak) * (bj + bk);
    const B vj = shfl(v, j), vk = shfl(v, k);
    const B u = w - vj - vk;
    const B uo = shfl(u, l == 1 ? j : k);  // lane 1 needs u_2 (its j), lane 2 needs u_1 (its k); lane 0 ignores it
    const B X = sel(l == 0, v, uo);
    const B Y = sel(l == 0, u, sel(l == 1, vj, vk));
    Fp3S r;
    r.c = X + sel(l == 2, Y, Y.template mul_small<NR>());
    return r;
  }
  // r_0 = a_0^2 + 2 NR a_1 a_2,  r_1 = 2 a_0 a_1 + NR a_2^2,  r_2 = a_1^2 + 2 a_0 a_2: six products, two per lane, every
  // result on its own lane (no exchange after the products)
  __device__ __forceinline__ Fp3S sqr() const {
    const int l = li();
    const int base = (int)(threadIdx.x & 31u) - l;
    const int j = base + (l == 2 ? 0 : l + 1), k = base + (l == 0 ? 2 : l - 1);
    const B aj = shfl(c, j), ak = shfl(c, k);
    // lane 0: p = a_0^2, q = a_1 a_2;  lane 1: p = a_2^2 (its j), q = a_0 a_1 (its k, own);  lane 2: p = a_1^2 (its k), q = a_2 a_0
    const B ps = sel(l == 0, c, sel(l == 1, aj, ak));
    const B p = ps * ps;
    const B qa = sel(l == 0, aj, c);
    const B qb = sel(l == 2, aj, ak);
    const B q2 = (qa * qb).dbl();
    Fp3S r;
    r.c = sel(l == 1, p.template mul_small<NR>(), p) + sel(l == 0, q2.template mul_small<NR>(), q2);
    return r;
  }
};
typedef Fp3S<FpR4, 5> Fq3S;
#endif
