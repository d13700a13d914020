// CPU build of the DEVICE arithmetic headers with emulated PTX carry primitives.
// TEST INFRASTRUCTURE ONLY: compiled by tests/ into tests/hostemu/_build/libhostemu.so and driven
// through ctypes so the exact instruction sequences of fp.cuh / fpx.cuh / ec.cuh can be compared
// with the oracle on a machine without a GPU.  Not linked into libpcdgpu.so.
#define PCDGPU_HOSTEMU 1
#include "fp.cuh"
#include <cstring>

template <class F> static F ld(const u32* p) { F r; memcpy(r.l, p, 40); return r; }
template <class F> static void st(u32* p, const F& a) { memcpy(p, a.l, 40); }

template <class F>
static void fp_op(int op, const u32* a, const u32* b, u32* out) {
  F x = ld<F>(a), y = ld<F>(b), r;
  switch (op) {
    case 0: r = x + y; break;
    case 1: r = x - y; break;
    case 2: r = x * y; break;
    case 3: r = x.sqr(); break;
    case 4: r = x.inverse(); break;
    case 5: r = x.neg(); break;
    case 6: r = x.from_mont(); break;
    case 7: r = x.to_mont(); break;
    case 8: r = x.template mul_small<17>(); break;
    case 9: r = x.template mul_small<11>(); break;
    case 10: r = x.dbl(); break;
    case 11: r = F::zero(); r.l[0] = x.lexicographically_largest(); break;
    case 12: r = F::mul_sc(x, y); break;
    default: r = F::zero();
  }
  st(out, r);
}

extern "C" void emu_fp_op(int field, int op, const u32* a, const u32* b, u32* out) {
  if (field == 0) fp_op<FpR4>(op, a, b, out); else fp_op<FpQ4>(op, a, b, out);
}

// ---- extension fields and curves ------------------------------------------------------------
#include "ec.cuh"
template <class F> static F ldx(const u32* p) { F r; memcpy((void*)&r, p, sizeof(F)); return r; }
template <class F> static void stx(u32* p, const F& a) { memcpy(p, (const void*)&a, sizeof(F)); }

template <class F>
static void fx_op(int op, const u32* a, const u32* b, u32* out) {
  F x = ldx<F>(a), y = ldx<F>(b), r;
  switch (op) {
    case 0: r = x + y; break;
    case 1: r = x - y; break;
    case 2: r = x * y; break;
    case 3: r = x.sqr(); break;
    case 4: r = x.inverse(); break;
    case 5: r = x.neg(); break;
    case 10: r = x.dbl(); break;
    case 11: r = F::zero(); ((u32*)&r)[0] = x.lexicographically_largest(); break;
    default: r = F::zero();
  }
  stx(out, r);
}
extern "C" void emu_fx_op(int kind, int op, const u32* a, const u32* b, u32* out) {
  if (kind == 2) fx_op<Fq2>(op, a, b, out); else fx_op<Fq3>(op, a, b, out);
}

template <class C>
static void ec_op(int op, const u32* p, const u32* q, const u32* k, int klimbs, u32* out) {
  typedef typename C::F F;
  AffinePoint<F> P = ldx<AffinePoint<F>>(p), Q = ldx<AffinePoint<F>>(q);
  XYZZ<C> r;
  switch (op) {
    case 0: r = XYZZ<C>::from_affine(P); r.madd(Q); break;                 // P + Q (zz = 1)
    case 1: r = XYZZ<C>::from_affine(P).dbl(); r.madd(Q); break;           // 2P + Q
    case 2: { r = XYZZ<C>::from_affine(P).dbl(); XYZZ<C> s = XYZZ<C>::from_affine(Q).dbl(); r.add(s); break; }  // 2P + 2Q
    case 3: r = XYZZ<C>::mul(XYZZ<C>::from_affine(P), k, klimbs); break;   // [k]P
    case 4: r = XYZZ<C>::dbl_affine(P); break;                              // 2P
    case 5: { r = XYZZ<C>::from_affine(P).dbl(); r = r.dbl(); break; }      // 4P
    case 6: { r = XYZZ<C>::from_affine(P).dbl(); XYZZ<C> s = XYZZ<C>::dbl_affine(P); r.add(s.neg()); break; }  // 2P - 2P
    case 7: { r = XYZZ<C>::from_affine(P).dbl(); XYZZ<C> s = XYZZ<C>::dbl_affine(P); r.add(s); break; }  // 2P + 2P via add
    case 8: { r = XYZZ<C>::inf(); for (int d = 0; d < klimbs; d++) { r.madd(P); (void)r.to_affine(); } break; }  // repeated madd of the same affine point (fixed-base table construction)
    default: r = XYZZ<C>::inf();
  }
  stx(out, r.to_affine());
}
extern "C" void emu_ec_op(int curve, int op, const u32* p, const u32* q, const u32* k, int klimbs, u32* out) {
  switch (curve) {
    case 0: ec_op<CurveMnt4G1>(op, p, q, k, klimbs, out); break;
    case 1: ec_op<CurveMnt4G2>(op, p, q, k, klimbs, out); break;
    case 2: ec_op<CurveMnt6G1>(op, p, q, k, klimbs, out); break;
    default: ec_op<CurveMnt6G2>(op, p, q, k, klimbs, out); break;
  }
}

// ---- radix-2^30 field and its G1 twins (fp30.cuh, ec.cuh: fast_affine / slow_xyzz) ---------------------
template <class P, class P30>
static void fp30_op(int op, const u32* a, const u32* b, u32* out) {
  Fp<P> x = ld<Fp<P>>(a), y = ld<Fp<P>>(b);
  Fp30<P30> fx = to_fast<P30>(x), fy = to_fast<P30>(y), r;
  switch (op) {
    case 0: r = fx + fy; break;
    case 1: r = fx - fy; break;
    case 2: r = fx * fy; break;
    case 3: r = fx.sqr(); break;
    case 5: r = fx.neg(); break;
    case 10: r = fx.dbl(); break;
    case 13: r = Fp30<P30>::zero(); r.l[0] = (fx == fy) ? 1 : 0; break;  // equality across representatives
    default: r = fx;
  }
  if (op == 13) { memcpy(out, r.l, 40); return; }
  st(out, from_fast<P>(r));
}
extern "C" void emu_fp30_op(int field, int op, const u32* a, const u32* b, u32* out) {
  if (field == 0) fp30_op<ParamsR4, Params30R4>(op, a, b, out); else fp30_op<ParamsQ4, Params30Q4>(op, a, b, out);
}
// buckets as the accumulate kernels build them: inf, then madd of the n given affine points (sign bit in
// signs[i] negates y), in the radix-2^30 twin; returned through slow_xyzz as an affine ABI point
template <class C>
static void ec_fast_acc(const u32* pts, const int* signs, int n, u32* out) {
  typedef typename C::F F;
  XYZZ<typename C::Fast> acc = XYZZ<typename C::Fast>::inf();
  for (int i = 0; i < n; i++) {
    AffinePoint<typename C::Fast::F> p = fast_affine<C>(ldx<AffinePoint<F>>(pts + (size_t)i * (sizeof(AffinePoint<F>) / 4)));
    if (signs[i]) p.y = p.y.neg();
    acc.madd(p);
  }
  stx(out, slow_xyzz<C>(acc).to_affine());
}
extern "C" void emu_ec_fast_acc(int curve, const u32* pts, const int* signs, int n, u32* out) {
  if (curve == 0) ec_fast_acc<CurveMnt4G1>(pts, signs, n, out); else ec_fast_acc<CurveMnt6G1>(pts, signs, n, out);
}
