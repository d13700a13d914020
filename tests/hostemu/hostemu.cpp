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
    case 12: r = x.inverse_fermat(); break;
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
