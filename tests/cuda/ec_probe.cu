// ec_probe.cu -- TEST INFRASTRUCTURE: runs the device field / curve operations of pcd_b200/csrc
// on the GPU, one thread, with the same op codes as tests/hostemu/hostemu.cpp, so that the exact
// instruction sequences can be compared with the oracle on real hardware.  Built into
// tests/cuda/_build/libecprobe.so by __graft_entry__.build(); not part of libpcdgpu.so.
#include <cuda_runtime.h>

#include <cstring>

#include "ec.cuh"

template <class C>
__global__ void ec_probe_kernel(int op, const u32* p, const u32* q, const u32* k, int klimbs, u32* out) {
  typedef typename C::F F;
  if (threadIdx.x != 0) return;
  AffinePoint<F> P, Q;
  for (int i = 0; i < (int)(sizeof(P) / 4); i++) {
    ((u32*)&P)[i] = p[i];
    ((u32*)&Q)[i] = q[i];
  }
  XYZZ<C> r;
  switch (op) {
    case 0: r = XYZZ<C>::from_affine(P); r.madd(Q); break;
    case 1: r = XYZZ<C>::from_affine(P).dbl(); r.madd(Q); break;
    case 2: { r = XYZZ<C>::from_affine(P).dbl(); XYZZ<C> s = XYZZ<C>::from_affine(Q).dbl(); r.add(s); break; }
    case 3: r = XYZZ<C>::mul(XYZZ<C>::from_affine(P), k, klimbs); break;
    case 4: r = XYZZ<C>::dbl_affine(P); break;
    case 5: { r = XYZZ<C>::from_affine(P).dbl(); r = r.dbl(); break; }
    case 6: { r = XYZZ<C>::from_affine(P).dbl(); XYZZ<C> s = XYZZ<C>::dbl_affine(P); r.add(s.neg()); break; }
    case 7: { r = XYZZ<C>::from_affine(P).dbl(); XYZZ<C> s = XYZZ<C>::dbl_affine(P); r.add(s); break; }
    case 8: { r = XYZZ<C>::inf(); for (int d = 0; d < klimbs; d++) r.madd(P); break; }
    case 9: { r = XYZZ<C>::inf(); for (int d = 0; d < klimbs; d++) { r.madd(P); AffinePoint<F> t = r.to_affine(); if (t.x.is_zero()) r.madd(P); } break; }
    case 10: { r = XYZZ<C>::from_affine(P).dbl(); r.add(XYZZ<C>::from_affine(Q)); break; }
    // 11-13: the inline bodies, G1 only (for the G2 curves every inlined copy of the group law costs minutes of ptxas)
    case 11: if constexpr (!C::OUTLINE) { r = XYZZ<C>::from_affine(P).dbl_impl(); r.madd(Q); } else r = XYZZ<C>::inf(); break;
    case 12: if constexpr (!C::OUTLINE) { r = XYZZ<C>::from_affine(P).dbl(); r.madd_impl(Q); } else r = XYZZ<C>::inf(); break;
    case 13: if constexpr (!C::OUTLINE) { r = XYZZ<C>::from_affine(P).dbl_impl(); r.madd_impl(Q); } else r = XYZZ<C>::inf(); break;
    case 14: { r = XYZZ<C>::from_affine(P).dbl(); AffinePoint<F> t = r.to_affine(); r = XYZZ<C>::from_affine(t); r.madd(Q); break; }
    default: r = XYZZ<C>::inf();
  }
  AffinePoint<F> a = r.to_affine();
  for (int i = 0; i < (int)(sizeof(a) / 4); i++) out[i] = ((u32*)&a)[i];
}

// Compiled once per curve (-DPROBE_CURVE=0..3, in parallel: the G2 curves take minutes of ptxas each) plus once as the
// dispatcher (-DPROBE_MAIN); __graft_entry__.build() links the five objects into libecprobe.so.
template <class C>
static int run_probe(size_t pb, int op, const void* p, const void* q, const void* k, int klimbs, void* out) {
  u32* d;
  if (cudaMalloc(&d, 3 * pb + 64) != cudaSuccess) return -1;
  u32 *dp = d, *dq = d + pb / 4, *dout = d + 2 * pb / 4, *dk = d + 3 * pb / 4;
  cudaMemcpy(dp, p, pb, cudaMemcpyHostToDevice);
  cudaMemcpy(dq, q, pb, cudaMemcpyHostToDevice);
  cudaMemcpy(dk, k, 40, cudaMemcpyHostToDevice);
  ec_probe_kernel<C><<<1, 32>>>(op, dp, dq, dk, klimbs, dout);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(out, dout, pb, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return e == cudaSuccess ? 0 : -(int)e;
}

#if defined(PROBE_MAIN)
extern "C" int probe_ec_op_0(int, const void*, const void*, const void*, int, void*);
extern "C" int probe_ec_op_1(int, const void*, const void*, const void*, int, void*);
extern "C" int probe_ec_op_2(int, const void*, const void*, const void*, int, void*);
extern "C" int probe_ec_op_3(int, const void*, const void*, const void*, int, void*);
extern "C" int probe_ec_op(int curve, int op, const void* p, const void* q, const void* k, int klimbs, void* out) {
  switch (curve) {
    case 0: return probe_ec_op_0(op, p, q, k, klimbs, out);
    case 1: return probe_ec_op_1(op, p, q, k, klimbs, out);
    case 2: return probe_ec_op_2(op, p, q, k, klimbs, out);
    default: return probe_ec_op_3(op, p, q, k, klimbs, out);
  }
}
#elif PROBE_CURVE == 0
extern "C" int probe_ec_op_0(int op, const void* p, const void* q, const void* k, int kl, void* out) { return run_probe<CurveMnt4G1>(80, op, p, q, k, kl, out); }
#elif PROBE_CURVE == 1
extern "C" int probe_ec_op_1(int op, const void* p, const void* q, const void* k, int kl, void* out) { return run_probe<CurveMnt4G2>(160, op, p, q, k, kl, out); }
#elif PROBE_CURVE == 2
extern "C" int probe_ec_op_2(int op, const void* p, const void* q, const void* k, int kl, void* out) { return run_probe<CurveMnt6G1>(80, op, p, q, k, kl, out); }
#else
extern "C" int probe_ec_op_3(int op, const void* p, const void* q, const void* k, int kl, void* out) { return run_probe<CurveMnt6G2>(240, op, p, q, k, kl, out); }
#endif
