// prims.cuh -- 32-bit multiply-add / carry-chain primitives for sm_100a.
//
// Every primitive is ONE PTX instruction.  ptxas fuses mad.lo.cc/madc.hi.cc pairs into
// IMAD.WIDE.U32(.X) (checked with cuobjdump -sass on sm_100a: a 10-limb Montgomery product is
// ~212 fma-pipe instructions + ~45 IADD3).  The asm statements are volatile so that the NVVM
// front end keeps the carry-flag order it cannot see.
//
// Host build: these primitives have NO host implementation in the product.  The CPU unit tests
// (tests/hostemu) define PCDGPU_HOSTEMU and provide a bit-exact emulation of the PTX carry flag so
// that the *same* field / curve / index code can be checked against the oracle without a GPU.
#pragma once
#include <cstdint>

typedef uint32_t u32;
typedef uint64_t u64;

#if defined(__CUDACC__)
#define PCD_HD __host__ __device__ __forceinline__
#define PCD_D __device__ __forceinline__
#define PCD_NOINLINE __host__ __device__ __noinline__
#else
#define PCD_NOINLINE
#define PCD_HD inline
#define PCD_D inline
#endif

#if defined(__CUDA_ARCH__)

namespace prims {
PCD_D u32 mul_lo(u32 a, u32 b) { u32 r; asm volatile("mul.lo.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PCD_D u32 mul_hi(u32 a, u32 b) { u32 r; asm volatile("mul.hi.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PCD_D u32 mad_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.lo.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PCD_D u32 mad_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.hi.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PCD_D u32 madc_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PCD_D u32 madc_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.cc.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PCD_D u32 madc_lo(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PCD_D u32 madc_hi(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.u32 %0,%1,%2,%3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PCD_D u32 add_cc(u32 a, u32 b) { u32 r; asm volatile("add.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PCD_D u32 addc_cc(u32 a, u32 b) { u32 r; asm volatile("addc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PCD_D u32 addc(u32 a, u32 b) { u32 r; asm volatile("addc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PCD_D u32 sub_cc(u32 a, u32 b) { u32 r; asm volatile("sub.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PCD_D u32 subc_cc(u32 a, u32 b) { u32 r; asm volatile("subc.cc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PCD_D u32 subc(u32 a, u32 b) { u32 r; asm volatile("subc.u32 %0,%1,%2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// (t1:t0) += a * b, the carry out of the 64-bit addition counted into c2.  One asm statement, carry flag
// internal and not volatile: ptxas emits IMAD.WIDE.U32 (no .X, full rate) + IADD3.X and may schedule
// independent instances freely.
PCD_D void mac_carry(u32& t0, u32& t1, u32& c2, u32 a, u32 b) {
  asm("mad.lo.cc.u32 %0,%3,%4,%0;\n\tmadc.hi.cc.u32 %1,%3,%4,%1;\n\taddc.u32 %2,%2,0;"
      : "+r"(t0), "+r"(t1), "+r"(c2)
      : "r"(a), "r"(b));
}
// (t1:t0) += a * b without a carry out (the caller knows it cannot overflow)
PCD_D void mac(u32& t0, u32& t1, u32 a, u32 b) {
  asm("mad.lo.cc.u32 %0,%2,%3,%0;\n\tmadc.hi.u32 %1,%2,%3,%1;" : "+r"(t0), "+r"(t1) : "r"(a), "r"(b));
}
// (t1:t0) = a * b
PCD_D void mul_wide(u32& t0, u32& t1, u32 a, u32 b) {
  asm("mul.lo.u32 %0,%2,%3;\n\tmul.hi.u32 %1,%2,%3;" : "=&r"(t0), "=r"(t1) : "r"(a), "r"(b));
}
// (t1:t0) += (x1:x0), the carry out counted into c2
PCD_D void add64_carry(u32& t0, u32& t1, u32& c2, u32 x0, u32 x1) {
  asm("add.cc.u32 %0,%0,%3;\n\taddc.cc.u32 %1,%1,%4;\n\taddc.u32 %2,%2,0;"
      : "+r"(t0), "+r"(t1), "+r"(c2)
      : "r"(x0), "r"(x1));
}
// (t1:t0) += (x1:x0) >> 30  (radix-2^30 column carry; the sum is known not to overflow 64 bits)
PCD_D void add_shr30(u32& t0, u32& t1, u32 x0, u32 x1) {
  const u32 s0 = __funnelshift_r(x0, x1, 30), s1 = x1 >> 30;
  asm("add.cc.u32 %0,%0,%2;\n\taddc.u32 %1,%1,%3;" : "+r"(t0), "+r"(t1) : "r"(s0), "r"(s1));
}
}  // namespace prims

#elif defined(PCDGPU_HOSTEMU)
#include "hostemu_prims.h"  // tests/hostemu -- CPU unit tests only
#else
// Host pass of nvcc over a .cu file: declarations only, so that __host__ __device__ templates
// parse; calling one on the host is a link error by design (no CPU fallback).
namespace prims {
u32 mul_lo(u32, u32); u32 mul_hi(u32, u32);
u32 mad_lo_cc(u32, u32, u32); u32 mad_hi_cc(u32, u32, u32);
u32 madc_lo_cc(u32, u32, u32); u32 madc_hi_cc(u32, u32, u32);
u32 madc_lo(u32, u32, u32); u32 madc_hi(u32, u32, u32);
u32 add_cc(u32, u32); u32 addc_cc(u32, u32); u32 addc(u32, u32);
u32 sub_cc(u32, u32); u32 subc_cc(u32, u32); u32 subc(u32, u32);
void mac_carry(u32&, u32&, u32&, u32, u32); void mac(u32&, u32&, u32, u32);
void mul_wide(u32&, u32&, u32, u32); void add64_carry(u32&, u32&, u32&, u32, u32);
void add_shr30(u32&, u32&, u32, u32);
}  // namespace prims
#endif
