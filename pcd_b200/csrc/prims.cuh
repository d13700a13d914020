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
}  // namespace prims
#endif
