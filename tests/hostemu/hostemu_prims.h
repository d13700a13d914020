// CPU emulation of the PTX carry-flag primitives in pcd_b200/csrc/prims.cuh.
// TEST INFRASTRUCTURE ONLY (tests/hostemu): lets the device field/curve/index code be compiled
// with g++ and checked against the oracle without a GPU.  Never part of libpcdgpu.so.
#pragma once
#include <cstdint>
namespace prims {
static thread_local uint32_t CC = 0;
inline u32 mul_lo(u32 a, u32 b) { return (u32)((u64)a * b); }
inline u32 mul_hi(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
inline u32 _acc(u64 t) { CC = (u32)(t >> 32) & 1; return (u32)t; }
inline u32 mad_lo_cc(u32 a, u32 b, u32 c) { return _acc((u64)mul_lo(a, b) + c); }
inline u32 mad_hi_cc(u32 a, u32 b, u32 c) { return _acc((u64)mul_hi(a, b) + c); }
inline u32 madc_lo_cc(u32 a, u32 b, u32 c) { return _acc((u64)mul_lo(a, b) + c + CC); }
inline u32 madc_hi_cc(u32 a, u32 b, u32 c) { return _acc((u64)mul_hi(a, b) + c + CC); }
inline u32 madc_lo(u32 a, u32 b, u32 c) { return (u32)((u64)mul_lo(a, b) + c + CC); }
inline u32 madc_hi(u32 a, u32 b, u32 c) { return (u32)((u64)mul_hi(a, b) + c + CC); }
inline u32 add_cc(u32 a, u32 b) { return _acc((u64)a + b); }
inline u32 addc_cc(u32 a, u32 b) { return _acc((u64)a + b + CC); }
inline u32 addc(u32 a, u32 b) { return (u32)((u64)a + b + CC); }
inline u32 _bor(u64 t) { CC = (u32)(t >> 63) & 1; return (u32)t; }  // CC = borrow
inline u32 sub_cc(u32 a, u32 b) { return _bor((u64)a - b); }
inline u32 subc_cc(u32 a, u32 b) { return _bor((u64)a - b - CC); }
inline u32 subc(u32 a, u32 b) { return (u32)((u64)a - b - CC); }
inline void mac_carry(u32& t0, u32& t1, u32& c2, u32 a, u32 b) {
  u64 p = (u64)a * b;
  u64 lo = (u64)t0 + (u32)p;
  u64 hi = (u64)t1 + (u32)(p >> 32) + (lo >> 32);
  t0 = (u32)lo;
  t1 = (u32)hi;
  c2 += (u32)(hi >> 32);
}
inline void mac(u32& t0, u32& t1, u32 a, u32 b) {
  u64 v = (((u64)t1 << 32) | t0) + (u64)a * b;
  t0 = (u32)v;
  t1 = (u32)(v >> 32);
}
inline void mul_wide(u32& t0, u32& t1, u32 a, u32 b) {
  u64 v = (u64)a * b;
  t0 = (u32)v;
  t1 = (u32)(v >> 32);
}
inline void add64_carry(u32& t0, u32& t1, u32& c2, u32 x0, u32 x1) {
  u64 lo = (u64)t0 + x0;
  u64 hi = (u64)t1 + x1 + (lo >> 32);
  t0 = (u32)lo;
  t1 = (u32)hi;
  c2 += (u32)(hi >> 32);
}
inline void add_shr30(u32& t0, u32& t1, u32 x0, u32 x1) {
  u64 v = (((u64)t1 << 32) | t0) + ((((u64)x1 << 32) | x0) >> 30);
  t0 = (u32)v;
  t1 = (u32)(v >> 32);
}
}  // namespace prims
