// vecio.cuh -- 16-byte vectorised loads/stores of points (affine: 80/160/240 B, xyzz: 160/320/480 B).
#pragma once
#include "ec.cuh"

// ---- vectorised point I/O -------------------------------------------------------------------
template <class T>
__device__ __forceinline__ T ld_vec(const void* base, size_t idx) {
  static_assert(sizeof(T) % 16 == 0, "16-byte multiples");
  T r;
  const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(base) + idx * sizeof(T));
  u32* q = reinterpret_cast<u32*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) {
    uint4 v = __ldg(p + i);
    q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
  }
  return r;
}
template <class T>
__device__ __forceinline__ T ld_vec_rw(const void* base, size_t idx) {
  T r;
  const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(base) + idx * sizeof(T));
  u32* q = reinterpret_cast<u32*>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) {
    uint4 v = p[i];
    q[4 * i] = v.x; q[4 * i + 1] = v.y; q[4 * i + 2] = v.z; q[4 * i + 3] = v.w;
  }
  return r;
}
template <class T>
__device__ __forceinline__ void st_vec(void* base, size_t idx, const T& v) {
  uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<char*>(base) + idx * sizeof(T));
  const u32* q = reinterpret_cast<const u32*>(&v);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 16); i++) p[i] = make_uint4(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
}


template <class C>
__device__ __forceinline__ AffinePoint<typename C::F> ld_aff(const void* base, size_t idx) {
  return ld_vec_rw<AffinePoint<typename C::F>>(base, idx);
}
template <class C>
__device__ __forceinline__ void st_aff(void* base, size_t idx, const AffinePoint<typename C::F>& p) {
  st_vec(base, idx, p);
}
template <class C>
__device__ __forceinline__ XYZZ<C> ld_xyzz(const void* base, size_t idx) {
  return ld_vec_rw<XYZZ<C>>(base, idx);
}
template <class C>
__device__ __forceinline__ void st_xyzz(void* base, size_t idx, const XYZZ<C>& p) {
  st_vec(base, idx, p);
}
