// rings.cuh -- device-side coefficient rings of the generic pass engine.
//
//   ZqRing   Z_q, 2 <= q < 2^32, canonical residues in a u32 (reference: class Zq, types.h:52-116;
//            the reference keeps (-q,q) and canonicalises at exit, zq.cpp:57-67 -- same residue)
//   I64Ring  wrapping int64 (the reference's hInt_t arithmetic)
//   F64Ring  double
//   C64Ring  complex double, operation order of class Complex (types.h:122-164)
#pragma once
#include <cstdint>

#include "lolb_internal.cuh"

namespace lolb {

struct ZqRing {
  typedef uint32_t T;    // in-kernel representation
  typedef int64_t IO;    // ABI representation
  uint32_t q;
  uint64_t mu;           // floor(2^64 / q)

  __device__ __forceinline__ static ZqRing make(const ZqConsts& c, int limb) { return ZqRing{c.q[limb], c.mu[limb]}; }
  __device__ __forceinline__ T zero() const { return 0u; }
  __device__ __forceinline__ T add(T a, T b) const
  {
    uint32_t s = a + b;                       // may wrap when q > 2^31
    return (s < a || s >= q) ? s - q : s;
  }
  __device__ __forceinline__ T sub(T a, T b) const { return a >= b ? a - b : a + (q - b); }
  // Barrett on the full 64-bit product: valid for every q < 2^32
  __device__ __forceinline__ T mul(T a, T b) const
  {
    uint64_t x = (uint64_t)a * b;
    uint64_t r = x - __umul64hi(x, mu) * q;   // in [0, 2q)
    if (r >= q) r -= q;
    if (r >= q) r -= q;
    return (T)r;
  }
  // any 64-bit x -> x mod q (same Barrett step as mul)
  __device__ __forceinline__ T reduce64(uint64_t x) const
  {
    uint64_t r = x - __umul64hi(x, mu) * q;   // in [0, 3q)
    if (r >= q) r -= q;
    if (r >= q) r -= q;
    return (T)r;
  }
  __device__ __forceinline__ T from_int(int i) const { return (uint32_t)i % q; }
  __device__ __forceinline__ T load(IO x) const
  {
    if ((uint64_t)x < (uint64_t)q) return (T)x;        // canonical input: the contract (Backend.hs)
    int64_t r = x % (int64_t)q;                        // tolerate anything else like `c % q` (types.h:62-66)
    return (T)(r < 0 ? r + (int64_t)q : r);
  }
  __device__ __forceinline__ IO store(T v) const { return (IO)v; }
};

struct I64Ring {
  typedef int64_t T;
  typedef int64_t IO;
  __device__ __forceinline__ T zero() const { return 0; }
  __device__ __forceinline__ T add(T a, T b) const { return (T)((uint64_t)a + (uint64_t)b); }
  __device__ __forceinline__ T sub(T a, T b) const { return (T)((uint64_t)a - (uint64_t)b); }
  __device__ __forceinline__ T mul(T a, T b) const { return (T)((uint64_t)a * (uint64_t)b); }
  __device__ __forceinline__ T from_int(int i) const { return (T)i; }
  __device__ __forceinline__ T load(IO x) const { return x; }
  __device__ __forceinline__ IO store(T v) const { return v; }
};

struct F64Ring {
  typedef double T;
  typedef double IO;
  __device__ __forceinline__ T zero() const { return 0.0; }
  // no FMA contraction: keep the reference's rounding points
  __device__ __forceinline__ T add(T a, T b) const { return __dadd_rn(a, b); }
  __device__ __forceinline__ T sub(T a, T b) const { return __dsub_rn(a, b); }
  __device__ __forceinline__ T mul(T a, T b) const { return __dmul_rn(a, b); }
  __device__ __forceinline__ T from_int(int i) const { return (double)i; }
  __device__ __forceinline__ T load(IO x) const { return x; }
  __device__ __forceinline__ IO store(T v) const { return v; }
};

struct C64Ring {
  typedef double2 T;
  typedef double2 IO;
  __device__ __forceinline__ T zero() const { return make_double2(0.0, 0.0); }
  __device__ __forceinline__ T add(T a, T b) const { return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y)); }
  __device__ __forceinline__ T sub(T a, T b) const { return make_double2(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y)); }
  // types.h:144-150: real = a.re*b.re - a.im*b.im ; imag = a.re*b.im + a.im*b.re
  __device__ __forceinline__ T mul(T a, T b) const
  {
    return make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)),
                        __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
  }
  __device__ __forceinline__ T from_int(int i) const { return make_double2((double)i, 0.0); }
  __device__ __forceinline__ T load(IO x) const { return x; }
  __device__ __forceinline__ IO store(T v) const { return v; }
};

}  // namespace lolb
