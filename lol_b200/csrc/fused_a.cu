// fused_a.cu -- fused Z_q CRT / CRT^-1 for m = 2^6 * 3^2 * 5^2 = 14400 (n = 3840 = 32 * 6 * 20), the index of
// BASELINE.json's headline metric: one HBM read and one HBM write per ring element, everything else on chip.
//
// The operator is the reference's (crt.cpp:518-581 on tensor.h:76-95): CRT_m = CRT_64 (x) CRT_9 (x) CRT_25 with the
// first factor on the fastest axis, CRT_{p^e} = (DFT_{p^(e-1)} (x) I_{p-1}) . That . (I_{p^(e-1)} (x) CRT_p).
// Factors on different axes commute, and over Z_q any exact evaluation order gives the same residues, so the
// schedule is chosen for the machine, not copied from the reference:
//
//   element X[i3][i2][i1], i1 < 32 (2^6 axis), i2 < 6 (3^2 axis), i3 < 20 (5^2 axis), flat j = i1 + 32*i2 + 192*i3
//
//   phase 1  thread (i1,i2) loads its 20 coefficients along i3 straight from HBM (each warp instruction reads
//            256 contiguous bytes), applies CRT_25 in registers, writes u32 to shared memory
//   phase 2  warp <-> i3: lane i1 reads 6 coefficients along i2 (conflict-free), applies CRT_9 in registers, then
//            CRT_64 across the 32 lanes with a register-exchange butterfly network (one shuffle per two
//            coefficients per radix-2 round), and stores int64 to HBM (128-byte contiguous runs)
//
// Diagonal twiddles (crtTwiddle) and the final mhat^-1 are folded into the small dense matrices on the host.
// Arithmetic (q*q*8 + 2q < 2^32, e.g. q = 14401): residues stay lazily in [0, 2q) as u32; a row of a dense
// stage is a plain 32-bit multiply-accumulate followed by ONE Barrett reduction; canonical [0,q) only at the store.
#include <cstdlib>

#include "fused.cuh"
#include "numtheory.h"

namespace lolb {

namespace {

struct FusedAConsts {
  uint32_t q, q2;
  uint32_t r0;              // ArithS: mu = floor(2^32 / q);  ArithM: -q^-1 mod 2^32
  uint32_t one;             // ArithS: 1;                     ArithM: 2^32 mod q (Montgomery form of 1)
  uint32_t r2;              // ArithS: unused;                ArithM: 2^64 mod q (data x data products, see mulv)
  uint32_t m5[5][4][4];     // fwd: (twiddle . CRT_5) per block i0;      inv: (CRT_5^-1' . twiddle) * mhat^-1
  uint32_t d5[5][5];        // DFT_5 over the block index (fwd or inverse roots)
  uint32_t m3[3][2][2];
  uint32_t d3[3][3];
  const uint32_t* lane_tw;  // device [kLaneRows][32] per-lane constants: 2^6-axis twiddles and folded m3 (see build_consts)
};

// Two arithmetic policies with one interface.  Residues are lazy u32 in [0,2q); `Acc` accumulates a row of a dense
// stage without intermediate reduction; red() brings an accumulator back to [0,2q).
//
// ArithS (8q^2 + 2q < 2^32, e.g. q = 14401): 32-bit multiply-accumulate, Barrett reduction, constants as plain residues.
struct ArithS {
  typedef uint32_t Acc;
  uint32_t q, q2, mu, nq;    // nq = 2^32 - q: "x - t*q" as one multiply-add without a negation
  __device__ __forceinline__ ArithS(const FusedAConsts& C) : q(C.q), q2(C.q2), mu(C.r0), nq(0u - C.q) {}
  __device__ __forceinline__ Acc mul(uint32_t c, uint32_t v) const { return c * v; }
  __device__ __forceinline__ Acc mad(Acc a, uint32_t c, uint32_t v) const { return a + c * v; }
  __device__ __forceinline__ Acc unit(uint32_t v) const { return v; }                                   // the "1 * v" term
  __device__ __forceinline__ uint32_t red(Acc x) const { return __umulhi(x, mu) * nq + x; }             // any x -> [0,2q)
  __device__ __forceinline__ uint32_t mulv(uint32_t x, uint32_t b) const { return red(x * b); }         // data x data, x < 2q, b < q
  __device__ __forceinline__ uint32_t fold(uint32_t x) const { return min(x, x - q2); }                 // [0,4q) -> [0,2q)
  __device__ __forceinline__ uint32_t canon(uint32_t x) const { return min(x, x - q); }                 // [0,2q) -> [0,q)
};

// ArithM (10q < 2^32, e.g. the 20-bit SymmSHE moduli 1008001, 1065601): 64-bit multiply-accumulate, ONE Montgomery
// reduction per row; constants are stored in Montgomery form c * 2^32 mod q, data stays in plain form, so
// REDC(sum c~_i v_i) = sum c_i v_i mod q.  acc < 10 q^2 and 10q < 2^32 give REDC(acc) < 2q.
struct ArithM {
  typedef uint64_t Acc;
  uint32_t q, q2, qinv, one, r2;
  __device__ __forceinline__ ArithM(const FusedAConsts& C) : q(C.q), q2(C.q2), qinv(C.r0), one(C.one), r2(C.r2) {}
  __device__ __forceinline__ Acc mul(uint32_t c, uint32_t v) const { return (uint64_t)c * v; }
  __device__ __forceinline__ Acc mad(Acc a, uint32_t c, uint32_t v) const { return a + (uint64_t)c * v; }
  __device__ __forceinline__ Acc unit(uint32_t v) const { return (uint64_t)one * v; }
  __device__ __forceinline__ uint32_t red(Acc x) const
  {
    const uint32_t m = (uint32_t)x * qinv;
    return (uint32_t)((x + (uint64_t)m * q) >> 32);
  }
  // data x data: both factors are in plain form, so REDC(x b) = x b 2^-32; a second REDC against 2^64 mod q restores it
  __device__ __forceinline__ uint32_t mulv(uint32_t x, uint32_t b) const { return red((uint64_t)r2 * red((uint64_t)x * b)); }
  __device__ __forceinline__ uint32_t fold(uint32_t x) const { return min(x, x - q2); }
  __device__ __forceinline__ uint32_t canon(uint32_t x) const { return min(x, x - q); }
};

// CTA shape of the Montgomery class (tuning: tools/build_variant.py).  Two warps balance both phases exactly (3 and 10 tasks
// per warp).  Measured on B200, CRT / CRT^-1 % of the HBM roofline, warps x CTAs per SM: one limb (q = 1008001) 2x12 66.2 / 61.6,
// 2x10 66.6 / 60.9, 3x8 63.8 / 58.8, 4x6 61.4 / 57.5, 5x4 56.9 / 52.6; tupSize 2 (config C) 2x6 62.5 / 55.7, 3x4 62.3 / 54.7,
// 4x3 58.2 / 53.0, 2x5 54.2 / 48.4, 5x2 46.1 / 40.4.  (The 32-bit class keeps 3x8: 2x14 measured 70 / 67 against 88 / 86.)
#ifndef LOLB_A_M_W
#define LOLB_A_M_W 2
#define LOLB_A_M_MB 12
#endif
#ifndef LOLB_A_K2_W
#define LOLB_A_K2_W 2
#define LOLB_A_K2_MB 6
#endif
constexpr int kLaneRows = 20;      // rows 0-7: 2^6-axis twiddles; rows 8-19: forward m3[i0][r][c] * crtTwiddle_64(lane)
constexpr int kN = 3840, kD1 = 32, kD2 = 6, kD3 = 20;

// CRT_25 / CRT_25^-1 on the 20 values of one (i1,i2) column, v[4*i0 + c]
template <bool INV, class AR>
__device__ __forceinline__ void axis5(uint32_t (&v)[20], const FusedAConsts& C, const AR& A)
{
  if (!INV) {
#pragma unroll
    for (int i0 = 0; i0 < 5; i0++) {
      uint32_t o[4];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        typename AR::Acc acc = A.mul(C.m5[i0][r][0], v[4 * i0]);
#pragma unroll
        for (int c = 1; c < 4; c++) acc = A.mad(acc, C.m5[i0][r][c], v[4 * i0 + c]);
        o[r] = A.red(acc);
      }
#pragma unroll
      for (int r = 0; r < 4; r++) v[4 * i0 + r] = o[r];
    }
  }
#pragma unroll
  for (int c = 0; c < 4; c++) {            // DFT_5 across the block index for residue column c
    uint32_t o[5];
    o[0] = A.red(A.unit(v[c] + v[4 + c] + v[8 + c] + v[12 + c] + v[16 + c]));
#pragma unroll
    for (int row = 1; row < 5; row++) {
      typename AR::Acc acc = A.unit(v[c]);
#pragma unroll
      for (int col = 1; col < 5; col++) acc = A.mad(acc, C.d5[row][col], v[4 * col + c]);
      o[row] = A.red(acc);
    }
#pragma unroll
    for (int row = 0; row < 5; row++) v[4 * row + c] = o[row];
  }
  if (INV) {
#pragma unroll
    for (int i0 = 0; i0 < 5; i0++) {
      uint32_t o[4];
#pragma unroll
      for (int r = 0; r < 4; r++) {
        typename AR::Acc acc = A.mul(C.m5[i0][r][0], v[4 * i0]);
#pragma unroll
        for (int c = 1; c < 4; c++) acc = A.mad(acc, C.m5[i0][r][c], v[4 * i0 + c]);
        o[r] = A.red(acc);
      }
#pragma unroll
      for (int r = 0; r < 4; r++) v[4 * i0 + r] = o[r];
    }
  }
}

// CRT_9 / CRT_9^-1 on the 6 values x[2*i0 + c] of one (i3, i1)
// Forward: the 2x2 blocks use the PER-LANE constants m3l = m3 * crtTwiddle_64(column = lane): the diagonal twiddle of
// the 2^6 axis is a scalar per lane, commutes with the 3^2 axis and costs nothing once folded in here.
template <bool INV, class AR>
__device__ __forceinline__ void axis3(uint32_t (&x)[6], const FusedAConsts& C, const AR& A, const uint32_t (&m3l)[12])
{
  if (!INV) {
#pragma unroll
    for (int i0 = 0; i0 < 3; i0++) {
      uint32_t a = x[2 * i0], b = x[2 * i0 + 1];
      x[2 * i0] = A.red(A.mad(A.mul(m3l[4 * i0], a), m3l[4 * i0 + 1], b));
      x[2 * i0 + 1] = A.red(A.mad(A.mul(m3l[4 * i0 + 2], a), m3l[4 * i0 + 3], b));
    }
  }
#pragma unroll
  for (int c = 0; c < 2; c++) {
    uint32_t a = x[c], b = x[2 + c], d = x[4 + c];
    x[c] = A.red(A.unit(a + b + d));
    x[2 + c] = A.red(A.mad(A.mad(A.unit(a), C.d3[1][1], b), C.d3[1][2], d));
    x[4 + c] = A.red(A.mad(A.mad(A.unit(a), C.d3[2][1], b), C.d3[2][2], d));
  }
  if (INV) {
#pragma unroll
    for (int i0 = 0; i0 < 3; i0++) {
      uint32_t a = x[2 * i0], b = x[2 * i0 + 1];
      x[2 * i0] = A.red(A.mad(A.mul(C.m3[i0][0][0], a), C.m3[i0][0][1], b));
      x[2 * i0 + 1] = A.red(A.mad(A.mul(C.m3[i0][1][0], a), C.m3[i0][1][1], b));
    }
  }
}

// One radix-2 round of the exchange network on lane bit `bit`.  Before: the lane owns (c0[j], c1[j]) for three
// rows; after: the two columns of one butterfly.  Forward: (u,t) -> (u+t, (u-t)*tw).  Inverse: (u,t) -> (u+t*tw, u-t*tw).
// TRIVIAL: every twiddle of the round is 1 (bit 4: i0 = column >> 5 = 0, crt.cpp:92-106 skips i0 = 0), so the
// multiplication and its reduction are dropped.
template <bool INV, class AR, bool TRIVIAL = false>
__device__ __forceinline__ void exchange_round(uint32_t (&c0)[3], uint32_t (&c1)[3], int lane, int bit, uint32_t tw, const AR& A)
{
  const bool hi = (lane >> bit) & 1;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const uint32_t send = hi ? c0[j] : c1[j];
    const uint32_t keep = hi ? c1[j] : c0[j];
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1 << bit);
    if (TRIVIAL) {
      const uint32_t u = hi ? recv : keep, t = hi ? keep : recv;
      c0[j] = A.fold(u + t);
      c1[j] = A.fold(u + A.q2 - t);
    } else if (!INV) {
      // u + t is symmetric; (u - t) = +-(keep - recv) and the sign lives in the lane's twiddle (host: q - tw for hi lanes)
      c0[j] = A.fold(keep + recv);
      c1[j] = A.red(A.mul(tw, keep + A.q2 - recv));
    } else {
      const uint32_t t = A.red(A.mul(tw, hi ? keep : recv));
      const uint32_t u = hi ? recv : keep;
      c0[j] = A.fold(u + t);
      c1[j] = A.fold(u + A.q2 - t);
    }
  }
}

// Last inverse round (lane bit 0) merged with the inverse crtTwiddle of the 2^6 axis that follows it: the lane ends up
// with columns 2c (c0) and 2c+1 (c1), whose twiddles a, b are per-lane constants, so
//   c0 = a (u + tw t) = a u + (a tw) t,   c1 = b (u - tw t) = b u + (-b tw) t
// are two 2-term rows with host-folded constants: one reduction each instead of three.
template <class AR>
__device__ __forceinline__ void exchange_last_inv(uint32_t (&c0)[3], uint32_t (&c1)[3], int lane, uint32_t a, uint32_t atw,
                                                  uint32_t b, uint32_t nbtw, const AR& A)
{
  const bool hi = lane & 1;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const uint32_t send = hi ? c0[j] : c1[j];
    const uint32_t keep = hi ? c1[j] : c0[j];
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
    const uint32_t t = hi ? keep : recv, u = hi ? recv : keep;
    c0[j] = A.red(A.mad(A.mul(a, u), atw, t));
    c1[j] = A.red(A.mad(A.mul(b, u), nbtw, t));
  }
}

// slow path for non-canonical input (outside the Haskell contract, tolerated like `c % q`, types.h:62-66)
__device__ __noinline__ uint32_t reduce_any(int64_t x, uint32_t q)
{
  int64_t r = x % (int64_t)q;
  return (uint32_t)(r < 0 ? r + q : r);
}

// K = compile-time tupSize (1: immediate address offsets, 128-bit stores) or 0 for a run-time k.
// One ring element per CTA iteration on WARPS warps (phase 1: 6 warp-tasks of 32 columns, phase 2: 20), one 15 KB tile, two
// barriers per element.  Coarser shapes (5 elements x 10 warps, double-buffered tiles, a register software pipeline, an L2
// prefetch of the next element) were all measured slower: DESIGN.md 4.1 keeps the numbers.
// MUL fuses the pointwise product with a second operand b (canonical residues, same layout; b_stride = 0 broadcasts one
// element): forward  y <- CRT(y) . b  (multiplied at the store), inverse  y <- CRT^-1(y . b)  (multiplied at the load).
template <bool INV, class AR, int K, int WARPS, int MINB, bool MUL = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_fused_a(int64_t* __restrict__ y, int64_t batch, int k_rt, int limb, const __grid_constant__ FusedAConsts C,
          const int64_t* __restrict__ bmul = nullptr, int64_t b_stride = 0)
{
  const int k = K ? K : k_rt;
  __shared__ __align__(16) uint32_t tile[kN];
  const AR A(C);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // per-lane twiddles of the 2^6 axis, loaded once
  uint32_t ltw[9], m3l[12];        // inverse: [7] = crtTwiddle(even column) * tw_0, [8] = -crtTwiddle(odd column) * tw_0
#pragma unroll
  for (int i = 0; i < 9; i++) ltw[i] = C.lane_tw[i * 32 + lane];
#pragma unroll
  for (int i = 0; i < 12; i++) m3l[i] = C.lane_tw[(8 + i) * 32 + lane];

  for (int64_t e = blockIdx.x; e < batch; e += gridDim.x) {
    int64_t* ebase = y + (size_t)e * kN * k + limb;
    // ---------------- phase 1: 5^2 axis; warp-task = i2, lane = i1
    for (int i2 = warp; i2 < kD2; i2 += WARPS) {
      const int col = i2 * 32 + lane;
      const int64_t* src = ebase + (size_t)col * k;
      // all 20 loads are issued before the first use; the canonical-range check is two OR/max reductions
      uint32_t v[20];
      uint32_t hi_or = 0, lo_max = 0;
#pragma unroll
      for (int a = 0; a < 20; a++) {
        const int64_t raw = __ldcs(src + (size_t)(a * 192) * k);
        v[a] = (uint32_t)raw;
        hi_or |= (uint32_t)((uint64_t)raw >> 32);
        lo_max = max(lo_max, v[a]);
      }
      if (hi_or != 0 || lo_max >= C.q) {
#pragma unroll 1
        for (int a = 0; a < 20; a++) v[a] = reduce_any(src[(size_t)(a * 192) * k], C.q);
      }
      if (MUL && INV) {
        const int64_t* bsrc = bmul + (size_t)e * b_stride + (size_t)col * k + limb;
        uint32_t w[20];
        uint32_t bh = 0, bm = 0;
#pragma unroll
        for (int a = 0; a < 20; a++) {
          const int64_t raw = __ldg(bsrc + (size_t)(a * 192) * k);
          w[a] = (uint32_t)raw;
          bh |= (uint32_t)((uint64_t)raw >> 32);
          bm = max(bm, w[a]);
        }
        if (bh != 0 || bm >= C.q) {
#pragma unroll
          for (int a = 0; a < 20; a++) w[a] = reduce_any(bsrc[(size_t)(a * 192) * k], C.q);
        }
#pragma unroll
        for (int a = 0; a < 20; a++) v[a] = A.mulv(v[a], w[a]);
      }
      axis5<INV, AR>(v, C, A);
#pragma unroll
      for (int a = 0; a < 20; a++) tile[a * 192 + col] = v[a];
    }
    __syncthreads();
    // ---------------- phase 2: 3^2 axis in the thread, 2^6 axis across the warp; warp-task = i3
    for (int i3 = warp; i3 < kD3; i3 += WARPS) {
      uint32_t x[6], c0[3], c1[3];
      const uint32_t* srow = tile + i3 * 192 + lane;
#pragma unroll
      for (int i2 = 0; i2 < 6; i2++) x[i2] = srow[i2 * 32];
      axis3<INV, AR>(x, C, A, m3l);
#pragma unroll
      for (int j = 0; j < 3; j++) { c0[j] = x[2 * j]; c1[j] = x[2 * j + 1]; }
      if (!INV) {
        // crtTwiddle of 2^6 is already in m3l
#pragma unroll
        for (int r = 0; r < 4; r++) exchange_round<false, AR>(c0, c1, lane, r, ltw[1 + r], A);
        exchange_round<false, AR, true>(c0, c1, lane, 4, 0u, A);
        // lane owns rows 2j + (lane&1), columns (lane>>1) and (lane>>1)+16
        const int pos = i3 * 192 + (lane & 1) * 32 + (lane >> 1);
        int64_t* out = ebase + (size_t)pos * k;
        if (MUL) {
          const int64_t* bo = bmul + (size_t)e * b_stride + (size_t)pos * k + limb;
          int64_t braw[6];
#pragma unroll
          for (int j = 0; j < 3; j++) { braw[2 * j] = __ldg(bo + (size_t)(j * 64) * k); braw[2 * j + 1] = __ldg(bo + (size_t)(j * 64 + 16) * k); }
#pragma unroll
          for (int j = 0; j < 6; j++) {
            const uint32_t bw = (uint64_t)braw[j] < (uint64_t)C.q ? (uint32_t)braw[j] : reduce_any(braw[j], C.q);
            if (j & 1) c1[j >> 1] = A.mulv(c1[j >> 1], bw); else c0[j >> 1] = A.mulv(c0[j >> 1], bw);
          }
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
          __stcs(out + (size_t)(j * 64) * k, (int64_t)A.canon(c0[j]));
          __stcs(out + (size_t)(j * 64 + 16) * k, (int64_t)A.canon(c1[j]));
        }
      } else {
        exchange_round<true, AR, true>(c0, c1, lane, 4, 0u, A);
#pragma unroll
        for (int r = 3; r >= 1; r--) exchange_round<true, AR>(c0, c1, lane, r, ltw[r], A);
        // last round merged with the inverse crtTwiddle: lane owns rows 2j + (lane>>4), columns 2*(lane&15) and +1
        exchange_last_inv<AR>(c0, c1, lane, ltw[5], ltw[7], ltw[6], ltw[8], A);
        int64_t* out = ebase + (size_t)(i3 * 192 + (lane >> 4) * 32 + 2 * (lane & 15)) * k;
#pragma unroll
        for (int j = 0; j < 3; j++) {
          const int64_t a = (int64_t)A.canon(c0[j]);
          const int64_t b = (int64_t)A.canon(c1[j]);
          if (K == 1) {
            __stcs(reinterpret_cast<longlong2*>(out + j * 64), make_longlong2(a, b));
          } else {
            __stcs(out + (size_t)(j * 64) * k, a);
            __stcs(out + (size_t)(j * 64 + 1) * k, b);
          }
        }
      }
    }
    __syncthreads();
  }
}

// tupSize = 2 (e.g. the SymmSHE key-switch modulus q = 1008001 * 1065601): both RNS limbs of a coefficient are one
// 16-byte unit in the ABI layout, so a thread loads and stores them together (128-bit accesses, every sector fully
// used) and runs the two limbs as independent instruction streams with their own constants.
struct FusedAConsts2 { FusedAConsts c[2]; };

// DIG mode of k_fused_a_k2 (TrivGad over two limbs, SymmSHE.hs:314 `fmap reduce <$> decompose`; she_stream.cu has the
// stand-alone pass and the citations): the load stage reads a Pow-basis element and forms digit d of it on the fly, so
// the CRT of the digits needs no decomposed copy in HBM.  Digit d is lift(x_d) reduced into both limbs: x_d itself in
// limb d, and in the other limb x_d or x_d - q_d + q_other according to the sign of the lift -- exact while
// q_other >= q_d / 2 (checked on the host), and three instructions per coefficient.  (A general BaseBGad digit in this
// compute-bound kernel measured slower than the separate streaming pass.)
__device__ __forceinline__ uint32_t triv_digit_other(uint32_t x, uint32_t q_own, uint32_t q_other)
{
  return (2 * (uint64_t)x < (uint64_t)q_own) ? x : x + (q_other - q_own);      // wraps through 2^32 to [0, q_other)
}

// DIG (forward only): y = the digit arrays [l * b_stride][n][2] (output), bmul = the Pow-basis source [b_stride][n][2];
// element e of y is digit e / b_stride of source element e % b_stride.
template <bool INV, class AR, int WARPS, int MINB, bool MUL = false, bool DIG = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_fused_a_k2(int64_t* __restrict__ y, int64_t batch, const __grid_constant__ FusedAConsts2 CC,
             const int64_t* __restrict__ bmul, int64_t b_stride)
{
  extern __shared__ __align__(16) uint32_t sm_dyn[];       // [2 limbs][kN]
  const AR A0(CC.c[0]), A1(CC.c[1]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t ltw[2][9], m3l[2][12];
#pragma unroll
  for (int l = 0; l < 2; l++) {
#pragma unroll
    for (int i = 0; i < 9; i++) ltw[l][i] = CC.c[l].lane_tw[i * 32 + lane];
#pragma unroll
    for (int i = 0; i < 12; i++) m3l[l][i] = CC.c[l].lane_tw[(8 + i) * 32 + lane];
  }

  for (int64_t e = blockIdx.x; e < batch; e += gridDim.x) {
    longlong2* ebase = reinterpret_cast<longlong2*>(y + (size_t)e * kN * 2);
    const int dsel = DIG ? (int)(e / b_stride) : 0;
    const longlong2* sbase = DIG ? reinterpret_cast<const longlong2*>(bmul + (size_t)(e % b_stride) * kN * 2) : ebase;
    // ---------------- phase 1: 5^2 axis, both limbs of column (i2, lane)
    for (int i2 = warp; i2 < kD2; i2 += WARPS) {
      const int col = i2 * 32 + lane;
      const longlong2* src = sbase + col;
      uint32_t v0[20], v1[20];
      uint32_t hi_or = 0, max0 = 0, max1 = 0;
#pragma unroll
      for (int half = 0; half < 2; half++) {         // two batches of ten 128-bit loads bound the registers in flight
        longlong2 raw[10];
#pragma unroll
        for (int a = 0; a < 10; a++) raw[a] = __ldcs(src + (half * 10 + a) * 192);
#pragma unroll
        for (int a = 0; a < 10; a++) {
          v0[half * 10 + a] = (uint32_t)raw[a].x;
          v1[half * 10 + a] = (uint32_t)raw[a].y;
          hi_or |= (uint32_t)((uint64_t)raw[a].x >> 32) | (uint32_t)((uint64_t)raw[a].y >> 32);
          max0 = max(max0, (uint32_t)raw[a].x);
          max1 = max(max1, (uint32_t)raw[a].y);
        }
      }
      if (hi_or != 0 || max0 >= CC.c[0].q || max1 >= CC.c[1].q) {
#pragma unroll
        for (int a = 0; a < 20; a++) {            // static indices: a rolled loop would push v0 / v1 into local memory
          const longlong2 raw = src[a * 192];
          v0[a] = reduce_any(raw.x, CC.c[0].q);
          v1[a] = reduce_any(raw.y, CC.c[1].q);
        }
      }
      if (DIG) {
#pragma unroll
        for (int a = 0; a < 20; a++) {
          if (dsel == 0) v1[a] = triv_digit_other(v0[a], CC.c[0].q, CC.c[1].q);
          else v0[a] = triv_digit_other(v1[a], CC.c[1].q, CC.c[0].q);
        }
      }
      if (MUL && INV) {
        const longlong2* bsrc = reinterpret_cast<const longlong2*>(bmul + (size_t)e * b_stride) + col;
#pragma unroll
        for (int part = 0; part < 4; part++) {       // four batches of five 128-bit loads
          longlong2 braw[5];
#pragma unroll
          for (int a = 0; a < 5; a++) braw[a] = __ldg(bsrc + (part * 5 + a) * 192);
#pragma unroll
          for (int a = 0; a < 5; a++) {
            const uint32_t w0 = (uint64_t)braw[a].x < (uint64_t)CC.c[0].q ? (uint32_t)braw[a].x : reduce_any(braw[a].x, CC.c[0].q);
            const uint32_t w1 = (uint64_t)braw[a].y < (uint64_t)CC.c[1].q ? (uint32_t)braw[a].y : reduce_any(braw[a].y, CC.c[1].q);
            v0[part * 5 + a] = A0.mulv(v0[part * 5 + a], w0);
            v1[part * 5 + a] = A1.mulv(v1[part * 5 + a], w1);
          }
        }
      }
      axis5<INV, AR>(v0, CC.c[0], A0);
#pragma unroll
      for (int a = 0; a < 20; a++) sm_dyn[a * 192 + col] = v0[a];
      axis5<INV, AR>(v1, CC.c[1], A1);
#pragma unroll
      for (int a = 0; a < 20; a++) sm_dyn[kN + a * 192 + col] = v1[a];
    }
    __syncthreads();
    // ---------------- phase 2: 3^2 axis in the thread, 2^6 axis across the warp, both limbs of row block i3
    for (int i3 = warp; i3 < kD3; i3 += WARPS) {
      uint32_t x[2][6], c0[2][3], c1[2][3];
#pragma unroll
      for (int l = 0; l < 2; l++)
#pragma unroll
        for (int i2 = 0; i2 < 6; i2++) x[l][i2] = sm_dyn[l * kN + i3 * 192 + i2 * 32 + lane];
      axis3<INV, AR>(x[0], CC.c[0], A0, m3l[0]);
      axis3<INV, AR>(x[1], CC.c[1], A1, m3l[1]);
      if (!INV) {
#pragma unroll
        for (int l = 0; l < 2; l++)
#pragma unroll
          for (int j = 0; j < 3; j++) { c0[l][j] = x[l][2 * j]; c1[l][j] = x[l][2 * j + 1]; }
#pragma unroll
        for (int r = 0; r < 4; r++) {
          exchange_round<false, AR>(c0[0], c1[0], lane, r, ltw[0][1 + r], A0);
          exchange_round<false, AR>(c0[1], c1[1], lane, r, ltw[1][1 + r], A1);
        }
        exchange_round<false, AR, true>(c0[0], c1[0], lane, 4, 0u, A0);
        exchange_round<false, AR, true>(c0[1], c1[1], lane, 4, 0u, A1);
        longlong2* out = ebase + i3 * 192 + (lane & 1) * 32 + (lane >> 1);
        if (MUL) {
          const longlong2* bo = reinterpret_cast<const longlong2*>(bmul + (size_t)e * b_stride) + i3 * 192 + (lane & 1) * 32 + (lane >> 1);
          longlong2 braw[6];
#pragma unroll
          for (int j = 0; j < 3; j++) { braw[2 * j] = __ldg(bo + j * 64); braw[2 * j + 1] = __ldg(bo + j * 64 + 16); }
#pragma unroll
          for (int j = 0; j < 6; j++) {
            const uint32_t w0 = (uint64_t)braw[j].x < (uint64_t)CC.c[0].q ? (uint32_t)braw[j].x : reduce_any(braw[j].x, CC.c[0].q);
            const uint32_t w1 = (uint64_t)braw[j].y < (uint64_t)CC.c[1].q ? (uint32_t)braw[j].y : reduce_any(braw[j].y, CC.c[1].q);
            if (j & 1) { c1[0][j >> 1] = A0.mulv(c1[0][j >> 1], w0); c1[1][j >> 1] = A1.mulv(c1[1][j >> 1], w1); }
            else { c0[0][j >> 1] = A0.mulv(c0[0][j >> 1], w0); c0[1][j >> 1] = A1.mulv(c0[1][j >> 1], w1); }
          }
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
          __stcs(out + j * 64, make_longlong2((int64_t)A0.canon(c0[0][j]), (int64_t)A1.canon(c0[1][j])));
          __stcs(out + j * 64 + 16, make_longlong2((int64_t)A0.canon(c1[0][j]), (int64_t)A1.canon(c1[1][j])));
        }
      } else {
#pragma unroll
        for (int l = 0; l < 2; l++)
#pragma unroll
          for (int j = 0; j < 3; j++) { c0[l][j] = x[l][2 * j]; c1[l][j] = x[l][2 * j + 1]; }
#pragma unroll
        exchange_round<true, AR, true>(c0[0], c1[0], lane, 4, 0u, A0);
        exchange_round<true, AR, true>(c0[1], c1[1], lane, 4, 0u, A1);
#pragma unroll
        for (int r = 3; r >= 1; r--) {
          exchange_round<true, AR>(c0[0], c1[0], lane, r, ltw[0][r], A0);
          exchange_round<true, AR>(c0[1], c1[1], lane, r, ltw[1][r], A1);
        }
        exchange_last_inv<AR>(c0[0], c1[0], lane, ltw[0][5], ltw[0][7], ltw[0][6], ltw[0][8], A0);     // last round + inverse crtTwiddle
        exchange_last_inv<AR>(c0[1], c1[1], lane, ltw[1][5], ltw[1][7], ltw[1][6], ltw[1][8], A1);
        longlong2* out = ebase + i3 * 192 + (lane >> 4) * 32 + 2 * (lane & 15);
#pragma unroll
        for (int j = 0; j < 3; j++) {
          __stcs(out + j * 64, make_longlong2((int64_t)A0.canon(c0[0][j]), (int64_t)A1.canon(c0[1][j])));
          __stcs(out + j * 64 + 1, make_longlong2((int64_t)A0.canon(c1[0][j]), (int64_t)A1.canon(c1[1][j])));
        }
      }
    }
    __syncthreads();
  }
}

// tupSize 3: every limb of one ring element in the same CTA iteration.  A launch per limb
// reads 8 of every 8k bytes it touches, so HBM moves the whole batch k times; here a warp walks the limbs of its column
// block back to back, the first limb's loads bring the sectors on chip and the others hit L1/L2, and the k partial-
// sector stores of a row meet in L2 before they are written back: one HBM read and one HBM write per element again.
// The limb index is a uniform loop counter, so one copy of the code reads its constants through c[limb].
constexpr int kMaxLimbsN = 3;      // serves tupSize 3 (the de-interleaving kernel below is faster from tupSize 4 on)
struct FusedAConstsN { FusedAConsts c[kMaxLimbsN]; };

template <bool INV, class AR, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_fused_a_kn(int64_t* __restrict__ y, int64_t batch, int k, const __grid_constant__ FusedAConstsN CC)
{
  extern __shared__ __align__(16) uint32_t sm_dyn[];       // [k][kN]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t e = blockIdx.x; e < batch; e += gridDim.x) {
    int64_t* ebase = y + (size_t)e * kN * k;
    // ---------------- phase 1: 5^2 axis; warp-task = column block i2, all limbs in turn
    // (work items (limb, task) are dealt round-robin in limb-major order, so the warps stay balanced whenever W divides
    // 6k and 20k, while the limb stays the uniform counter of the outer loop)
#pragma unroll 1
    for (int limb = 0; limb < k; limb++) {
      const FusedAConsts& C = CC.c[limb];
      const AR A(C);
      for (int i2 = (warp + WARPS * kD2 - limb * kD2 % WARPS) % WARPS; i2 < kD2; i2 += WARPS) {
        const int col = i2 * 32 + lane;
        const int64_t* src = ebase + (size_t)col * k + limb;
        uint32_t v[20];
        uint32_t hi_or = 0, lo_max = 0;
#pragma unroll
        for (int a = 0; a < 20; a++) {
          const int64_t raw = src[(size_t)(a * 192) * k];
          v[a] = (uint32_t)raw;
          hi_or |= (uint32_t)((uint64_t)raw >> 32);
          lo_max = max(lo_max, v[a]);
        }
        if (hi_or != 0 || lo_max >= C.q) {
#pragma unroll
          for (int a = 0; a < 20; a++) v[a] = reduce_any(src[(size_t)(a * 192) * k], C.q);
        }
        axis5<INV, AR>(v, C, A);
        uint32_t* dst = sm_dyn + limb * kN + col;
#pragma unroll
        for (int a = 0; a < 20; a++) dst[a * 192] = v[a];
      }
    }
    __syncthreads();
    // ---------------- phase 2: 3^2 axis in the thread, 2^6 axis across the warp; warp-task = row block i3, all limbs in turn
#pragma unroll 1
    for (int limb = 0; limb < k; limb++) {
      const FusedAConsts& C = CC.c[limb];
      const AR A(C);
      const uint32_t* lt = C.lane_tw + lane;
      for (int i3 = (warp + WARPS * kD3 - limb * kD3 % WARPS) % WARPS; i3 < kD3; i3 += WARPS) {
        uint32_t x[6], c0[3], c1[3];
#pragma unroll
        for (int j = 0; j < 6; j++) x[j] = sm_dyn[limb * kN + i3 * 192 + j * 32 + lane];
        if (!INV) {
          uint32_t m3l[12], ltw[4];
#pragma unroll
          for (int i = 0; i < 12; i++) m3l[i] = __ldg(lt + (8 + i) * 32);
#pragma unroll
          for (int i = 0; i < 4; i++) ltw[i] = __ldg(lt + (1 + i) * 32);
          axis3<false, AR>(x, C, A, m3l);
#pragma unroll
          for (int j = 0; j < 3; j++) { c0[j] = x[2 * j]; c1[j] = x[2 * j + 1]; }
#pragma unroll
          for (int r = 0; r < 4; r++) exchange_round<false, AR>(c0, c1, lane, r, ltw[r], A);
          exchange_round<false, AR, true>(c0, c1, lane, 4, 0u, A);
          int64_t* out = ebase + (size_t)(i3 * 192 + (lane & 1) * 32 + (lane >> 1)) * k + limb;
#pragma unroll
          for (int j = 0; j < 3; j++) {
            out[(size_t)(j * 64) * k] = (int64_t)A.canon(c0[j]);
            out[(size_t)(j * 64 + 16) * k] = (int64_t)A.canon(c1[j]);
          }
        } else {
          uint32_t m3l[12] = {}, ltw[9];
#pragma unroll
          for (int i = 1; i < 9; i++) ltw[i] = (i == 4) ? 0u : __ldg(lt + i * 32);
          axis3<true, AR>(x, C, A, m3l);
#pragma unroll
          for (int j = 0; j < 3; j++) { c0[j] = x[2 * j]; c1[j] = x[2 * j + 1]; }
          exchange_round<true, AR, true>(c0, c1, lane, 4, 0u, A);
#pragma unroll
          for (int r = 3; r >= 1; r--) exchange_round<true, AR>(c0, c1, lane, r, ltw[r], A);
          exchange_last_inv<AR>(c0, c1, lane, ltw[5], ltw[7], ltw[6], ltw[8], A);
          int64_t* out = ebase + (size_t)(i3 * 192 + (lane >> 4) * 32 + 2 * (lane & 15)) * k + limb;
#pragma unroll
          for (int j = 0; j < 3; j++) {
            out[(size_t)(j * 64) * k] = (int64_t)A.canon(c0[j]);
            out[(size_t)(j * 64 + 1) * k] = (int64_t)A.canon(c1[j]);
          }
        }
      }
    }
    __syncthreads();
  }
}

// tupSize 2 .. 8, de-interleaving schedule: the element [n][K] is read and written ONLY with coalesced 128-bit accesses of
// the whole 8 n K byte block (every sector fully used, for any K), the limbs are separated on the way into a [K][n] u32
// tile, and both transform phases work tile -> registers -> tile:
//   phase 0  all threads: 16-byte loads of the element, canonical check, u32 words to tile[limb][j]
//   phase 1  warp-task (limb, i2): 5^2 axis on the 20-value column of a lane, in place in the tile
//   phase 2  warp-task (limb, i3): 3^2 axis in the thread, 2^6 axis across the warp, canonical residues back to the tile
//            (a row block is touched by its own warp-task only: the writes follow the last shuffle, which follows every
//            lane's reads)
//   phase 3  all threads: tile -> interleaved int64, 16-byte stores
// With 2K warps every warp has exactly 3 tasks in phase 1 and 10 in phase 2.  The limb is a run-time index into the
// per-limb constants (LDC), so one copy of the code serves every limb; per-lane constants come from L1.
// (k_fused_a_kn, which this replaces for K >= 3, read 8 of every 8K bytes it touched per instruction: 49 % -> 31 % of
// the HBM roofline for K = 3 .. 7.)
constexpr int kMaxLimbsD = 8;
template <int K> struct FusedAConstsK { FusedAConsts c[K]; };
__host__ __device__ constexpr int kd_pad(int K) { return K == 4 ? 8 : K == 5 ? 4 : K == 6 ? 3 : K == 7 ? 4 : K == 8 ? 4 : 0; }   // limb stride kN + pad: fewest bank conflicts in phases 0 / 3

template <bool INV, class AR, int K, int MINB>
__global__ void __launch_bounds__(64 * K, MINB)
k_fused_a_kd(int64_t* __restrict__ y, int64_t batch, const __grid_constant__ FusedAConstsK<K> CC)
{
  constexpr int WARPS = 2 * K, T = 32 * WARPS, LS = kN + kd_pad(K), NPIECE = kN * K / 2, PPT = NPIECE / T, UB = 10;
  static_assert(NPIECE % T == 0 && PPT % UB == 0, "pieces must tile the CTA");
  extern __shared__ __align__(16) uint32_t tile[];       // [K][LS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t e = blockIdx.x; e < batch; e += gridDim.x) {
    longlong2* e16 = reinterpret_cast<longlong2*>(y + (size_t)e * kN * K);
    // ---------------- phase 0: coalesced load, limbs separated
#pragma unroll 1
    for (int b0 = 0; b0 < PPT; b0 += UB) {
      longlong2 raw[UB];
#pragma unroll
      for (int u = 0; u < UB; u++) raw[u] = __ldcs(e16 + tid + (b0 + u) * T);
#pragma unroll
      for (int u = 0; u < UB; u++) {
        const int f = 2 * (tid + (b0 + u) * T);
        const int j0 = f / K, l0 = f - j0 * K, j1 = (f + 1) / K, l1 = (f + 1) - j1 * K;
        const uint32_t q0 = CC.c[l0].q, q1 = CC.c[l1].q;
        uint32_t w0 = (uint32_t)raw[u].x, w1 = (uint32_t)raw[u].y;
        if ((uint64_t)raw[u].x >= (uint64_t)q0) w0 = reduce_any(raw[u].x, q0);      // outside the Haskell contract: like `c % q`
        if ((uint64_t)raw[u].y >= (uint64_t)q1) w1 = reduce_any(raw[u].y, q1);
        tile[l0 * LS + j0] = w0;
        tile[l1 * LS + j1] = w1;
      }
    }
    __syncthreads();
    // ---------------- phase 1: 5^2 axis, warp-task (limb, i2)
#pragma unroll 1
    for (int item = warp; item < K * kD2; item += WARPS) {
      const int limb = item / kD2, i2 = item - limb * kD2;
      const FusedAConsts& C = CC.c[limb];
      const AR A(C);
      uint32_t* colp = tile + limb * LS + i2 * 32 + lane;
      uint32_t v[20];
#pragma unroll
      for (int a = 0; a < 20; a++) v[a] = colp[a * 192];
      axis5<INV, AR>(v, C, A);
#pragma unroll
      for (int a = 0; a < 20; a++) colp[a * 192] = v[a];
    }
    __syncthreads();
    // ---------------- phase 2: 3^2 axis in the thread, 2^6 axis across the warp, warp-task (limb, i3)
#pragma unroll 1
    for (int item = warp; item < K * kD3; item += WARPS) {
      const int limb = item / kD3, i3 = item - limb * kD3;
      const FusedAConsts& C = CC.c[limb];
      const AR A(C);
      const uint32_t* lt = C.lane_tw + lane;
      uint32_t* row = tile + limb * LS + i3 * 192;
      uint32_t x[6], c0[3], c1[3];
#pragma unroll
      for (int j = 0; j < 6; j++) x[j] = row[j * 32 + lane];
      if (!INV) {
        uint32_t m3l[12], ltw[4];
#pragma unroll
        for (int i = 0; i < 12; i++) m3l[i] = __ldg(lt + (8 + i) * 32);
#pragma unroll
        for (int i = 0; i < 4; i++) ltw[i] = __ldg(lt + (1 + i) * 32);
        axis3<false, AR>(x, C, A, m3l);
#pragma unroll
        for (int j = 0; j < 3; j++) { c0[j] = x[2 * j]; c1[j] = x[2 * j + 1]; }
#pragma unroll
        for (int r = 0; r < 4; r++) exchange_round<false, AR>(c0, c1, lane, r, ltw[r], A);
        exchange_round<false, AR, true>(c0, c1, lane, 4, 0u, A);
        uint32_t* out = row + (lane & 1) * 32 + (lane >> 1);
#pragma unroll
        for (int j = 0; j < 3; j++) {
          out[j * 64] = A.canon(c0[j]);
          out[j * 64 + 16] = A.canon(c1[j]);
        }
      } else {
        uint32_t m3l[12] = {}, ltw[9];
#pragma unroll
        for (int i = 1; i < 9; i++) ltw[i] = (i == 4) ? 0u : __ldg(lt + i * 32);
        axis3<true, AR>(x, C, A, m3l);
#pragma unroll
        for (int j = 0; j < 3; j++) { c0[j] = x[2 * j]; c1[j] = x[2 * j + 1]; }
        exchange_round<true, AR, true>(c0, c1, lane, 4, 0u, A);
#pragma unroll
        for (int r = 3; r >= 1; r--) exchange_round<true, AR>(c0, c1, lane, r, ltw[r], A);
        exchange_last_inv<AR>(c0, c1, lane, ltw[5], ltw[7], ltw[6], ltw[8], A);
        uint32_t* out = row + (lane >> 4) * 32 + 2 * (lane & 15);      // (the limb stride may be odd: no 64-bit store)
#pragma unroll
        for (int j = 0; j < 3; j++) {
          out[j * 64] = A.canon(c0[j]);
          out[j * 64 + 1] = A.canon(c1[j]);
        }
      }
    }
    __syncthreads();
    // ---------------- phase 3: coalesced store
#pragma unroll 5
    for (int b0 = 0; b0 < PPT; b0++) {
      const int f = 2 * (tid + b0 * T);
      const int j0 = f / K, l0 = f - j0 * K, j1 = (f + 1) / K, l1 = (f + 1) - j1 * K;
      __stcs(e16 + tid + b0 * T, make_longlong2((int64_t)tile[l0 * LS + j0], (int64_t)tile[l1 * LS + j1]));
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ host: constants from the plan's root tables

enum ArithClass { ARITH_NONE = 0, ARITH_S, ARITH_M };

// which policy a modulus admits (see ArithS / ArithM)
inline ArithClass arith_class(uint64_t q)
{
  if (8 * q * q + 2 * q < ((uint64_t)1 << 32)) return ARITH_S;
  if (10 * q < ((uint64_t)1 << 32) && (q & 1)) return ARITH_M;      // Montgomery needs odd q
  return ARITH_NONE;
}

struct FusedA {
  bool ok_fwd = false, ok_inv = false;
  std::vector<int> cls;                    // ArithClass per limb
  std::vector<FusedAConsts> fwd, inv;      // per limb
  uint32_t* d_lane_tw = nullptr;           // [k][2][kLaneRows][32]
};

inline uint64_t rd(const std::vector<int64_t>& tab, int64_t j, int k, int limb, uint64_t q)
{
  int64_t v = tab[(size_t)j * k + limb] % (int64_t)q;
  if (v < 0) v += q;
  return (uint64_t)v;
}

void build_consts(const lolb_plan* pl, bool inverse, int limb, int cls, FusedAConsts* C, uint32_t* lane_tw /* [kLaneRows][32] */)
{
  const int k = pl->k;
  const uint64_t q = (uint64_t)pl->qs[limb];
  const auto& T = inverse ? pl->ruinv : pl->ru;
  auto r64 = [&](int64_t j) { return rd(T[0], j % 64, k, limb, q); };
  auto r9 = [&](int64_t j) { return rd(T[1], j % 9, k, limb, q); };
  auto r25 = [&](int64_t j) { return rd(T[2], j % 25, k, limb, q); };
  C->q = (uint32_t)q; C->q2 = (uint32_t)(2 * q); C->r0 = (uint32_t)((((uint64_t)1) << 32) / q); C->one = 1;
  const uint64_t scale = inverse ? (uint64_t)(((pl->mhatinv[limb] % (int64_t)q) + (int64_t)q) % (int64_t)q) : 1;
  for (int row = 0; row < 5; row++) for (int col = 0; col < 5; col++) C->d5[row][col] = (uint32_t)r25(5 * ((row * col) % 5));
  for (int row = 0; row < 3; row++) for (int col = 0; col < 3; col++) C->d3[row][col] = (uint32_t)r9(3 * ((row * col) % 3));
  for (int i0 = 0; i0 < 5; i0++)
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) {
        uint64_t v;
        if (!inverse) {   // crtTwiddle(i0, r) * CRT_5[r][c]   (crt.cpp:60-79, 272-295)
          const uint64_t tw = i0 ? r25((int64_t)i0 * (r + 1)) : 1;
          v = mulmod64(tw, r25(5 * (((r + 1) * c) % 5)), q);
        } else {          // (w^-r(c+1) - w^(c+1)) * crtTwiddle(i0, c) * mhat^-1   (crt.cpp:376-399)
          const uint64_t tw = i0 ? r25((int64_t)i0 * (c + 1)) : 1;
          const uint64_t mat = (r25(5 * ((r * (c + 1)) % 5)) + q - r25(5 * (5 - c - 1))) % q;
          v = mulmod64(mulmod64(tw, mat, q), scale, q);
        }
        C->m5[i0][r][c] = (uint32_t)v;
      }
  for (int i0 = 0; i0 < 3; i0++)
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 2; c++) {
        uint64_t v;
        if (!inverse) {
          const uint64_t tw = i0 ? r9((int64_t)i0 * (r + 1)) : 1;
          v = mulmod64(tw, r9(3 * (((r + 1) * c) % 3)), q);
        } else {
          const uint64_t tw = i0 ? r9((int64_t)i0 * (c + 1)) : 1;
          const uint64_t mat = (r9(3 * ((r * (c + 1)) % 3)) + q - r9(3 * (3 - c - 1))) % q;
          v = mulmod64(tw, mat, q);
        }
        C->m3[i0][r][c] = (uint32_t)v;
      }
  // 2^6 axis.  Forward rows: [0] crtTwiddle by column = lane; [1+r] round r, i0 = lane >> (r+1).
  // Inverse rows: [r] round r, i0 = (lane >> r) & (2^(4-r) - 1); [5],[6] crtTwiddle of columns 2*(lane&15) + {0,1}.
  for (int i = 0; i < kLaneRows * 32; i++) lane_tw[i] = 1;
  for (int lane = 0; lane < 32; lane++) {
    if (!inverse) {
      const uint64_t tw0 = lane ? r64(digit_rev(2, 5, lane)) : 1;           // crtTwiddle of 2^6 (crt.cpp:43-58), column = lane
      lane_tw[0 * 32 + lane] = (uint32_t)tw0;
      for (int i0 = 0; i0 < 3; i0++)
        for (int r = 0; r < 2; r++)
          for (int c = 0; c < 2; c++)
            lane_tw[(8 + 4 * i0 + 2 * r + c) * 32 + lane] = (uint32_t)mulmod64(C->m3[i0][r][c], tw0, q);
      for (int r = 0; r < 5; r++) {
        const int i0 = lane >> (r + 1);
        const uint64_t tw = i0 ? r64(digit_rev(2, 4 - r, i0) * (2 << r)) : 1;
        // lanes whose bit r is set hold (t, u) instead of (u, t): they multiply (t - u) by -tw
        lane_tw[(1 + r) * 32 + lane] = (uint32_t)(((lane >> r) & 1) ? (q - tw) % q : tw);
      }
    } else {
      for (int r = 0; r < 5; r++) {
        const int i0 = (lane >> r) & ((1 << (4 - r)) - 1);
        lane_tw[r * 32 + lane] = i0 ? (uint32_t)r64(digit_rev(2, 4 - r, i0) * (2 << r)) : 1;
      }
      for (int s = 0; s < 2; s++) {
        const int col = 2 * (lane & 15) + s;
        lane_tw[(5 + s) * 32 + lane] = col ? (uint32_t)r64(digit_rev(2, 5, col)) : 1;
      }
      // the last round (r = 0) merged with the crtTwiddle above: a * tw_0 and -(b * tw_0)  (exchange_last_inv)
      lane_tw[7 * 32 + lane] = (uint32_t)mulmod64(lane_tw[5 * 32 + lane], lane_tw[0 * 32 + lane], q);
      lane_tw[8 * 32 + lane] = (uint32_t)((q - mulmod64(lane_tw[6 * 32 + lane], lane_tw[0 * 32 + lane], q)) % q);
    }
  }
  if (cls == ARITH_M) {
    // constants to Montgomery form c * 2^32 mod q; r0 = -q^-1 mod 2^32 (Newton iteration on the odd q)
    auto mont = [&](uint32_t c) { return (uint32_t)((((uint64_t)c) << 32) % q); };
    uint32_t inv = (uint32_t)q;                       // q * inv == 1 mod 2^3 initially
    for (int i = 0; i < 5; i++) inv *= 2u - (uint32_t)q * inv;
    C->r0 = 0u - inv;
    C->one = mont(1);
    C->r2 = mont(mont(1));
    for (auto& a : C->m5) for (auto& b : a) for (auto& c : b) c = mont(c);
    for (auto& a : C->d5) for (auto& c : a) c = mont(c);
    for (auto& a : C->m3) for (auto& b : a) for (auto& c : b) c = mont(c);
    for (auto& a : C->d3) for (auto& c : a) c = mont(c);
    for (int i = 0; i < kLaneRows * 32; i++) lane_tw[i] = mont(lane_tw[i]);
  }
}

bool shape_is_a(const lolb_plan* pl)
{
  if (pl->kind != PLAN_RQ || pl->pe.size() != 3) return false;
  const PrimeExponent want[3] = {{2, 6}, {3, 2}, {5, 2}};
  for (int i = 0; i < 3; i++) if (pl->pe[i].prime != want[i].prime || pl->pe[i].exponent != want[i].exponent) return false;
  for (int64_t q : pl->qs)
    if (arith_class((uint64_t)q) == ARITH_NONE) return false;
  return true;
}

}  // namespace

int fused_a_select(lolb_plan* pl, void** slot)
{
  FusedA* F = (FusedA*)*slot;
  if (!shape_is_a(pl)) { return LOLB_OK; }
  if (!F) { F = new FusedA(); *slot = F; }
  const int k = pl->k;
  std::vector<uint32_t> lt((size_t)k * 2 * kLaneRows * 32, 1u);
  F->cls.assign(k, ARITH_NONE);
  for (int t = 0; t < k; t++) F->cls[t] = arith_class((uint64_t)pl->qs[t]);
  // mixed classes: every odd modulus of the small class also fits the Montgomery class, and one class for all limbs
  // lets the multi-limb kernels (k_fused_a_k2, k_fused_a_kn) take the element in one pass
  bool any_m = false, all_odd = true;
  for (int t = 0; t < k; t++) { any_m |= F->cls[t] == ARITH_M; all_odd &= (pl->qs[t] & 1) != 0; }
  if (any_m && all_odd) F->cls.assign(k, ARITH_M);
  F->fwd.assign(k, FusedAConsts{});
  F->inv.assign(k, FusedAConsts{});
  F->ok_fwd = pl->has_fwd && pl->ru.size() == 3;
  F->ok_inv = pl->has_inv && pl->ruinv.size() == 3 && (int)pl->mhatinv.size() == k;
  for (int t = 0; t < k; t++) {
    if (F->ok_fwd) build_consts(pl, false, t, F->cls[t], &F->fwd[t], lt.data() + ((size_t)t * 2 + 0) * kLaneRows * 32);
    if (F->ok_inv) build_consts(pl, true, t, F->cls[t], &F->inv[t], lt.data() + ((size_t)t * 2 + 1) * kLaneRows * 32);
  }
  if (F->d_lane_tw) { cudaFree(F->d_lane_tw); F->d_lane_tw = nullptr; }
  LOLB_CUDA(cudaMalloc((void**)&F->d_lane_tw, lt.size() * sizeof(uint32_t)));
  LOLB_CUDA(cudaMemcpy(F->d_lane_tw, lt.data(), lt.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  for (int t = 0; t < k; t++) {
    F->fwd[t].lane_tw = F->d_lane_tw + ((size_t)t * 2 + 0) * kLaneRows * 32;
    F->inv[t].lane_tw = F->d_lane_tw + ((size_t)t * 2 + 1) * kLaneRows * 32;
  }
  return LOLB_OK;
}

void fused_a_release(void* slot)
{
  FusedA* F = (FusedA*)slot;
  if (!F) return;
  if (F->d_lane_tw) cudaFree(F->d_lane_tw);
  delete F;
}

bool fused_a_available(const void* slot, bool inverse)
{
  const FusedA* F = (const FusedA*)slot;
  return F && (inverse ? F->ok_inv : F->ok_fwd);
}

// the default launch: 3 warps x 8 CTAs/SM (80 registers, no spills), grid-stride over the batch
template <bool INV, class AR, int K>
static int launch_a(const lolb_plan* pl, int64_t* y, int64_t batch, int limb, const FusedAConsts& C, cudaStream_t st)
{
  constexpr bool MONT = sizeof(typename AR::Acc) == 8;
  constexpr int W = MONT ? LOLB_A_M_W : 3, MB = MONT ? LOLB_A_M_MB : 8;
  int64_t grid = (int64_t)pl->num_sms * MB;
  if (grid > batch) grid = batch;
  k_fused_a<INV, AR, K, W, MB><<<(int)grid, W * 32, 0, st>>>(y, batch, pl->k, limb, C, nullptr, 0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_fused_a");
  count_launch();
  return LOLB_OK;
}

// y <- CRT(y) . b   /   y <- CRT^-1(y . b)   in one pass (b_batch = 1 broadcasts b)
template <bool INV, class AR, int K, int MINB>
static int launch_a_mul(const lolb_plan* pl, int64_t* y, const int64_t* b, int64_t batch, int64_t b_batch, int limb,
                        const FusedAConsts& C, cudaStream_t st)
{
  int64_t grid = (int64_t)pl->num_sms * MINB;
  if (grid > batch) grid = batch;
  const int64_t b_stride = b_batch == 1 ? 0 : (int64_t)kN * pl->k;
  k_fused_a<INV, AR, K, 3, MINB, true><<<(int)grid, 96, 0, st>>>(y, batch, pl->k, limb, C, b, b_stride);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_fused_a<MUL>");
  count_launch();
  return LOLB_OK;
}

int fused_a_crt_mul(const lolb_plan* pl, const void* slot, bool inverse, int64_t* y, const int64_t* b, int64_t batch,
                    int64_t b_batch, cudaStream_t st)
{
  const FusedA* F = (const FusedA*)slot;
  if (!fused_a_available(slot, inverse)) return LOLB_FUSED_UNAVAILABLE;
  if (batch <= 0) return LOLB_OK;
  if (pl->k == 2 && F->cls[0] == F->cls[1]) {      // both limbs in one 128-bit kernel (one strided launch per limb was measured slower than the unfused pair)
    FusedAConsts2 CC;
    CC.c[0] = inverse ? F->inv[0] : F->fwd[0];
    CC.c[1] = inverse ? F->inv[1] : F->fwd[1];
    constexpr int W = 3, MB = 5;
    const size_t smem = 2 * kN * sizeof(uint32_t);
    int64_t grid = (int64_t)pl->num_sms * MB;
    if (grid > batch) grid = batch;
    const int64_t b_stride = b_batch == 1 ? 0 : (int64_t)kN * 2;
    if (F->cls[0] == ARITH_M) {      // the Montgomery class needs the registers: CTA shape of fused_a_crt
      constexpr int W2 = LOLB_A_K2_W, MB2 = LOLB_A_K2_MB;
      int64_t g4 = (int64_t)pl->num_sms * MB2;
      if (g4 > batch) g4 = batch;
      if (inverse) k_fused_a_k2<true, ArithM, W2, MB2, true><<<(int)g4, W2 * 32, smem, st>>>(y, batch, CC, b, b_stride);
      else k_fused_a_k2<false, ArithM, W2, MB2, true><<<(int)g4, W2 * 32, smem, st>>>(y, batch, CC, b, b_stride);
    } else {
      if (inverse) k_fused_a_k2<true, ArithS, W, MB, true><<<(int)grid, W * 32, smem, st>>>(y, batch, CC, b, b_stride);
      else k_fused_a_k2<false, ArithS, W, MB, true><<<(int)grid, W * 32, smem, st>>>(y, batch, CC, b, b_stride);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_fused_a_k2<MUL>");
    count_launch();
    return LOLB_OK;
  }
  for (int t = 0; t < pl->k; t++) {
    const FusedAConsts& C = inverse ? F->inv[t] : F->fwd[t];
    int rc;
#define LM(AR, KK, MB) (inverse ? launch_a_mul<true, AR, KK, MB>(pl, y, b, batch, b_batch, t, C, st) \
                                : launch_a_mul<false, AR, KK, MB>(pl, y, b, batch, b_batch, t, C, st))
    if (F->cls[t] == ARITH_M) rc = pl->k == 1 ? LM(ArithM, 1, 8) : LM(ArithM, 0, 8);
    else rc = pl->k == 1 ? LM(ArithS, 1, 8) : LM(ArithS, 0, 8);
#undef LM
    if (rc) return rc;
  }
  return LOLB_OK;
}

// digits[d] <- CRT(reduce(decompose(x)[d])), d = 0, 1, for TrivGad over two limbs and x in the powerful basis: the
// decomposition happens in the load stage of the tupSize-2 kernel
int fused_a_decompose_crt(const lolb_plan* pl, const void* slot, const int64_t* x, int64_t* digits, int64_t batch, int64_t base,
                          cudaStream_t st)
{
  const FusedA* F = (const FusedA*)slot;
  if (!fused_a_available(slot, false) || pl->k != 2 || F->cls[0] != F->cls[1] || base != 0) return LOLB_FUSED_UNAVAILABLE;
  if (2 * pl->qs[0] < pl->qs[1] || 2 * pl->qs[1] < pl->qs[0]) return LOLB_FUSED_UNAVAILABLE;      // see triv_digit_other
  if (batch <= 0) return LOLB_OK;
  FusedAConsts2 CC;
  CC.c[0] = F->fwd[0];
  CC.c[1] = F->fwd[1];
  const int64_t total = batch * 2;
  constexpr int W = 3;
  const size_t smem = 2 * kN * sizeof(uint32_t);
  if (F->cls[0] == ARITH_M) {
    constexpr int W2 = LOLB_A_K2_W, MB2 = LOLB_A_K2_MB;
    int64_t gg = (int64_t)pl->num_sms * MB2;
    if (gg > total) gg = total;
    k_fused_a_k2<false, ArithM, W2, MB2, false, true><<<(int)gg, W2 * 32, smem, st>>>(digits, total, CC, x, batch);
  } else {
    int64_t gg = (int64_t)pl->num_sms * 5;
    if (gg > total) gg = total;
    k_fused_a_k2<false, ArithS, W, 5, false, true><<<(int)gg, W * 32, smem, st>>>(digits, total, CC, x, batch);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_fused_a_k2<DIG>");
  count_launch();
  return LOLB_OK;
}

int fused_a_crt(const lolb_plan* pl, const void* slot, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  const FusedA* F = (const FusedA*)slot;
  if (!fused_a_available(slot, inverse)) return LOLB_FUSED_UNAVAILABLE;
  if (batch <= 0) return LOLB_OK;
  if (pl->k >= 2 && pl->k <= kMaxLimbsD) {
    bool same = true;
    for (int t = 1; t < pl->k; t++) same &= F->cls[t] == F->cls[0];
    // Measured on B200 (CRT / CRT^-1, % of the HBM roofline, Montgomery class): tupSize 2  k_fused_a_k2 62 / 55, this kernel
    // 54 / 48 (ncu: 26 % more instructions for the two extra shared-memory stages);  tupSize 3  k_fused_a_kn 49 / 44, this
    // kernel 44 / 42;  tupSize 4 / 5 / 7 / 8  this kernel 47 / 46, 41 / 39, 37 / 36, 36 / 36 against 41 / 38, 38 / -, 31 / -
    // and one launch per limb.  LOLB_FUSED_A_KD=1 forces this kernel for tupSize 2 and 3 too (tests, tuning).
    const char* kd_env = getenv("LOLB_FUSED_A_KD");
    if (same && (pl->k >= 4 || (kd_env && kd_env[0] == '1'))) {
      cudaError_t e = cudaSuccess;
#define KD(AR, KK, MB)                                                                                             \
      do {                                                                                                          \
        FusedAConstsK<KK> CC;                                                                                       \
        for (int t = 0; t < KK; t++) CC.c[t] = inverse ? F->inv[t] : F->fwd[t];                                     \
        const size_t smem = (size_t)KK * (kN + kd_pad(KK)) * sizeof(uint32_t);                                      \
        auto kern = inverse ? k_fused_a_kd<true, AR, KK, MB> : k_fused_a_kd<false, AR, KK, MB>;                     \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                     \
        int64_t gg = (int64_t)pl->num_sms * MB;                                                                     \
        if (gg > batch) gg = batch;                                                                                 \
        if (e == cudaSuccess) kern<<<(int)gg, 64 * KK, smem, st>>>(y, batch, CC);                                   \
      } while (0)
#define KDS(AR)                                                                                                    \
      switch (pl->k) {                                                                                              \
        case 2: KD(AR, 2, 6); break;                                                                                \
        case 3: KD(AR, 3, 4); break;                                                                                \
        case 4: KD(AR, 4, 3); break;                                                                                \
        case 5: KD(AR, 5, 2); break;                                                                                \
        case 6: KD(AR, 6, 2); break;                                                                                \
        case 7: KD(AR, 7, 2); break;                                                                                \
        default: KD(AR, 8, 1); break;                                                                               \
      }
      if (F->cls[0] == ARITH_M) { KDS(ArithM) } else { KDS(ArithS) }
#undef KDS
#undef KD
      if (e == cudaSuccess) e = cudaGetLastError();
      if (e != cudaSuccess) return cuda_fail(e, "k_fused_a_kd");
      count_launch();
      return LOLB_OK;
    }
  }
  if (pl->k == 2 && F->cls[0] == F->cls[1]) {
    FusedAConsts2 CC;
    CC.c[0] = inverse ? F->inv[0] : F->fwd[0];
    CC.c[1] = inverse ? F->inv[1] : F->fwd[1];
    constexpr int W = 3, MB = 5;
    const size_t smem = 2 * kN * sizeof(uint32_t);
    int64_t grid = (int64_t)pl->num_sms * MB;
    if (grid > batch) grid = batch;
    if (F->cls[0] == ARITH_M) {
      // ArithM with both limbs in flight needs more than the 128 registers that 5 CTAs/SM leave: measured (config C
      // moduli, % of HBM peak forward / inverse) 5 CTAs x 128 registers (spills) 57.5 / 52.8, 4 x 168 62.4 / 54.8; the
      // per-lane constants in shared memory (5 / 6 CTAs per SM) were slower still
      constexpr int W2 = LOLB_A_K2_W, MB2 = LOLB_A_K2_MB;
      int64_t gg = (int64_t)pl->num_sms * MB2;
      if (gg > batch) gg = batch;
      if (inverse) k_fused_a_k2<true, ArithM, W2, MB2><<<(int)gg, W2 * 32, smem, st>>>(y, batch, CC, nullptr, 0);
      else k_fused_a_k2<false, ArithM, W2, MB2><<<(int)gg, W2 * 32, smem, st>>>(y, batch, CC, nullptr, 0);
    } else {
      if (inverse) k_fused_a_k2<true, ArithS, W, MB><<<(int)grid, W * 32, smem, st>>>(y, batch, CC, nullptr, 0);
      else k_fused_a_k2<false, ArithS, W, MB><<<(int)grid, W * 32, smem, st>>>(y, batch, CC, nullptr, 0);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_fused_a_k2");
    count_launch();
    return LOLB_OK;
  }
  if (pl->k == 3 && F->cls[1] == F->cls[0] && F->cls[2] == F->cls[0]) {      // tupSize 3: limb loop over strided 8-byte accesses
    FusedAConstsN CC;
    for (int t = 0; t < 3; t++) CC.c[t] = inverse ? F->inv[t] : F->fwd[t];
    const size_t smem = (size_t)3 * kN * sizeof(uint32_t);
    constexpr int W = 6, MB = 4;      // 6 warps: 3 phase-1 and 10 phase-2 tasks per warp; 4 CTAs per SM next to the 45 KB tile
    cudaError_t e = cudaSuccess;
#define KN(AR)                                                                                                    \
    do {                                                                                                           \
      auto kern = inverse ? k_fused_a_kn<true, AR, W, MB> : k_fused_a_kn<false, AR, W, MB>;                        \
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                      \
      int64_t gg = (int64_t)pl->num_sms * MB;                                                                      \
      if (gg > batch) gg = batch;                                                                                  \
      if (e == cudaSuccess) kern<<<(int)gg, W * 32, smem, st>>>(y, batch, 3, CC);                                  \
    } while (0)
    if (F->cls[0] == ARITH_M) KN(ArithM); else KN(ArithS);
#undef KN
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_fused_a_kn");
    count_launch();
    return LOLB_OK;
  }
  for (int t = 0; t < pl->k; t++) {
    const FusedAConsts& C = inverse ? F->inv[t] : F->fwd[t];
    int rc;
#define LA(AR, KK) (inverse ? launch_a<true, AR, KK>(pl, y, batch, t, C, st) : launch_a<false, AR, KK>(pl, y, batch, t, C, st))
    if (F->cls[t] == ARITH_M) rc = pl->k == 1 ? LA(ArithM, 1) : LA(ArithM, 0);
    else rc = pl->k == 1 ? LA(ArithS, 1) : LA(ArithS, 0);
#undef LA
    if (rc) return rc;
  }
  return LOLB_OK;
}

}  // namespace lolb
