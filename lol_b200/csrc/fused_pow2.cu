// fused_pow2.cu -- fused Z_q CRT / CRT^-1 for power-of-two index m = 2^e (n = 2^(e-1) <= 32768), the shape of
// BASELINE.json config B (m = 2^16 over four ~30-bit primes).  One CTA per (ring element, RNS limb): the limb lives
// in shared memory as u32 (128 KB at n = 32768), one HBM read and one HBM write per coefficient.
//
// Operator (crt.cpp:518-538 with p = 2): y[i] *= w^{rev(i)} (crtTwiddle, crt.cpp:43-58), then e-1 radix-2 rounds
// r = 0 .. e-2 of { butterfly at stride 2^r (crt.cpp:137-149) ; positions with bit r set *= w^{rev(i0) 2^(r+1)},
// i0 = pos >> (r+1) (dftTwiddle, crt.cpp:92-106) }.  The inverse runs the rounds backwards with inverse roots
// (crt.cpp:488-516), then the inverse crtTwiddle and mhat^-1 (crt.cpp:573-579), folded into one table here.
//
// Schedule: rounds are grouped into passes of S <= 5; in a pass a thread owns the 2^S coefficients that differ in
// bits [r, r+S) and runs S rounds in registers.  Shared memory is XOR-swizzled (word ^ ((word >> 5) & 31)) so the
// stride-1, stride-32 and stride-1024 accesses of the three passes are all bank-conflict free.  The last forward
// pass stores straight to HBM, the first inverse pass loads straight from HBM.
//
// Arithmetic (4q < 2^32): lazy residues in [0,2q); twiddles in Montgomery form w * 2^32 mod q, so one
// multiplication is IMAD.WIDE, IMAD, IMAD.WIDE with no companion table; canonical [0,q) only at the store.
#include <cstdlib>

#include "fused.cuh"
#include "numtheory.h"

namespace lolb {

namespace {

struct Pow2Limb {
  uint32_t q, q2, qinv;          // qinv = -q^-1 mod 2^32
  const uint32_t* crt_tw;        // [n]  fwd: mont(w^rev(i));  inv: mont(w^-rev(i) * mhat^-1)
  const uint32_t* round_tw;      // rounds 0 .. e-2 back to back: round r has 2^(e-2-r) entries mont(w^{+-rev(i0) 2^(r+1)})
};

struct Pow2Params {
  int32_t e, n, k;
  int32_t npass;
  int32_t pass_r[8], pass_s[8];  // forward order; the inverse walks it backwards
  int32_t round_off[32];         // offset of round r inside round_tw
  Pow2Limb limb[kMaxLimbs];
};

struct Mont {
  uint32_t q, q2, qinv;
  // x any u32, w < q  ->  x*w*2^-32 mod q  in [0, 2q)
  __device__ __forceinline__ uint32_t mul(uint32_t x, uint32_t w) const
  {
    const uint64_t p = (uint64_t)x * w;
    const uint32_t m = (uint32_t)p * qinv;
    return (uint32_t)((p + (uint64_t)m * q) >> 32);
  }
  __device__ __forceinline__ uint32_t fold(uint32_t x) const { return min(x, x - q2); }
  __device__ __forceinline__ uint32_t canon(uint32_t x) const { return min(x, x - q); }
};

__device__ __forceinline__ int swz(int pos) { return pos ^ ((pos >> 5) & 31); }

__device__ __noinline__ uint32_t reduce_any64(int64_t x, uint32_t q)
{
  int64_t r = x % (int64_t)q;
  return (uint32_t)(r < 0 ? r + q : r);
}

// S rounds on the 2^S registers of one block.  Twiddle of the pair (j0, j0 | 2^a) in round r + a:
// index (j0 >> (a+1)) + (H << (S-1-a)) of that round's table.
template <int S, bool INV>
__device__ __forceinline__ void rounds_in_regs(uint32_t (&v)[1 << S], const Mont& M, const uint32_t* __restrict__ round_tw,
                                               const int32_t* round_off, int r, int H)
{
  if (!INV) {
#pragma unroll
    for (int a = 0; a < S; a++) {
      const uint32_t* tw = round_tw + round_off[r + a] + (H << (S - 1 - a));
#pragma unroll
      for (int j0 = 0; j0 < (1 << S); j0++) {
        if (j0 & (1 << a)) continue;
        const int j1 = j0 | (1 << a);
        const uint32_t w = __ldg(tw + (j0 >> (a + 1)));
        const uint32_t u = v[j0], t = v[j1];
        v[j0] = M.fold(u + t);
        v[j1] = M.mul(u + M.q2 - t, w);
      }
    }
  } else {
#pragma unroll
    for (int a = S - 1; a >= 0; a--) {
      const uint32_t* tw = round_tw + round_off[r + a] + (H << (S - 1 - a));
#pragma unroll
      for (int j0 = 0; j0 < (1 << S); j0++) {
        if (j0 & (1 << a)) continue;
        const int j1 = j0 | (1 << a);
        const uint32_t w = __ldg(tw + (j0 >> (a + 1)));
        const uint32_t u = v[j0], t = M.mul(v[j1], w);
        v[j0] = M.fold(u + t);
        v[j1] = M.fold(u + M.q2 - t);
      }
    }
  }
}

// one pass over the whole limb held in shared memory.  FROM_GLOBAL / TO_GLOBAL fuse the HBM load (inverse, first
// pass) or store (forward, last pass): those passes have r + S = e - 1, so pos(j) = low + (j << r) with `low` = block
// index and consecutive threads touch consecutive coefficients.
template <int S, bool INV, bool FROM_GLOBAL, bool TO_GLOBAL>
__device__ __forceinline__ void run_pass(uint32_t* sm, int64_t* gbase, int k, const Pow2Params& P, const Pow2Limb& L, const Mont& M, int r)
{
  const int n = P.n;
  const int blocks = n >> S;
  for (int b = threadIdx.x; b < blocks; b += blockDim.x) {
    const int low = b & ((1 << r) - 1), H = b >> r;
    const int base = low + (H << (r + S));
    uint32_t v[1 << S];
    if (FROM_GLOBAL) {
      uint32_t hi_or = 0, lo_max = 0;
#pragma unroll
      for (int j = 0; j < (1 << S); j++) {
        const int64_t raw = __ldcs(gbase + (size_t)(base + (j << r)) * k);
        v[j] = (uint32_t)raw;
        hi_or |= (uint32_t)((uint64_t)raw >> 32);
        lo_max = max(lo_max, v[j]);
      }
      if (hi_or != 0 || lo_max >= L.q) {
#pragma unroll 1
        for (int j = 0; j < (1 << S); j++) v[j] = reduce_any64(gbase[(size_t)(base + (j << r)) * k], L.q);
      }
    } else {
#pragma unroll
      for (int j = 0; j < (1 << S); j++) v[j] = sm[swz(base + (j << r))];
    }
    rounds_in_regs<S, INV>(v, M, L.round_tw, P.round_off, r, H);
    if (TO_GLOBAL) {
#pragma unroll
      for (int j = 0; j < (1 << S); j++) __stcs(gbase + (size_t)(base + (j << r)) * k, (int64_t)M.canon(v[j]));
    } else {
#pragma unroll
      for (int j = 0; j < (1 << S); j++) sm[swz(base + (j << r))] = v[j];
    }
  }
}

template <bool INV, bool FROM_GLOBAL, bool TO_GLOBAL>
__device__ __forceinline__ void run_pass_s(int S, uint32_t* sm, int64_t* gbase, int k, const Pow2Params& P, const Pow2Limb& L, const Mont& M, int r)
{
  switch (S) {
    case 5: run_pass<5, INV, FROM_GLOBAL, TO_GLOBAL>(sm, gbase, k, P, L, M, r); break;
    case 4: run_pass<4, INV, FROM_GLOBAL, TO_GLOBAL>(sm, gbase, k, P, L, M, r); break;
    default: run_pass<3, INV, FROM_GLOBAL, TO_GLOBAL>(sm, gbase, k, P, L, M, r); break;
  }
}

template <bool INV>
__global__ void __launch_bounds__(1024, 1)
k_pow2(int64_t* __restrict__ y, int64_t batch, const __grid_constant__ Pow2Params P)
{
  extern __shared__ __align__(16) uint32_t sm[];
  const int n = P.n, k = P.k;
  const int64_t items = batch * k;
  for (int64_t w = blockIdx.x; w < items; w += gridDim.x) {
    const int64_t el = w / k;
    const int limb = (int)(w - el * k);
    const Pow2Limb& L = P.limb[limb];
    const Mont M{L.q, L.q2, L.qinv};
    int64_t* gbase = y + (size_t)el * n * k + limb;
    if (!INV) {
      // load + crtTwiddle, coalesced, 8 loads in flight per thread
      for (int i0 = threadIdx.x; i0 < n; i0 += blockDim.x * 8) {
        int64_t raw[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int pos = i0 + u * blockDim.x;
          raw[u] = pos < n ? __ldcs(gbase + (size_t)pos * k) : 0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int pos = i0 + u * blockDim.x;
          if (pos < n) {
            uint32_t x = (uint64_t)raw[u] < (uint64_t)L.q ? (uint32_t)raw[u] : reduce_any64(raw[u], L.q);
            sm[swz(pos)] = M.mul(x, __ldg(L.crt_tw + pos));
          }
        }
      }
      __syncthreads();
      for (int p = 0; p < P.npass; p++) {
        if (p == P.npass - 1) run_pass_s<false, false, true>(P.pass_s[p], sm, gbase, k, P, L, M, P.pass_r[p]);
        else run_pass_s<false, false, false>(P.pass_s[p], sm, gbase, k, P, L, M, P.pass_r[p]);
        __syncthreads();
      }
    } else {
      for (int p = P.npass - 1; p >= 0; p--) {
        if (p == P.npass - 1) run_pass_s<true, true, false>(P.pass_s[p], sm, gbase, k, P, L, M, P.pass_r[p]);
        else run_pass_s<true, false, false>(P.pass_s[p], sm, gbase, k, P, L, M, P.pass_r[p]);
        __syncthreads();
      }
      // inverse crtTwiddle * mhat^-1, canonical store, coalesced
      for (int pos = threadIdx.x; pos < n; pos += blockDim.x)
        __stcs(gbase + (size_t)pos * k, (int64_t)M.canon(M.mul(sm[swz(pos)], __ldg(L.crt_tw + pos))));
      __syncthreads();
    }
  }
}

// ---- compile-time specialisation for E = 16 (n = 32768, passes (0,5) (5,5) (10,5)): pass geometry, swizzle and
// table offsets fold into immediates; 512 threads x 2 blocks per pass keeps 2 x 32 coefficients in 128 registers.
constexpr int kE16 = 16, kN16 = 1 << 15, kT16 = 512;
__host__ __device__ constexpr int round_off16(int r) { return (1 << 14) * 2 - (1 << (15 - r)); }   // sum_{r'<r} 2^(14-r')

// the 31 twiddles of one block: round R+a needs entries [H << (4-a), (H+1) << (4-a)) of its table -- contiguous, so
// they are fetched with 128/64/32-bit loads up front (9 instructions) instead of one dependent load per butterfly
__device__ __forceinline__ void load_tw16(uint32_t (&tw)[31], const uint32_t* __restrict__ round_tw, int R, int H)
{
  const uint32_t* t0 = round_tw + round_off16(R) + (H << 4);
  const uint32_t* t1 = round_tw + round_off16(R + 1) + (H << 3);
  const uint32_t* t2 = round_tw + round_off16(R + 2) + (H << 2);
  const uint32_t* t3 = round_tw + round_off16(R + 3) + (H << 1);
  const uint32_t* t4 = round_tw + round_off16(R + 4) + H;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(t0) + i);
    tw[4 * i] = x.x; tw[4 * i + 1] = x.y; tw[4 * i + 2] = x.z; tw[4 * i + 3] = x.w;
  }
#pragma unroll
  for (int i = 0; i < 2; i++) {
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(t1) + i);
    tw[16 + 4 * i] = x.x; tw[16 + 4 * i + 1] = x.y; tw[16 + 4 * i + 2] = x.z; tw[16 + 4 * i + 3] = x.w;
  }
  {
    const uint4 x = __ldg(reinterpret_cast<const uint4*>(t2));
    tw[24] = x.x; tw[25] = x.y; tw[26] = x.z; tw[27] = x.w;
    const uint2 z = __ldg(reinterpret_cast<const uint2*>(t3));
    tw[28] = z.x; tw[29] = z.y;
    tw[30] = __ldg(t4);
  }
}

// H0: the block has H = 0 (the pass on bits [10,15)), so the twiddle index of a pair is just j0 >> (a+1), known at
// compile time; index 0 is the twiddle 1 (crt.cpp:92-106 starts at i0 = 1) and those 31 of 80 multiplications are dropped.
template <bool INV, bool H0 = false>
__device__ __forceinline__ void rounds16(uint32_t (&v)[32], const Mont& M, const uint32_t (&tw)[31])
{
#pragma unroll
  for (int aa = 0; aa < 5; aa++) {
    const int a = INV ? 4 - aa : aa;
    const int toff = 32 - (32 >> a);          // 0, 16, 24, 28, 30
#pragma unroll
    for (int j0 = 0; j0 < 32; j0++) {
      if (j0 & (1 << a)) continue;
      const int j1 = j0 | (1 << a);
      const uint32_t w = tw[toff + (j0 >> (a + 1))];
      const bool trivial = H0 && (j0 >> (a + 1)) == 0;
      if (trivial) {
        const uint32_t u = v[j0], t = v[j1];
        v[j0] = M.fold(u + t);
        v[j1] = M.fold(u + M.q2 - t);
      } else if (!INV) {
        const uint32_t u = v[j0], t = v[j1];
        v[j0] = M.fold(u + t);
        v[j1] = M.mul(u + M.q2 - t, w);
      } else {
        const uint32_t u = v[j0], t = M.mul(v[j1], w);
        v[j0] = M.fold(u + t);
        v[j1] = M.fold(u + M.q2 - t);
      }
    }
  }
}

// physical (swizzled) shared-memory word of coefficient j of block (low, H) in the pass starting at bit R
template <int R>
__device__ __forceinline__ int phys16(int low, int H, int j)
{
  if (R == 0) return (j ^ (H & 31)) + (H << 5);                       // pos = j + 32 H
  if (R == 5) return (low ^ j) + (j << 5) + (H << 10);                // pos = low + 32 j + 1024 H
  return (low ^ ((low >> 5) & 31)) + (j << 10);                       // pos = low + 1024 j
}

// one block (32 coefficients) of the pass starting at bit R, shared memory -> registers -> shared memory
template <int R, bool INV>
__device__ __forceinline__ void block16_smem(uint32_t* sm, const Pow2Limb& L, const Mont& M, int low, int H)
{
  uint32_t v[32], tw[31];
  load_tw16(tw, L.round_tw, R, H);
#pragma unroll
  for (int j = 0; j < 32; j++) v[j] = sm[phys16<R>(low, H, j)];
  rounds16<INV>(v, M, tw);
#pragma unroll
  for (int j = 0; j < 32; j++) sm[phys16<R>(low, H, j)] = v[j];
}

// Dependency structure used below: the pass on bits [0,5) works inside aligned 32-blocks and the pass on bits
// [5,10) inside aligned 1024-chunks, so a warp that owns whole 1024-chunks (load/store, pass 0 and pass 1 of that
// chunk) needs only __syncwarp between them.  Only the pass on bits [10,15) couples all chunks: two CTA barriers
// per ring element, and warps drift apart so that one warp's HBM traffic overlaps another warp's arithmetic.
template <bool INV, int K>
__global__ void __launch_bounds__(kT16, 1)
k_pow2_e16(int64_t* __restrict__ y, int64_t batch, const __grid_constant__ Pow2Params P)
{
  extern __shared__ __align__(16) uint32_t sm[];
  const int k = K ? K : P.k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kT16 / 32, kChunks = kN16 >> 10;
  const int64_t items = batch * k;
  for (int64_t w = blockIdx.x; w < items; w += gridDim.x) {
    const int64_t el = w / k;
    const int limb = (int)(w - el * k);
    const Pow2Limb& L = P.limb[limb];
    const Mont M{L.q, L.q2, L.qinv};
    int64_t* gbase = y + (size_t)el * kN16 * k + limb;
    if (!INV) {
#pragma unroll 1
      for (int H = warp; H < kChunks; H += kWarps) {
        // load chunk H (coalesced) with crtTwiddle
        {
          // all 32 loads of the chunk in flight at once (one HBM latency per chunk instead of four)
          int64_t raw[32];
          uint32_t hi_or = 0, lo_max = 0;
#pragma unroll
          for (int u = 0; u < 32; u++) raw[u] = __ldcs(gbase + (size_t)((H << 10) + (u << 5) + lane) * k);
#pragma unroll
          for (int u = 0; u < 32; u++) { hi_or |= (uint32_t)((uint64_t)raw[u] >> 32); lo_max = max(lo_max, (uint32_t)raw[u]); }
          const bool odd_input = hi_or != 0 || lo_max >= L.q;
#pragma unroll
          for (int u = 0; u < 32; u++) {
            const int pos = (H << 10) + (u << 5) + lane;
            const uint32_t x = odd_input ? reduce_any64(raw[u], L.q) : (uint32_t)raw[u];
            sm[swz(pos)] = M.mul(x, __ldg(L.crt_tw + pos));
          }
        }
        __syncwarp();
        block16_smem<0, false>(sm, L, M, 0, (H << 5) + lane);
        __syncwarp();
        block16_smem<5, false>(sm, L, M, lane, H);
      }
      __syncthreads();
      // bits [10,15): block b = coefficients b + 1024 j, straight to HBM
#pragma unroll 1
      for (int b = threadIdx.x; b < 1024; b += kT16) {
        uint32_t v[32], tw[31];
        load_tw16(tw, L.round_tw, 10, 0);
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = sm[phys16<10>(b, 0, j)];
        rounds16<false, true>(v, M, tw);
#pragma unroll
        for (int j = 0; j < 32; j++) __stcs(gbase + (size_t)(b + (j << 10)) * k, (int64_t)M.canon(v[j]));
      }
      __syncthreads();
    } else {
#pragma unroll 1
      for (int b = threadIdx.x; b < 1024; b += kT16) {
        uint32_t v[32], tw[31];
        load_tw16(tw, L.round_tw, 10, 0);
        uint32_t hi_or = 0, lo_max = 0;
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const int64_t raw = __ldcs(gbase + (size_t)(b + (j << 10)) * k);
          v[j] = (uint32_t)raw;
          hi_or |= (uint32_t)((uint64_t)raw >> 32);
          lo_max = max(lo_max, v[j]);
        }
        if (hi_or != 0 || lo_max >= L.q) {
#pragma unroll 1
          for (int j = 0; j < 32; j++) v[j] = reduce_any64(gbase[(size_t)(b + (j << 10)) * k], L.q);
        }
        rounds16<true, true>(v, M, tw);
#pragma unroll
        for (int j = 0; j < 32; j++) sm[phys16<10>(b, 0, j)] = v[j];
      }
      __syncthreads();
#pragma unroll 1
      for (int H = warp; H < kChunks; H += kWarps) {
        block16_smem<5, true>(sm, L, M, lane, H);
        __syncwarp();
        block16_smem<0, true>(sm, L, M, 0, (H << 5) + lane);
        __syncwarp();
        // inverse crtTwiddle * mhat^-1, canonical store of chunk H
#pragma unroll 8
        for (int i = 0; i < 32; i++) {
          const int pos = (H << 10) + (i << 5) + lane;
          __stcs(gbase + (size_t)pos * k, (int64_t)M.canon(M.mul(sm[swz(pos)], __ldg(L.crt_tw + pos))));
        }
      }
      __syncthreads();
    }
  }
}

// ---- tupSize > 1 at E = 16: one CTA per ring ELEMENT.  The RNS limbs of a coefficient are interleaved in the ABI
// layout (32 bytes per coefficient at k = 4), so a per-limb CTA touches every sector of the element for a quarter of
// its bytes.  Here the element is read ONCE with full-width coalesced loads and split into a per-CTA scratch
// [k][n] of u32 in global memory (512 KB per CTA at k = 4: it lives in the 126 MB L2, never meant to reach HBM);
// each limb is then transformed from / to its contiguous u32 row, and the element is re-interleaved and written
// ONCE with full-width stores.  HBM sees exactly one fully coalesced read and write of the element.
template <bool INV>
__device__ __forceinline__ void limb16_from_scratch(uint32_t* sm, uint32_t* row, const Pow2Limb& L, const Mont& M, int lane, int warp)
{
  constexpr int kWarps = kT16 / 32, kChunks = kN16 >> 10;
  if (!INV) {
#pragma unroll 1
    for (int H = warp; H < kChunks; H += kWarps) {
#pragma unroll 8
      for (int i = 0; i < 32; i++) {
        const int pos = (H << 10) + (i << 5) + lane;
        sm[swz(pos)] = M.mul(row[pos], __ldg(L.crt_tw + pos));
      }
      __syncwarp();
      block16_smem<0, false>(sm, L, M, 0, (H << 5) + lane);
      __syncwarp();
      block16_smem<5, false>(sm, L, M, lane, H);
    }
    __syncthreads();
#pragma unroll 1
    for (int b = threadIdx.x; b < 1024; b += kT16) {
      uint32_t v[32], tw[31];
      load_tw16(tw, L.round_tw, 10, 0);
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] = sm[phys16<10>(b, 0, j)];
      rounds16<false, true>(v, M, tw);
#pragma unroll
      for (int j = 0; j < 32; j++) row[b + (j << 10)] = M.canon(v[j]);
    }
    __syncthreads();
  } else {
#pragma unroll 1
    for (int b = threadIdx.x; b < 1024; b += kT16) {
      uint32_t v[32], tw[31];
      load_tw16(tw, L.round_tw, 10, 0);
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] = row[b + (j << 10)];
      rounds16<true, true>(v, M, tw);
#pragma unroll
      for (int j = 0; j < 32; j++) sm[phys16<10>(b, 0, j)] = v[j];
    }
    __syncthreads();
#pragma unroll 1
    for (int H = warp; H < kChunks; H += kWarps) {
      block16_smem<5, true>(sm, L, M, lane, H);
      __syncwarp();
      block16_smem<0, true>(sm, L, M, 0, (H << 5) + lane);
      __syncwarp();
#pragma unroll 8
      for (int i = 0; i < 32; i++) {
        const int pos = (H << 10) + (i << 5) + lane;
        row[pos] = M.canon(M.mul(sm[swz(pos)], __ldg(L.crt_tw + pos)));
      }
    }
    __syncthreads();
  }
}

template <bool INV, int K>
__global__ void __launch_bounds__(kT16, 1)
k_pow2_e16_elem(int64_t* __restrict__ y, int64_t batch, const __grid_constant__ Pow2Params P, uint32_t* __restrict__ scratch)
{
  extern __shared__ __align__(16) uint32_t sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* mine = scratch + (size_t)blockIdx.x * K * kN16;
  for (int64_t el = blockIdx.x; el < batch; el += gridDim.x) {
    int64_t* ebase = y + (size_t)el * kN16 * K;
    // split: coefficient `pos` -> scratch[l][pos]
#pragma unroll 2
    for (int pos = threadIdx.x; pos < kN16; pos += kT16) {
      int64_t c[K];
      if (K % 2 == 0) {
#pragma unroll
        for (int h = 0; h < K / 2; h++) {
          const longlong2 raw = __ldcs(reinterpret_cast<const longlong2*>(ebase + (size_t)pos * K) + h);
          c[2 * h] = raw.x; c[2 * h + 1] = raw.y;
        }
      } else {
#pragma unroll
        for (int l = 0; l < K; l++) c[l] = __ldcs(ebase + (size_t)pos * K + l);
      }
#pragma unroll
      for (int l = 0; l < K; l++) {
        const uint32_t q = P.limb[l].q;
        mine[(size_t)l * kN16 + pos] = (uint64_t)c[l] < (uint64_t)q ? (uint32_t)c[l] : reduce_any64(c[l], q);
      }
    }
    __syncthreads();
#pragma unroll 1
    for (int l = 0; l < K; l++) {
      const Pow2Limb& L = P.limb[l];
      const Mont M{L.q, L.q2, L.qinv};
      limb16_from_scratch<INV>(sm, mine + (size_t)l * kN16, L, M, lane, warp);
    }
    // merge: scratch[l][pos] -> element
#pragma unroll 2
    for (int pos = threadIdx.x; pos < kN16; pos += kT16) {
      int64_t c[K];
#pragma unroll
      for (int l = 0; l < K; l++) c[l] = (int64_t)mine[(size_t)l * kN16 + pos];
      if (K % 2 == 0) {
#pragma unroll
        for (int h = 0; h < K / 2; h++)
          __stcs(reinterpret_cast<longlong2*>(ebase + (size_t)pos * K) + h, make_longlong2(c[2 * h], c[2 * h + 1]));
      } else {
#pragma unroll
        for (int l = 0; l < K; l++) __stcs(ebase + (size_t)pos * K + l, c[l]);
      }
    }
    __syncthreads();
  }
}

struct FusedPow2 {
  bool ok_fwd = false, ok_inv = false;
  Pow2Params fwd{}, inv{};
  uint32_t* d_tab = nullptr;       // all tables of all limbs, both directions
  int threads = 1024;
};

bool shape_is_pow2(const lolb_plan* pl)
{
  if (pl->kind != PLAN_RQ || pl->pe.size() != 1 || pl->pe[0].prime != 2) return false;
  const int e = pl->pe[0].exponent;
  if (e < 7 || e > 16) return false;                    // n = 64 .. 32768 (<= 128 KB of shared memory)
  for (int64_t q : pl->qs) if (!(q & 1) || 4 * (uint64_t)q >= ((uint64_t)1 << 32)) return false;
  return true;
}

uint32_t neg_inv32(uint32_t q)
{
  uint32_t inv = q;
  for (int i = 0; i < 5; i++) inv *= 2u - q * inv;
  return 0u - inv;
}

}  // namespace

int fused_pow2_select(lolb_plan* pl, void** slot)
{
  if (!shape_is_pow2(pl)) return LOLB_OK;
  FusedPow2* F = (FusedPow2*)*slot;
  if (!F) { F = new FusedPow2(); *slot = F; }
  const int e = pl->pe[0].exponent, n = pl->n, k = pl->k, rounds = e - 1;
  F->ok_fwd = pl->has_fwd && pl->ru.size() == 1;
  F->ok_inv = pl->has_inv && pl->ruinv.size() == 1 && (int)pl->mhatinv.size() == k;
  Pow2Params P{};
  P.e = e; P.n = n; P.k = k;
  // passes of at most 5 rounds, as even as possible, at least 3 rounds each (template instances 3, 4, 5)
  int npass = (rounds + 4) / 5;
  if (rounds / npass < 3 && npass > 1) npass--;
  P.npass = npass;
  for (int p = 0, r = 0; p < npass; p++) {
    int s = rounds / npass + (p < rounds % npass ? 1 : 0);
    P.pass_r[p] = r; P.pass_s[p] = s; r += s;
  }
  if (rounds < 3 || P.pass_s[npass - 1] < 3 || P.pass_s[0] > 5) { F->ok_fwd = F->ok_inv = false; return LOLB_OK; }
  int32_t off = 0;
  for (int r = 0; r < rounds; r++) { P.round_off[r] = off; off += 1 << (e - 2 - r); }
  const size_t per_dir = (((size_t)n + (size_t)off) + 3) & ~(size_t)3;   // crt table + round tables, 16-byte multiple
  std::vector<uint32_t> host((size_t)k * 2 * per_dir, 0u);
  const int64_t m = pl->m;
  for (int t = 0; t < k; t++) {
    const uint64_t q = (uint64_t)pl->qs[t];
    auto mont = [&](uint64_t c) { return (uint32_t)(((c % q) << 32) % q); };
    for (int dir = 0; dir < 2; dir++) {
      if (dir == 0 ? !F->ok_fwd : !F->ok_inv) continue;
      const std::vector<int64_t>& T = dir == 0 ? pl->ru[0] : pl->ruinv[0];
      auto root = [&](int64_t j) { int64_t v = T[(size_t)(j % m) * k + t] % (int64_t)q; return (uint64_t)(v < 0 ? v + (int64_t)q : v); };
      uint32_t* crt = host.data() + ((size_t)t * 2 + dir) * per_dir;
      uint32_t* rnd = crt + n;
      const uint64_t scale = dir == 1 ? (uint64_t)(((pl->mhatinv[t] % (int64_t)q) + (int64_t)q) % (int64_t)q) : 1;
      for (int i = 0; i < n; i++) {
        const uint64_t w = i ? root(digit_rev(2, e - 1, i)) : 1;          // crt.cpp:43-58
        crt[i] = mont(mulmod64(w, scale, q));
      }
      for (int r = 0; r < rounds; r++)
        for (int i0 = 0; i0 < (1 << (e - 2 - r)); i0++)                   // crt.cpp:92-106, twidRuStride = 2^(r+1)
          rnd[P.round_off[r] + i0] = mont(i0 ? root(digit_rev(2, e - 2 - r, i0) * ((int64_t)2 << r)) : 1);
    }
  }
  if (F->d_tab) { cudaFree(F->d_tab); F->d_tab = nullptr; }
  LOLB_CUDA(cudaMalloc((void**)&F->d_tab, host.size() * sizeof(uint32_t)));
  LOLB_CUDA(cudaMemcpy(F->d_tab, host.data(), host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  F->fwd = P; F->inv = P;
  for (int t = 0; t < k; t++) {
    const uint32_t q = (uint32_t)pl->qs[t];
    for (int dir = 0; dir < 2; dir++) {
      Pow2Limb& L = (dir == 0 ? F->fwd : F->inv).limb[t];
      L.q = q; L.q2 = 2 * q; L.qinv = neg_inv32(q);
      L.crt_tw = F->d_tab + ((size_t)t * 2 + dir) * per_dir;
      L.round_tw = L.crt_tw + n;
    }
  }
  // one block of 2^S coefficients per thread in the widest pass, at most 1024 threads
  int blocks = n >> 5;
  for (int p = 0; p < npass; p++) if ((n >> P.pass_s[p]) > blocks) blocks = n >> P.pass_s[p];
  F->threads = blocks > 1024 ? 1024 : (blocks < 32 ? 32 : blocks);
  return LOLB_OK;
}

void fused_pow2_release(void* slot)
{
  FusedPow2* F = (FusedPow2*)slot;
  if (!F) return;
  if (F->d_tab) cudaFree(F->d_tab);
  delete F;
}

bool fused_pow2_available(const void* slot, bool inverse)
{
  const FusedPow2* F = (const FusedPow2*)slot;
  return F && (inverse ? F->ok_inv : F->ok_fwd);
}

int fused_pow2_crt(const lolb_plan* pl, const void* slot, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  const FusedPow2* F = (const FusedPow2*)slot;
  if (!fused_pow2_available(slot, inverse)) return LOLB_FUSED_UNAVAILABLE;
  if (batch <= 0) return LOLB_OK;
  const size_t smem = (size_t)pl->n * sizeof(uint32_t);
  static PerDeviceOnce once[2];
  if (smem > 48 * 1024 && once[inverse].first()) {
    cudaError_t e = inverse ? cudaFuncSetAttribute(k_pow2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                            : cudaFuncSetAttribute(k_pow2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_pow2)");
  }
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  const int by_threads = 2048 / F->threads;
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 16) per_sm = 16;
  int64_t grid = (int64_t)pl->num_sms * per_sm;
  const int64_t items = batch * pl->k;
  if (grid > items) grid = items;
  if (pl->pe[0].exponent == kE16 && !getenv("LOLB_POW2_GENERIC")) {
    static PerDeviceOnce once16;
    if (once16.first()) {
      cudaFuncSetAttribute(k_pow2_e16<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(k_pow2_e16<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(k_pow2_e16<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(k_pow2_e16<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(k_pow2_e16<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(k_pow2_e16<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(k_pow2_e16<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      cudaFuncSetAttribute(k_pow2_e16<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    }
    if (pl->k == 1) {
      if (inverse) k_pow2_e16<true, 1><<<(int)grid, kT16, smem, st>>>(y, batch, F->inv);
      else k_pow2_e16<false, 1><<<(int)grid, kT16, smem, st>>>(y, batch, F->fwd);
    } else if (!getenv("LOLB_POW2_PER_LIMB") && ((pl->k == 4 && !inverse) || ((pl->k == 2 || pl->k == 3 || pl->k == 4) && getenv("LOLB_POW2_ELEM")))) {
      // element-per-CTA with an L2 scratch: measured faster only for the forward transform at tupSize 4
      // (28.7 % vs 24.8 % of HBM peak; inverse 28.3 % vs 30.8 %); part of the scratch is written back to HBM
      int64_t g2 = pl->num_sms;
      if (g2 > batch) g2 = batch;
      uint32_t* scratch = (uint32_t*)plan_ws(pl, st, (size_t)g2 * pl->k * kN16 * sizeof(uint32_t));
      if (!scratch) return LOLB_ERR_CUDA;
      static PerDeviceOnce once_el;
      if (once_el.first()) {
        cudaFuncSetAttribute(k_pow2_e16_elem<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_pow2_e16_elem<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_pow2_e16_elem<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_pow2_e16_elem<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_pow2_e16_elem<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_pow2_e16_elem<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      }
#define LE(KK) (inverse ? k_pow2_e16_elem<true, KK><<<(int)g2, kT16, smem, st>>>(y, batch, F->inv, scratch) \
                        : k_pow2_e16_elem<false, KK><<<(int)g2, kT16, smem, st>>>(y, batch, F->fwd, scratch))
      if (pl->k == 2) LE(2); else if (pl->k == 3) LE(3); else LE(4);
#undef LE
    } else if (pl->k == 2) {
      if (inverse) k_pow2_e16<true, 2><<<(int)grid, kT16, smem, st>>>(y, batch, F->inv);
      else k_pow2_e16<false, 2><<<(int)grid, kT16, smem, st>>>(y, batch, F->fwd);
    } else if (pl->k == 4) {
      if (inverse) k_pow2_e16<true, 4><<<(int)grid, kT16, smem, st>>>(y, batch, F->inv);
      else k_pow2_e16<false, 4><<<(int)grid, kT16, smem, st>>>(y, batch, F->fwd);
    } else {
      if (inverse) k_pow2_e16<true, 0><<<(int)grid, kT16, smem, st>>>(y, batch, F->inv);
      else k_pow2_e16<false, 0><<<(int)grid, kT16, smem, st>>>(y, batch, F->fwd);
    }
  } else if (inverse) k_pow2<true><<<(int)grid, F->threads, smem, st>>>(y, batch, F->inv);
  else k_pow2<false><<<(int)grid, F->threads, smem, st>>>(y, batch, F->fwd);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_pow2");
  count_launch();
  return LOLB_OK;
}

}  // namespace lolb
