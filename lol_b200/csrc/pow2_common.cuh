// pow2_common.cuh -- shared by the power-of-two CRT kernels (fused_pow2_df.cu: dataflow kernels with the L2 exchange ring;
// fused_pow2_res.cu: warp- and element-resident kernels): per-limb constants, lazy Montgomery arithmetic, and the
// register passes of a 1024-residue unit.
//
// Operator.  For p = 2 the reference's  crtTwiddle ; {dftp ; dftTwiddle} x (e-1)  (crt.cpp:43-58, 137-149, 92-106,
// 459-486, 518-538) evaluates  f(x) = sum_i y[i] x^rev(i)  at  psi^(2 pos + 1),  pos = 0 .. n-1, psi = ru[0][1]
// (pinned on the CPU by tests/test_*_pinning.py::test_pow2_crt_is_negacyclic_evaluation).  Over Z_q every exact
// evaluation order gives the same residues, so the kernels use the twist-free Cooley-Tukey form of the same map:
//   round r = 0 .. e-2, pairs (pos, pos + 2^r) with bit r of pos clear, p = pos mod 2^r:
//     forward   (u, t) -> (u + T t, u - T t),      T = psi^((2p+1) n / 2^(r+1))          rounds ascending
//     inverse   (u, t) -> (u + t, (u - t) / T),    rounds descending, then * mhat^-1      (crt.cpp:488-516, 573-579)
// Arithmetic (odd q, 4q < 2^32): lazy residues in [0,4q) (forward) / [0,2q) (inverse), twiddles in Montgomery form
// (IMAD.WIDE, IMAD, IMAD.HI), fold = one VIADDMNMX, canonical [0,q) only at the final store.
#pragma once
#include "fused.cuh"

namespace lolb {
namespace pow2 {

constexpr int kDfUnit = 1024 + 32 + 8;   // words per (chunk, limb) unit in shared memory: +1 per 32 (padding), +8 (bank shift per unit)
constexpr int kDfMaxK = 4;
constexpr int kDfCtrHead = 16;           // ctr[0] = task counter; per-element counters start here

struct DfLimb {
  uint32_t q, q2, qinv;
  uint32_t sA, sB;         // inverse: mont(mhat^-1), mont(mhat^-1 / T_0)
  uint32_t c0[31];         // Montgomery twiddles of rounds 0..4: entry (2^a - 1) + p
  const uint32_t* tw;      // all rounds: entry (2^r - 1) + p, p < 2^r, Montgomery form
};

struct DfParams {
  int32_t n, k;
  int32_t ring, lag;       // exchange-ring slots; distance (in elements) between the two task kinds in the queue
  DfLimb limb[kDfMaxK];
};

struct Mont {
  uint32_t q, q2, qinv;    // qinv = -q^-1 mod 2^32
  // x any u32, w < q in Montgomery form  ->  x * w mod q  in [0, 2q)
  __device__ __forceinline__ uint32_t mul(uint32_t x, uint32_t w) const
  {
    const uint64_t p = (uint64_t)x * w;
    const uint32_t m = (uint32_t)p * qinv;
    return (uint32_t)((p + (uint64_t)m * q) >> 32);
  }
  __device__ __forceinline__ uint32_t fold(uint32_t x) const { return min(x, x - q2); }    // [0,4q) -> [0,2q)
  __device__ __forceinline__ uint32_t canon(uint32_t x) const { return min(x, x - q); }    // [0,2q) -> [0,q)
};

__device__ __noinline__ static uint32_t df_reduce_any64(int64_t x, uint32_t q)
{
  int64_t r = x % (int64_t)q;
  return (uint32_t)(r < 0 ? r + q : r);
}

// S forward rounds on the 2^S registers of one block; tw(a, jj) = twiddle of the pairs with j0 mod 2^a = jj
template <int S, bool CANON_IN, class TW>
__device__ __forceinline__ void ct_rounds(uint32_t (&v)[1 << S], const Mont& M, TW tw)
{
#pragma unroll
  for (int a = 0; a < S; a++) {
#pragma unroll
    for (int j0 = 0; j0 < (1 << S); j0++) {
      if (j0 & (1 << a)) continue;
      const int j1 = j0 | (1 << a);
      const uint32_t w = tw(a, j0 & ((1 << a) - 1));
      const uint32_t u = (CANON_IN && a == 0) ? v[j0] : M.fold(v[j0]);
      const uint32_t t = M.mul(v[j1], w);
      v[j0] = u + t;
      v[j1] = u + M.q2 - t;
    }
  }
}

// inverse rounds S-1 .. LOW on residues in [0,2q)
template <int S, int LOW, class TW>
__device__ __forceinline__ void gs_rounds(uint32_t (&v)[1 << S], const Mont& M, TW tw)
{
#pragma unroll
  for (int a = S - 1; a >= LOW; a--) {
#pragma unroll
    for (int j0 = 0; j0 < (1 << S); j0++) {
      if (j0 & (1 << a)) continue;
      const int j1 = j0 | (1 << a);
      const uint32_t w = tw(a, j0 & ((1 << a) - 1));
      const uint32_t u = v[j0], t = v[j1];
      v[j0] = M.fold(u + t);
      v[j1] = M.mul(u + M.q2 - t, w);
    }
  }
}

// rounds 0-4 with a run-time limb: one copy of the code, twiddles fetched with LDC
template <bool INV>
__device__ __forceinline__ void unit_rounds_0_4_rt(int limb, uint32_t* Uu, const DfParams& P, int lane)
{
  const DfLimb& L = P.limb[limb];
  const Mont M{L.q, L.q2, L.qinv};
  uint32_t* base = Uu + 33 * lane;
  uint32_t v[32];
#pragma unroll
  for (int j = 0; j < 32; j++) v[j] = base[j];
  if (!INV) {
    ct_rounds<5, true>(v, M, [&](int a, int jj) { return L.c0[(1 << a) - 1 + jj]; });
  } else {
    gs_rounds<5, 1>(v, M, [&](int a, int jj) { return L.c0[(1 << a) - 1 + jj]; });
    const uint32_t sA = L.sA, sB = L.sB;
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += 2) {
      const uint32_t u = v[j0], t = v[j0 + 1];
      v[j0] = M.canon(M.mul(u + t, sA));
      v[j0 + 1] = M.canon(M.mul(u + M.q2 - t, sB));
    }
  }
#pragma unroll
  for (int j = 0; j < 32; j++) base[j] = v[j];
}


// host-side state of the power-of-two kernels of one plan (tables in Montgomery form, both directions)
struct FusedPow2Df {
  bool ok_fwd = false, ok_inv = false;
  DfParams fwd{}, inv{};
  uint32_t* d_tab = nullptr;
  int top = 0;                 // e - 11
};

// fused_pow2_res.cu: m = 2^10 .. 2^14 with the element (or a limb) resident on chip; LOLB_FUSED_UNAVAILABLE when the
// shape has no resident kernel
int pow2_resident_crt(const lolb_plan* pl, const FusedPow2Df* F, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);

// fused_pow2_split.cu: m = 2^14 .. 2^16 as two streaming kernels per sub-batch on two streams (no CTA waits for another)
int pow2_split_crt(const lolb_plan* pl, const FusedPow2Df* F, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);
void pow2_split_release(const lolb_plan* pl);


// fused_pow2_cl.cu: m = 2^16 with the element resident across a thread-block cluster (one kernel, DSMEM exchange)
int pow2_cluster_crt(const lolb_plan* pl, const FusedPow2Df* F, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);

}  // namespace pow2
}  // namespace lolb
