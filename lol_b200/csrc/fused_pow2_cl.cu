// fused_pow2_cl.cu -- Z_q CRT / CRT^-1 for m = 2^16 with the whole ring element resident on chip across a thread-block
// cluster: ONE kernel, one HBM round trip (16 bytes per coefficient), no intermediate in global memory.
//
// Operator, arithmetic and tables: pow2_common.cuh (crt.cpp:43-58, 92-106, 137-149, 459-538).  Schedule:
//   * a cluster of 2 K CTAs owns one ring element (K = tupSize limbs, n = 2^15 coefficients each); CTA (limb l, half h)
//     keeps the 2^14 residues of limb l at positions pos = 2 i + h as u32 words in 66 KB of shared memory (3 CTAs per SM);
//   * load: every CTA reads 1/(2K) of the element's interleaved [n][K] int64 words (sector-complete 8-byte loads), does
//     round 0 -- the only round that pairs the two halves, pos and pos + 1 -- in registers (forward), and scatters the
//     residues to their owners through distributed shared memory (st.shared::cluster);
//   * rounds 1 .. 14 are local to a CTA: three register passes over its words (5 + 5 + 4 rounds), twiddles from the plan's
//     tables (entry (2^r - 1) + p, p = pos mod 2^r = 2 (i mod 2^(r-1)) + h);
//   * store: every CTA gathers 1/(2K) of the element from its owners (ld.shared::cluster), does round 0 and the mhat^-1
//     scaling (inverse), and writes canonical int64 words.
// Cluster barriers: after the scatter, after the passes, after the gather (a CTA must not exit while its shared memory is
// being read).  No CTA waits on global memory state; clusters of different elements overlap on an SM.
#include <cooperative_groups.h>

#include <cstdlib>

#include "pow2_common.cuh"

namespace cg = cooperative_groups;

namespace lolb {
namespace pow2 {

namespace {

constexpr int kClThreads = 256;
constexpr int kClWords = 16384;                       // residues per CTA
constexpr int kClSmemWords = kClWords + kClWords / 32;      // + 1 word per 32: conflict-free in all three passes
constexpr int kClLogN = 15;

__device__ __forceinline__ int cl_pad(int i) { return i + (i >> 5); }

template <bool INV>
__device__ __forceinline__ void cl_passes(uint32_t* S, const DfLimb& L, const int h, const int tid)
{
  const Mont M{L.q, L.q2, L.qinv};
  const int lane = tid & 31;
  auto pass1 = [&]() {      // i bits 0-4 (rounds 1-5): task t owns words 32 t .. 32 t + 31; twiddles uniform
#pragma unroll 1
    for (int t = tid; t < kClWords / 32; t += kClThreads) {
      uint32_t* base = S + 33 * t;
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] = base[j];
      const uint32_t* tw = L.tw + h;
      if (!INV) ct_rounds<5, false>(v, M, [&](int a, int jj) { return __ldg(tw + ((2 << a) - 1 + 2 * jj)); });
      else gs_rounds<5, 0>(v, M, [&](int a, int jj) { return __ldg(tw + ((2 << a) - 1 + 2 * jj)); });
#pragma unroll
      for (int j = 0; j < 32; j++) base[j] = v[j];
    }
  };
  auto pass2 = [&]() {      // i bits 5-9 (rounds 6-10): task (hi = i >> 10, lane = i & 31)
#pragma unroll 1
    for (int t = tid; t < kClWords / 32; t += kClThreads) {
      uint32_t* base = S + 1056 * (t >> 5) + lane;
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] = base[33 * j];
      const uint32_t* tw = L.tw + 2 * lane + h;
      if (!INV) ct_rounds<5, false>(v, M, [&](int a, int jj) { return __ldg(tw + ((64 << a) - 1 + 64 * jj)); });
      else gs_rounds<5, 0>(v, M, [&](int a, int jj) { return __ldg(tw + ((64 << a) - 1 + 64 * jj)); });
#pragma unroll
      for (int j = 0; j < 32; j++) base[33 * j] = v[j];
    }
  };
  auto pass3 = [&]() {      // i bits 10-13 (rounds 11-14): task t = i & 1023, 16 words at stride 1024
#pragma unroll 1
    for (int t = tid; t < 1024; t += kClThreads) {
      uint32_t* base = S + t + (t >> 5);
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; j++) v[j] = base[1056 * j];
      const uint32_t* tw = L.tw + 2 * t + h;
      if (!INV) {
        ct_rounds<4, false>(v, M, [&](int a, int jj) { return __ldg(tw + ((2048 << a) - 1 + 2048 * jj)); });
#pragma unroll
        for (int j = 0; j < 16; j++) base[1056 * j] = M.canon(M.fold(v[j]));
      } else {
        gs_rounds<4, 0>(v, M, [&](int a, int jj) { return __ldg(tw + ((2048 << a) - 1 + 2048 * jj)); });
#pragma unroll
        for (int j = 0; j < 16; j++) base[1056 * j] = v[j];
      }
    }
  };
  if (!INV) { pass1(); __syncthreads(); pass2(); __syncthreads(); pass3(); }
  else { pass3(); __syncthreads(); pass2(); __syncthreads(); pass1(); }
}

#ifndef LOLB_CL_MINB
#define LOLB_CL_MINB 3
#endif
#ifndef LOLB_CL_UNROLL
#define LOLB_CL_UNROLL 8
#endif

template <bool INV, int K>
__global__ void __launch_bounds__(kClThreads, LOLB_CL_MINB)
k_pow2_cl(int64_t* __restrict__ y, const __grid_constant__ DfParams P)
{
  constexpr int CL = 2 * K;                             // CTAs per cluster = (limb, half) pairs of one element
  constexpr int n = 1 << kClLogN;
  constexpr int PAIRS = (n / 2) / CL;                   // coefficient pairs (pos, pos + 1) this CTA loads and stores
  constexpr int TASKS = PAIRS * K / kClThreads;         // (pair, limb) tasks per thread
  static_assert(kClThreads % K == 0 && TASKS % LOLB_CL_UNROLL == 0, "geometry");
  extern __shared__ __align__(16) uint32_t S[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int64_t el = blockIdx.x / CL;
  const int tid = threadIdx.x;
  int64_t* ebase = y + (size_t)el * ((size_t)K * n);
  const int my_limb = rank >> 1, my_h = rank & 1;

  // ---- thread-task (pair g, limb l): task index T = tid + 256 it;  l = T % K is fixed per thread
  const int l = tid % K;
  const DfLimb& Ll = P.limb[l];
  const Mont Ml{Ll.q, Ll.q2, Ll.qinv};
#ifdef LOLB_CL_EXP_LOCAL      // timing experiment only (wrong results): no distributed shared memory traffic
  uint32_t* R0 = S;
  uint32_t* R1 = S;
#else
  uint32_t* R0 = cluster.map_shared_rank(S, 2 * l);     // owner of (limb l, even positions)
  uint32_t* R1 = cluster.map_shared_rank(S, 2 * l + 1);
#endif
  const int g0 = rank * PAIRS + tid / K;                // global pair index of task it = 0; + (256 / K) per task
  constexpr int GSTEP = kClThreads / K;
  int64_t* gp = ebase + (size_t)2 * g0 * K + l;         // word of (pos = 2 g, limb l); pos + 1 is K words later

  {
    const uint32_t w0 = Ll.c0[0];                       // forward: T of round 0 (a single twiddle: p = 0)
#pragma unroll 1
    for (int it0 = 0; it0 < TASKS; it0 += LOLB_CL_UNROLL) {
      int64_t ra[LOLB_CL_UNROLL], rb[LOLB_CL_UNROLL];
#pragma unroll
      for (int u = 0; u < LOLB_CL_UNROLL; u++) {
        const int64_t* p = gp + (size_t)(it0 + u) * GSTEP * 2 * K;
        ra[u] = __ldcs(p);
        rb[u] = __ldcs(p + K);
      }
#pragma unroll
      for (int u = 0; u < LOLB_CL_UNROLL; u++) {
        uint32_t a = (uint32_t)ra[u], b = (uint32_t)rb[u];
        if ((((uint64_t)ra[u] | (uint64_t)rb[u]) >> 32) != 0 || a >= Ll.q || b >= Ll.q) {      // outside [0,q): like the reference's c % q
          a = df_reduce_any64(ra[u], Ll.q);
          b = df_reduce_any64(rb[u], Ll.q);
        }
        const int i = cl_pad(g0 + (it0 + u) * GSTEP);
        if (!INV) {
          const uint32_t t = Ml.mul(b, w0);
          R0[i] = a + t;
          R1[i] = a + Ml.q2 - t;
        } else {
          R0[i] = a;
          R1[i] = b;
        }
      }
    }
  }
#ifdef LOLB_CL_EXP_NOSYNC
#define CL_SYNC() __syncthreads()
#else
#define CL_SYNC() cluster.sync()
#endif
  CL_SYNC();
#ifndef LOLB_CL_EXP_NOPASS
  cl_passes<INV>(S, P.limb[my_limb], my_h, tid);
#endif
  CL_SYNC();
  {
    const uint32_t sA = Ll.sA, sB = Ll.sB;
#pragma unroll 1
    for (int it0 = 0; it0 < TASKS; it0 += LOLB_CL_UNROLL) {
      uint32_t a[LOLB_CL_UNROLL], b[LOLB_CL_UNROLL];
#pragma unroll
      for (int u = 0; u < LOLB_CL_UNROLL; u++) {
        const int i = cl_pad(g0 + (it0 + u) * GSTEP);
        a[u] = R0[i];
        b[u] = R1[i];
      }
#pragma unroll
      for (int u = 0; u < LOLB_CL_UNROLL; u++) {
        int64_t* p = gp + (size_t)(it0 + u) * GSTEP * 2 * K;
        if (!INV) {
          __stcs(p, (int64_t)a[u]);
          __stcs(p + K, (int64_t)b[u]);
        } else {      // round 0 of the inverse with mhat^-1 folded in (crt.cpp:573-579)
          __stcs(p, (int64_t)Ml.canon(Ml.mul(a[u] + b[u], sA)));
          __stcs(p + K, (int64_t)Ml.canon(Ml.mul(a[u] + Ml.q2 - b[u], sB)));
        }
      }
    }
  }
  CL_SYNC();      // nobody leaves while its shared memory may still be read
}

template <bool INV, int K>
int launch_cl(const lolb_plan* pl, const FusedPow2Df* F, int64_t* y, int64_t batch, cudaStream_t st)
{
  (void)pl;
  constexpr int CL = 2 * K;
  constexpr size_t smem = (size_t)kClSmemWords * sizeof(uint32_t);
  static PerDeviceOnce once;
  if (once.first()) LOLB_CUDA(cudaFuncSetAttribute(k_pow2_cl<INV, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const DfParams& P = INV ? F->inv : F->fwd;
  const int64_t kMaxEl = 0x7fffffff / CL;
  for (int64_t done = 0; done < batch;) {
    const int64_t cnt = batch - done < kMaxEl ? batch - done : kMaxEl;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cnt * CL), 1, 1);
    cfg.blockDim = dim3(kClThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_pow2_cl<INV, K>, y + (size_t)done * K * (1 << kClLogN), P);
    if (e != cudaSuccess) return cuda_fail(e, "k_pow2_cl");
    count_launch();
    done += cnt;
  }
  return LOLB_OK;
}

}  // namespace

// m = 2^16 (top == 5), tupSize 1 / 2 / 4; LOLB_FUSED_UNAVAILABLE otherwise
int pow2_cluster_crt(const lolb_plan* pl, const FusedPow2Df* F, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  if (F->top != 5) return LOLB_FUSED_UNAVAILABLE;
  if (batch <= 0) return LOLB_OK;
  switch (pl->k) {
    case 1: return inverse ? launch_cl<true, 1>(pl, F, y, batch, st) : launch_cl<false, 1>(pl, F, y, batch, st);
    case 2: return inverse ? launch_cl<true, 2>(pl, F, y, batch, st) : launch_cl<false, 2>(pl, F, y, batch, st);
    case 4: return inverse ? launch_cl<true, 4>(pl, F, y, batch, st) : launch_cl<false, 4>(pl, F, y, batch, st);
  }
  return LOLB_FUSED_UNAVAILABLE;
}

}  // namespace pow2
}  // namespace lolb
