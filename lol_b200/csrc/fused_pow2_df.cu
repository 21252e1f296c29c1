// fused_pow2_df.cu -- Z_q CRT / CRT^-1 for power-of-two index m = 2^e, 13 <= e <= 16, as ONE persistent dataflow kernel
// with an L2-resident exchange ring (BASELINE.json config B: m = 2^16, four ~30-bit primes), plus the plan-level
// selection and dispatch of all power-of-two kernels (operator and arithmetic: pow2_common.cuh; m <= 2^14 with the
// element resident on chip: fused_pow2_res.cu).
//
// Schedule.  The limb (n <= 32768 coefficients, 128 KB as u32) is never resident in one SM.  Position bits [0,10)
// ("chunk" rounds) and bits [10, e-1) ("column" rounds) are two task kinds of 128 threads each:
//   chunk task   one contiguous 32 KB piece of the element (1024 coefficients x K limbs x 4/K chunks): coalesced
//                128-bit loads, limbs de-interleaved through 17 KB of shared memory, rounds 0-4 by the thread
//                that owns 32 consecutive coefficients, rounds 5-9 by the lane that owns stride-32 coefficients
//   column task  128 consecutive (coefficient, limb) pairs x all 2^(e-11) chunks: rounds 10 .. e-2 in registers,
//                no shared memory; every int64 store / load of the element is part of a fully used 32-byte sector
// The two kinds exchange u32 residues through a ring of element slots in global memory that is sized to stay in the
// 126 MB L2 (72 slots = 37 MB at config B), so HBM sees one read and one write of the element.  CTAs are persistent
// and take tasks from an atomic counter; per-element counters order  first kind -> second kind -> slot reuse.  A task
// only waits for tasks with a smaller index, which are already running: no deadlock.
#include <cstdlib>

#include "numtheory.h"
#include "pow2_common.cuh"

namespace lolb {

using namespace pow2;

namespace {

// rounds 0-4 of one unit: lane owns coefficients 32*lane .. 32*lane+31 (padded word 33*lane + j, conflict free).
// LIMB is a template parameter so the 31 twiddles and q, q', 2q are constant-bank operands, not registers.
template <bool INV, int LIMB>
__device__ __forceinline__ void unit_rounds_0_4(uint32_t* Uu, const DfParams& P, int lane)
{
  const DfLimb& L = P.limb[LIMB];
  const Mont M{L.q, L.q2, L.qinv};
  uint32_t* base = Uu + 33 * lane;
  uint32_t v[32];
#pragma unroll
  for (int j = 0; j < 32; j++) v[j] = base[j];
  if (!INV) {
    ct_rounds<5, true>(v, M, [&](int a, int jj) { return L.c0[(1 << a) - 1 + jj]; });
  } else {
    gs_rounds<5, 1>(v, M, [&](int a, int jj) { return L.c0[(1 << a) - 1 + jj]; });
    // round 0 with mhat^-1 folded in (crt.cpp:573-579), then canonical
#pragma unroll
    for (int j0 = 0; j0 < 32; j0 += 2) {
      const uint32_t u = v[j0], t = v[j0 + 1];
      v[j0] = M.canon(M.mul(u + t, L.sA));
      v[j0 + 1] = M.canon(M.mul(u + M.q2 - t, L.sB));
    }
  }
#pragma unroll
  for (int j = 0; j < 32; j++) base[j] = v[j];
}

template <bool INV>
__device__ __forceinline__ void unit_rounds_0_4_any(int limb, uint32_t* Uu, const DfParams& P, int lane)
{
  switch (limb) {
    case 0: unit_rounds_0_4<INV, 0>(Uu, P, lane); break;
    case 1: unit_rounds_0_4<INV, 1>(Uu, P, lane); break;
    case 2: unit_rounds_0_4<INV, 2>(Uu, P, lane); break;
    default: unit_rounds_0_4<INV, 3>(Uu, P, lane); break;
  }
}

#ifndef LOLB_DF_NW
#define LOLB_DF_NW 4            // warps per CTA = (chunk, limb) units per chunk task
#endif
#ifndef LOLB_DF_MINB
#define LOLB_DF_MINB (640 / (32 * LOLB_DF_NW))   // CTAs per SM the register allocation must allow
#endif
#ifndef LOLB_DF_SWITCH
#define LOLB_DF_SWITCH 0        // 1: rounds 0-4 specialised per limb (twiddles as constant-bank operands); 0: one copy, LDC
#endif
constexpr int kDfWarps = LOLB_DF_NW;
constexpr int kDfThreads = 32 * kDfWarps;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// (A TMA-staged variant -- cp.async.bulk + mbarrier filling a staging buffer one task ahead -- was measured slower, 46 % / 48 %
// against 56 % / 54 %: 4 CTAs/SM instead of 5 and the copy runs only one 1.5 us task ahead; removed in round 2)

// exchange-ring accesses: L2 only (the reader is another SM), optionally tagged evict_last so the ring stays in L2
// while the streamed element data (ld.cs / st.cs) passes through
#ifndef LOLB_DF_L2HINT
#define LOLB_DF_L2HINT 0      // measured: evict_last on the ring does not reduce DRAM write-back, and costs reads
#endif
__device__ __forceinline__ uint64_t ring_policy()
{
  uint64_t pol = 0;
#if LOLB_DF_L2HINT
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#endif
  return pol;
}
__device__ __forceinline__ uint32_t ring_ld(const uint32_t* p, uint64_t pol)
{
#if LOLB_DF_L2HINT
  uint32_t v;
  asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
  return v;
#else
  return __ldcg(p);
#endif
}
__device__ __forceinline__ void ring_st(uint32_t* p, uint32_t v, uint64_t pol)
{
#if LOLB_DF_L2HINT
  asm volatile("st.global.cg.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
#else
  *p = v;
#endif
}

// task index -> (element, kind, index inside the element); the queue interleaves, per element e, the first-kind
// tasks of e with the second-kind tasks of e - lag
template <int NT_A, int NT_B>
struct TaskId {
  int el, task;
  bool first, valid;
  __device__ __forceinline__ TaskId(unsigned t, int batch, int lag)
  {
    const unsigned grp = t / (unsigned)(NT_A + NT_B);
    const int r = (int)(t - grp * (unsigned)(NT_A + NT_B));
    first = r < NT_A;
    el = first ? (int)grp : (int)grp - lag;
    task = first ? r : r - NT_A;
    valid = el >= 0 && el < batch;
  }
};

template <int K, int TOP>
struct DfGeom {
  static constexpr int NCH = 1 << TOP;                 // chunks per limb
  static constexpr int N = 1024 << TOP;                // coefficients per limb
  static constexpr int G = kDfWarps / K;               // chunks per chunk task (kDfWarps units of 1024 residues)
  static constexpr int NT_CHUNK = NCH / G;             // chunk tasks per element
  static constexpr int NT_COL = (1024 * K) / kDfThreads;   // column tasks per element
  static constexpr int NV = 1 << TOP;                  // residues per thread in a column task
  static constexpr int PIECES = (kDfWarps * 1024) / (2 * kDfThreads);   // 16-byte pieces per thread in a chunk task
  static constexpr int STEP = (2 * kDfThreads) / K;    // coefficients between consecutive pieces of a thread
  static constexpr int CHUNK_BYTES = kDfWarps * 1024 * 8;
  static constexpr int COL_BYTES = NV * kDfThreads * 8;
  static constexpr int SMEM_BYTES = kDfWarps * kDfUnit * 4 + 64;
  // shared-memory word of piece ii of a thread, relative to  U + l0 * kDfUnit + c0 + (c0 >> 5),  c0 = 2 tid / K
  static __host__ __device__ constexpr int piece_off(int ii)
  {
    return ((STEP * ii) >> 10) * K * kDfUnit + ((STEP * ii) & 1023) + (((STEP * ii) & 1023) >> 5);
  }
};

// K = tupSize (1, 2 or 4); TOP = e - 11 = rounds above bit 10 (2..5)
template <bool INV, int K, int TOP>
__global__ void __launch_bounds__(kDfThreads, LOLB_DF_MINB)
k_pow2_df(int64_t* __restrict__ y, int batch, const __grid_constant__ DfParams P, uint32_t* __restrict__ ring,
          unsigned* __restrict__ ctr)
{
  typedef DfGeom<K, TOP> Geo;
  constexpr int N = Geo::N, G = Geo::G, NV = Geo::NV, PIECES = Geo::PIECES;
  constexpr int NT_A = INV ? Geo::NT_COL : Geo::NT_CHUNK;     // first kind (reads the element from HBM)
  constexpr int NT_B = INV ? Geo::NT_CHUNK : Geo::NT_COL;     // second kind (writes the element to HBM)
  static_assert(G >= 1 && Geo::NCH % G == 0, "chunk tasks must tile the element");
  static_assert(Geo::STEP % 32 == 0 && 1024 % Geo::STEP == 0, "piece addressing");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* U = reinterpret_cast<uint32_t*>(smem_raw);                    // kDfWarps units of u32 residues
  unsigned* mail = U + kDfWarps * kDfUnit;                               // [2][4]: task, ready, element, ring slot

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t pol = ring_policy();
  unsigned* cnt_a = ctr + kDfCtrHead;           // finished first-kind tasks per element
  unsigned* cnt_b = cnt_a + batch;              // finished second-kind tasks per element
  const unsigned total = (unsigned)(batch + P.lag) * (unsigned)(NT_A + NT_B);
  typedef TaskId<NT_A, NT_B> Tid;

  // what a task waits for: first kind -> its ring slot is free (the element `ring` before it is consumed);
  // second kind -> every first-kind task of its element is done.  Counters only grow, so a value read early
  // that already satisfies the condition stays valid.
  auto dep_ptr = [&](const Tid& id) -> const unsigned* {
    if (!id.valid) return nullptr;
    if (id.first) return id.el >= P.ring ? cnt_b + (id.el - P.ring) : nullptr;
    return cnt_a + id.el;
  };
  auto dep_target = [&](const Tid& id) -> unsigned { return id.first ? NT_B : NT_A; };

  // Thread 0 runs two tasks ahead: the atomic that hands out task i+2, the counter read for task i+1 and the copy
  // of task i+1's input are in flight while task i is computed.  One CTA barrier per task hands over the decoded
  // task through a double-buffered mailbox; the completion signal of task i (fence + atomic) is issued by thread 0
  // AFTER that barrier, while the other threads already work on task i+1.  `ready` = the dependency was already
  // satisfied when it was read ahead (the normal case); otherwise thread 0 signals first (so it never waits while
  // holding back its own completion) and then spins.
  unsigned t_next = 0;
  unsigned* pending = nullptr;                  // thread 0: completion counter of the task that just ended
  if (tid == 0) {
    const unsigned t0 = atomicAdd(ctr, 1u);
    t_next = atomicAdd(ctr, 1u);
    const Tid id0(t0, batch, P.lag);
    mail[0] = t0; mail[1] = 0u; mail[2] = (unsigned)id0.el; mail[3] = id0.valid ? (unsigned)id0.el % (unsigned)P.ring : 0u;
  }
  __syncthreads();

  for (int it = 0;; it++) {
    const unsigned* mc = mail + 4 * (it & 1);
    unsigned* mn = mail + 4 * ((it & 1) ^ 1);
    const unsigned t = mc[0];
    if (t >= total) break;
    const Tid id(t, batch, P.lag);
    if (tid == 0 && pending) {
      __threadfence();
      atomicAdd(pending, 1u);
      pending = nullptr;
    }
    if (!mc[1]) {
      if (tid == 0) {
        const unsigned* dp = dep_ptr(id);
        if (dp) while (ld_acquire(dp) < dep_target(id)) __nanosleep(64);
      }
      __syncthreads();
    }
    unsigned t_after = 0, dep_next = 0, dep_need = 0;
    if (tid == 0) {
      t_after = atomicAdd(ctr, 1u);
      mn[0] = t_next;
      if (t_next < total) {
        const Tid idn(t_next, batch, P.lag);
        mn[2] = (unsigned)idn.el;
        mn[3] = idn.valid ? (unsigned)idn.el % (unsigned)P.ring : 0u;
        const unsigned* dp = dep_ptr(idn);
        if (dp) { dep_next = ld_acquire(dp); dep_need = dep_target(idn); }
      }
    }

    if (id.valid) {
    const int el = id.el;
    const int task = id.task;
    const bool chunk_task = (id.first != INV);
    uint32_t* slot = ring + (size_t)mc[3] * ((size_t)K * N);
    int64_t* ebase = y + (size_t)el * ((size_t)K * N);
    unsigned* done = id.first ? cnt_a + el : cnt_b + el;

    if (chunk_task) {
      // ---------------------------------------------------------------- chunk task: bits [0,10)
      const int chunk0 = task * G;
      int64_t* gpiece = ebase + (size_t)chunk0 * 1024 * K;              // kDfWarps * 1024 contiguous int64
      const int unit = warp;                                             // (chunk_in_task, limb) of this warp
      const int uch = unit / K, limb = unit % K;
      uint32_t* Uu = U + unit * kDfUnit;
      const DfLimb& L = P.limb[limb];
      const Mont M{L.q, L.q2, L.qinv};
      uint32_t* srow = slot + (size_t)limb * N + (size_t)(chunk0 + uch) * 1024 + lane;
      // the pieces of this thread: int64 pairs (2 tid + 2 kDfThreads ii, +1); the limb of each half is fixed per
      // thread and the shared-memory word of piece ii is a compile-time offset from `ubase`
      const int l0 = (2 * tid) % K, c0 = (2 * tid) / K;
      uint32_t* ubase = U + l0 * kDfUnit + c0 + (c0 >> 5);
      constexpr int second = K == 1 ? 1 : kDfUnit;                      // the other half: next coefficient / next limb
      if (!INV) {
        const uint32_t q0 = P.limb[l0].q, q1 = P.limb[K == 1 ? 0 : l0 + 1].q;
        const longlong2* src = reinterpret_cast<const longlong2*>(gpiece) + tid;
        // coalesced read of the piece (all 16-byte loads in flight at once), limbs de-interleaved into the units
        {
          longlong2 raw[PIECES];
#pragma unroll
          for (int ii = 0; ii < PIECES; ii++) raw[ii] = __ldcs(src + kDfThreads * ii);
          uint32_t hi_or = 0, max0 = 0, max1 = 0;
#pragma unroll
          for (int ii = 0; ii < PIECES; ii++) {
            hi_or |= (uint32_t)((uint64_t)raw[ii].x >> 32) | (uint32_t)((uint64_t)raw[ii].y >> 32);
            max0 = max(max0, (uint32_t)raw[ii].x);
            max1 = max(max1, (uint32_t)raw[ii].y);
            ubase[Geo::piece_off(ii)] = (uint32_t)raw[ii].x;
            ubase[Geo::piece_off(ii) + second] = (uint32_t)raw[ii].y;
          }
          if (hi_or != 0 || max0 >= q0 || max1 >= q1) {
            // outside the Haskell contract (values not in [0,q)): redo this thread's pieces like the reference's c % q
#pragma unroll 1
            for (int ii = 0; ii < PIECES; ii++) {
              const longlong2 r = src[kDfThreads * ii];
              const int off = Geo::piece_off(ii);
              ubase[off] = df_reduce_any64(r.x, q0);
              ubase[off + second] = df_reduce_any64(r.y, q1);
            }
          }
        }
        __syncthreads();
#if LOLB_DF_SWITCH
        unit_rounds_0_4_any<false>(limb, Uu, P, lane);
#else
        unit_rounds_0_4_rt<false>(limb, Uu, P, lane);
#endif
        __syncwarp();
        // rounds 5-9: lane owns coefficients lane + 32 j
        {
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = Uu[lane + 33 * j];
          const uint32_t* twl = L.tw + lane;
          ct_rounds<5, false>(v, M, [&](int a, int jj) { return __ldg(twl + ((32 << a) - 1 + 32 * jj)); });
#pragma unroll
          for (int j = 0; j < 32; j++) ring_st(srow + 32 * j, v[j], pol);
        }
      } else {
        {
          uint32_t v[32];
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = ring_ld(srow + 32 * j, pol);
          const uint32_t* twl = L.tw + lane;
          gs_rounds<5, 0>(v, M, [&](int a, int jj) { return __ldg(twl + ((32 << a) - 1 + 32 * jj)); });
#pragma unroll
          for (int j = 0; j < 32; j++) Uu[lane + 33 * j] = v[j];
        }
        __syncwarp();
#if LOLB_DF_SWITCH
        unit_rounds_0_4_any<true>(limb, Uu, P, lane);
#else
        unit_rounds_0_4_rt<true>(limb, Uu, P, lane);
#endif
        __syncthreads();
        // canonical residues -> interleaved int64, coalesced 128-bit stores
#pragma unroll
        for (int ii = 0; ii < PIECES; ii++) {
          const uint32_t x0 = ubase[Geo::piece_off(ii)], x1 = ubase[Geo::piece_off(ii) + second];
          __stcs(reinterpret_cast<longlong2*>(gpiece) + tid + kDfThreads * ii, make_longlong2((int64_t)x0, (int64_t)x1));
        }
      }
    } else {
      // ---------------------------------------------------------------- column task: bits [10, 10 + TOP)
      const int f = task * kDfThreads + tid;                             // (coefficient b, limb) pair, ABI order
      const int b = f / K, limb = f % K;
      const DfLimb& L = P.limb[limb];
      const Mont M{L.q, L.q2, L.qinv};
      uint32_t* scol = slot + (size_t)limb * N + b;
      int64_t* gcol = ebase + f;
      const uint32_t* twb = L.tw + b;
      uint32_t v[NV];
      if (!INV) {
#pragma unroll
        for (int j = 0; j < NV; j++) v[j] = ring_ld(scol + 1024 * j, pol);
        ct_rounds<TOP, false>(v, M, [&](int a, int jj) { return __ldg(twb + ((1024 << a) - 1 + 1024 * jj)); });
#pragma unroll
        for (int j = 0; j < NV; j++) __stcs(gcol + (size_t)1024 * K * j, (int64_t)M.canon(M.fold(v[j])));
      } else {
        uint32_t hi_or = 0, lo_max = 0;
#pragma unroll
        for (int j = 0; j < NV; j++) {
          const int64_t raw = __ldcs(gcol + (size_t)1024 * K * j);
          v[j] = (uint32_t)raw;
          hi_or |= (uint32_t)((uint64_t)raw >> 32);
          lo_max = max(lo_max, v[j]);
        }
        if (hi_or != 0 || lo_max >= L.q) {      // outside the Haskell contract: reduce like the reference's c % q
#pragma unroll
          for (int j = 0; j < NV; j++) v[j] = df_reduce_any64(gcol[(size_t)1024 * K * j], L.q);
        }
        gs_rounds<TOP, 0>(v, M, [&](int a, int jj) { return __ldg(twb + ((1024 << a) - 1 + 1024 * jj)); });
#pragma unroll
        for (int j = 0; j < NV; j++) ring_st(scol + 1024 * j, v[j], pol);
      }
    }
    if (tid == 0) pending = done;
    }
    if (tid == 0) {
      mn[1] = dep_next >= dep_need ? 1u : 0u;
      t_next = t_after;
    }
    __syncthreads();      // the task's stores are issued, U may be overwritten, the mailbox of the next task is visible
  }
  if (tid == 0 && pending) {
    __threadfence();
    atomicAdd(pending, 1u);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Paired schedule (default): one queue entry = one chunk task of element e PLUS the matching column task(s) of
// element e - lag, executed back to back by the same CTA.  Per entry there is ONE CTA barrier (after the
// de-interleave, forward; before the interleaved store, inverse); it also publishes the next entry, claimed by
// thread 0 at the top of the current one.  Shared memory U is double buffered by entry parity, so a warp that is
// done moves on without waiting for the others.  Every warp signals its own part (fence + atomic by lane 0,
// issued after the loads of its next part, so the fence waits together with them) and reads the counters it
// depends on one part ahead.  Counters count warp-parts: NT * kDfWarps per element and kind.
template <bool INV, int K, int TOP>
__global__ void __launch_bounds__(kDfThreads, LOLB_DF_MINB)
k_pow2_dfm(int64_t* __restrict__ y, int batch, const __grid_constant__ DfParams P, uint32_t* __restrict__ ring,
           unsigned* __restrict__ ctr)
{
  typedef DfGeom<K, TOP> Geo;
  constexpr int N = Geo::N, G = Geo::G, NV = Geo::NV, PIECES = Geo::PIECES;
  constexpr int NT = Geo::NT_CHUNK;                     // queue entries per element
  constexpr int CPI = Geo::NT_COL / Geo::NT_CHUNK;      // column tasks per entry (1 at e = 16)
  constexpr unsigned FULL = (unsigned)NT * kDfWarps;    // warp-parts per element and kind
  static_assert(G >= 1 && Geo::NCH % G == 0 && Geo::NT_COL % Geo::NT_CHUNK == 0 && CPI >= 1, "entry geometry");

  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* U2 = reinterpret_cast<uint32_t*>(smem_raw);                 // [2][kDfWarps units]
  unsigned* mail = U2 + 2 * kDfWarps * kDfUnit;                         // [2] queue entry

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t pol = ring_policy();
  unsigned* cnt_a = ctr + kDfCtrHead;           // finished first-kind warp-parts per element
  unsigned* cnt_b = cnt_a + batch;              // finished second-kind warp-parts per element
  const unsigned total = (unsigned)(batch + P.lag) * (unsigned)NT;

  // Completion signal of the part this warp finished last (warp-uniform pointer).  EVERY lane fences its own ring /
  // element stores before lane 0 bumps the counter: a fence by lane 0 alone does not order the other lanes' stores.
  unsigned* pend = nullptr;
  auto flush = [&]() {
    if (pend) {
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(pend, 1u);
      pend = nullptr;
    }
  };
  // seen: value read ahead by lane 0.  A warp never waits while it holds back its own completion signal (the
  // counter it waits for may, through a chain of other CTAs, depend on it), so the slow path signals first.
  auto wait_dep = [&](const unsigned* dp, unsigned seen) {
    if (dp != nullptr) {
      const bool slow = __shfl_sync(0xffffffffu, (int)(seen < FULL), 0) != 0;
      if (slow) {
        flush();
        if (lane == 0) do { __nanosleep(64); seen = ld_acquire(dp); } while (seen < FULL);
        __syncwarp();
      }
    }
  };

  if (tid == 0) mail[0] = atomicAdd(ctr, 1u);
  __syncthreads();

  // the unit of this warp in a chunk task, and this thread's pieces in the (de)interleave
  const int unit = warp, uch = unit / K, limb_u = unit % K;
  const int l0 = (2 * tid) % K, c0 = (2 * tid) / K;
  constexpr int second = K == 1 ? 1 : kDfUnit;

  for (int it = 0;; it++) {
    const unsigned c = mail[it & 1];
    if (c >= total) break;
    uint32_t* U = U2 + (it & 1) * (kDfWarps * kDfUnit);
    uint32_t* Uu = U + unit * kDfUnit;
    uint32_t* ubase = U + l0 * kDfUnit + c0 + (c0 >> 5);
    const int el1 = (int)(c / (unsigned)NT), task = (int)(c % (unsigned)NT), el2 = el1 - P.lag;
    const bool v1 = el1 < batch, v2 = el2 >= 0 && el2 < batch;
    unsigned c_new = 0;
    if (tid == 0) c_new = atomicAdd(ctr, 1u);
    // counters this warp depends on, read ahead by lane 0: first kind -> ring slot free, second kind -> element ready
    const unsigned* dp1 = (v1 && el1 >= P.ring) ? cnt_b + (el1 - P.ring) : nullptr;
    const unsigned* dp2 = v2 ? cnt_a + el2 : nullptr;
    unsigned seen1 = FULL, seen2 = FULL;
    if (lane == 0) {
      if (dp1) seen1 = ld_acquire(dp1);
      if (dp2) seen2 = ld_acquire(dp2);
    }
    uint32_t* slot1 = ring + (size_t)((unsigned)el1 % (unsigned)P.ring) * ((size_t)K * N);
    uint32_t* slot2 = ring + (size_t)((unsigned)(v2 ? el2 : 0) % (unsigned)P.ring) * ((size_t)K * N);
    int64_t* ebase1 = y + (size_t)el1 * ((size_t)K * N);
    int64_t* ebase2 = y + (size_t)(v2 ? el2 : 0) * ((size_t)K * N);
    const DfLimb& Lu = P.limb[limb_u];
    const Mont Mu{Lu.q, Lu.q2, Lu.qinv};

    if (!INV) {
      // ================================================================ forward
      // ---- chunk task of element el1: bits [0,10)
      const int chunk0 = task * G;
      if (v1) {
        const longlong2* src = reinterpret_cast<const longlong2*>(ebase1 + (size_t)chunk0 * 1024 * K) + tid;
        const uint32_t q0 = P.limb[l0].q, q1 = P.limb[K == 1 ? 0 : l0 + 1].q;
        longlong2 raw[PIECES];
#pragma unroll
        for (int ii = 0; ii < PIECES; ii++) raw[ii] = __ldcs(src + kDfThreads * ii);
        flush();
        uint32_t hi_or = 0, max0 = 0, max1 = 0;
#pragma unroll
        for (int ii = 0; ii < PIECES; ii++) {
          hi_or |= (uint32_t)((uint64_t)raw[ii].x >> 32) | (uint32_t)((uint64_t)raw[ii].y >> 32);
          max0 = max(max0, (uint32_t)raw[ii].x);
          max1 = max(max1, (uint32_t)raw[ii].y);
          ubase[Geo::piece_off(ii)] = (uint32_t)raw[ii].x;
          ubase[Geo::piece_off(ii) + second] = (uint32_t)raw[ii].y;
        }
        if (hi_or != 0 || max0 >= q0 || max1 >= q1) {
          // outside the Haskell contract (values not in [0,q)): redo this thread's pieces like the reference's c % q
#pragma unroll 1
          for (int ii = 0; ii < PIECES; ii++) {
            const longlong2 r = src[kDfThreads * ii];
            const int off = Geo::piece_off(ii);
            ubase[off] = df_reduce_any64(r.x, q0);
            ubase[off + second] = df_reduce_any64(r.y, q1);
          }
        }
      } else {
        flush();
      }
      if (tid == 0) mail[(it & 1) ^ 1] = c_new;
      __syncthreads();
      if (v1) {
        unit_rounds_0_4_rt<false>(limb_u, Uu, P, lane);
        __syncwarp();
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = Uu[lane + 33 * j];
        const uint32_t* twl = Lu.tw + lane;
        ct_rounds<5, false>(v, Mu, [&](int a, int jj) { return __ldg(twl + ((32 << a) - 1 + 32 * jj)); });
        wait_dep(dp1, seen1);
        uint32_t* srow = slot1 + (size_t)limb_u * N + (size_t)(chunk0 + uch) * 1024 + lane;
#pragma unroll
        for (int j = 0; j < 32; j++) ring_st(srow + 32 * j, v[j], pol);
        pend = cnt_a + el1;
      }
      // ---- column task(s) of element el2: bits [10, 10 + TOP)
      if (v2) {
        wait_dep(dp2, seen2);
#pragma unroll 1
        for (int sub = 0; sub < CPI; sub++) {
          const int f = (task * CPI + sub) * kDfThreads + tid;           // (coefficient b, limb) pair, ABI order
          const int b = f / K, limb = f % K;
          const DfLimb& L = P.limb[limb];
          const Mont M{L.q, L.q2, L.qinv};
          const uint32_t* scol = slot2 + (size_t)limb * N + b;
          int64_t* gcol = ebase2 + f;
          const uint32_t* twb = L.tw + b;
          uint32_t v[NV];
#pragma unroll
          for (int j = 0; j < NV; j++) v[j] = ring_ld(scol + 1024 * j, pol);
          if (sub == 0) flush();
          ct_rounds<TOP, false>(v, M, [&](int a, int jj) { return __ldg(twb + ((1024 << a) - 1 + 1024 * jj)); });
#pragma unroll
          for (int j = 0; j < NV; j++) __stcs(gcol + (size_t)1024 * K * j, (int64_t)M.canon(M.fold(v[j])));
        }
        pend = cnt_b + el2;
      }
    } else {
      // ================================================================ inverse
      // ---- column task(s) of element el1: bits [10, 10 + TOP), reads the element
      if (v1) {
#pragma unroll 1
        for (int sub = 0; sub < CPI; sub++) {
          const int f = (task * CPI + sub) * kDfThreads + tid;
          const int b = f / K, limb = f % K;
          const DfLimb& L = P.limb[limb];
          const Mont M{L.q, L.q2, L.qinv};
          uint32_t* scol = slot1 + (size_t)limb * N + b;
          const int64_t* gcol = ebase1 + f;
          const uint32_t* twb = L.tw + b;
          uint32_t v[NV];
          uint32_t hi_or = 0, lo_max = 0;
#pragma unroll
          for (int j = 0; j < NV; j++) {
            const int64_t raw = __ldcs(gcol + (size_t)1024 * K * j);
            v[j] = (uint32_t)raw;
            hi_or |= (uint32_t)((uint64_t)raw >> 32);
          }
          if (sub == 0) flush();
#pragma unroll
          for (int j = 0; j < NV; j++) lo_max = max(lo_max, v[j]);
          if (hi_or != 0 || lo_max >= L.q) {      // outside the Haskell contract: reduce like the reference's c % q
#pragma unroll
            for (int j = 0; j < NV; j++) v[j] = df_reduce_any64(gcol[(size_t)1024 * K * j], L.q);
          }
          gs_rounds<TOP, 0>(v, M, [&](int a, int jj) { return __ldg(twb + ((1024 << a) - 1 + 1024 * jj)); });
          if (sub == 0) wait_dep(dp1, seen1);
#pragma unroll
          for (int j = 0; j < NV; j++) ring_st(scol + 1024 * j, v[j], pol);
        }
        pend = cnt_a + el1;
      }
      // ---- chunk task of element el2: bits [0,10), writes the element
      const int chunk0 = task * G;
      if (v2) {
        wait_dep(dp2, seen2);
        const uint32_t* srow = slot2 + (size_t)limb_u * N + (size_t)(chunk0 + uch) * 1024 + lane;
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = ring_ld(srow + 32 * j, pol);
        flush();
        const uint32_t* twl = Lu.tw + lane;
        gs_rounds<5, 0>(v, Mu, [&](int a, int jj) { return __ldg(twl + ((32 << a) - 1 + 32 * jj)); });
#pragma unroll
        for (int j = 0; j < 32; j++) Uu[lane + 33 * j] = v[j];
        __syncwarp();
        unit_rounds_0_4_rt<true>(limb_u, Uu, P, lane);
      } else {
        flush();
      }
      if (tid == 0) mail[(it & 1) ^ 1] = c_new;
      __syncthreads();
      if (v2) {
        longlong2* dst = reinterpret_cast<longlong2*>(ebase2 + (size_t)chunk0 * 1024 * K) + tid;
#pragma unroll
        for (int ii = 0; ii < PIECES; ii++) {
          const uint32_t x0 = ubase[Geo::piece_off(ii)], x1 = ubase[Geo::piece_off(ii) + second];
          __stcs(dst + kDfThreads * ii, make_longlong2((int64_t)x0, (int64_t)x1));
        }
        pend = cnt_b + el2;
      }
    }
  }
  flush();
}

uint32_t neg_inv32(uint32_t q)
{
  uint32_t inv = q;
  for (int i = 0; i < 5; i++) inv *= 2u - q * inv;
  return 0u - inv;
}

bool shape_ok(const lolb_plan* pl)
{
  if (pl->kind != PLAN_RQ || pl->pe.size() != 1 || pl->pe[0].prime != 2) return false;
  const int e = pl->pe[0].exponent;
  if (e < 10 || e > 16) return false;
  if (pl->k != 1 && pl->k != 2 && pl->k != 4) return false;
  if (e >= 14 && ((1 << (e - 11)) * pl->k) % kDfWarps != 0) return false;      // dataflow: chunk tasks of kDfWarps units must tile the element
  for (int64_t q : pl->qs) if (!(q & 1) || 4 * (uint64_t)q >= ((uint64_t)1 << 32)) return false;
  return true;
}

template <bool INV, int K, int TOP>
int launch_df(const lolb_plan* pl, const FusedPow2Df* F, int64_t* y, int64_t batch, cudaStream_t st)
{
  if constexpr (((1 << TOP) * K) % kDfWarps != 0) return LOLB_FUSED_UNAVAILABLE;
  else {
  typedef DfGeom<K, TOP> Geo;
  static int per_sm = 0;
  static PerDeviceOnce once;      // the attribute is per device; the occupancy is the same on every B200
  if (once.first() || !per_sm) {
    LOLB_CUDA(cudaFuncSetAttribute(k_pow2_df<INV, K, TOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo::SMEM_BYTES));
    LOLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pow2_df<INV, K, TOP>, kDfThreads, Geo::SMEM_BYTES));
    if (per_sm < 1) per_sm = 1;
  }
  DfParams P = INV ? F->inv : F->fwd;
  const int64_t tasks_per_el = (int64_t)(((1 << TOP) * K) / kDfWarps + (1024 * K) / kDfThreads);
  int64_t grid = (int64_t)pl->num_sms * per_sm;
  if (grid > batch * tasks_per_el) grid = batch * tasks_per_el;
  // every CTA holds up to 3 tasks (one running, two claimed ahead): the second kind must trail the first by more
  // than that window or it would wait for tasks that have not started, and a slot is reused a window after that
  const int64_t window = (3 * grid + tasks_per_el - 1) / tasks_per_el;
  P.lag = (int32_t)(window + 2);
  P.ring = 2 * P.lag + 2;
  const char* s;
  if ((s = getenv("LOLB_DF_RING")) != nullptr && atoi(s) > 0) P.ring = atoi(s);
  if ((s = getenv("LOLB_DF_LAG")) != nullptr && atoi(s) > 0) P.lag = atoi(s);
  if (P.ring > batch) P.ring = (int32_t)batch;
  if (P.ring < 2) P.ring = 2;
  if (P.lag >= P.ring) P.lag = P.ring - 1;
  const size_t slot_bytes = (size_t)K * pl->n * sizeof(uint32_t);
  const size_t ring_bytes = (size_t)P.ring * slot_bytes;
  const size_t ctr_bytes = ((size_t)kDfCtrHead + 2 * (size_t)batch) * sizeof(unsigned);
  uint32_t* ring = (uint32_t*)plan_ws(pl, st, ring_bytes + ctr_bytes);      // per stream: concurrent launches never share the ring or its counters
  if (!ring) return LOLB_ERR_CUDA;
  unsigned* ctr = (unsigned*)((char*)ring + ring_bytes);
  LOLB_CUDA(cudaMemsetAsync(ctr, 0, ctr_bytes, st));
  k_pow2_df<INV, K, TOP><<<(int)grid, kDfThreads, Geo::SMEM_BYTES, st>>>(y, (int)batch, P, ring, ctr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_pow2_df");
  count_launch();
  return LOLB_OK;
  }
}

template <bool INV, int K, int TOP>
int launch_dfm(const lolb_plan* pl, const FusedPow2Df* F, int64_t* y, int64_t batch, cudaStream_t st)
{
  if constexpr (((1 << TOP) * K) % kDfWarps != 0) return LOLB_FUSED_UNAVAILABLE;
  else {
  typedef DfGeom<K, TOP> Geo;
  constexpr int smem = 2 * kDfWarps * kDfUnit * 4 + 64;
  static int per_sm = 0;
  static PerDeviceOnce once;      // the attribute is per device; the occupancy is the same on every B200
  if (once.first() || !per_sm) {
    LOLB_CUDA(cudaFuncSetAttribute(k_pow2_dfm<INV, K, TOP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    LOLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pow2_dfm<INV, K, TOP>, kDfThreads, smem));
    if (per_sm < 1) per_sm = 1;
  }
  DfParams P = INV ? F->inv : F->fwd;
  const int64_t nt = Geo::NT_CHUNK;
  int64_t grid = (int64_t)pl->num_sms * per_sm;
  if (grid > batch * nt) grid = batch * nt;
  // a CTA holds two queue entries (one running, one claimed): the second kind trails the first by more than that
  // window, and a ring slot is reused one more window later
  const int64_t window = (2 * grid + nt - 1) / nt;
  P.lag = (int32_t)(window + 2);
  P.ring = 2 * P.lag + 2;
  const char* s;
  if ((s = getenv("LOLB_DF_RING")) != nullptr && atoi(s) > 0) P.ring = atoi(s);
  if ((s = getenv("LOLB_DF_LAG")) != nullptr && atoi(s) > 0) P.lag = atoi(s);
  if (P.ring > batch) P.ring = (int32_t)batch;
  if (P.ring < 2) P.ring = 2;
  if (P.lag >= P.ring) P.lag = P.ring - 1;
  const size_t ring_bytes = (size_t)P.ring * K * pl->n * sizeof(uint32_t);
  const size_t ctr_bytes = ((size_t)kDfCtrHead + 2 * (size_t)batch) * sizeof(unsigned);
  uint32_t* ring = (uint32_t*)plan_ws(pl, st, ring_bytes + ctr_bytes);      // per stream: concurrent launches never share the ring or its counters
  if (!ring) return LOLB_ERR_CUDA;
  unsigned* ctr = (unsigned*)((char*)ring + ring_bytes);
  LOLB_CUDA(cudaMemsetAsync(ctr, 0, ctr_bytes, st));
  k_pow2_dfm<INV, K, TOP><<<(int)grid, kDfThreads, smem, st>>>(y, (int)batch, P, ring, ctr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_pow2_dfm");
  count_launch();
  return LOLB_OK;
  }
}

template <bool INV, int K>
int launch_df_top(const lolb_plan* pl, const FusedPow2Df* F, int64_t* y, int64_t batch, cudaStream_t st)
{
  {
    const int rc = pow2_resident_crt(pl, F, INV, y, batch, st);      // m <= 2^14: element-resident kernels where they win
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  // Measured on B200, CRT / CRT^-1 as % of the HBM roofline:
  //   e = 16: tupSize 4  split 56.4 / 56.1, unpaired 55.9 / 54.6, paired 54 / 54;  tupSize 1  split 65.6 / 60.5, paired 62.7 / 59.5;
  //           tupSize 2  paired 60.2 / 57.8, split 58.9 / 56.4
  //   e = 15, tupSize 4: split 61.6 / 59.4, unpaired 51.9 / 49.1;   e = 14, tupSize 4: split 61.8 / 61.5, unpaired 42 / 38
  // so the split schedule (two plain kernels, the u32 intermediate through a workspace) serves everything except tupSize 2 at
  // e = 16; keeping the rounds 5-9 twiddles in shared memory instead of L1 was measured too: no gain.
  const char* sched = getenv("LOLB_DF_SCHEDULE");      // "split" / "paired" / "unpaired" / "cluster" override
  if (sched && sched[0] == 'c') {
    const int rc = pow2_cluster_crt(pl, F, INV, y, batch, st);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  if (sched ? sched[0] == 's' : !(K == 2 && F->top == 5)) {
    const int rc = pow2_split_crt(pl, F, INV, y, batch, st);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  const bool paired = sched ? sched[0] == 'p' : K != 4;
  if (paired) {
    switch (F->top) {
      case 2: return launch_dfm<INV, K, 2>(pl, F, y, batch, st);
      case 3: return launch_dfm<INV, K, 3>(pl, F, y, batch, st);
      case 4: return launch_dfm<INV, K, 4>(pl, F, y, batch, st);
      case 5: return launch_dfm<INV, K, 5>(pl, F, y, batch, st);
    }
    return LOLB_FUSED_UNAVAILABLE;
  }
  switch (F->top) {
    case 2: return launch_df<INV, K, 2>(pl, F, y, batch, st);
    case 3: return launch_df<INV, K, 3>(pl, F, y, batch, st);
    case 4: return launch_df<INV, K, 4>(pl, F, y, batch, st);
    case 5: return launch_df<INV, K, 5>(pl, F, y, batch, st);
  }
  return LOLB_FUSED_UNAVAILABLE;
}

}  // namespace

int fused_pow2_df_select(lolb_plan* pl, void** slot)
{
  if (!shape_ok(pl)) return LOLB_OK;
  FusedPow2Df* F = (FusedPow2Df*)*slot;
  if (!F) { F = new FusedPow2Df(); *slot = F; }
  const int e = pl->pe[0].exponent, n = pl->n, k = pl->k, rounds = e - 1;
  F->top = e - 11;
  F->ok_fwd = pl->has_fwd && pl->ru.size() == 1;
  F->ok_inv = pl->has_inv && pl->ruinv.size() == 1 && (int)pl->mhatinv.size() == k;
  const size_t per_dir = ((size_t)n + 3) & ~(size_t)3;      // n - 1 entries, padded
  std::vector<uint32_t> host((size_t)k * 2 * per_dir, 0u);
  const int64_t m = pl->m;
  DfParams P[2]{};
  for (int dir = 0; dir < 2; dir++) {
    P[dir].n = n; P[dir].k = k;
    P[dir].ring = 48; P[dir].lag = 12;
  }
  for (int t = 0; t < k; t++) {
    const uint64_t q = (uint64_t)pl->qs[t];
    auto mont = [&](uint64_t c) { return (uint32_t)(((c % q) << 32) % q); };
    for (int dir = 0; dir < 2; dir++) {
      if (dir == 0 ? !F->ok_fwd : !F->ok_inv) continue;
      const std::vector<int64_t>& T = dir == 0 ? pl->ru[0] : pl->ruinv[0];
      auto root = [&](int64_t j) { int64_t v = T[(size_t)(j % m) * k + t] % (int64_t)q; return (uint64_t)(v < 0 ? v + (int64_t)q : v); };
      uint32_t* tw = host.data() + ((size_t)t * 2 + dir) * per_dir;
      for (int r = 0; r < rounds; r++)
        for (int64_t p = 0; p < ((int64_t)1 << r); p++)
          tw[((size_t)1 << r) - 1 + p] = mont(root((2 * p + 1) * (n >> (r + 1))));
      DfLimb& L = P[dir].limb[t];
      L.q = (uint32_t)q; L.q2 = 2 * (uint32_t)q; L.qinv = neg_inv32((uint32_t)q);
      for (int i = 0; i < 31; i++) L.c0[i] = tw[i];
      if (dir == 1) {
        const uint64_t s = (uint64_t)(((pl->mhatinv[t] % (int64_t)q) + (int64_t)q) % (int64_t)q);
        L.sA = mont(s);
        L.sB = mont(mulmod64(s, root(n >> 1), q));
      }
    }
  }
  if (F->d_tab) { cudaFree(F->d_tab); F->d_tab = nullptr; }
  LOLB_CUDA(cudaMalloc((void**)&F->d_tab, host.size() * sizeof(uint32_t)));
  LOLB_CUDA(cudaMemcpy(F->d_tab, host.data(), host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  for (int t = 0; t < k; t++)
    for (int dir = 0; dir < 2; dir++) P[dir].limb[t].tw = F->d_tab + ((size_t)t * 2 + dir) * per_dir;
  F->fwd = P[0]; F->inv = P[1];
  return LOLB_OK;
}

void fused_pow2_df_release(void* slot)
{
  FusedPow2Df* F = (FusedPow2Df*)slot;
  if (!F) return;
  if (F->d_tab) cudaFree(F->d_tab);
  delete F;
}

bool fused_pow2_df_available(const void* slot, bool inverse)
{
  const FusedPow2Df* F = (const FusedPow2Df*)slot;
  if (!F || getenv("LOLB_POW2_NO_DF")) return false;
  return inverse ? F->ok_inv : F->ok_fwd;
}

int fused_pow2_df_crt(const lolb_plan* pl, const void* slot, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  const FusedPow2Df* F = (const FusedPow2Df*)slot;
  if (!fused_pow2_df_available(slot, inverse)) return LOLB_FUSED_UNAVAILABLE;
  if (batch <= 0) return LOLB_OK;
  // one launch per slab of at most 2^16 elements: bounds the per-element counters (2 x 4 bytes per element) and the
  // 32-bit task index; a slab is >= 2 GiB of data at the smallest supported shape, so the extra launches are noise
  const int64_t slab = 65536;
  for (int64_t off = 0; off < batch; off += slab) {
    const int64_t cnt = batch - off < slab ? batch - off : slab;
    int64_t* ys = y + (size_t)off * pl->n * pl->k;
    int rc;
    switch (pl->k) {
      case 1: rc = inverse ? launch_df_top<true, 1>(pl, F, ys, cnt, st) : launch_df_top<false, 1>(pl, F, ys, cnt, st); break;
      case 2: rc = inverse ? launch_df_top<true, 2>(pl, F, ys, cnt, st) : launch_df_top<false, 2>(pl, F, ys, cnt, st); break;
      case 4: rc = inverse ? launch_df_top<true, 4>(pl, F, ys, cnt, st) : launch_df_top<false, 4>(pl, F, ys, cnt, st); break;
      default: return LOLB_FUSED_UNAVAILABLE;
    }
    if (rc) return rc;
  }
  return LOLB_OK;
}

}  // namespace lolb
