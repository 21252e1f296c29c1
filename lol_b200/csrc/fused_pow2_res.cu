// fused_pow2_res.cu -- Z_q CRT / CRT^-1 for power-of-two index m = 2^10 .. 2^14 with the ring element resident on chip
// (see pow2_common.cuh for the operator and the arithmetic; fused_pow2_df.cu for the larger indices).
#include <cstdlib>

#include "pow2_common.cuh"

namespace lolb {
namespace pow2 {

namespace {

// ---------------------------------------------------------------------------------------------------------------
// Small indices m = 2^10, 2^11 (n = 512, 1024: the reference's own benchmark parameters, lol/.../Benchmarks/
// Default.hs:41-46): a limb fits one warp -- 32 residues per lane -- so the whole transform is the two register passes
// of a chunk task with a warp-private transposition through 4 KB of shared memory.  No queue, no ring, no counters.
// tupSize 1: every warp is independent (own loads, __syncwarp only).  tupSize 2, 4: the CTA de-interleaves a 32 KB piece
// cooperatively (three CTA barriers per piece).  n = 512: a warp holds two limbs (16 + 16 residues per lane in the
// second pass).
template <bool INV, int K, int E>
__global__ void __launch_bounds__(128, 5)
k_pow2_small(int64_t* __restrict__ y, int64_t batch, const __grid_constant__ DfParams P)
{
  constexpr int n = 1 << (E - 1);                       // 512 or 1024
  constexpr int UPW = 1024 / n;                         // (element, limb) units per warp
  constexpr int UW = n + n / 32 + (E == 11 ? 8 : 0);    // words per unit: 1064 / 528 (528 = 16 mod 32: the two units of a warp hit disjoint banks)
  constexpr int S1 = E - 6;                             // rounds of the second pass: 5 or 4
  constexpr int V1 = 1 << S1;                           // residues per lane and unit in the second pass
  constexpr int LG = K == 1 ? 32 : 128;                 // threads that load one contiguous piece together
  constexpr int UPG = K == 1 ? UPW : 4 * UPW;           // units per loader group
  constexpr int EPG = UPG / K;                          // ring elements per loader group
  constexpr int PIECES = (UPG * n) / (2 * LG);          // 16-byte pieces per thread (= 16)
  constexpr int STEP = (2 * LG) / K;                    // coefficients between consecutive pieces of a thread
  static_assert(EPG >= 1 && PIECES == 16 && n % STEP == 0 && STEP % 32 == 0, "geometry");
  __shared__ __align__(16) uint32_t U[4 * UPW * UW];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tg = K == 1 ? lane : tid;                   // index inside the loader group
  const int grp = K == 1 ? warp : 0;                    // loader group inside the CTA
  const int groups_per_cta = K == 1 ? 4 : 1;
  const int l0 = (2 * tg) % K, c0 = (2 * tg) / K;
  uint32_t* ubase = U + (grp * UPG + l0) * UW + c0 + (c0 >> 5);
  constexpr int second = K == 1 ? 1 : UW;
  auto piece_off = [](int ii) { return ((STEP * ii) / n) * K * UW + ((STEP * ii) % n) + (((STEP * ii) % n) >> 5); };
  auto piece_el = [](int ii) { return (STEP * ii) / n; };                 // element (inside the group) of piece ii
  auto sync_group = [&]() { if (K == 1) __syncwarp(); else __syncthreads(); };

  const int64_t ngroups = (batch + EPG - 1) / EPG;
  for (int64_t g = (int64_t)blockIdx.x * groups_per_cta + grp; g < ngroups; g += (int64_t)gridDim.x * groups_per_cta) {
    const int64_t e0 = g * EPG;
    const int cnt = (int)(batch - e0 < EPG ? batch - e0 : EPG);           // ring elements of this group that exist
    longlong2* gp = reinterpret_cast<longlong2*>(y + (size_t)e0 * n * K) + tg;
    // this warp's units: u = first + h, element u / K, limb u % K
    const int ufirst = K == 1 ? 0 : warp * UPW;                           // inside the group
    uint32_t* Uw = U + (grp * UPG + ufirst) * UW;                         // this warp's first unit

    if (!INV || K > 1) {
      // contiguous piece -> units, limbs de-interleaved
      const uint32_t q0 = P.limb[l0].q, q1 = P.limb[K == 1 ? 0 : l0 + 1].q;
      longlong2 raw[PIECES];
#pragma unroll
      for (int ii = 0; ii < PIECES; ii++) raw[ii] = piece_el(ii) < cnt ? __ldcs(gp + LG * ii) : make_longlong2(0, 0);
      uint32_t hi_or = 0, max0 = 0, max1 = 0;
#pragma unroll
      for (int ii = 0; ii < PIECES; ii++) {
        hi_or |= (uint32_t)((uint64_t)raw[ii].x >> 32) | (uint32_t)((uint64_t)raw[ii].y >> 32);
        max0 = max(max0, (uint32_t)raw[ii].x);
        max1 = max(max1, (uint32_t)raw[ii].y);
        ubase[piece_off(ii)] = (uint32_t)raw[ii].x;
        ubase[piece_off(ii) + second] = (uint32_t)raw[ii].y;
      }
      if (hi_or != 0 || max0 >= q0 || max1 >= q1) {      // outside the Haskell contract: reduce like the reference's c % q
#pragma unroll 1
        for (int ii = 0; ii < PIECES; ii++) {
          if (piece_el(ii) >= cnt) continue;
          const longlong2 r = gp[LG * ii];
          ubase[piece_off(ii)] = df_reduce_any64(r.x, q0);
          ubase[piece_off(ii) + second] = df_reduce_any64(r.y, q1);
        }
      }
      sync_group();
    }

    // ---- the two register passes on this warp's unit(s)
    if (!INV) {
      // rounds 0-4: lane owns 32 consecutive residues (n = 512: lanes 0-15 the first unit, 16-31 the second)
      {
        const int hu = UPW == 1 ? 0 : lane >> 4, blk = UPW == 1 ? lane : lane & 15;
        const int limb = (ufirst + hu) % K;
        unit_rounds_0_4_rt<false>(limb, Uw + hu * UW - 33 * lane + 33 * blk, P, lane);      // base + 33 * blk
      }
      __syncwarp();
      // rounds 5 .. e-2: lane owns residues lane + 32 j of each unit
#pragma unroll
      for (int h = 0; h < UPW; h++) {
        const int u = ufirst + h, limb = u % K, el = u / K;
        const DfLimb& L = P.limb[limb];
        const Mont M{L.q, L.q2, L.qinv};
        uint32_t* ub = Uw + h * UW + lane;
        uint32_t v[V1];
#pragma unroll
        for (int j = 0; j < V1; j++) v[j] = ub[33 * j];
        const uint32_t* twl = L.tw + lane;
        ct_rounds<S1, false>(v, M, [&](int a, int jj) { return __ldg(twl + ((32 << a) - 1 + 32 * jj)); });
        if (K == 1) {
          if (el < cnt) {
            int64_t* out = y + (size_t)(e0 + el) * n + lane;
#pragma unroll
            for (int j = 0; j < V1; j++) __stcs(out + 32 * j, (int64_t)M.canon(M.fold(v[j])));
          }
        } else {
#pragma unroll
          for (int j = 0; j < V1; j++) ub[33 * j] = M.canon(M.fold(v[j]));
        }
      }
    } else {
#pragma unroll
      for (int h = 0; h < UPW; h++) {
        const int u = ufirst + h, limb = u % K, el = u / K;
        const DfLimb& L = P.limb[limb];
        const Mont M{L.q, L.q2, L.qinv};
        uint32_t* ub = Uw + h * UW + lane;
        uint32_t v[V1];
        if (K == 1) {
          const int64_t* in = y + (size_t)(e0 + (el < cnt ? el : 0)) * n + lane;
          uint32_t hi_or = 0, lo_max = 0;
#pragma unroll
          for (int j = 0; j < V1; j++) {
            const int64_t raw = __ldcs(in + 32 * j);
            v[j] = (uint32_t)raw;
            hi_or |= (uint32_t)((uint64_t)raw >> 32);
            lo_max = max(lo_max, v[j]);
          }
          if (hi_or != 0 || lo_max >= L.q) {
#pragma unroll
            for (int j = 0; j < V1; j++) v[j] = df_reduce_any64(in[32 * j], L.q);
          }
        } else {
#pragma unroll
          for (int j = 0; j < V1; j++) v[j] = ub[33 * j];
        }
        const uint32_t* twl = L.tw + lane;
        gs_rounds<S1, 0>(v, M, [&](int a, int jj) { return __ldg(twl + ((32 << a) - 1 + 32 * jj)); });
#pragma unroll
        for (int j = 0; j < V1; j++) ub[33 * j] = v[j];
      }
      __syncwarp();
      {
        const int hu = UPW == 1 ? 0 : lane >> 4, blk = UPW == 1 ? lane : lane & 15;
        const int limb = (ufirst + hu) % K;
        unit_rounds_0_4_rt<true>(limb, Uw + hu * UW - 33 * lane + 33 * blk, P, lane);
      }
    }

    if (INV || K > 1) {
      // units -> contiguous piece (canonical residues), coalesced 128-bit stores
      sync_group();
#pragma unroll
      for (int ii = 0; ii < PIECES; ii++) {
        if (piece_el(ii) < cnt) {
          const uint32_t x0 = ubase[piece_off(ii)], x1 = ubase[piece_off(ii) + second];
          __stcs(gp + LG * ii, make_longlong2((int64_t)x0, (int64_t)x1));
        }
      }
    }
    sync_group();      // U is reused by the next group
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Middle indices m = 2^12, 2^13 (tupSize 1, 2, 4) and 2^14 (tupSize 1): one ring element (all limbs, 17-68 KB as u32) is
// resident in the shared memory of a 128-thread CTA.  The (chunk, limb) units of 1024 residues go through the two
// register passes of a chunk task, one warp per unit; the remaining 1-3 rounds couple the chunks and run with a thread
// per coefficient (2-4 residues at stride 1024, eight coefficients per thread in flight).  Four CTA barriers per
// element, no queue and no ring; 5 CTAs per SM overlap each other's phases.
template <bool INV, int K, int E>
__global__ void __launch_bounds__(128, 5)
k_pow2_mid(int64_t* __restrict__ y, int64_t batch, const __grid_constant__ DfParams P)
{
  constexpr int n = 1 << (E - 1);                       // 2048 .. 8192
  constexpr int NCH = n / 1024;                         // chunks per limb: 2, 4, 8
  constexpr int T = E - 11;                             // top rounds: 1 .. 3
  constexpr int XI = T <= 2 ? 8 : 4;                    // coefficients per thread in flight in the top pass (<= 32 residues)
  constexpr int NV = 1 << T;
  constexpr int UW = kDfUnit;
  constexpr int EPC = NCH * K >= 4 ? 1 : 4 / (NCH * K);   // ring elements per CTA iteration: every warp gets a unit (2 at m = 2^12, tupSize 1)
  constexpr int UNITS = NCH * K * EPC;
  constexpr int PIECES = (n * K * EPC) / (2 * 128);     // 16-byte pieces per thread: 16 or 32
  constexpr int STEP = 256 / K;
  static_assert(UNITS * UW * 4 <= 72 * 1024 && 1024 % STEP == 0 && STEP % 32 == 0, "geometry");
  extern __shared__ __align__(16) uint32_t U[];         // [UNITS][UW], unit = (element * NCH + chunk) * K + limb

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int l0 = (2 * tid) % K, c0 = (2 * tid) / K;
  uint32_t* ubase = U + l0 * UW + c0 + (c0 >> 5);
  constexpr int second = K == 1 ? 1 : UW;
  auto piece_off = [](int ii) { return ((STEP * ii) >> 10) * K * UW + ((STEP * ii) & 1023) + (((STEP * ii) & 1023) >> 5); };

  // rounds 10 .. e-2 on the residues x + 1024 j of limb l, eight coefficients x = tid + 128 i per thread
  auto top_pass = [&](int64_t* gbase, int cnt) {
#pragma unroll 1
    for (int el_l = 0; el_l < EPC * K; el_l++) {
      const int l = el_l % K, eo = el_l / K;
      if (eo >= cnt) break;
      int64_t* ebase = gbase + (size_t)eo * n * K;
      uint32_t* Ue = U + eo * NCH * K * UW;
      const DfLimb& L = P.limb[l];
      const Mont M{L.q, L.q2, L.qinv};
#pragma unroll 1
      for (int ib = 0; ib < 8; ib += XI) {
      uint32_t v[XI][NV];
      const int xt = tid + 128 * ib;                      // first coefficient of this thread in this block
      if (INV && K == 1) {                              // straight from HBM: lanes are consecutive coefficients
        uint32_t hi_or = 0, lo_max = 0;
#pragma unroll
        for (int i = 0; i < XI; i++)
#pragma unroll
          for (int j = 0; j < NV; j++) {
            const int64_t raw = __ldcs(ebase + xt + 128 * i + 1024 * j);
            v[i][j] = (uint32_t)raw;
            hi_or |= (uint32_t)((uint64_t)raw >> 32);
            lo_max = max(lo_max, v[i][j]);
          }
        if (hi_or != 0 || lo_max >= L.q) {
#pragma unroll
          for (int i = 0; i < XI; i++)
#pragma unroll
            for (int j = 0; j < NV; j++) v[i][j] = df_reduce_any64(ebase[xt + 128 * i + 1024 * j], L.q);
        }
      } else {
#pragma unroll
        for (int i = 0; i < XI; i++) {
          const int x = xt + 128 * i;
#pragma unroll
          for (int j = 0; j < NV; j++) v[i][j] = Ue[(j * K + l) * UW + x + (x >> 5)];
        }
      }
#pragma unroll
      for (int i = 0; i < XI; i++) {
        const uint32_t* twx = L.tw + xt + 128 * i;
        if (!INV) ct_rounds<T, false>(v[i], M, [&](int a, int jj) { return __ldg(twx + ((1024 << a) - 1 + 1024 * jj)); });
        else gs_rounds<T, 0>(v[i], M, [&](int a, int jj) { return __ldg(twx + ((1024 << a) - 1 + 1024 * jj)); });
      }
      if (!INV && K == 1) {                             // straight to HBM
#pragma unroll
        for (int i = 0; i < XI; i++)
#pragma unroll
          for (int j = 0; j < NV; j++) __stcs(ebase + xt + 128 * i + 1024 * j, (int64_t)M.canon(M.fold(v[i][j])));
      } else {
#pragma unroll
        for (int i = 0; i < XI; i++) {
          const int x = xt + 128 * i;
#pragma unroll
          for (int j = 0; j < NV; j++) Ue[(j * K + l) * UW + x + (x >> 5)] = INV ? v[i][j] : M.canon(M.fold(v[i][j]));
        }
      }
      }
    }
  };

  // the two register passes of every (chunk, limb) unit, one warp per unit
  auto unit_passes = [&]() {
#pragma unroll 1
    for (int u = warp; u < UNITS; u += 4) {
      const int limb = u % K;
      const DfLimb& L = P.limb[limb];
      const Mont M{L.q, L.q2, L.qinv};
      uint32_t* Uu = U + u * UW;
      const uint32_t* twl = L.tw + lane;
      if (!INV) {
        unit_rounds_0_4_rt<false>(limb, Uu, P, lane);
        __syncwarp();
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = Uu[lane + 33 * j];
        ct_rounds<5, false>(v, M, [&](int a, int jj) { return __ldg(twl + ((32 << a) - 1 + 32 * jj)); });
#pragma unroll
        for (int j = 0; j < 32; j++) Uu[lane + 33 * j] = v[j];
      } else {
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = Uu[lane + 33 * j];
        gs_rounds<5, 0>(v, M, [&](int a, int jj) { return __ldg(twl + ((32 << a) - 1 + 32 * jj)); });
#pragma unroll
        for (int j = 0; j < 32; j++) Uu[lane + 33 * j] = v[j];
        __syncwarp();
        unit_rounds_0_4_rt<true>(limb, Uu, P, lane);
      }
    }
  };

  auto piece_el = [](int ii) { return (STEP * ii) / n; };                 // element (inside the CTA's group) of piece ii
  const int64_t ngroups = (batch + EPC - 1) / EPC;
  for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const int64_t e = g * EPC;
    const int cnt = (int)(batch - e < EPC ? batch - e : EPC);
    int64_t* ebase = y + (size_t)e * n * K;
    longlong2* gp = reinterpret_cast<longlong2*>(ebase) + tid;
    if (!(INV && K == 1)) {
      // element -> units, limbs de-interleaved (at most 16 pieces in flight per thread)
      const uint32_t q0 = P.limb[l0].q, q1 = P.limb[K == 1 ? 0 : l0 + 1].q;
#pragma unroll
      for (int part = 0; part < (PIECES + 15) / 16; part++) {
        constexpr int PP = PIECES < 16 ? PIECES : 16;
        longlong2 raw[PP];
#pragma unroll
        for (int i = 0; i < PP; i++) raw[i] = piece_el(part * 16 + i) < cnt ? __ldcs(gp + 128 * (part * 16 + i)) : make_longlong2(0, 0);
        uint32_t hi_or = 0, max0 = 0, max1 = 0;
#pragma unroll
        for (int i = 0; i < PP; i++) {
          hi_or |= (uint32_t)((uint64_t)raw[i].x >> 32) | (uint32_t)((uint64_t)raw[i].y >> 32);
          max0 = max(max0, (uint32_t)raw[i].x);
          max1 = max(max1, (uint32_t)raw[i].y);
          ubase[piece_off(part * 16 + i)] = (uint32_t)raw[i].x;
          ubase[piece_off(part * 16 + i) + second] = (uint32_t)raw[i].y;
        }
        if (hi_or != 0 || max0 >= q0 || max1 >= q1) {      // outside the Haskell contract: reduce like the reference's c % q
#pragma unroll 1
          for (int i = 0; i < PP; i++) {
            if (piece_el(part * 16 + i) >= cnt) continue;
            const longlong2 r = gp[128 * (part * 16 + i)];
            ubase[piece_off(part * 16 + i)] = df_reduce_any64(r.x, q0);
            ubase[piece_off(part * 16 + i) + second] = df_reduce_any64(r.y, q1);
          }
        }
      }
      __syncthreads();
    }
    if (!INV) {
      unit_passes();
      __syncthreads();
      top_pass(ebase, cnt);
    } else {
      top_pass(ebase, cnt);
      __syncthreads();
      unit_passes();
    }
    if (!(!INV && K == 1)) {
      // units -> element (canonical residues), coalesced 128-bit stores
      __syncthreads();
#pragma unroll
      for (int ii = 0; ii < PIECES; ii++) {
        if (piece_el(ii) < cnt) {
          const uint32_t x0 = ubase[piece_off(ii)], x1 = ubase[piece_off(ii) + second];
          __stcs(gp + 128 * ii, make_longlong2((int64_t)x0, (int64_t)x1));
        }
      }
    }
    __syncthreads();      // U is reused by the next group
  }
}

template <bool INV, int K, int E>
int launch_small(const lolb_plan* pl, const FusedPow2Df* F, int64_t* y, int64_t batch, cudaStream_t st)
{
  static int per_sm = 0;
  static PerDeviceOnce once;      // the attribute is per device; the occupancy is the same on every B200
  if (once.first() || !per_sm) {
    LOLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pow2_small<INV, K, E>, 128, 0));
    if (per_sm < 1) per_sm = 1;
  }
  constexpr int n = 1 << (E - 1), upw = 1024 / n;
  const int64_t el_per_cta = K == 1 ? 4 * upw : (4 * upw) / K;      // ring elements one CTA iteration covers
  int64_t grid = (int64_t)pl->num_sms * per_sm;
  const int64_t need = (batch + el_per_cta - 1) / el_per_cta;
  if (grid > need) grid = need;
  k_pow2_small<INV, K, E><<<(int)grid, 128, 0, st>>>(y, batch, INV ? F->inv : F->fwd);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_pow2_small");
  count_launch();
  return LOLB_OK;
}

template <bool INV, int K, int E>
int launch_mid(const lolb_plan* pl, const FusedPow2Df* F, int64_t* y, int64_t batch, cudaStream_t st)
{
  constexpr int units1 = ((1 << (E - 1)) / 1024) * K, epc = units1 >= 4 ? 1 : 4 / units1;
  constexpr int smem = units1 * epc * kDfUnit * 4;
  static int per_sm = 0;
  static PerDeviceOnce once;      // the attribute is per device; the occupancy is the same on every B200
  if (once.first() || !per_sm) {
    LOLB_CUDA(cudaFuncSetAttribute(k_pow2_mid<INV, K, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    LOLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pow2_mid<INV, K, E>, 128, smem));
    if (per_sm < 1) per_sm = 1;
  }
  int64_t grid = (int64_t)pl->num_sms * per_sm;
  if (grid > (batch + epc - 1) / epc) grid = (batch + epc - 1) / epc;
  k_pow2_mid<INV, K, E><<<(int)grid, 128, smem, st>>>(y, batch, INV ? F->inv : F->fwd);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_pow2_mid");
  count_launch();
  return LOLB_OK;
}

}  // namespace

int pow2_resident_crt(const lolb_plan* pl, const FusedPow2Df* F, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  const int k = pl->k, top = F->top;
#define RES(FN, E)                                                                                     \
  do {                                                                                                 \
    if (k == 1) return inverse ? FN<true, 1, E>(pl, F, y, batch, st) : FN<false, 1, E>(pl, F, y, batch, st); \
    if (k == 2) return inverse ? FN<true, 2, E>(pl, F, y, batch, st) : FN<false, 2, E>(pl, F, y, batch, st); \
    if (k == 4) return inverse ? FN<true, 4, E>(pl, F, y, batch, st) : FN<false, 4, E>(pl, F, y, batch, st); \
  } while (0)
  if (top == -1) RES(launch_small, 10);
  if (top == 0) RES(launch_small, 11);
  if (top == 1) RES(launch_mid, 12);
  // measured (B200, % of HBM peak forward / inverse, element-resident vs dataflow): m = 2^13: tupSize 1 82 / 78 vs 62 / 51,
  // tupSize 2 69 / 67 vs 56 / -, tupSize 4 (68 KB, 3 CTAs/SM) 55 / 56 vs 30 / 27; m = 2^14: tupSize 1 70 / 58 vs 65 / 56,
  // tupSize 2 (68 KB) 55 / 54 vs 63 / 54 -> dataflow from there on
  if (!getenv("LOLB_POW2_MID_OFF")) {
    if (top == 2) RES(launch_mid, 13);
    if (top == 3 && k == 1) return inverse ? launch_mid<true, 1, 14>(pl, F, y, batch, st) : launch_mid<false, 1, 14>(pl, F, y, batch, st);
  }
#undef RES
  return LOLB_FUSED_UNAVAILABLE;
}

}  // namespace pow2
}  // namespace lolb
