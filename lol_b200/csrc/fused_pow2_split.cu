// fused_pow2_split.cu -- Z_q CRT / CRT^-1 for m = 2^e, 14 <= e <= 16, as TWO plain streaming kernels per sub-batch that
// overlap through the hardware's own block scheduler (BASELINE.json config B: m = 2^16, four ~30-bit primes).
//
// Same operator, arithmetic and task bodies as the dataflow kernel in fused_pow2_df.cu (pow2_common.cuh has the
// citations): position bits [0,10) are "chunk" work (one contiguous 32 KB piece of the element per CTA, limbs
// de-interleaved through shared memory, rounds 0-9 in registers), bits [10, e-1) are "column" work (128 (coefficient,
// limb) pairs x all 2^(e-11) chunks per CTA, rounds 10 .. e-2 in registers, no shared memory).  The two kinds exchange
// u32 residues through a ring in global memory that stays in L2.
//
// What differs is who orders the work.  There a persistent kernel claims tasks from an atomic queue and spins on
// per-element counters (ncu: 25 % of the stalls at its hand-over barrier, 1.07x DRAM traffic, 55 % of HBM peak at tupSize
// 4).  Here the batch is cut into sub-batches of S elements; kernel A(i) (reads the element from HBM, writes the ring) runs
// on the caller's stream, kernel B(i) (reads the ring, writes the element) on an auxiliary stream behind an event, and
// A(i + R) waits for B(i) before it reuses ring slot i mod R.  No CTA waits for another CTA: the block scheduler fills the
// tail of B(i) with CTAs of A(i + 1), and each kernel gets its own register allocation (the column kernel needs half
// the registers of the chunk kernel).
#include <cstdlib>
#include <mutex>
#include <vector>

#include "pow2_common.cuh"

namespace lolb {

using namespace pow2;

namespace {

constexpr int kSpWarps = 4, kSpThreads = 32 * kSpWarps;

template <int K, int TOP>
struct SpGeom {
  static constexpr int NCH = 1 << TOP;                 // chunks per limb
  static constexpr int N = 1024 << TOP;                // coefficients per limb
  static constexpr int G = kSpWarps / K;               // chunks per chunk CTA (kSpWarps units of 1024 residues)
  static constexpr int NT_CHUNK = NCH / G;             // chunk CTAs per element
  static constexpr int NT_COL = (1024 * K) / kSpThreads;   // column CTAs per element
  static constexpr int NV = 1 << TOP;                  // residues per thread in a column CTA
  static constexpr int PIECES = (kSpWarps * 1024) / (2 * kSpThreads);   // 16-byte pieces per thread in a chunk CTA
  static constexpr int STEP = (2 * kSpThreads) / K;    // coefficients between consecutive pieces of a thread
  // shared-memory word of piece ii of a thread, relative to  U + l0 * kDfUnit + c0 + (c0 >> 5),  c0 = 2 tid / K
  static __host__ __device__ constexpr int piece_off(int ii)
  {
    return ((STEP * ii) >> 10) * K * kDfUnit + ((STEP * ii) & 1023) + (((STEP * ii) & 1023) >> 5);
  }
};

#ifndef LOLB_SP_CHUNK_MINB
#define LOLB_SP_CHUNK_MINB 5
#endif
#ifndef LOLB_SP_COL_MINB
#define LOLB_SP_COL_MINB 8
#endif

// chunk kernel: bits [0,10).  Forward: HBM element -> ring.  Inverse: ring -> HBM element (canonical, mhat^-1 folded in).
template <bool INV, int K, int TOP>
__device__ __forceinline__ void chunk_body(uint32_t* U, const int bid, int64_t* __restrict__ y, const DfParams& P, uint32_t* __restrict__ ring,
                                           const int rev_n)
{
  typedef SpGeom<K, TOP> Geo;
  constexpr int N = Geo::N, G = Geo::G, PIECES = Geo::PIECES;
  static_assert(G >= 1 && Geo::NCH % G == 0, "chunk CTAs must tile the element");
  static_assert(Geo::STEP % 32 == 0 && 1024 % Geo::STEP == 0, "piece addressing");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int el_l = bid / Geo::NT_CHUNK, task = bid - el_l * Geo::NT_CHUNK;
  const int el = rev_n ? rev_n - 1 - el_l : el_l;      // the second kernel of a sub-batch walks it backwards: the ring words written last are still in L2
  uint32_t* slot = ring + (size_t)el * ((size_t)K * N);
  int64_t* ebase = y + (size_t)el * ((size_t)K * N);

  const int chunk0 = task * G;
  int64_t* gpiece = ebase + (size_t)chunk0 * 1024 * K;              // kSpWarps * 1024 contiguous int64
  const int unit = warp;                                             // (chunk in CTA, limb) of this warp
  const int uch = unit / K, limb = unit % K;
  uint32_t* Uu = U + unit * kDfUnit;
  const DfLimb& L = P.limb[limb];
  const Mont M{L.q, L.q2, L.qinv};
  uint32_t* srow = slot + (size_t)limb * N + (size_t)(chunk0 + uch) * 1024 + lane;
  // the pieces of this thread: int64 pairs (2 tid + 2 kSpThreads ii, +1); the limb of each half is fixed per thread
  // and the shared-memory word of piece ii is a compile-time offset from `ubase`
  const int l0 = (2 * tid) % K, c0 = (2 * tid) / K;
  uint32_t* ubase = U + l0 * kDfUnit + c0 + (c0 >> 5);
  constexpr int second = K == 1 ? 1 : kDfUnit;                      // the other half: next coefficient / next limb
  if (!INV) {
    const uint32_t q0 = P.limb[l0].q, q1 = P.limb[K == 1 ? 0 : l0 + 1].q;
    const longlong2* src = reinterpret_cast<const longlong2*>(gpiece) + tid;
    {
      // coalesced read of the piece (all 16-byte loads in flight at once), limbs de-interleaved into the units
      longlong2 raw[PIECES];
#pragma unroll
      for (int ii = 0; ii < PIECES; ii++) raw[ii] = __ldcs(src + kSpThreads * ii);
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
          const longlong2 r = src[kSpThreads * ii];
          const int off = Geo::piece_off(ii);
          ubase[off] = df_reduce_any64(r.x, q0);
          ubase[off + second] = df_reduce_any64(r.y, q1);
        }
      }
    }
    __syncthreads();
    unit_rounds_0_4_rt<false>(limb, Uu, P, lane);
    __syncwarp();
    // rounds 5-9: lane owns coefficients lane + 32 j
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; j++) v[j] = Uu[lane + 33 * j];
    const uint32_t* twl = L.tw + lane;
    ct_rounds<5, false>(v, M, [&](int a, int jj) { return __ldg(twl + ((32 << a) - 1 + 32 * jj)); });
#pragma unroll
    for (int j = 0; j < 32; j++) srow[32 * j] = v[j];
  } else {
    {
      uint32_t v[32];
#pragma unroll
      for (int j = 0; j < 32; j++) v[j] = __ldcg(srow + 32 * j);
      const uint32_t* twl = L.tw + lane;
      gs_rounds<5, 0>(v, M, [&](int a, int jj) { return __ldg(twl + ((32 << a) - 1 + 32 * jj)); });
#pragma unroll
      for (int j = 0; j < 32; j++) Uu[lane + 33 * j] = v[j];
    }
    __syncwarp();
    unit_rounds_0_4_rt<true>(limb, Uu, P, lane);
    __syncthreads();
    // canonical residues -> interleaved int64, coalesced 128-bit stores
#pragma unroll
    for (int ii = 0; ii < PIECES; ii++) {
      const uint32_t x0 = ubase[Geo::piece_off(ii)], x1 = ubase[Geo::piece_off(ii) + second];
      __stcs(reinterpret_cast<longlong2*>(gpiece) + tid + kSpThreads * ii, make_longlong2((int64_t)x0, (int64_t)x1));
    }
  }
}

template <bool INV, int K, int TOP>
__global__ void __launch_bounds__(kSpThreads, LOLB_SP_CHUNK_MINB)
k_pow2_chunk(int64_t* __restrict__ y, const __grid_constant__ DfParams P, uint32_t* __restrict__ ring, const int rev_n)
{
  __shared__ __align__(16) uint32_t U[kSpWarps * kDfUnit];
  chunk_body<INV, K, TOP>(U, (int)blockIdx.x, y, P, ring, rev_n);
}

// column kernel: bits [10, 10 + TOP).  Forward: ring -> HBM element (canonical).  Inverse: HBM element -> ring.
template <bool INV, int K, int TOP>
__device__ __forceinline__ void col_body(const int bid, int64_t* __restrict__ y, const DfParams& P, uint32_t* __restrict__ ring, const int rev_n)
{
  typedef SpGeom<K, TOP> Geo;
  constexpr int N = Geo::N, NV = Geo::NV;
  const int tid = threadIdx.x;
  const int el_l = bid / Geo::NT_COL, task = bid - el_l * Geo::NT_COL;
  const int el = rev_n ? rev_n - 1 - el_l : el_l;
  uint32_t* slot = ring + (size_t)el * ((size_t)K * N);
  int64_t* ebase = y + (size_t)el * ((size_t)K * N);
  const int f = task * kSpThreads + tid;                             // (coefficient b, limb) pair, ABI order
  const int b = f / K, limb = f % K;
  const DfLimb& L = P.limb[limb];
  const Mont M{L.q, L.q2, L.qinv};
  uint32_t* scol = slot + (size_t)limb * N + b;
  int64_t* gcol = ebase + f;
  const uint32_t* twb = L.tw + b;
  uint32_t v[NV];
  if (!INV) {
#pragma unroll
    for (int j = 0; j < NV; j++) v[j] = __ldcg(scol + 1024 * j);
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
    for (int j = 0; j < NV; j++) scol[1024 * j] = v[j];
  }
}

template <bool INV, int K, int TOP>
__global__ void __launch_bounds__(kSpThreads, LOLB_SP_COL_MINB)
k_pow2_col(int64_t* __restrict__ y, const __grid_constant__ DfParams P, uint32_t* __restrict__ ring, const int rev_n)
{
  col_body<INV, K, TOP>((int)blockIdx.x, y, P, ring, rev_n);
}

// Mixed launch: the first-kind work of sub-batch i + 1 and the second-kind work of sub-batch i in ONE grid, interleaved by
// block index, so that consecutive launches on one stream need neither events nor a second stream: launch i's second-kind
// CTAs read what launch i - 1's first-kind CTAs wrote (stream order), and two ring slots alternate.  Compute-heavy chunk
// CTAs and latency-bound column CTAs share every SM.
template <bool INV, int K, int TOP>
__global__ void __launch_bounds__(kSpThreads, LOLB_SP_CHUNK_MINB)
k_pow2_mix(int64_t* __restrict__ yA, uint32_t* __restrict__ ringA, const int nA, int64_t* __restrict__ yB, uint32_t* __restrict__ ringB,
           const int nB, const int rev_nB, const __grid_constant__ DfParams P)
{
  __shared__ __align__(16) uint32_t U[kSpWarps * kDfUnit];
  // blocks alternate A, B, A, B ... while both kinds have work left
  const int b = (int)blockIdx.x, m = nA < nB ? nA : nB;
  bool isA;
  int idx;
  if (b < 2 * m) { isA = (b & 1) == 0; idx = b >> 1; }
  else { isA = nA > nB; idx = b - m; }
  if (isA) {
    if (!INV) chunk_body<false, K, TOP>(U, idx, yA, P, ringA, 0);
    else col_body<true, K, TOP>(idx, yA, P, ringA, 0);
  } else {
    if (!INV) col_body<false, K, TOP>(idx, yB, P, ringB, rev_nB);
    else chunk_body<true, K, TOP>(U, idx, yB, P, ringB, rev_nB);
  }
}

// ---- graph-scheduled variant: the element base pointer comes from a device cell (the graph is built once per batch size and
// replayed for any y), the sub-batch is an element offset
template <bool INV, int K, int TOP>
__global__ void __launch_bounds__(kSpThreads, LOLB_SP_CHUNK_MINB)
k_pow2_chunk_g(int64_t* const* __restrict__ cell, const int64_t el0, const __grid_constant__ DfParams P, uint32_t* __restrict__ ring, const int rev_n,
               const int total)
{
  __shared__ __align__(16) uint32_t U[kSpWarps * kDfUnit];
  int64_t* y = *cell + (size_t)el0 * ((size_t)K * SpGeom<K, TOP>::N);
  for (int bid = (int)blockIdx.x; bid < total; bid += (int)gridDim.x) {      // a throttled grid walks its tasks
    chunk_body<INV, K, TOP>(U, bid, y, P, ring, rev_n);
    __syncthreads();
  }
}
template <bool INV, int K, int TOP>
__global__ void __launch_bounds__(kSpThreads, LOLB_SP_COL_MINB)
k_pow2_col_g(int64_t* const* __restrict__ cell, const int64_t el0, const __grid_constant__ DfParams P, uint32_t* __restrict__ ring, const int rev_n,
             const int total)
{
  int64_t* y = *cell + (size_t)el0 * ((size_t)K * SpGeom<K, TOP>::N);
  for (int bid = (int)blockIdx.x; bid < total; bid += (int)gridDim.x) col_body<INV, K, TOP>(bid, y, P, ring, rev_n);
}
__global__ void k_set_cell(int64_t** cell, int64_t* y) { *cell = y; }

// one instantiated CUDA graph per (plan, direction, batch, sub-batch, ring depth): nodes A(i) (first kind of sub-batch i) and
// B(i) (second kind), edges A(i) -> B(i) and B(i) -> A(i + R) (ring slot reuse).  Nothing else orders the nodes, so up to R
// first-kind kernels and their second-kind partners are in flight at once: no kernel waits on a stream neighbour, the
// dependency latency of one pair hides behind the others, and the ring (R x sub-batch) stays in L2.
struct SplitGraph {
  const lolb_plan* plan;
  cudaStream_t st;       // an executable graph (and its cell / ring) serves one stream: launches on it are ordered by the stream
  int inv, k, top;
  int64_t batch, S;
  int R;
  int device;
  cudaGraphExec_t exec = nullptr;
  cudaGraph_t graph = nullptr;
  uint32_t* ring = nullptr;
  int64_t** cell = nullptr;
};
std::mutex g_graph_mu;
std::vector<SplitGraph> g_graphs;

template <bool INV, int K, int TOP>
int build_split_graph(const lolb_plan* pl, const DfParams& P, int64_t batch, int64_t S, int R, SplitGraph* out)
{
  typedef SpGeom<K, TOP> Geo;
  const size_t el_words = (size_t)K * Geo::N;
  const int64_t nsub = (batch + S - 1) / S;
  if (R > nsub) R = (int)nsub;
  out->R = R;
  LOLB_CUDA(cudaMalloc((void**)&out->ring, (size_t)R * S * el_words * sizeof(uint32_t)));
  LOLB_CUDA(cudaMalloc((void**)&out->cell, sizeof(int64_t*)));
  LOLB_CUDA(cudaGraphCreate(&out->graph, 0));
  int prio_lo = 0, prio_hi = 0, prio_b = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  { const char* pe = getenv("LOLB_SPLIT_PRIO"); if (pe && pe[0] == '1') prio_b = prio_hi; }      // measured: no effect
  int agrid = 0, bgrid = 0;      // tuning: cap the grid of one kernel kind (CTAs), 0 = one CTA per task
  { const char* ge = getenv("LOLB_SPLIT_AGRID"); if (ge) agrid = atoi(ge); ge = getenv("LOLB_SPLIT_BGRID"); if (ge) bgrid = atoi(ge); }
  std::vector<cudaGraphNode_t> nodeA((size_t)nsub), nodeB((size_t)nsub);
  DfParams Pc = P;
  for (int64_t i = 0; i < nsub; i++) {
    const int64_t cnt = batch - i * S < S ? batch - i * S : S;
    int64_t el0 = i * S;
    uint32_t* rs = out->ring + (size_t)(i % R) * S * el_words;
    int zero = 0, rev = (int)cnt;
    int64_t** cell = out->cell;
    int totA = (int)(cnt * (INV ? Geo::NT_COL : Geo::NT_CHUNK)), totB = (int)(cnt * (INV ? Geo::NT_CHUNK : Geo::NT_COL));
    void* argsA[6] = {&cell, &el0, &Pc, &rs, &zero, &totA};
    void* argsB[6] = {&cell, &el0, &Pc, &rs, &rev, &totB};
    cudaKernelNodeParams kp{};
    kp.blockDim = dim3(kSpThreads); kp.sharedMemBytes = 0; kp.extra = nullptr;
    // A(i): depends on B(i - R)
    kp.func = INV ? (void*)k_pow2_col_g<true, K, TOP> : (void*)k_pow2_chunk_g<false, K, TOP>;
    kp.gridDim = dim3((unsigned)(agrid > 0 && agrid < totA ? agrid : totA));
    kp.kernelParams = argsA;
    cudaGraphNode_t depA[1];
    size_t ndepA = 0;
    if (i >= R) depA[ndepA++] = nodeB[(size_t)(i - R)];
    LOLB_CUDA(cudaGraphAddKernelNode(&nodeA[(size_t)i], out->graph, depA, ndepA, &kp));
    // B(i): depends on A(i)
    kp.func = INV ? (void*)k_pow2_chunk_g<true, K, TOP> : (void*)k_pow2_col_g<false, K, TOP>;
    kp.gridDim = dim3((unsigned)(bgrid > 0 && bgrid < totB ? bgrid : totB));
    kp.kernelParams = argsB;
    cudaGraphNode_t depB[1] = {nodeA[(size_t)i]};
    LOLB_CUDA(cudaGraphAddKernelNode(&nodeB[(size_t)i], out->graph, depB, 1, &kp));
    if (prio_b != 0) {      // the finishing kernels first: they free ring slots, and their CTAs are not crowded out by ready first-kind work
      cudaKernelNodeAttrValue v{};
      v.priority = prio_b;
      LOLB_CUDA(cudaGraphKernelNodeSetAttribute(nodeB[(size_t)i], cudaKernelNodeAttributePriority, &v));
    }
  }
  LOLB_CUDA(cudaGraphInstantiate(&out->exec, out->graph, 0));
  return LOLB_OK;
}

void free_split_graph(SplitGraph& g)
{
  if (g.exec) cudaGraphExecDestroy(g.exec);
  if (g.graph) cudaGraphDestroy(g.graph);
  if (g.ring) cudaFree(g.ring);
  if (g.cell) cudaFree(g.cell);
  g.exec = nullptr; g.graph = nullptr; g.ring = nullptr; g.cell = nullptr;
}

template <bool INV, int K, int TOP>
int launch_split_graph(const lolb_plan* pl, const DfParams& P, int64_t* y, int64_t batch, int64_t S, int R, cudaStream_t st)
{
  std::lock_guard<std::mutex> lock(g_graph_mu);
  SplitGraph* g = nullptr;
  for (auto& e : g_graphs)
    if (e.plan == pl && e.st == st && e.inv == (INV ? 1 : 0) && e.k == K && e.top == TOP && e.batch == batch && e.S == S && e.device == pl->device) { g = &e; break; }
  if (!g) {
    if (g_graphs.size() >= 16) { free_split_graph(g_graphs.front()); g_graphs.erase(g_graphs.begin()); }
    SplitGraph ng{};
    ng.plan = pl; ng.st = st; ng.inv = INV ? 1 : 0; ng.k = K; ng.top = TOP; ng.batch = batch; ng.S = S; ng.device = pl->device;
    int rc = build_split_graph<INV, K, TOP>(pl, P, batch, S, R, &ng);
    if (rc) { free_split_graph(ng); return rc; }
    g_graphs.push_back(ng);
    g = &g_graphs.back();
  }
  k_set_cell<<<1, 1, 0, st>>>(g->cell, y);
  LOLB_CUDA(cudaGraphLaunch(g->exec, st));
  const int64_t nsub = (batch + S - 1) / S;
  count_launch((int)(2 * nsub + 1));
  return LOLB_OK;
}

constexpr int kSpRing = 3;      // sub-batches of ring in flight: one being written, one being read, one of slack

template <bool INV, int K, int TOP>
int launch_split(const lolb_plan* pl, const FusedPow2Df* F, int64_t* y, int64_t batch, cudaStream_t st)
{
  if constexpr ((1 << TOP) * K < kSpWarps || ((1 << TOP) * K) % kSpWarps != 0) return LOLB_FUSED_UNAVAILABLE;
  else {
    typedef SpGeom<K, TOP> Geo;
    const DfParams P = INV ? F->inv : F->fwd;
    const size_t el_words = (size_t)K * Geo::N;
    // Sub-batch size.  Measured at config B (CRT / CRT^-1, % of HBM roofline): 4 MiB 16 / 26, 8 MiB 36 / 38, 16 MiB 38 / 50,
    // 32 MiB 47 / 46, 64 MiB 52 / 50, 128 MiB 54 / 53, 256 MiB 55 / 54, 512 MiB 56.4 / 56.1: every kernel boundary costs
    // about 10 us of drain, fill and cross-stream dependency latency, more than an L2-resident ring wins back, so the
    // sub-batch is as large as a 512 MiB workspace allows and the intermediate mostly travels through HBM (ncu: both kernels
    // run at 78-82 % of the measured HBM peak on 1.5x the algorithmic traffic).
    const char* mb_env = getenv("LOLB_SPLIT_MB");      // tuning / test override of the sub-batch size
    const char* gr_env = getenv("LOLB_SPLIT_GRAPH");   // ring depth of the graph schedule; "0" = off; unset = the measured policy
    // Graph schedule (launch_split_graph): measured at config B, % of HBM roofline CRT / CRT^-1: 2 MiB x 16 slots 61.5 / 52.7,
    // 8 MiB x 6 59.9 / 53.2, 16 MiB x 3 59.8 / 53.6 -- against 56.4 / 57.8 for one 512 MiB sub-batch.  The forward transform
    // (compute-heavy producer, light consumer) gains 5 points from the L2-resident ring; the inverse (light producer,
    // compute-heavy consumer) loses 4, with or without node priorities or throttled grids.  tupSize 1: 67.9 / 59.7 against
    // 65.6 / 60.5.  So: forward transforms at e = 16 take the graph, everything else one large sub-batch.
    auto sub_batch = [&](int64_t mib) {
      int64_t v = (mib << 20) / (int64_t)(el_words * sizeof(uint32_t));
      if (v < 1) v = 1;
      return v > batch ? batch : v;
    };
    const int graph_r = gr_env ? atoi(gr_env) : ((!INV && TOP == 5 && !mb_env) ? 16 : 0);
    if (graph_r > 0) {
      const int64_t Sg = sub_batch(mb_env && atoi(mb_env) > 0 ? atoi(mb_env) : 2);
      if ((batch + Sg - 1) / Sg >= 2) return launch_split_graph<INV, K, TOP>(pl, P, y, batch, Sg, graph_r, st);
    }
    const int64_t S = sub_batch(mb_env && atoi(mb_env) > 0 ? atoi(mb_env) : 512);
    const int64_t nsub = (batch + S - 1) / S;
    const char* mix_env = getenv("LOLB_SPLIT_MIX");
    if (mix_env && mix_env[0] == '1' && nsub >= 2) {
      uint32_t* ring2 = (uint32_t*)plan_ws(pl, st, (size_t)2 * S * el_words * sizeof(uint32_t));
      if (!ring2) return LOLB_ERR_CUDA;
      constexpr int NTA = INV ? Geo::NT_COL : Geo::NT_CHUNK, NTB = INV ? Geo::NT_CHUNK : Geo::NT_COL;
      for (int64_t i = 0; i <= nsub; i++) {      // launch i: first kind of sub-batch i, second kind of sub-batch i - 1
        const int64_t cntA = i < nsub ? (batch - i * S < S ? batch - i * S : S) : 0;
        const int64_t cntB = i > 0 ? (batch - (i - 1) * S < S ? batch - (i - 1) * S : S) : 0;
        int64_t* yA = y + (size_t)i * S * el_words;
        int64_t* yB = y + (size_t)(i > 0 ? i - 1 : 0) * S * el_words;
        uint32_t* rA = ring2 + (size_t)(i & 1) * S * el_words;
        uint32_t* rB = ring2 + (size_t)((i + 1) & 1) * S * el_words;
        const int nA = (int)(cntA * NTA), nB = (int)(cntB * NTB);
        k_pow2_mix<INV, K, TOP><<<(unsigned)(nA + nB), kSpThreads, 0, st>>>(yA, rA, nA, yB, rB, nB, (int)cntB, P);
      }
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return cuda_fail(e, "k_pow2_mix");
      count_launch((int)(nsub + 1));
      return LOLB_OK;
    }
    const int R = (int)(nsub < kSpRing ? nsub : kSpRing);
    uint32_t* ring = (uint32_t*)plan_ws(pl, st, (size_t)R * S * el_words * sizeof(uint32_t));
    if (!ring) return LOLB_ERR_CUDA;
    cudaStream_t aux = nullptr;
    cudaEvent_t* ev = nullptr;      // [0 .. kSpRing): A done, [kSpRing .. 2 kSpRing): B done
    int rc = plan_ws_aux(pl, st, &aux, &ev);
    if (rc) return rc;
    for (int64_t i = 0; i < nsub; i++) {
      const int s = (int)(i % R);
      const int64_t cnt = batch - i * S < S ? batch - i * S : S;
      int64_t* ys = y + (size_t)i * S * el_words;
      uint32_t* rs = ring + (size_t)s * S * el_words;
      if (i >= R) LOLB_CUDA(cudaStreamWaitEvent(st, ev[kSpRing + s], 0));      // B(i - R) has drained this ring slot
      if (!INV) k_pow2_chunk<false, K, TOP><<<(unsigned)(cnt * Geo::NT_CHUNK), kSpThreads, 0, st>>>(ys, P, rs, 0);
      else k_pow2_col<true, K, TOP><<<(unsigned)(cnt * Geo::NT_COL), kSpThreads, 0, st>>>(ys, P, rs, 0);
      LOLB_CUDA(cudaEventRecord(ev[s], st));
      LOLB_CUDA(cudaStreamWaitEvent(aux, ev[s], 0));
      if (!INV) k_pow2_col<false, K, TOP><<<(unsigned)(cnt * Geo::NT_COL), kSpThreads, 0, aux>>>(ys, P, rs, (int)cnt);
      else k_pow2_chunk<true, K, TOP><<<(unsigned)(cnt * Geo::NT_CHUNK), kSpThreads, 0, aux>>>(ys, P, rs, (int)cnt);
      LOLB_CUDA(cudaEventRecord(ev[kSpRing + s], aux));
    }
    LOLB_CUDA(cudaStreamWaitEvent(st, ev[kSpRing + (int)((nsub - 1) % R)], 0));      // join: aux is in order, the last B covers all
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_pow2_chunk / k_pow2_col");
    count_launch((int)(2 * nsub));
    return LOLB_OK;
  }
}

}  // namespace

// drop the graphs of a plan that is being destroyed
void pow2::pow2_split_release(const lolb_plan* pl)
{
  std::lock_guard<std::mutex> lock(g_graph_mu);
  for (size_t i = 0; i < g_graphs.size();) {
    if (g_graphs[i].plan == pl) { free_split_graph(g_graphs[i]); g_graphs.erase(g_graphs.begin() + (long)i); }
    else i++;
  }
}

// LOLB_FUSED_UNAVAILABLE when the shape is not served here (e < 14: the element-resident kernels win)
int pow2::pow2_split_crt(const lolb_plan* pl, const FusedPow2Df* F, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  const int k = pl->k;
#define SP(TOPV)                                                                                                  \
  do {                                                                                                             \
    if (k == 1) return inverse ? launch_split<true, 1, TOPV>(pl, F, y, batch, st) : launch_split<false, 1, TOPV>(pl, F, y, batch, st); \
    if (k == 2) return inverse ? launch_split<true, 2, TOPV>(pl, F, y, batch, st) : launch_split<false, 2, TOPV>(pl, F, y, batch, st); \
    if (k == 4) return inverse ? launch_split<true, 4, TOPV>(pl, F, y, batch, st) : launch_split<false, 4, TOPV>(pl, F, y, batch, st); \
  } while (0)
  switch (F->top) {
    case 3: SP(3); break;
    case 4: SP(4); break;
    case 5: SP(5); break;
    default: break;
  }
#undef SP
  return LOLB_FUSED_UNAVAILABLE;
}

}  // namespace lolb
