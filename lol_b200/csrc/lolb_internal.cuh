// lolb_internal.cuh -- shared declarations of libctensor_b200 (plan, pass lists, device rings).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lol_b200.h"

namespace lolb {

constexpr int kMaxLimbs = 16;     // RNS limbs per element (Haskell tuples in the reference tests use <= 3)
constexpr int kMaxPasses = 96;    // enough for m < 2^31: <= 2 passes per prime-power digit
constexpr int kEngineThreads = 256;
constexpr size_t kSmemBudget = 200 * 1024;   // dynamic shared memory the engine may ask for (<= 227 KB per CTA)

// ---------------------------------------------------------------- passes
// One pass is  I_L (x) A (x) I_R  on the element buffer, A of dimension d, L = n / (d*R).
enum PassKind : int32_t {
  PASS_DFT = 0,      // A = DFT_p            (crt.cpp:131-246)   d = p
  PASS_CRT,          // A = CRT_p            (crt.cpp:248-346)   d = p-1
  PASS_CRTINV,       // A = CRT_p^{-1}'      (crt.cpp:349-457)   d = p-1
  PASS_DIAG,         // A = diag(table)      (crt.cpp:35-126: crtTwiddle / dftTwiddle as a length-d table)
  PASS_L,            // l.cpp:28-57
  PASS_LINV,         // l.cpp:67-98
  PASS_GPOW,         // g.cpp:16-35
  PASS_GDEC,         // g.cpp:37-58
  PASS_GINVPOW,      // g.cpp:60-90
  PASS_GINVDEC,      // g.cpp:92-123
  PASS_NORMSQ,       // norm.cpp:15-37
  PASS_GAUSS,        // random.cpp:19-50   d = p-1
};

struct Pass {
  int32_t kind;
  int32_t p;         // prime
  int32_t d;         // dimension of A
  int32_t R;         // right stride (rts of the reference)
  int32_t rustride;  // stride into the root table for dense passes
  int32_t tab;       // offset of the root / diagonal table inside the per-limb table block
};

struct PassList {
  int32_t count;
  int32_t needs_alt;   // some pass writes out of place
  Pass pass[kMaxPasses];
};

// ---------------------------------------------------------------- per-limb Zq constants
struct ZqConsts {
  uint32_t q[kMaxLimbs];
  uint64_t mu[kMaxLimbs];       // floor(2^64 / q)
  uint32_t scale[kMaxLimbs];    // final per-limb scalar (mhat^-1 or rad_odd^-1), 1 when unused
};

// ---------------------------------------------------------------- host-side plan
enum PlanKind { PLAN_RQ = 1, PLAN_C = 2 };

struct FusedAInfo;   // fused_a.cu

}  // namespace lolb

struct lolb_plan {
  int kind = 0;
  std::vector<PrimeExponent> pe;
  int64_t m = 1;
  int32_t n = 1;        // totient
  int32_t k = 1;        // tupSize
  int64_t odd_rad = 1;
  bool force_generic = false;
  int device = 0;
  int num_sms = 148;

  // ---- Zq
  std::vector<int64_t> qs;
  bool has_fwd = false, has_inv = false;   // CRT tables present (given or derived) per direction
  bool ginv_ok = false;             // rad_odd invertible modulo every q
  std::vector<std::vector<int64_t>> ru, ruinv;   // ABI layout [p^e][k]
  std::vector<int64_t> mhatinv;
  lolb::ZqConsts zq_plain{}, zq_mhat{}, zq_radinv{};   // scale = 1 / mhat^-1 / rad_odd^-1
  uint32_t* d_tab_fwd = nullptr;    // [k][tab_stride_fwd] u32: root tables then diagonal tables
  uint32_t* d_tab_inv = nullptr;
  uint32_t* d_tab_fwd_m = nullptr;  // the same table blocks in Montgomery form (x 2^32 mod q) for engine_axis; nullptr when a modulus is even or >= 2^28
  uint32_t* d_tab_inv_m = nullptr;
  int32_t tab_stride_fwd = 0, tab_stride_inv = 0;
  int64_t* d_gcrt = nullptr;        // [n][k]
  int64_t* d_gcrtinv = nullptr;

  // ---- complex
  double2* d_ctab_fwd = nullptr;    // [k][ctab_stride_fwd]
  double2* d_ctab_inv = nullptr;
  int32_t ctab_stride_fwd = 0, ctab_stride_inv = 0;
  std::vector<std::vector<lolb_complex>> cru, cruinv;   // ABI layout [p^e][k]
  double2 c_mhatinv[lolb::kMaxLimbs];

  // ---- pass lists (shared shapes: CRT lists index d_tab_* / d_ctab_* identically)
  lolb::PassList crt_fwd{}, crt_inv{};
  lolb::PassList line[12]{};        // indexed by PassKind for PASS_L .. PASS_GAUSS (k folded into R for modulus-free rings: see line_folded)
  lolb::PassList line_folded[12]{}; // same passes with R scaled by k, for CTA-per-element execution

  // ---- fused kernels (selected at plan creation; nullptr / 0 = not available)
  void* fused = nullptr;            // lolb::FusedInfo*, owned
  // ---- kernel workspaces (exchange ring + counters of the dataflow kernel, spilled elements of the generic engine):
  // one per CUDA stream the plan has been used on, so that calls on different streams (or from different host threads
  // on different streams) never share mutable device state; calls on ONE stream are ordered by the stream (plan_ws)
  struct WsSlot {
    cudaStream_t st; void* p; size_t bytes;
    cudaStream_t aux = nullptr;       // second stream of the split power-of-two schedule, forked from / joined to `st` by events
    cudaEvent_t ev[8] = {};
  };
  mutable std::vector<WsSlot> ws;
  mutable std::mutex ws_mu;
  // ---- staging for the drop-in (host pointer) entry points and host-batched calls
  mutable void* d_stage = nullptr;
  mutable size_t stage_bytes = 0;
  mutable cudaStream_t streams[3] = {nullptr, nullptr, nullptr};
  mutable cudaEvent_t events[8] = {};
  // identity of caller-supplied tables (drop-in entry points re-use a plan while these match)
  uint64_t fwd_hash = 0, inv_hash = 0;
};

namespace lolb {

// error plumbing (capi.cu)
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);
#define LOLB_CUDA(call)                                                     \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess) return ::lolb::cuda_fail(e__, #call);           \
  } while (0)

// Function attributes (the dynamic shared memory opt-in) and occupancy are per device: a call-site flag that is true the
// first time it is asked on each device
struct PerDeviceOnce {
  std::atomic<unsigned char> done[64] = {};
  bool first()
  {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    return done[d].exchange(1) == 0;
  }
};

// Shape of a CTA of the shared-memory tile kernels (k_line_tile, k_plain_tile): every axis pass deals epb * n / (p - 1) lines to
// the CTA's threads, so line counts that are not multiples of the thread count idle most of the CTA in the last round of a
// pass (n = 1152, three elements, 256 threads: 288 lines of 12 = two rounds for 1.125 rounds of work).  Choose (threads, elements
// per CTA) with the least idle work -- an axis weighs (p - 1) per line, (p - 1)^2 for dense matrices -- and, within 2 %, the
// smallest tile (more CTAs per SM overlap one CTA's loads with another's passes), then the larger CTA.
struct TileShape { int threads, epb; };
inline TileShape choose_tile_shape(int64_t n, const int* p, int cnt, size_t value_bytes, size_t cap_bytes, bool quad, int lines_per_thread = 1, int max_threads = 256)
{
  int64_t cap = (int64_t)(cap_bytes / ((size_t)n * value_bytes));
  if (cap < 1) cap = 1;
  if (cap > 64) cap = 64;
  TileShape best{max_threads, 1};
  double best_w = 1e30;
  for (int64_t e = 1; e <= cap; e++)
    for (int t = max_threads; t >= 128; t -= 32) {
      double used = 0.0, work = 0.0;
      for (int i = 0; i < cnt; i++) {
        const double w = quad ? (double)(p[i] - 1) * (p[i] - 1) : (double)(p[i] - 1);
        const int64_t lines = e * (n / (p[i] - 1));
        const int64_t step = (int64_t)t * lines_per_thread;      // lines a CTA takes per round
        used += w * (double)((lines + step - 1) / step * step);
        work += w * (double)lines;
      }
      const double W = used / work;
      if (W < best_w * 0.98) { best_w = W; best = TileShape{t, (int)e}; }      // e ascending, t descending: ties keep the smaller tile, then the larger CTA
    }
  return best;
}

// plan.cu
int plan_build_common(lolb_plan* pl, const PrimeExponent* pe, int npe, int k);
int plan_derive_rq_roots(lolb_plan* pl);              // ZqBasic.hs:144-171 -> pl->ru, ruinv, mhatinv (LOLB_ERR_NO_CRT if none)
int plan_upload_rq_dir(lolb_plan* pl, bool inverse);  // pl->ru / pl->ruinv -> device root + diagonal tables, pass list
int plan_upload_rq_gcrt(lolb_plan* pl);               // gCRT / gInvCRT vectors from pl->ru (Tensor.hs:319-337)
void plan_derive_c_roots(lolb_plan* pl);              // CRTrans.hs:88-95 -> pl->cru, cruinv, c_mhatinv
int plan_upload_c_dir(lolb_plan* pl, bool inverse);
uint64_t hash_bytes(const void* p, size_t bytes, uint64_t seed);
void* plan_ws(const lolb_plan* pl, cudaStream_t st, size_t bytes);
int plan_ws_aux(const lolb_plan* pl, cudaStream_t st, cudaStream_t* aux, cudaEvent_t** events /* [8], untimed */);   // the stream's workspace, grown to `bytes`; nullptr + error set on failure
int plan_reserve_stage(const lolb_plan* pl, size_t bytes);

// engine.cu -- generic pass engine
enum RingId { RING_ZQ = 0, RING_I64, RING_F64, RING_C64 };
enum Finish : int32_t {
  FIN_NONE = 0,
  FIN_SCALE,        // Zq: * consts.scale[limb];  C64: * cscale[limb] (complex);  F64 unused
  FIN_DIV_EXACT,    // I64: all entries must be multiples of `divisor`; ok[element] says so
  FIN_REAL_SCALE,   // C64: * rscale (real)
};
int engine_crt_zq(const lolb_plan* pl, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);
// engine_axis.cu -- one prime power at a time in registers; -1 = shape not supported, use engine_crt_*
int engine_axis_crt_zq(const lolb_plan* pl, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);
int engine_axis_crt_c(const lolb_plan* pl, bool inverse, double2* y, int64_t batch, cudaStream_t st);
bool engine_axis_supported(const lolb_plan* pl, bool inverse);
int engine_crt_c(const lolb_plan* pl, bool inverse, double2* y, int64_t batch, cudaStream_t st);
int engine_line_zq(const lolb_plan* pl, int kind, const ZqConsts& zc, bool scale, int64_t* y, int64_t batch, cudaStream_t st);
int engine_line_i64(const lolb_plan* pl, int kind, int64_t divisor, int16_t* ok, int64_t* y, int64_t batch, cudaStream_t st);
int engine_line_f64(const lolb_plan* pl, int kind, double* y, int64_t batch, cudaStream_t st);
int engine_line_c64(const lolb_plan* pl, int kind, double rscale, double2* y, int64_t batch, cudaStream_t st);
int engine_gauss(const lolb_plan* pl, double* y, int64_t batch, cudaStream_t st);
int engine_normsq_i64(const lolb_plan* pl, const int64_t* y, int64_t* out, int64_t batch, cudaStream_t st);
int engine_normsq_f64(const lolb_plan* pl, const double* y, double* out, int64_t batch, cudaStream_t st);
int engine_mul_zq(const lolb_plan* pl, int64_t* a, const int64_t* b, int64_t batch, int64_t b_batch, cudaStream_t st);
int engine_mul_c(const lolb_plan* pl, double2* a, const double2* b, int64_t batch, int64_t b_batch, cudaStream_t st);

}  // namespace lolb
