// capi.cu -- the C ABI of libctensor_b200.so (include/lol_b200.h): plan management, the batched
// device-resident operators, the host-batched pipeline and the 29 drop-in symbols of
// lol-cpp/Crypto/Lol/Cyclotomic/Tensor/CPP/Backend.hs:304-337.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>

#include "fused.cuh"
#include "lolb_internal.cuh"
#include "numtheory.h"

using namespace lolb;

// ------------------------------------------------------------------ errors / counters
namespace lolb {

static thread_local std::string t_last_error;
static std::atomic<int64_t> g_launches{0};

void set_error(const std::string& msg) { t_last_error = msg; }

int cuda_fail(cudaError_t e, const char* what)
{
  t_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return LOLB_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace lolb

extern "C" const char* lolb_last_error(void) { return t_last_error.c_str(); }
extern "C" int64_t lolb_kernel_launch_count(void) { return g_launches.load(); }

extern "C" int lolb_device_available(void)
{
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return count > 0 ? 1 : 0;
}

// ------------------------------------------------------------------ plans

static void fill_zq_consts(lolb_plan* pl)
{
  for (int t = 0; t < pl->k; t++) {
    const uint64_t q = (uint64_t)pl->qs[t];
    const uint64_t mu = (uint64_t)((((u128)1) << 64) / q);
    ZqConsts* all[3] = {&pl->zq_plain, &pl->zq_mhat, &pl->zq_radinv};
    for (ZqConsts* c : all) { c->q[t] = (uint32_t)q; c->mu[t] = mu; c->scale[t] = 1u % (uint32_t)q; }
  }
  pl->ginv_ok = true;
  for (int t = 0; t < pl->k; t++) {
    const int64_t inv = mod_inverse(pl->qs[t], pl->odd_rad % pl->qs[t]);   // g.cpp:193-199
    if (inv == 0) pl->ginv_ok = false;
    pl->zq_radinv.scale[t] = (uint32_t)inv;
  }
}

static int copy_tables(const lolb_plan* pl, hInt_t* const* src, std::vector<std::vector<int64_t>>* dst)
{
  const int npe = (int)pl->pe.size();
  dst->assign(npe, {});
  for (int i = 0; i < npe; i++) {
    if (!src[i]) { set_error("plan: NULL root table"); return LOLB_ERR_ARG; }
    const size_t cnt = (size_t)ipow64(pl->pe[i].prime, pl->pe[i].exponent) * pl->k;
    (*dst)[i].assign(src[i], src[i] + cnt);
  }
  return LOLB_OK;
}

static uint64_t hash_tables(const lolb_plan* pl, const void* const* tabs, size_t elem_bytes, const void* extra, size_t extra_bytes)
{
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < pl->pe.size(); i++)
    h = hash_bytes(tabs[i], (size_t)ipow64(pl->pe[i].prime, pl->pe[i].exponent) * pl->k * elem_bytes, h);
  if (extra) h = hash_bytes(extra, extra_bytes, h);
  return h;
}

extern "C" int lolb_plan_create_rq(lolb_plan** out, const PrimeExponent* peArr, hShort_t sizeOfPE, hShort_t tupSize,
                                   const hInt_t* qs, hInt_t* const* ru, hInt_t* const* ruinv, const hInt_t* mhatInv)
{
  if (!out || !qs) { set_error("lolb_plan_create_rq: NULL argument"); return LOLB_ERR_ARG; }
  *out = nullptr;
  lolb_plan* pl = new lolb_plan();
  pl->kind = PLAN_RQ;
  int rc = plan_build_common(pl, peArr, sizeOfPE, tupSize);
  if (rc) { delete pl; return rc; }
  pl->qs.assign(qs, qs + tupSize);
  for (int t = 0; t < tupSize; t++)
    if (qs[t] < 2 || qs[t] >= ((int64_t)1 << 32)) { set_error("lolb_plan_create_rq: modulus out of range [2, 2^32)"); delete pl; return LOLB_ERR_ARG; }
  fill_zq_consts(pl);
  if (ru || ruinv) {
    if (ru) { rc = copy_tables(pl, ru, &pl->ru); if (!rc) rc = plan_upload_rq_dir(pl, false); }
    if (!rc && ruinv) {
      if (!mhatInv) { set_error("lolb_plan_create_rq: ruinv without mhatInv"); rc = LOLB_ERR_ARG; }
      if (!rc) rc = copy_tables(pl, ruinv, &pl->ruinv);
      if (!rc) { pl->mhatinv.assign(mhatInv, mhatInv + tupSize); rc = plan_upload_rq_dir(pl, true); }
    }
    if (!rc && ru) rc = plan_upload_rq_gcrt(pl);
  } else {
    rc = plan_derive_rq_roots(pl);
    if (rc == LOLB_OK) {
      rc = plan_upload_rq_dir(pl, false);
      if (!rc) rc = plan_upload_rq_dir(pl, true);
      if (!rc) rc = plan_upload_rq_gcrt(pl);
    } else if (rc == LOLB_ERR_NO_CRT) {
      rc = LOLB_OK;   // still serves L / G / mul
    }
  }
  if (!rc) rc = fused_select(pl);
  if (rc) { lolb_plan_destroy(pl); return rc; }
  *out = pl;
  return LOLB_OK;
}

static int plan_set_c_tables(lolb_plan* pl, lolb_complex* const* ru, lolb_complex* const* ruinv, const lolb_complex* mhatInv)
{
  const int npe = (int)pl->pe.size();
  int rc = LOLB_OK;
  if (ru) {
    pl->cru.assign(npe, {});
    for (int i = 0; i < npe; i++) {
      const size_t cnt = (size_t)ipow64(pl->pe[i].prime, pl->pe[i].exponent) * pl->k;
      pl->cru[i].assign(ru[i], ru[i] + cnt);
    }
    rc = plan_upload_c_dir(pl, false);
  }
  if (!rc && ruinv) {
    pl->cruinv.assign(npe, {});
    for (int i = 0; i < npe; i++) {
      const size_t cnt = (size_t)ipow64(pl->pe[i].prime, pl->pe[i].exponent) * pl->k;
      pl->cruinv[i].assign(ruinv[i], ruinv[i] + cnt);
    }
    if (mhatInv) for (int t = 0; t < pl->k; t++) pl->c_mhatinv[t] = make_double2(mhatInv[t].real, mhatInv[t].imag);
    rc = plan_upload_c_dir(pl, true);
  }
  if (!rc) rc = fused_select(pl);
  return rc;
}

extern "C" int lolb_plan_create_c(lolb_plan** out, const PrimeExponent* peArr, hShort_t sizeOfPE, hShort_t tupSize)
{
  if (!out) { set_error("lolb_plan_create_c: NULL argument"); return LOLB_ERR_ARG; }
  *out = nullptr;
  lolb_plan* pl = new lolb_plan();
  pl->kind = PLAN_C;
  int rc = plan_build_common(pl, peArr, sizeOfPE, tupSize);
  if (!rc) {
    plan_derive_c_roots(pl);
    rc = plan_upload_c_dir(pl, false);
    if (!rc) rc = plan_upload_c_dir(pl, true);
    if (!rc) rc = fused_select(pl);
  }
  if (rc) { lolb_plan_destroy(pl); return rc; }
  *out = pl;
  return LOLB_OK;
}

extern "C" void lolb_plan_destroy(lolb_plan* pl)
{
  if (!pl) return;
  fused_release(pl);
  void* ptrs[] = {pl->d_tab_fwd, pl->d_tab_inv, pl->d_gcrt, pl->d_gcrtinv, pl->d_ctab_fwd, pl->d_ctab_inv, pl->d_stage,
                  pl->d_tab_fwd_m, pl->d_tab_inv_m};
  for (void* p : ptrs) if (p) cudaFree(p);
  for (auto& w : pl->ws) {
    if (w.p) cudaFree(w.p);
    if (w.aux) cudaStreamDestroy(w.aux);
    for (auto& e : w.ev) if (e) cudaEventDestroy(e);
  }
  for (auto& s : pl->streams) if (s) cudaStreamDestroy(s);
  for (auto& e : pl->events) if (e) cudaEventDestroy(e);
  delete pl;
}

extern "C" int32_t lolb_plan_totient(const lolb_plan* pl) { return pl ? pl->n : 0; }
extern "C" int32_t lolb_plan_tupsize(const lolb_plan* pl) { return pl ? pl->k : 0; }

extern "C" int lolb_plan_get_ru_rq(const lolb_plan* pl, int inverse, int pp_index, hInt_t* out)
{
  if (!pl || pl->kind != PLAN_RQ || !out) { set_error("lolb_plan_get_ru_rq: bad argument"); return LOLB_ERR_ARG; }
  const auto& tabs = inverse ? pl->ruinv : pl->ru;
  if (pp_index < 0 || pp_index >= (int)tabs.size()) return LOLB_ERR_NO_CRT;
  memcpy(out, tabs[pp_index].data(), tabs[pp_index].size() * sizeof(int64_t));
  return LOLB_OK;
}

extern "C" int lolb_plan_get_mhatinv_rq(const lolb_plan* pl, hInt_t* out)
{
  if (!pl || pl->kind != PLAN_RQ || !out) { set_error("lolb_plan_get_mhatinv_rq: bad argument"); return LOLB_ERR_ARG; }
  if ((int)pl->mhatinv.size() != pl->k) return LOLB_ERR_NO_CRT;
  memcpy(out, pl->mhatinv.data(), sizeof(int64_t) * pl->k);
  return LOLB_OK;
}

extern "C" const hInt_t* lolb_plan_gcrt_dev(const lolb_plan* pl, int inverse)
{
  if (!pl) return nullptr;
  return inverse ? pl->d_gcrtinv : pl->d_gcrt;
}

extern "C" void lolb_plan_set_force_generic(lolb_plan* pl, int on) { if (pl) pl->force_generic = on != 0; }

extern "C" const char* lolb_plan_kernel_name(const lolb_plan* pl, const char* op)
{
  if (!pl || !op) return "none";
  if (pl->force_generic) return "generic";
  return fused_kernel_name(pl, op);
}

// ------------------------------------------------------------------ batched operators

// a plan's tables live on the device that was current when it was created: calls from another device are refused
static bool plan_device_ok(const lolb_plan* plan, const char* fn)
{
  int d = -1;
  if (cudaGetDevice(&d) == cudaSuccess && d == plan->device) return true;
  set_error(std::string(fn) + ": plan was created on device " + std::to_string(plan->device) + ", current device is " + std::to_string(d));
  return false;
}

#define REQUIRE_PLAN(KIND)                                                                 \
  if (!plan || plan->kind != (KIND)) { set_error(std::string(__func__) + ": wrong or NULL plan"); return LOLB_ERR_ARG; } \
  if (batch < 0 || (batch > 0 && !y)) { set_error(std::string(__func__) + ": bad batch / NULL data"); return LOLB_ERR_ARG; } \
  if (!plan_device_ok(plan, __func__)) return LOLB_ERR_ARG;

static int crt_rq(const lolb_plan* plan, bool inverse, hInt_t* y, int64_t batch, void* stream)
{
  if (!(inverse ? plan->has_inv : plan->has_fwd)) {
    set_error("no CRT over this modulus / index (ZqBasic.hs:159-165) or tables not supplied");
    return LOLB_ERR_NO_CRT;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (!plan->force_generic) {
    int rc = fused_crt_rq(plan, inverse, y, batch, st);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
    rc = engine_axis_crt_zq(plan, inverse, y, batch, st);
    if (rc != -1) return rc;
  }
  return engine_crt_zq(plan, inverse, y, batch, st);
}

static int line_rq(const lolb_plan* plan, int kind, bool ginv, hInt_t* y, int64_t batch, void* stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  if (ginv && !plan->ginv_ok) return LOLB_ERR_NOT_INVERTIBLE;
  const ZqConsts& zc = ginv ? plan->zq_radinv : plan->zq_plain;
  if (!plan->force_generic) {
    int rc = fused_line_rq(plan, kind, zc, ginv, y, batch, st);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  return engine_line_zq(plan, kind, zc, ginv, y, batch, st);
}

extern "C" int lolb_tensorCRTRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_RQ); return crt_rq(plan, false, y, batch, stream); }
extern "C" int lolb_tensorCRTInvRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_RQ); return crt_rq(plan, true, y, batch, stream); }
extern "C" int lolb_tensorLRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_RQ); return line_rq(plan, PASS_L, false, y, batch, stream); }
extern "C" int lolb_tensorLInvRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_RQ); return line_rq(plan, PASS_LINV, false, y, batch, stream); }
extern "C" int lolb_tensorGPowRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_RQ); return line_rq(plan, PASS_GPOW, false, y, batch, stream); }
extern "C" int lolb_tensorGDecRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_RQ); return line_rq(plan, PASS_GDEC, false, y, batch, stream); }
extern "C" int lolb_tensorGInvPowRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_RQ); return line_rq(plan, PASS_GINVPOW, true, y, batch, stream); }
extern "C" int lolb_tensorGInvDecRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_RQ); return line_rq(plan, PASS_GINVDEC, true, y, batch, stream); }

extern "C" int lolb_mulRq(const lolb_plan* plan, hInt_t* y, const hInt_t* b, int64_t batch, int64_t b_batch, void* stream)
{
  REQUIRE_PLAN(PLAN_RQ);
  if (!b || (b_batch != 1 && b_batch != batch)) { set_error("lolb_mulRq: b_batch must be 1 or batch"); return LOLB_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  if (!plan->force_generic) {
    int rc = fused_mul_rq(plan, y, b, batch, b_batch, st);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  return engine_mul_zq(plan, y, b, batch, b_batch, st);
}

// Fused pair (the callers either side of the CRT).  Cyc's ring product converts a Pow/Dec
// operand to the CRT basis and multiplies pointwise (Cyc.hs:276-297, UCyc.hs:232); key switching multiplies in the CRT
// basis and converts back (SymmSHE.hs:302-314).  One HBM pass instead of two where a fused kernel exists (m = 14400);
// otherwise the same two CUDA kernels back to back.
extern "C" int lolb_crtMulRq(const lolb_plan* plan, hInt_t* y, const hInt_t* b, int64_t batch, int64_t b_batch, void* stream)
{
  REQUIRE_PLAN(PLAN_RQ);
  if (!b || (b_batch != 1 && b_batch != batch)) { set_error("lolb_crtMulRq: b_batch must be 1 or batch"); return LOLB_ERR_ARG; }
  if (!plan->has_fwd) { set_error("no CRT over this modulus / index (ZqBasic.hs:159-165) or tables not supplied"); return LOLB_ERR_NO_CRT; }
  if (!plan->force_generic && b != y) {
    int rc = fused_crt_mul_rq(plan, false, y, b, batch, b_batch, (cudaStream_t)stream);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  int rc = crt_rq(plan, false, y, batch, stream);
  return rc ? rc : lolb_mulRq(plan, y, b, batch, b_batch, stream);
}

extern "C" int lolb_mulCrtInvRq(const lolb_plan* plan, hInt_t* y, const hInt_t* b, int64_t batch, int64_t b_batch, void* stream)
{
  REQUIRE_PLAN(PLAN_RQ);
  if (!b || (b_batch != 1 && b_batch != batch)) { set_error("lolb_mulCrtInvRq: b_batch must be 1 or batch"); return LOLB_ERR_ARG; }
  if (!plan->has_inv) { set_error("no CRT over this modulus / index (ZqBasic.hs:159-165) or tables not supplied"); return LOLB_ERR_NO_CRT; }
  if (!plan->force_generic && b != y) {
    int rc = fused_crt_mul_rq(plan, true, y, b, batch, b_batch, (cudaStream_t)stream);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  int rc = lolb_mulRq(plan, y, b, batch, b_batch, stream);
  return rc ? rc : crt_rq(plan, true, y, batch, stream);
}

// SymmSHE's steps between the CRTs (BASELINE.json configs[3]).  Device pointers, canonical
// residues, the batch layout of every other entry point.
extern "C" int lolb_ctMulRq(const lolb_plan* plan, const hInt_t* a0, const hInt_t* a1, const hInt_t* b0, const hInt_t* b1,
                            hInt_t* d0, hInt_t* d1, hInt_t* d2, int64_t batch, int mul_g, void* stream)
{
  if (!plan || plan->kind != PLAN_RQ) { set_error(std::string(__func__) + ": wrong or NULL plan"); return LOLB_ERR_ARG; }
  if (batch < 0) { set_error(std::string(__func__) + ": bad batch"); return LOLB_ERR_ARG; }
  if (batch > 0 && (!a0 || !a1 || !b0 || !b1 || !d0 || !d1 || !d2)) { set_error("lolb_ctMulRq: null operand"); return LOLB_ERR_ARG; }
  if (mul_g && !plan->d_gcrt) { set_error("lolb_ctMulRq: no gCRT vector (no CRT over this modulus / index)"); return LOLB_ERR_NO_CRT; }
  return she_ct_mul(plan, a0, a1, b0, b1, mul_g ? plan->d_gcrt : nullptr, d0, d1, d2, batch, (cudaStream_t)stream);
}

extern "C" int lolb_gadgetLength(const lolb_plan* plan, int64_t base)
{
  if (!plan || plan->kind != PLAN_RQ) { set_error("lolb_gadgetLength: needs an Rq plan"); return -1; }
  return she_gadget_length(plan, base);
}

extern "C" int lolb_decomposeRq(const lolb_plan* plan, const hInt_t* x, hInt_t* digits, int64_t batch, int64_t base, void* stream)
{
  if (!plan || plan->kind != PLAN_RQ) { set_error(std::string(__func__) + ": wrong or NULL plan"); return LOLB_ERR_ARG; }
  if (batch < 0) { set_error(std::string(__func__) + ": bad batch"); return LOLB_ERR_ARG; }
  if (batch > 0 && (!x || !digits || x == digits)) { set_error("lolb_decomposeRq: x and digits must be distinct device arrays"); return LOLB_ERR_ARG; }
  return she_decompose(plan, x, digits, batch, base, (cudaStream_t)stream);
}

extern "C" int lolb_decomposeCrtRq(const lolb_plan* plan, const hInt_t* x, hInt_t* digits, int64_t batch, int64_t base, void* stream)
{
  if (!plan || plan->kind != PLAN_RQ) { set_error(std::string(__func__) + ": wrong or NULL plan"); return LOLB_ERR_ARG; }
  if (batch < 0) { set_error(std::string(__func__) + ": bad batch"); return LOLB_ERR_ARG; }
  if (batch > 0 && (!x || !digits || x == digits)) { set_error("lolb_decomposeCrtRq: x and digits must be distinct device arrays"); return LOLB_ERR_ARG; }
  if (!plan->has_fwd) { set_error("no CRT over this modulus / index (ZqBasic.hs:159-165) or tables not supplied"); return LOLB_ERR_NO_CRT; }
  int nd[kMaxLimbs], shift = -1;
  const int ell = she_gadget_digits(plan, base, nd, &shift);
  if (ell < 0) return LOLB_ERR_ARG;
  if (!plan->force_generic) {
    int rc = fused_decompose_crt_rq(plan, x, digits, batch, base, (cudaStream_t)stream);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  int rc = she_decompose(plan, x, digits, batch, base, (cudaStream_t)stream);
  return rc ? rc : crt_rq(plan, false, digits, batch * ell, stream);
}

extern "C" int lolb_knapsackRq(const lolb_plan* plan, const hInt_t* digits, int ell, const hInt_t* hints, hInt_t* c0, hInt_t* c1,
                               int64_t batch, void* stream)
{
  if (!plan || plan->kind != PLAN_RQ) { set_error(std::string(__func__) + ": wrong or NULL plan"); return LOLB_ERR_ARG; }
  if (batch < 0) { set_error(std::string(__func__) + ": bad batch"); return LOLB_ERR_ARG; }
  if (batch > 0 && ell > 0 && (!digits || !hints || !c0 || !c1 || c0 == c1)) { set_error("lolb_knapsackRq: null or aliased operand"); return LOLB_ERR_ARG; }
  return she_knapsack(plan, digits, ell, hints, c0, c1, batch, (cudaStream_t)stream);
}

// modulus-free rings: streaming kernel when the index has one or two small odd primes, generic engine otherwise
static int plain_line(const lolb_plan* plan, int ring, int kind, void* y, int64_t batch, double rscale, void* stream)
{
  cudaStream_t st = (cudaStream_t)stream;
  if (!plan->force_generic) {
    int rc = fused_plain_line(plan, ring, kind, y, batch, rscale, st);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  if (ring == RING_I64) return engine_line_i64(plan, kind, 0, nullptr, (int64_t*)y, batch, st);
  if (ring == RING_F64) return engine_line_f64(plan, kind, (double*)y, batch, st);
  return engine_line_c64(plan, kind, rscale, (double2*)y, batch, st);
}

extern "C" int lolb_tensorLR(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return plain_line(plan, RING_I64, PASS_L, y, batch, 0.0, stream); }
extern "C" int lolb_tensorLInvR(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return plain_line(plan, RING_I64, PASS_LINV, y, batch, 0.0, stream); }
extern "C" int lolb_tensorGPowR(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return plain_line(plan, RING_I64, PASS_GPOW, y, batch, 0.0, stream); }
extern "C" int lolb_tensorGDecR(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return plain_line(plan, RING_I64, PASS_GDEC, y, batch, 0.0, stream); }
extern "C" int lolb_tensorGInvPowR(const lolb_plan* plan, hInt_t* y, hShort_t* ok, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return engine_line_i64(plan, PASS_GINVPOW, plan->odd_rad, ok, y, batch, (cudaStream_t)stream); }
extern "C" int lolb_tensorGInvDecR(const lolb_plan* plan, hInt_t* y, hShort_t* ok, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return engine_line_i64(plan, PASS_GINVDEC, plan->odd_rad, ok, y, batch, (cudaStream_t)stream); }
extern "C" int lolb_tensorLDouble(const lolb_plan* plan, double* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return plain_line(plan, RING_F64, PASS_L, y, batch, 0.0, stream); }
extern "C" int lolb_tensorLInvDouble(const lolb_plan* plan, double* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return plain_line(plan, RING_F64, PASS_LINV, y, batch, 0.0, stream); }
extern "C" int lolb_tensorLC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return plain_line(plan, RING_C64, PASS_L, y, batch, 0.0, stream); }
extern "C" int lolb_tensorLInvC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return plain_line(plan, RING_C64, PASS_LINV, y, batch, 0.0, stream); }
extern "C" int lolb_tensorGPowC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return plain_line(plan, RING_C64, PASS_GPOW, y, batch, 0.0, stream); }
extern "C" int lolb_tensorGDecC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return plain_line(plan, RING_C64, PASS_GDEC, y, batch, 0.0, stream); }
extern "C" int lolb_tensorGInvPowC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return plain_line(plan, RING_C64, PASS_GINVPOW, y, batch, 1.0 / (double)plan->odd_rad, stream); }
extern "C" int lolb_tensorGInvDecC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream)
{ REQUIRE_PLAN(PLAN_C); return plain_line(plan, RING_C64, PASS_GINVDEC, y, batch, 1.0 / (double)plan->odd_rad, stream); }

extern "C" int lolb_tensorCRTC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream)
{
  REQUIRE_PLAN(PLAN_C);
  if (!plan->has_fwd) return LOLB_ERR_NO_CRT;
  if (!plan->force_generic) {
    int rc = fused_crt_c(plan, false, (double2*)y, batch, (cudaStream_t)stream);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
    rc = engine_axis_crt_c(plan, false, (double2*)y, batch, (cudaStream_t)stream);
    if (rc != -1) return rc;
  }
  return engine_crt_c(plan, false, (double2*)y, batch, (cudaStream_t)stream);
}
extern "C" int lolb_tensorCRTInvC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream)
{
  REQUIRE_PLAN(PLAN_C);
  if (!plan->has_inv) return LOLB_ERR_NO_CRT;
  if (!plan->force_generic) {
    int rc = fused_crt_c(plan, true, (double2*)y, batch, (cudaStream_t)stream);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
    rc = engine_axis_crt_c(plan, true, (double2*)y, batch, (cudaStream_t)stream);
    if (rc != -1) return rc;
  }
  return engine_crt_c(plan, true, (double2*)y, batch, (cudaStream_t)stream);
}
extern "C" int lolb_mulC(const lolb_plan* plan, lolb_complex* y, const lolb_complex* b, int64_t batch, int64_t b_batch, void* stream)
{
  REQUIRE_PLAN(PLAN_C);
  if (!b || (b_batch != 1 && b_batch != batch)) { set_error("lolb_mulC: b_batch must be 1 or batch"); return LOLB_ERR_ARG; }
  return engine_mul_c(plan, (double2*)y, (const double2*)b, batch, b_batch, (cudaStream_t)stream);
}
extern "C" int lolb_tensorGaussianDec(const lolb_plan* plan, double* y, int64_t batch, void* stream)
{
  REQUIRE_PLAN(PLAN_C);
  if (!plan->has_fwd) return LOLB_ERR_NO_CRT;
  if (!plan->force_generic) {
    int rc = fused_plain_gauss(plan, y, batch, (cudaStream_t)stream);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  return engine_gauss(plan, y, batch, (cudaStream_t)stream);
}
// GaussRandom.hs:52-59 on the device: y[batch][n] i.i.d. Gaussians of scaled variance svar (true variance svar / (2 pi))
extern "C" int lolb_realGaussians(double svar, uint64_t seed, uint64_t first_element, double* y, int64_t n, int64_t batch, void* stream)
{
  if (n < 0 || batch < 0 || (n > 0 && batch > 0 && !y) || !(svar >= 0.0)) { set_error("lolb_realGaussians: bad argument"); return LOLB_ERR_ARG; }
  return fused_plain_real_gaussians(nullptr, y, n, batch, seed, first_element, svar / 3.14159265358979323846, (cudaStream_t)stream);
}

// tGaussianDec (Tensor.hs:143; CPP.hs:376-389): n reals of scaled variance v * m / rad(m), then the E_m transform
// (tensorGaussianDec), both on the device -- one pass where the streaming kernel serves the index, two otherwise
extern "C" int lolb_tGaussianDec(const lolb_plan* plan, double v, uint64_t seed, uint64_t first_element, double* y, int64_t batch, void* stream)
{
  REQUIRE_PLAN(PLAN_C);
  if (!plan->has_fwd) return LOLB_ERR_NO_CRT;
  if (!(v >= 0.0)) { set_error("lolb_tGaussianDec: negative variance"); return LOLB_ERR_ARG; }
  int64_t rad = 1;
  for (const PrimeExponent& pe : plan->pe) rad *= pe.prime;
  const double var2 = v * (double)(plan->m / rad) / 3.14159265358979323846;      // twice the true variance
  cudaStream_t st = (cudaStream_t)stream;
  if (!plan->force_generic) {
    int rc = fused_plain_gauss_gen(plan, y, batch, st, true, seed, first_element, var2);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  int rc = fused_plain_real_gaussians(plan, y, (int64_t)plan->n * plan->k, batch, seed, first_element, var2, st);
  return rc ? rc : lolb_tensorGaussianDec(plan, y, batch, stream);
}

extern "C" int lolb_tensorNormSqR(const lolb_plan* plan, const hInt_t* y, hInt_t* out, int64_t batch, void* stream)
{
  REQUIRE_PLAN(PLAN_C);
  if (batch > 0 && !out) { set_error("lolb_tensorNormSqR: NULL out"); return LOLB_ERR_ARG; }
  if (!plan->force_generic) {
    int rc = fused_plain_normsq_i64(plan, y, out, batch, (cudaStream_t)stream);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  return engine_normsq_i64(plan, y, out, batch, (cudaStream_t)stream);
}
extern "C" int lolb_tensorNormSqD(const lolb_plan* plan, const double* y, double* out, int64_t batch, void* stream)
{
  REQUIRE_PLAN(PLAN_C);
  if (batch > 0 && !out) { set_error("lolb_tensorNormSqD: NULL out"); return LOLB_ERR_ARG; }
  if (!plan->force_generic) {
    int rc = fused_plain_normsq_f64(plan, y, out, batch, (cudaStream_t)stream);
    if (rc != LOLB_FUSED_UNAVAILABLE) return rc;
  }
  return engine_normsq_f64(plan, y, out, batch, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ host-batched pipeline

// Device memory for callers without a CUDA binding of their own (the Haskell instance keeps ring elements in GPU memory behind a
// ForeignPtr whose finalizer is lolb_dev_free).  Plain cudaMalloc / cudaFree / cudaMemcpy on the current device.
extern "C" void* lolb_dev_alloc(uint64_t bytes)
{
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { cuda_fail(cudaGetLastError(), "cudaMalloc"); return nullptr; }
  return p;
}
extern "C" void lolb_dev_free(void* p) { if (p) cudaFree(p); }
extern "C" int lolb_dev_upload(void* dst_dev, const void* src_host, uint64_t bytes, void* stream)
{
  if (bytes && (!dst_dev || !src_host)) { set_error("lolb_dev_upload: NULL pointer"); return LOLB_ERR_ARG; }
  LOLB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return LOLB_OK;
}
extern "C" int lolb_dev_download(void* dst_host, const void* src_dev, uint64_t bytes, void* stream)
{
  if (bytes && (!dst_host || !src_dev)) { set_error("lolb_dev_download: NULL pointer"); return LOLB_ERR_ARG; }
  LOLB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  LOLB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));      // the host reads the result next
  return LOLB_OK;
}
extern "C" int lolb_dev_copy(void* dst_dev, const void* src_dev, uint64_t bytes, void* stream)
{
  if (bytes && (!dst_dev || !src_dev)) { set_error("lolb_dev_copy: NULL pointer"); return LOLB_ERR_ARG; }
  LOLB_CUDA(cudaMemcpyAsync(dst_dev, src_dev, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return LOLB_OK;
}

extern "C" void* lolb_host_alloc(uint64_t bytes)
{
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
extern "C" void lolb_host_free(void* p) { if (p) cudaFreeHost(p); }

static int apply_named_rq(const lolb_plan* plan, const std::string& op, hInt_t* y, int64_t batch, cudaStream_t st)
{
  if (op == "CRT") return lolb_tensorCRTRq(plan, y, batch, st);
  if (op == "CRTInv") return lolb_tensorCRTInvRq(plan, y, batch, st);
  if (op == "L") return lolb_tensorLRq(plan, y, batch, st);
  if (op == "LInv") return lolb_tensorLInvRq(plan, y, batch, st);
  if (op == "GPow") return lolb_tensorGPowRq(plan, y, batch, st);
  if (op == "GDec") return lolb_tensorGDecRq(plan, y, batch, st);
  if (op == "GInvPow") return lolb_tensorGInvPowRq(plan, y, batch, st);
  if (op == "GInvDec") return lolb_tensorGInvDecRq(plan, y, batch, st);
  if (op == "MulGCRT") return lolb_mulRq(plan, y, plan->d_gcrt, batch, 1, st);
  if (op == "DivGCRT") return lolb_mulRq(plan, y, plan->d_gcrtinv, batch, 1, st);
  set_error("lolb_rq_apply_host: unknown operator '" + op + "'");
  return LOLB_ERR_ARG;
}

// u32 wire format <-> the int64 ABI layout on the device (lolb_rq_apply_host_u32)
__global__ void k_widen_u32(const uint32_t* __restrict__ src, int64_t* __restrict__ dst, int64_t n4)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldcs(reinterpret_cast<const uint4*>(src) + i);
    longlong2* o = reinterpret_cast<longlong2*>(dst) + 2 * i;
    o[0] = make_longlong2((int64_t)v.x, (int64_t)v.y);
    o[1] = make_longlong2((int64_t)v.z, (int64_t)v.w);
  }
}
__global__ void k_narrow_u32(const int64_t* __restrict__ src, uint32_t* __restrict__ dst, int64_t n4)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const longlong2* in = reinterpret_cast<const longlong2*>(src) + 2 * i;
    const longlong2 a = in[0], b = in[1];
    __stcs(reinterpret_cast<uint4*>(dst) + i, make_uint4((uint32_t)a.x, (uint32_t)a.y, (uint32_t)b.x, (uint32_t)b.y));
  }
}

// Three-slot ring: slot s carries chunk c = s (mod 3) through  H2D -> kernels -> D2H  on its own stream, so
// the copy engines (one per direction) and the SMs all stay busy.  WIRE = bytes per coefficient on the host side.
template <class HostT>
static int apply_host_pipeline(const lolb_plan* plan, const char* fn, const char* ops, HostT* y, int64_t batch)
{
  if (!plan || plan->kind != PLAN_RQ || !ops || batch < 0 || (batch > 0 && !y)) { set_error(std::string(fn) + ": bad argument"); return LOLB_ERR_ARG; }
  if (!plan_device_ok(plan, fn)) return LOLB_ERR_ARG;
  if (batch == 0) return LOLB_OK;
  constexpr bool NARROW = sizeof(HostT) == 4;
  std::vector<std::string> names;
  {
    std::string cur;
    for (const char* c = ops;; c++) {
      if (*c == ',' || *c == 0) { if (!cur.empty()) names.push_back(cur); cur.clear(); if (*c == 0) break; }
      else if (*c != ' ') cur.push_back(*c);
    }
  }
  const size_t elem_words = (size_t)plan->n * plan->k;
  const size_t elem_bytes = elem_words * sizeof(int64_t), wire_bytes = elem_words * sizeof(HostT);
  if (NARROW && elem_words % 4 != 0) { set_error(std::string(fn) + ": the u32 wire format needs totm * tupSize to be a multiple of 4"); return LOLB_ERR_ARG; }
  const int slots = 3;
  static size_t slot_mib = 0;                                          // LOLB_STAGE_MIB: tuning override
  if (!slot_mib) { const char* v = getenv("LOLB_STAGE_MIB"); slot_mib = v && atoi(v) > 0 ? (size_t)atoi(v) : 96; }
  int64_t chunk = (int64_t)((slot_mib << 20) / elem_bytes);            // ~96 MiB of int64 elements per slot
  if (chunk < 1) chunk = 1;
  if (chunk > batch) chunk = batch;
  const size_t slot_bytes = (size_t)chunk * (elem_bytes + (NARROW ? wire_bytes : 0));
  int rc = plan_reserve_stage(plan, (size_t)slots * slot_bytes);
  if (rc) return rc;
  for (int s = 0; s < slots; s++)
    if (!plan->streams[s]) LOLB_CUDA(cudaStreamCreateWithFlags(&plan->streams[s], cudaStreamNonBlocking));
  int64_t done = 0;
  for (int c = 0; done < batch; c++) {
    const int s = c % slots;
    const int64_t cnt = (batch - done < chunk) ? batch - done : chunk;
    hInt_t* dev = (hInt_t*)((char*)plan->d_stage + (size_t)s * slot_bytes);
    HostT* wire = NARROW ? (HostT*)((char*)dev + (size_t)chunk * elem_bytes) : (HostT*)dev;
    HostT* host = y + (size_t)done * elem_words;
    cudaStream_t st = plan->streams[s];       // stream order makes slot reuse safe
    LOLB_CUDA(cudaMemcpyAsync(wire, host, (size_t)cnt * wire_bytes, cudaMemcpyHostToDevice, st));
    const int64_t n4 = (int64_t)(cnt * elem_words / 4);
    const int blocks = (int)std::min<int64_t>((n4 + 255) / 256, (int64_t)plan->num_sms * 8);
    if (NARROW && !names.empty()) { k_widen_u32<<<blocks, 256, 0, st>>>((const uint32_t*)wire, dev, n4); count_launch(); }
    for (const std::string& op : names) {      // kernel workspaces are per stream (plan_ws): the slots never share one
      rc = apply_named_rq(plan, op, dev, cnt, st);
      if (rc) break;
    }
    if (rc) { cudaDeviceSynchronize(); return rc; }
    if (NARROW && !names.empty()) { k_narrow_u32<<<blocks, 256, 0, st>>>(dev, (uint32_t*)wire, n4); count_launch(); }
    LOLB_CUDA(cudaMemcpyAsync(host, wire, (size_t)cnt * wire_bytes, cudaMemcpyDeviceToHost, st));
    done += cnt;
  }
  for (int s = 0; s < slots; s++) LOLB_CUDA(cudaStreamSynchronize(plan->streams[s]));
  LOLB_CUDA(cudaGetLastError());
  return LOLB_OK;
}

extern "C" int lolb_rq_apply_host(const lolb_plan* plan, const char* ops, hInt_t* y, int64_t batch)
{ return apply_host_pipeline<hInt_t>(plan, __func__, ops, y, batch); }

extern "C" int lolb_rq_apply_host_u32(const lolb_plan* plan, const char* ops, uint32_t* y, int64_t batch)
{ return apply_host_pipeline<uint32_t>(plan, __func__, ops, y, batch); }

// ------------------------------------------------------------------ drop-in symbols (host pointers, one element)

namespace {

std::mutex g_mutex;
std::map<std::string, lolb_plan*> g_plans;

[[noreturn]] void die(const char* sym, int rc)
{
  fprintf(stderr, "libctensor_b200: %s failed (status %d): %s\n", sym, rc, lolb_last_error());
  fprintf(stderr, "libctensor_b200 has no CPU path; a CUDA device (sm_100a) is required.\n");
  abort();
}

std::string plan_key(int kind, hShort_t k, const PrimeExponent* pe, hShort_t npe, const hInt_t* qs)
{
  std::string key((const char*)&kind, sizeof(kind));
  int dev = 0;
  cudaGetDevice(&dev);                               // plans (tables, staging) are per device
  key.append((const char*)&dev, sizeof(dev));
  key.append((const char*)&k, sizeof(k));
  if (npe > 0) key.append((const char*)pe, sizeof(PrimeExponent) * (size_t)npe);
  if (qs) key.append((const char*)qs, sizeof(hInt_t) * (size_t)k);
  return key;
}

void check_totm(const char* sym, const lolb_plan* pl, hDim_t totm)
{
  if (pl->n != totm) {
    set_error("totm does not match the totient of the prime powers");
    die(sym, LOLB_ERR_ARG);
  }
}

// plan without CRT tables (L, G, mul need none); tables are attached lazily by the CRT symbols
lolb_plan* legacy_plan_rq(const char* sym, hShort_t k, const PrimeExponent* pe, hShort_t npe, const hInt_t* qs, hDim_t totm)
{
  const std::string key = plan_key(PLAN_RQ, k, pe, npe, qs);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) { check_totm(sym, it->second, totm); return it->second; }
  lolb_plan* pl = new lolb_plan();
  pl->kind = PLAN_RQ;
  int rc = plan_build_common(pl, pe, npe, k);
  if (rc) die(sym, rc);
  pl->qs.assign(qs, qs + k);
  for (int t = 0; t < k; t++)
    if (qs[t] < 2 || qs[t] >= ((int64_t)1 << 32)) { set_error("modulus out of range [2, 2^32)"); die(sym, LOLB_ERR_ARG); }
  fill_zq_consts(pl);
  rc = fused_select(pl);
  if (rc) die(sym, rc);
  check_totm(sym, pl, totm);
  g_plans[key] = pl;
  return pl;
}

lolb_plan* legacy_plan_c(const char* sym, hShort_t k, const PrimeExponent* pe, hShort_t npe, hDim_t totm)
{
  const std::string key = plan_key(PLAN_C, k, pe, npe, nullptr);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) { check_totm(sym, it->second, totm); return it->second; }
  lolb_plan* pl = nullptr;
  int rc = lolb_plan_create_c(&pl, pe, npe, k);
  if (rc) die(sym, rc);
  check_totm(sym, pl, totm);
  g_plans[key] = pl;
  return pl;
}

// stage `bytes` of host data on the device, run `body(dev)`, copy back `back_bytes`
template <class Body>
void with_staged(const char* sym, const lolb_plan* pl, void* host, size_t bytes, size_t back_bytes, Body body)
{
  int rc = plan_reserve_stage(pl, bytes);
  if (rc) die(sym, rc);
  cudaError_t e = cudaMemcpy(pl->d_stage, host, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) die(sym, cuda_fail(e, "cudaMemcpy H2D"));
  rc = body(pl->d_stage);
  if (rc) { cudaDeviceSynchronize(); die(sym, rc); }
  e = cudaMemcpy(host, pl->d_stage, back_bytes, cudaMemcpyDeviceToHost);   // synchronises with the null stream
  if (e != cudaSuccess) die(sym, cuda_fail(e, "cudaMemcpy D2H"));
}

}  // namespace

#define LEGACY_LOCK std::lock_guard<std::mutex> lock__(g_mutex)

extern "C" void tensorCRTRq(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, hInt_t** ru, hInt_t* qs)
{
  LEGACY_LOCK;
  lolb_plan* pl = legacy_plan_rq(__func__, tupSize, peArr, sizeOfPE, qs, totm);
  const uint64_t h = hash_tables(pl, (const void* const*)ru, sizeof(hInt_t), nullptr, 0);
  if (!pl->has_fwd || pl->fwd_hash != h) {
    int rc = copy_tables(pl, ru, &pl->ru);
    if (!rc) rc = plan_upload_rq_dir(pl, false);
    if (!rc) rc = fused_select(pl);
    if (rc) die(__func__, rc);
    pl->fwd_hash = h;
  }
  const size_t bytes = (size_t)totm * tupSize * sizeof(hInt_t);
  with_staged(__func__, pl, y, bytes, bytes, [&](void* d) { return lolb_tensorCRTRq(pl, (hInt_t*)d, 1, nullptr); });
}

extern "C" void tensorCRTInvRq(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, hInt_t** ruinv, hInt_t* mhatInv, hInt_t* qs)
{
  LEGACY_LOCK;
  lolb_plan* pl = legacy_plan_rq(__func__, tupSize, peArr, sizeOfPE, qs, totm);
  const uint64_t h = hash_tables(pl, (const void* const*)ruinv, sizeof(hInt_t), mhatInv, sizeof(hInt_t) * (size_t)tupSize);
  if (!pl->has_inv || pl->inv_hash != h) {
    int rc = copy_tables(pl, ruinv, &pl->ruinv);
    pl->mhatinv.assign(mhatInv, mhatInv + tupSize);
    if (!rc) rc = plan_upload_rq_dir(pl, true);
    if (!rc) rc = fused_select(pl);
    if (rc) die(__func__, rc);
    pl->inv_hash = h;
  }
  const size_t bytes = (size_t)totm * tupSize * sizeof(hInt_t);
  with_staged(__func__, pl, y, bytes, bytes, [&](void* d) { return lolb_tensorCRTInvRq(pl, (hInt_t*)d, 1, nullptr); });
}

extern "C" void tensorCRTC(hShort_t tupSize, lolb_complex* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, lolb_complex** ru)
{
  LEGACY_LOCK;
  lolb_plan* pl = legacy_plan_c(__func__, tupSize, peArr, sizeOfPE, totm);
  const uint64_t h = hash_tables(pl, (const void* const*)ru, sizeof(lolb_complex), nullptr, 0);
  if (pl->fwd_hash != h) {
    int rc = plan_set_c_tables(pl, ru, nullptr, nullptr);
    if (rc) die(__func__, rc);
    pl->fwd_hash = h;
  }
  const size_t bytes = (size_t)totm * tupSize * sizeof(lolb_complex);
  with_staged(__func__, pl, y, bytes, bytes, [&](void* d) { return lolb_tensorCRTC(pl, (lolb_complex*)d, 1, nullptr); });
}

extern "C" void tensorCRTInvC(hShort_t tupSize, lolb_complex* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, lolb_complex** ruinv, lolb_complex* mhatInv)
{
  LEGACY_LOCK;
  lolb_plan* pl = legacy_plan_c(__func__, tupSize, peArr, sizeOfPE, totm);
  const uint64_t h = hash_tables(pl, (const void* const*)ruinv, sizeof(lolb_complex), mhatInv, sizeof(lolb_complex) * (size_t)tupSize);
  if (pl->inv_hash != h) {
    int rc = plan_set_c_tables(pl, nullptr, ruinv, mhatInv);
    if (rc) die(__func__, rc);
    pl->inv_hash = h;
  }
  const size_t bytes = (size_t)totm * tupSize * sizeof(lolb_complex);
  with_staged(__func__, pl, y, bytes, bytes, [&](void* d) { return lolb_tensorCRTInvC(pl, (lolb_complex*)d, 1, nullptr); });
}

extern "C" void tensorGaussianDec(hShort_t tupSize, double* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, lolb_complex** ru)
{
  LEGACY_LOCK;
  lolb_plan* pl = legacy_plan_c(__func__, tupSize, peArr, sizeOfPE, totm);
  const uint64_t h = hash_tables(pl, (const void* const*)ru, sizeof(lolb_complex), nullptr, 0);
  if (pl->fwd_hash != h) {
    int rc = plan_set_c_tables(pl, ru, nullptr, nullptr);
    if (rc) die(__func__, rc);
    pl->fwd_hash = h;
  }
  const size_t bytes = (size_t)totm * tupSize * sizeof(double);
  with_staged(__func__, pl, y, bytes, bytes, [&](void* d) { return lolb_tensorGaussianDec(pl, (double*)d, 1, nullptr); });
}

#define LEGACY_RQ(NAME)                                                                                               \
  extern "C" void tensor##NAME##Rq(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, hInt_t* qs) \
  {                                                                                                                   \
    LEGACY_LOCK;                                                                                                      \
    lolb_plan* pl = legacy_plan_rq(__func__, tupSize, peArr, sizeOfPE, qs, totm);                                     \
    const size_t bytes = (size_t)totm * tupSize * sizeof(hInt_t);                                                     \
    with_staged(__func__, pl, y, bytes, bytes, [&](void* d) { return lolb_tensor##NAME##Rq(pl, (hInt_t*)d, 1, nullptr); }); \
  }
LEGACY_RQ(L)
LEGACY_RQ(LInv)
LEGACY_RQ(GPow)
LEGACY_RQ(GDec)

#define LEGACY_GINV_RQ(NAME)                                                                                          \
  extern "C" hShort_t tensor##NAME##Rq(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, hInt_t* qs) \
  {                                                                                                                   \
    LEGACY_LOCK;                                                                                                      \
    lolb_plan* pl = legacy_plan_rq(__func__, tupSize, peArr, sizeOfPE, qs, totm);                                     \
    if (!pl->ginv_ok) return 0;   /* g.cpp:196-198: reciprocal == 0 */                                               \
    const size_t bytes = (size_t)totm * tupSize * sizeof(hInt_t);                                                     \
    with_staged(__func__, pl, y, bytes, bytes, [&](void* d) { return lolb_tensor##NAME##Rq(pl, (hInt_t*)d, 1, nullptr); }); \
    return 1;                                                                                                         \
  }
LEGACY_GINV_RQ(GInvPow)
LEGACY_GINV_RQ(GInvDec)

#define LEGACY_PLAIN(NAME, TAG, T, CALLT)                                                                             \
  extern "C" void tensor##NAME##TAG(hShort_t tupSize, T* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE)     \
  {                                                                                                                   \
    LEGACY_LOCK;                                                                                                      \
    lolb_plan* pl = legacy_plan_c(__func__, tupSize, peArr, sizeOfPE, totm);                                          \
    const size_t bytes = (size_t)totm * tupSize * sizeof(T);                                                          \
    with_staged(__func__, pl, y, bytes, bytes, [&](void* d) { return lolb_tensor##NAME##TAG(pl, (CALLT*)d, 1, nullptr); }); \
  }
LEGACY_PLAIN(L, R, hInt_t, hInt_t)
LEGACY_PLAIN(LInv, R, hInt_t, hInt_t)
LEGACY_PLAIN(GPow, R, hInt_t, hInt_t)
LEGACY_PLAIN(GDec, R, hInt_t, hInt_t)
LEGACY_PLAIN(L, Double, double, double)
LEGACY_PLAIN(LInv, Double, double, double)
LEGACY_PLAIN(L, C, lolb_complex, lolb_complex)
LEGACY_PLAIN(LInv, C, lolb_complex, lolb_complex)
LEGACY_PLAIN(GPow, C, lolb_complex, lolb_complex)
LEGACY_PLAIN(GDec, C, lolb_complex, lolb_complex)

#define LEGACY_GINV_C(NAME)                                                                                           \
  extern "C" hShort_t tensor##NAME##C(hShort_t tupSize, lolb_complex* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE) \
  {                                                                                                                   \
    LEGACY_LOCK;                                                                                                      \
    lolb_plan* pl = legacy_plan_c(__func__, tupSize, peArr, sizeOfPE, totm);                                          \
    const size_t bytes = (size_t)totm * tupSize * sizeof(lolb_complex);                                               \
    with_staged(__func__, pl, y, bytes, bytes, [&](void* d) { return lolb_tensor##NAME##C(pl, (lolb_complex*)d, 1, nullptr); }); \
    return 1;                                                                                                         \
  }
LEGACY_GINV_C(GInvPow)
LEGACY_GINV_C(GInvDec)

// element at [0, bytes), one int16 status right behind it (8-byte aligned)
#define LEGACY_GINV_R(NAME)                                                                                           \
  extern "C" hShort_t tensor##NAME##R(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE) \
  {                                                                                                                   \
    LEGACY_LOCK;                                                                                                      \
    lolb_plan* pl = legacy_plan_c(__func__, tupSize, peArr, sizeOfPE, totm);                                          \
    const size_t bytes = (size_t)totm * tupSize * sizeof(hInt_t);                                                     \
    int rc = plan_reserve_stage(pl, bytes + 8);                                                                       \
    if (rc) die(__func__, rc);                                                                                        \
    hShort_t* ok_dev = (hShort_t*)((char*)pl->d_stage + bytes);                                                       \
    hShort_t ok = 0;                                                                                                  \
    with_staged(__func__, pl, y, bytes, bytes, [&](void* d) { return lolb_tensor##NAME##R(pl, (hInt_t*)d, ok_dev, 1, nullptr); }); \
    cudaError_t e = cudaMemcpy(&ok, ok_dev, sizeof(ok), cudaMemcpyDeviceToHost);                                      \
    if (e != cudaSuccess) die(__func__, cuda_fail(e, "cudaMemcpy status"));                                           \
    return ok;                                                                                                        \
  }
LEGACY_GINV_R(GInvPow)
LEGACY_GINV_R(GInvDec)

#define LEGACY_NORM(TAG, T)                                                                                           \
  extern "C" void tensorNormSq##TAG(hShort_t tupSize, T* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE)     \
  {                                                                                                                   \
    LEGACY_LOCK;                                                                                                      \
    lolb_plan* pl = legacy_plan_c(__func__, tupSize, peArr, sizeOfPE, totm);                                          \
    const size_t bytes = (size_t)totm * tupSize * sizeof(T);                                                          \
    const size_t out_bytes = (size_t)tupSize * sizeof(T);                                                             \
    int rc = plan_reserve_stage(pl, bytes + out_bytes);                                                               \
    if (rc) die(__func__, rc);                                                                                        \
    cudaError_t e = cudaMemcpy(pl->d_stage, y, bytes, cudaMemcpyHostToDevice);                                        \
    if (e != cudaSuccess) die(__func__, cuda_fail(e, "cudaMemcpy H2D"));                                              \
    T* out_dev = (T*)((char*)pl->d_stage + bytes);                                                                    \
    rc = lolb_tensorNormSq##TAG(pl, (const T*)pl->d_stage, out_dev, 1, nullptr);                                      \
    if (rc) die(__func__, rc);                                                                                        \
    e = cudaMemcpy(y, out_dev, out_bytes, cudaMemcpyDeviceToHost);                                                    \
    if (e != cudaSuccess) die(__func__, cuda_fail(e, "cudaMemcpy D2H"));                                              \
  }
LEGACY_NORM(R, hInt_t)
LEGACY_NORM(D, double)

extern "C" void mulRq(hShort_t tupSize, hInt_t* a, hInt_t* b, hDim_t totm, hInt_t* qs)
{
  LEGACY_LOCK;
  // mul.cpp:27-30 takes no prime powers: a one-factor "plan" keyed on (totm, qs) carries the moduli
  const std::string key = plan_key(-PLAN_RQ, tupSize, nullptr, 0, qs) + std::string((const char*)&totm, sizeof(totm));
  lolb_plan* pl;
  auto it = g_plans.find(key);
  if (it != g_plans.end()) pl = it->second;
  else {
    pl = new lolb_plan();
    pl->kind = PLAN_RQ;
    int rc = plan_build_common(pl, nullptr, 0, tupSize);
    if (rc) die(__func__, rc);
    pl->n = totm;
    pl->qs.assign(qs, qs + tupSize);
    for (int t = 0; t < tupSize; t++)
      if (qs[t] < 2 || qs[t] >= ((int64_t)1 << 32)) { set_error("modulus out of range [2, 2^32)"); die(__func__, LOLB_ERR_ARG); }
    fill_zq_consts(pl);
    g_plans[key] = pl;
  }
  const size_t bytes = (size_t)totm * tupSize * sizeof(hInt_t);
  int rc = plan_reserve_stage(pl, 2 * bytes);
  if (rc) die(__func__, rc);
  hInt_t* bdev = (hInt_t*)((char*)pl->d_stage + bytes);
  cudaError_t e = cudaMemcpy(bdev, b, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) die(__func__, cuda_fail(e, "cudaMemcpy H2D"));
  with_staged(__func__, pl, a, bytes, bytes, [&](void* d) { return lolb_mulRq(pl, (hInt_t*)d, bdev, 1, 1, nullptr); });
}

extern "C" void mulC(hShort_t tupSize, lolb_complex* a, lolb_complex* b, hDim_t totm)
{
  LEGACY_LOCK;
  const std::string key = plan_key(-PLAN_C, tupSize, nullptr, 0, nullptr) + std::string((const char*)&totm, sizeof(totm));
  lolb_plan* pl;
  auto it = g_plans.find(key);
  if (it != g_plans.end()) pl = it->second;
  else {
    pl = new lolb_plan();
    pl->kind = PLAN_C;
    int rc = plan_build_common(pl, nullptr, 0, tupSize);
    if (rc) die(__func__, rc);
    pl->n = totm;
    g_plans[key] = pl;
  }
  const size_t bytes = (size_t)totm * tupSize * sizeof(lolb_complex);
  int rc = plan_reserve_stage(pl, 2 * bytes);
  if (rc) die(__func__, rc);
  lolb_complex* bdev = (lolb_complex*)((char*)pl->d_stage + bytes);
  cudaError_t e = cudaMemcpy(bdev, b, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) die(__func__, cuda_fail(e, "cudaMemcpy H2D"));
  with_staged(__func__, pl, a, bytes, bytes, [&](void* d) { return lolb_mulC(pl, (lolb_complex*)d, bdev, 1, 1, nullptr); });
}
