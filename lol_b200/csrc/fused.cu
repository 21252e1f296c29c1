// fused.cu -- selection and dispatch of the fused kernels (see fused.cuh).
#include "fused.cuh"

#include <cstring>

namespace lolb {

// fused_a.cu
int fused_a_select(lolb_plan* pl, void** slot);
void fused_a_release(void* slot);
bool fused_a_available(const void* slot, bool inverse);
int fused_a_crt(const lolb_plan* pl, const void* slot, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);
int fused_a_decompose_crt(const lolb_plan* pl, const void* slot, const int64_t* x, int64_t* digits, int64_t batch, int64_t base,
                          cudaStream_t st);
int fused_a_crt_mul(const lolb_plan* pl, const void* slot, bool inverse, int64_t* y, const int64_t* b, int64_t batch,
                    int64_t b_batch, cudaStream_t st);

// fused_w_impl.cuh (fused_w_p0..p4.cu)
int fused_w_select(lolb_plan* pl, void** slot);
void fused_w_release(void* slot);
bool fused_w_available(const void* slot, bool inverse);
int fused_w_crt(const lolb_plan* pl, const void* slot, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);
int fused_wc_select(lolb_plan* pl, void** slot);
void fused_wc_release(void* slot);
bool fused_wc_available(const void* slot, bool inverse);
int fused_wc_crt(const lolb_plan* pl, const void* slot, bool inverse, double2* y, int64_t batch, cudaStream_t st);

// fused_pow2.cu
int fused_pow2_select(lolb_plan* pl, void** slot);
void fused_pow2_release(void* slot);
bool fused_pow2_available(const void* slot, bool inverse);
int fused_pow2_crt(const lolb_plan* pl, const void* slot, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);

// fused_pow2_df.cu
int fused_pow2_df_select(lolb_plan* pl, void** slot);
void fused_pow2_df_release(void* slot);
bool fused_pow2_df_available(const void* slot, bool inverse);
int fused_pow2_df_crt(const lolb_plan* pl, const void* slot, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);

// fused_ac.cu
int fused_ac_select(lolb_plan* pl, void** slot);
void fused_ac_release(void* slot);
bool fused_ac_available(const void* slot, bool inverse);
int fused_ac_crt(const lolb_plan* pl, const void* slot, bool inverse, double2* y, int64_t batch, cudaStream_t st);

// fused_pow2c.cu
int fused_pow2c_select(lolb_plan* pl, void** slot);
void fused_pow2c_release(void* slot);
bool fused_pow2c_available(const void* slot, bool inverse);
int fused_pow2c_crt(const lolb_plan* pl, const void* slot, bool inverse, double2* y, int64_t batch, cudaStream_t st);

// fused_stream.cu
const char* fused_stream_line_name(const lolb_plan* pl, bool ginv);
int fused_stream_line(const lolb_plan* pl, int kind, const ZqConsts& zc, bool scale, int64_t* y, int64_t batch, cudaStream_t st);
int fused_stream_mul(const lolb_plan* pl, int64_t* a, const int64_t* b, int64_t batch, int64_t b_batch, cudaStream_t st);

namespace {
struct FusedSet {
  void* a = nullptr;      // m = 14400 CRT / CRT^-1
  void* pow2 = nullptr;   // m = 2^e CRT / CRT^-1, limb resident in shared memory (e <= 12, tupSize 3, ...)
  void* ac = nullptr;       // m = 14400 complex CRT / CRT^-1
  void* pow2_df = nullptr;  // m = 2^e CRT / CRT^-1, dataflow kernel with an L2 exchange ring (13 <= e <= 16)
  void* w = nullptr;        // m = 2^a x odd prime powers (1728, 5184, 2912, 728, 3640, 2016, ...) CRT / CRT^-1
  void* wc = nullptr;       // the same schedule over complex doubles (tensorCRTC / tensorCRTInvC of those indices)
  void* pow2c = nullptr;    // m = 2^e complex CRT / CRT^-1 (tupSize 1, e <= 14)
};
FusedSet* set_of(const lolb_plan* pl) { return (FusedSet*)pl->fused; }
}  // namespace

int fused_select(lolb_plan* pl)
{
  if (pl->kind == PLAN_C) {
    if (!pl->fused) pl->fused = new FusedSet();
    int rc = fused_ac_select(pl, &set_of(pl)->ac);
    if (!rc) rc = fused_wc_select(pl, &set_of(pl)->wc);
    if (!rc) rc = fused_pow2c_select(pl, &set_of(pl)->pow2c);
    return rc;
  }
  if (pl->kind != PLAN_RQ) return LOLB_OK;
  if (!pl->fused) pl->fused = new FusedSet();
  int rc = fused_a_select(pl, &set_of(pl)->a);
  if (!rc) rc = fused_w_select(pl, &set_of(pl)->w);
  if (!rc) rc = fused_pow2_select(pl, &set_of(pl)->pow2);
  if (!rc) rc = fused_pow2_df_select(pl, &set_of(pl)->pow2_df);
  return rc;
}

namespace pow2 { void pow2_split_release(const lolb_plan* pl); }

void fused_release(lolb_plan* pl)
{
  pow2::pow2_split_release(pl);      // CUDA graphs of the split power-of-two schedule are keyed by the plan
  FusedSet* s = set_of(pl);
  if (!s) return;
  fused_a_release(s->a);
  fused_w_release(s->w);
  fused_wc_release(s->wc);
  fused_pow2_release(s->pow2);
  fused_pow2_df_release(s->pow2_df);
  fused_ac_release(s->ac);
  fused_pow2c_release(s->pow2c);
  delete s;
  pl->fused = nullptr;
}

const char* fused_kernel_name(const lolb_plan* pl, const char* op)
{
  const FusedSet* s = set_of(pl);
  if (s) {
    if (!strcmp(op, "CRTMul") && fused_a_available(s->a, false)) return "fused_a+mul";
    if (!strcmp(op, "MulCRTInv") && fused_a_available(s->a, true)) return "fused_a+mul";
    if (!strcmp(op, "CRTC") && fused_ac_available(s->ac, false)) return "fused_ac";
    if (!strcmp(op, "CRTInvC") && fused_ac_available(s->ac, true)) return "fused_ac";
    if (!strcmp(op, "CRTC") && fused_wc_available(s->wc, false)) return "fused_w";
    if (!strcmp(op, "CRTInvC") && fused_wc_available(s->wc, true)) return "fused_w";
    if (!strcmp(op, "CRTC") && fused_pow2c_available(s->pow2c, false)) return "fused_pow2c";
    if (!strcmp(op, "CRTInvC") && fused_pow2c_available(s->pow2c, true)) return "fused_pow2c";
    if (!strcmp(op, "CRT") && fused_a_available(s->a, false)) return "fused_a";
    if (!strcmp(op, "CRTInv") && fused_a_available(s->a, true)) return "fused_a";
    if (!strcmp(op, "CRT") && fused_w_available(s->w, false)) return "fused_w";
    if (!strcmp(op, "CRTInv") && fused_w_available(s->w, true)) return "fused_w";
    if (!strcmp(op, "CRT") && fused_pow2_df_available(s->pow2_df, false)) return "fused_pow2_df";
    if (!strcmp(op, "CRTInv") && fused_pow2_df_available(s->pow2_df, true)) return "fused_pow2_df";
    if (!strcmp(op, "CRT") && fused_pow2_available(s->pow2, false)) return "fused_pow2";
    if (!strcmp(op, "CRTInv") && fused_pow2_available(s->pow2, true)) return "fused_pow2";
  }
  if ((!strcmp(op, "CRT") || !strcmp(op, "CRTC")) && engine_axis_supported(pl, false)) return "engine_axis";
  if ((!strcmp(op, "CRTInv") || !strcmp(op, "CRTInvC")) && engine_axis_supported(pl, true)) return "engine_axis";
  if (!strcmp(op, "mulRq") || !strcmp(op, "MulGCRT") || !strcmp(op, "DivGCRT")) return ((int64_t)pl->n * pl->k) % 2 == 0 ? "mul_stream" : "generic";
  if (pl->kind == PLAN_RQ && (!strcmp(op, "L") || !strcmp(op, "LInv") || !strcmp(op, "GPow") || !strcmp(op, "GDec") ||
                              !strcmp(op, "GInvPow") || !strcmp(op, "GInvDec")))
    return fused_stream_line_name(pl, !strncmp(op, "GInv", 4));
  if (pl->kind == PLAN_C) {
    if (!strcmp(op, "GaussianDec")) return fused_plain_name(pl, true, false);
    for (const char* nm : {"LR", "LInvR", "GPowR", "GDecR", "LDouble", "LInvDouble"}) if (!strcmp(op, nm)) return fused_plain_name(pl, false, false);
    for (const char* nm : {"LC", "LInvC", "GPowC", "GDecC", "GInvPowC", "GInvDecC"}) if (!strcmp(op, nm)) return fused_plain_name(pl, false, true);
  }
  return "generic";
}

int fused_crt_rq(const lolb_plan* pl, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  const FusedSet* s = set_of(pl);
  if (!s) return LOLB_FUSED_UNAVAILABLE;
  int rc = fused_a_crt(pl, s->a, inverse, y, batch, st);
  if (rc == LOLB_FUSED_UNAVAILABLE) rc = fused_w_crt(pl, s->w, inverse, y, batch, st);
  if (rc == LOLB_FUSED_UNAVAILABLE) rc = fused_pow2_df_crt(pl, s->pow2_df, inverse, y, batch, st);
  if (rc == LOLB_FUSED_UNAVAILABLE) rc = fused_pow2_crt(pl, s->pow2, inverse, y, batch, st);
  return rc;
}

int fused_decompose_crt_rq(const lolb_plan* pl, const int64_t* x, int64_t* digits, int64_t batch, int64_t base, cudaStream_t st)
{
  const FusedSet* s = set_of(pl);
  if (!s) return LOLB_FUSED_UNAVAILABLE;
  return fused_a_decompose_crt(pl, s->a, x, digits, batch, base, st);
}

int fused_crt_mul_rq(const lolb_plan* pl, bool inverse, int64_t* y, const int64_t* b, int64_t batch, int64_t b_batch, cudaStream_t st)
{
  const FusedSet* s = set_of(pl);
  if (!s) return LOLB_FUSED_UNAVAILABLE;
  return fused_a_crt_mul(pl, s->a, inverse, y, b, batch, b_batch, st);
}

int fused_crt_c(const lolb_plan* pl, bool inverse, double2* y, int64_t batch, cudaStream_t st)
{
  const FusedSet* s = set_of(pl);
  if (!s) return LOLB_FUSED_UNAVAILABLE;
  int rc = fused_ac_crt(pl, s->ac, inverse, y, batch, st);
  if (rc == LOLB_FUSED_UNAVAILABLE) rc = fused_wc_crt(pl, s->wc, inverse, y, batch, st);
  if (rc == LOLB_FUSED_UNAVAILABLE) rc = fused_pow2c_crt(pl, s->pow2c, inverse, y, batch, st);
  return rc;
}

int fused_line_rq(const lolb_plan* pl, int kind, const ZqConsts& zc, bool scale, int64_t* y, int64_t batch, cudaStream_t st)
{
  return fused_stream_line(pl, kind, zc, scale, y, batch, st);
}

int fused_mul_rq(const lolb_plan* pl, int64_t* a, const int64_t* b, int64_t batch, int64_t b_batch, cudaStream_t st)
{
  return fused_stream_mul(pl, a, b, batch, b_batch, st);
}

}  // namespace lolb
