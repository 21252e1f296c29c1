// fused.cu -- selection and dispatch of the fused kernels (see fused.cuh).
#include "fused.cuh"

#include <cstring>

namespace lolb {

// fused_a.cu
int fused_a_select(lolb_plan* pl, void** slot);
void fused_a_release(void* slot);
bool fused_a_available(const void* slot, bool inverse);
int fused_a_crt(const lolb_plan* pl, const void* slot, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);

namespace {
struct FusedSet {
  void* a = nullptr;      // m = 14400 CRT / CRT^-1
};
FusedSet* set_of(const lolb_plan* pl) { return (FusedSet*)pl->fused; }
}  // namespace

int fused_select(lolb_plan* pl)
{
  if (pl->kind != PLAN_RQ) return LOLB_OK;
  if (!pl->fused) pl->fused = new FusedSet();
  return fused_a_select(pl, &set_of(pl)->a);
}

void fused_release(lolb_plan* pl)
{
  FusedSet* s = set_of(pl);
  if (!s) return;
  fused_a_release(s->a);
  delete s;
  pl->fused = nullptr;
}

const char* fused_kernel_name(const lolb_plan* pl, const char* op)
{
  const FusedSet* s = set_of(pl);
  if (s) {
    if (!strcmp(op, "CRT") && fused_a_available(s->a, false)) return "fused_a";
    if (!strcmp(op, "CRTInv") && fused_a_available(s->a, true)) return "fused_a";
  }
  return "generic";
}

int fused_crt_rq(const lolb_plan* pl, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  const FusedSet* s = set_of(pl);
  if (!s) return LOLB_FUSED_UNAVAILABLE;
  return fused_a_crt(pl, s->a, inverse, y, batch, st);
}

int fused_line_rq(const lolb_plan*, int, const ZqConsts&, bool, int64_t*, int64_t, cudaStream_t) { return LOLB_FUSED_UNAVAILABLE; }
int fused_mul_rq(const lolb_plan*, int64_t*, const int64_t*, int64_t, int64_t, cudaStream_t) { return LOLB_FUSED_UNAVAILABLE; }

}  // namespace lolb
