// fused.cu -- dispatch of the fused kernels (see fused.cuh).
#include "fused.cuh"

namespace lolb {

int fused_select(lolb_plan*) { return LOLB_OK; }
void fused_release(lolb_plan*) {}
const char* fused_kernel_name(const lolb_plan*, const char*) { return "generic"; }
int fused_crt_rq(const lolb_plan*, bool, int64_t*, int64_t, cudaStream_t) { return LOLB_FUSED_UNAVAILABLE; }
int fused_line_rq(const lolb_plan*, int, const ZqConsts&, bool, int64_t*, int64_t, cudaStream_t) { return LOLB_FUSED_UNAVAILABLE; }
int fused_mul_rq(const lolb_plan*, int64_t*, const int64_t*, int64_t, int64_t, cudaStream_t) { return LOLB_FUSED_UNAVAILABLE; }

}  // namespace lolb
