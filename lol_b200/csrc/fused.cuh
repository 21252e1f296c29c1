// fused.cuh -- interface of the fused (single HBM round trip, shape-specialised) kernels.
// Every entry returns LOLB_FUSED_UNAVAILABLE when the plan's shape has no fused kernel; the caller then
// runs the generic pass engine (engine.cu).  Both paths are CUDA; neither is a CPU fallback.
#pragma once
#include "lolb_internal.cuh"

namespace lolb {

constexpr int LOLB_FUSED_UNAVAILABLE = -1;

int fused_select(lolb_plan* pl);          // (re)build fused-kernel tables after the plan's root tables changed
void fused_release(lolb_plan* pl);
const char* fused_kernel_name(const lolb_plan* pl, const char* op);
int fused_crt_rq(const lolb_plan* pl, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);
int fused_crt_mul_rq(const lolb_plan* pl, bool inverse, int64_t* y, const int64_t* b, int64_t batch, int64_t b_batch, cudaStream_t st);
int fused_decompose_crt_rq(const lolb_plan* pl, const int64_t* x, int64_t* digits, int64_t batch, int64_t base, cudaStream_t st);
int fused_crt_c(const lolb_plan* pl, bool inverse, double2* y, int64_t batch, cudaStream_t st);
int fused_line_rq(const lolb_plan* pl, int kind, const ZqConsts& zc, bool scale, int64_t* y, int64_t batch, cudaStream_t st);
int fused_mul_rq(const lolb_plan* pl, int64_t* a, const int64_t* b, int64_t batch, int64_t b_batch, cudaStream_t st);
// coefficient-wise steps of SymmSHE's ciphertext multiply and key switch (she_stream.cu)
int she_gadget_length(const lolb_plan* pl, int64_t base);      // -1 on a bad base
int she_gadget_digits(const lolb_plan* pl, int64_t base, int* nd /* [tupSize] */, int* shift);      // digits per limb; returns l or -1
int she_ct_mul(const lolb_plan* pl, const int64_t* a0, const int64_t* a1, const int64_t* b0, const int64_t* b1, const int64_t* g,
               int64_t* d0, int64_t* d1, int64_t* d2, int64_t batch, cudaStream_t st);
int she_decompose(const lolb_plan* pl, const int64_t* x, int64_t* digits, int64_t batch, int64_t base, cudaStream_t st);
int she_knapsack(const lolb_plan* pl, const int64_t* digits, int ell, const int64_t* hints, int64_t* c0, int64_t* c1, int64_t batch,
                 cudaStream_t st);
// modulus-free rings (fused_plain.cu); ring = RING_I64 / RING_F64 / RING_C64
int fused_plain_line(const lolb_plan* pl, int ring, int kind, void* y, int64_t batch, double rscale, cudaStream_t st);
int fused_plain_gauss(const lolb_plan* pl, double* y, int64_t batch, cudaStream_t st);
const char* fused_plain_name(const lolb_plan* pl, bool gauss, bool cplx);
// on-device Gaussian source: y[batch][n] i.i.d. N(0, var2 / 2) (pl may be NULL), and tGaussianDec in one pass (draw + transform)
int fused_plain_real_gaussians(const lolb_plan* pl, double* y, int64_t n, int64_t batch, uint64_t seed, uint64_t first, double var2, cudaStream_t st);
int fused_plain_gauss_gen(const lolb_plan* pl, double* y, int64_t batch, cudaStream_t st, bool gen, uint64_t seed, uint64_t first, double var2);
int fused_plain_normsq_i64(const lolb_plan* pl, const int64_t* y, int64_t* out, int64_t batch, cudaStream_t st);
int fused_plain_normsq_f64(const lolb_plan* pl, const double* y, double* out, int64_t batch, cudaStream_t st);

}  // namespace lolb
