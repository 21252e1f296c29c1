/*
 * lol_b200.h -- C ABI of libctensor_b200.so, the B200 (sm_100a) back end for
 * Lol's cyclotomic `Tensor` hot path.
 *
 * Two groups of entry points:
 *
 *  (1) DROP-IN SYMBOLS.  The 29 `extern "C"` functions that
 *      lol-cpp/Crypto/Lol/Cyclotomic/Tensor/CPP/Backend.hs:304-337 imports with
 *      `foreign import ccall unsafe`, with the reference's exact names and C
 *      signatures (crt.cpp:562-598, l.cpp:109-180, g.cpp:125-273,
 *      norm.cpp:39,61, random.cpp:61, mul.cpp:27,32).  Host pointers, one ring
 *      element per call, in place, the caller keeps ownership, nothing is
 *      retained after return.  Each call stages the element to the GPU, runs
 *      the same CUDA kernels as group (2) and copies the result back.  There is
 *      no CPU implementation behind them: without a usable CUDA device they
 *      print a diagnostic and abort(), as the reference's ASSERT does
 *      (types.h:36-41).
 *
 *  (2) BATCHED, DEVICE-RESIDENT ENTRY POINTS (`lolb_*`).  What a
 *      `Crypto.Lol.Cyclotomic.Tensor.CUDA` instance binds in place of the
 *      per-element calls: a plan holds the per-(m, moduli) tables on the
 *      device; every operator takes a device pointer to `batch` ring elements
 *      laid out back to back in the reference's element layout and a CUDA
 *      stream, launches asynchronously and returns a status.
 *      Concurrency: a plan belongs to the device that was current when it was
 *      created (calls made with another current device return LOLB_ERR_ARG).
 *      It may be used from several streams and host threads at once: kernels
 *      that need scratch memory (exchange ring, counters, spilled elements)
 *      take a workspace private to the stream they are launched on, and calls
 *      on one stream are ordered by that stream.
 *
 * Element layout (both groups; reference: tensor.h:69,91 and the tuple
 * `Storable` instance at Backend.hs:80-90): coefficient j of RNS limb t of
 * element b is  y[(b*totm + j)*tupSize + t];  Zq coefficients are int64 in
 * [0, q_t) on entry and exit (zq.cpp:57-67), complex coefficients are
 * {double re, double im} (types.h:122-126).  The tensor index is
 * j = i_1 + phi_1*(i_2 + phi_2*(...)) with prime powers in increasing prime
 * order, the first one fastest (tensor.h:44-73; FactoredDefs.hs:92-94).
 */
#ifndef LOL_B200_H_
#define LOL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ types */

typedef int64_t hInt_t;   /* reference: types.h:21 */
typedef int32_t hDim_t;   /* reference: types.h:22 (Haskell passes Int64; the low 32 bits are read) */
typedef int16_t hShort_t; /* reference: types.h:23 */

/* reference: types.h:27-31; Haskell side `type CPP = (Int16, Int16)` (Backend.hs:78) */
typedef struct { hShort_t prime; hShort_t exponent; } PrimeExponent;

/* reference: types.h:122-126 (class Complex is {double real; double imag;}) */
typedef struct { double real; double imag; } lolb_complex;

/* ------------------------------------------------- (1) drop-in symbols ---- */

/* replaces crt.cpp:562-566 */
void tensorCRTRq(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, hInt_t** ru, hInt_t* qs);
/* replaces crt.cpp:569-581 */
void tensorCRTInvRq(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, hInt_t** ruinv, hInt_t* mhatInv, hInt_t* qs);
/* replaces crt.cpp:583-586 */
void tensorCRTC(hShort_t tupSize, lolb_complex* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, lolb_complex** ru);
/* replaces crt.cpp:589-598 */
void tensorCRTInvC(hShort_t tupSize, lolb_complex* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, lolb_complex** ruinv, lolb_complex* mhatInv);

/* replace l.cpp:109-115, 125-139, 150-156, 166-180 */
void tensorLRq(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, hInt_t* qs);
void tensorLInvRq(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, hInt_t* qs);
void tensorLR(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);
void tensorLInvR(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);
void tensorLDouble(hShort_t tupSize, double* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);
void tensorLInvDouble(hShort_t tupSize, double* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);
void tensorLC(hShort_t tupSize, lolb_complex* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);
void tensorLInvC(hShort_t tupSize, lolb_complex* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);

/* replace norm.cpp:39-59, 61-80: result in y[0 .. tupSize); the rest of y is
 * left untouched (the reference leaves its scratch transform there; the
 * Haskell caller reads index 0 only, CPP.hs:344-346) */
void tensorNormSqR(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);
void tensorNormSqD(hShort_t tupSize, double* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);

/* replace g.cpp:125-155 */
void tensorGPowR(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);
void tensorGPowRq(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, hInt_t* qs);
void tensorGPowC(hShort_t tupSize, lolb_complex* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);
void tensorGDecR(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);
void tensorGDecRq(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, hInt_t* qs);
void tensorGDecC(hShort_t tupSize, lolb_complex* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);

/* replace g.cpp:169-273.  Return 1 = ok, 0 = not divisible by g (R), or
 * rad_odd(m) not invertible mod some q (Rq) -> Haskell `Nothing`
 * (CPP.hs:309-323).  The R and C variants implement the documented intent
 * (exact / real division by rad_odd(m)), not the defects at g.cpp:175-182 and
 * g.cpp:213-218; see DESIGN.md "reference quirks". */
hShort_t tensorGInvPowR(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);
hShort_t tensorGInvPowRq(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, hInt_t* qs);
hShort_t tensorGInvPowC(hShort_t tupSize, lolb_complex* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);
hShort_t tensorGInvDecR(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);
hShort_t tensorGInvDecRq(hShort_t tupSize, hInt_t* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, hInt_t* qs);
hShort_t tensorGInvDecC(hShort_t tupSize, lolb_complex* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE);

/* replaces random.cpp:61-64 */
void tensorGaussianDec(hShort_t tupSize, double* y, hDim_t totm, PrimeExponent* peArr, hShort_t sizeOfPE, lolb_complex** ru);

/* replace mul.cpp:27-30, 32-35 */
void mulRq(hShort_t tupSize, hInt_t* a, hInt_t* b, hDim_t totm, hInt_t* qs);
void mulC(hShort_t tupSize, lolb_complex* a, lolb_complex* b, hDim_t totm);

/* ------------------------------------- (2) batched, device-resident API --- */

typedef struct lolb_plan lolb_plan;

/* status codes of every lolb_* function returning int */
#define LOLB_OK            0
#define LOLB_ERR_ARG       1   /* bad argument (NULL, totm mismatch, q out of range ...) */
#define LOLB_ERR_NO_CRT    2   /* some q is not prime or m does not divide q-1 (ZqBasic.hs:159-165) */
#define LOLB_ERR_CUDA      3   /* CUDA runtime error; lolb_last_error() has the text */
#define LOLB_ERR_NOT_INVERTIBLE 4 /* rad_odd(m) not invertible mod some q (g.cpp:193-199): Haskell `Nothing` */

/* text of the last error on the calling thread ("" if none) */
const char* lolb_last_error(void);
/* number of CUDA kernels launched by this library in this process so far */
int64_t lolb_kernel_launch_count(void);
/* 1 if a CUDA device is usable, else 0 (never falls back to the CPU) */
int lolb_device_available(void);

/*
 * Plan over Z_q1 x ... x Z_qk for index m = prod peArr (prime powers in
 * increasing prime order, as `ppsFact` gives them).  Root tables:
 *   ru / ruinv / mhatInv  as the drop-in symbols take them (host pointers,
 *   ru[i][j*tupSize + t] = w_t^{+-j*m/p_i^e_i}, CPP.hs:422-442), or all NULL to
 *   let the library derive them exactly as the Haskell side does
 *   (smallest generator of Z_q^*, ZqBasic.hs:144-171).  With NULL tables and a
 *   modulus without a CRT the plan is still created and serves L / G / mul;
 *   the CRT entry points then return LOLB_ERR_NO_CRT.
 */
int lolb_plan_create_rq(lolb_plan** out, const PrimeExponent* peArr, hShort_t sizeOfPE, hShort_t tupSize,
                        const hInt_t* qs, hInt_t* const* ru, hInt_t* const* ruinv, const hInt_t* mhatInv);
/* Plan for the modulus-free rings (int64 "R", double, complex): roots are cis(2 pi j / p^e) (CRTrans.hs:88-95). */
int lolb_plan_create_c(lolb_plan** out, const PrimeExponent* peArr, hShort_t sizeOfPE, hShort_t tupSize);
void lolb_plan_destroy(lolb_plan* plan);

int32_t lolb_plan_totient(const lolb_plan* plan);   /* n = phi(m) */
int32_t lolb_plan_tupsize(const lolb_plan* plan);
/* copy the plan's own tables out (host buffers sized like the drop-in arguments); for cross-checking */
int lolb_plan_get_ru_rq(const lolb_plan* plan, int inverse, int pp_index, hInt_t* out /* p^e * tupSize */);
int lolb_plan_get_mhatinv_rq(const lolb_plan* plan, hInt_t* out /* tupSize */);
/* gCRT / gInvCRT vectors (Tensor.hs:290-337; CPP.hs:451-454) resident on the device: [totm][tupSize] int64 */
const hInt_t* lolb_plan_gcrt_dev(const lolb_plan* plan, int inverse);
/* force the generic pass engine instead of a fused kernel (testing / profiling) */
void lolb_plan_set_force_generic(lolb_plan* plan, int on);
/* name of the kernel family an operator will use: "fused_a", "pow2", "generic", ... */
const char* lolb_plan_kernel_name(const lolb_plan* plan, const char* op);

/*
 * Batched operators.  `y` is a DEVICE pointer to batch*totm*tupSize
 * coefficients, transformed in place.  `stream` is a cudaStream_t (NULL =
 * legacy default stream).  Asynchronous: the call returns after the launch.
 * Same operator, same result bit for bit, as `batch` calls of the drop-in
 * symbol of the same name.
 */
int lolb_tensorCRTRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream);
int lolb_tensorCRTInvRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream);
int lolb_tensorLRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream);
int lolb_tensorLInvRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream);
int lolb_tensorGPowRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream);
int lolb_tensorGDecRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream);
int lolb_tensorGInvPowRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream);   /* LOLB_ERR_NOT_INVERTIBLE = reference's 0 */
int lolb_tensorGInvDecRq(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream);
/* a[i] <- a[i] * b[i]; b holds `b_batch` elements (1 = broadcast one element, e.g. gCRT; else == batch) */
int lolb_mulRq(const lolb_plan* plan, hInt_t* a, const hInt_t* b, int64_t batch, int64_t b_batch, void* stream);
/* Fused pairs around the CRT (the callers either side of the path: the ring product of lol/Crypto/Lol/Cyclotomic/Cyc.hs:276-297
 * = toCRT then zipWith (*) (UCyc.hs:232), and the CRT-basis products of key switching, lol-apps SymmSHE.hs:302-314):
 *   lolb_crtMulRq     y <- tensorCRTRq(y) . b        (crt.cpp:562-566 then mul.cpp:27-30)
 *   lolb_mulCrtInvRq  y <- tensorCRTInvRq(y . b)     (mul.cpp:27-30 then crt.cpp:569-581)
 * b as in lolb_mulRq, must not alias y.  Results are those of the two calls in sequence; one HBM pass where a fused
 * kernel exists (m = 14400), the two kernels back to back otherwise. */
int lolb_crtMulRq(const lolb_plan* plan, hInt_t* y, const hInt_t* b, int64_t batch, int64_t b_batch, void* stream);
int lolb_mulCrtInvRq(const lolb_plan* plan, hInt_t* y, const hInt_t* b, int64_t batch, int64_t b_batch, void* stream);

/* The coefficient-wise steps of SymmSHE's ciphertext multiply and quadratic key switch, which the reference runs on the
 * host between its FFI calls (lol-apps/Crypto/Lol/Applications/SymmSHE.hs); here one streaming pass each over the
 * device-resident batch.  Operands are [batch][n][tupSize] arrays of canonical residues.
 *   lolb_ctMulRq      (d0,d1,d2) <- (a0 b0, a0 b1 + a1 b0, a1 b1) [. gCRT when mul_g]: the product of two linear
 *                     ciphertexts whose components are in the CRT basis, SymmSHE.hs:443-449 (`mulG <$> c1 * c2`;
 *                     zipWithT (*) of UCyc.hs:232; mulGCRT of CPP.hs:230).  Outputs may alias inputs.
 *   lolb_gadgetLength number of gadget digits l over this plan's moduli: base 0 = TrivGad (one per limb,
 *                     ZqBasic.hs:227-232), base b >= 2 = BaseBGad b (gadlen, ZqBasic.hs:241-243); tuples concatenate
 *                     (Gadget.hs:92-101).  -1 on a bad argument.
 *   lolb_decomposeRq  digits[d] <- reduce(decompose(x)[d]) for a Pow-basis x (SymmSHE.hs:314; Cyc.hs:603; lift =
 *                     decode', ZqBasic.hs:92-94; decomp / divModCent, Numeric.hs:202-205, 227-234); digits is
 *                     [l][batch][n][tupSize].
 *   lolb_knapsackRq   c0 += sum_d digits[d] . hints[d][0],  c1 += sum_d digits[d] . hints[d][1]  (SymmSHE.hs:302-305 and
 *                     :372); digits in the CRT basis, hints is [l][2][n][tupSize] (one ring element each). */
int lolb_ctMulRq(const lolb_plan* plan, const hInt_t* a0, const hInt_t* a1, const hInt_t* b0, const hInt_t* b1,
                 hInt_t* d0, hInt_t* d1, hInt_t* d2, int64_t batch, int mul_g, void* stream);
int lolb_gadgetLength(const lolb_plan* plan, int64_t base);
int lolb_decomposeRq(const lolb_plan* plan, const hInt_t* x, hInt_t* digits, int64_t batch, int64_t base, void* stream);
/* digits[d] <- tensorCRTRq(reduce(decompose(x)[d])): lolb_decomposeRq followed by the CRT of every digit (`adviseCRT <$> xs`,
 * SymmSHE.hs:305), the decomposition folded into the CRT kernel's load stage where that pays (m = 14400, tupSize 2, TrivGad) */
int lolb_decomposeCrtRq(const lolb_plan* plan, const hInt_t* x, hInt_t* digits, int64_t batch, int64_t base, void* stream);
int lolb_knapsackRq(const lolb_plan* plan, const hInt_t* digits, int ell, const hInt_t* hints, hInt_t* c0, hInt_t* c1,
                    int64_t batch, void* stream);

/* modulus-free rings; plan from lolb_plan_create_c */
int lolb_tensorLR(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream);
int lolb_tensorLInvR(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream);
int lolb_tensorGPowR(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream);
int lolb_tensorGDecR(const lolb_plan* plan, hInt_t* y, int64_t batch, void* stream);
/* ok[b] (device, int16) <- 1 if element b was divisible by g and has been divided, else 0 (element then holds the undivided transform) */
int lolb_tensorGInvPowR(const lolb_plan* plan, hInt_t* y, hShort_t* ok, int64_t batch, void* stream);
int lolb_tensorGInvDecR(const lolb_plan* plan, hInt_t* y, hShort_t* ok, int64_t batch, void* stream);
int lolb_tensorLDouble(const lolb_plan* plan, double* y, int64_t batch, void* stream);
int lolb_tensorLInvDouble(const lolb_plan* plan, double* y, int64_t batch, void* stream);
int lolb_tensorLC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream);
int lolb_tensorLInvC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream);
int lolb_tensorGPowC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream);
int lolb_tensorGDecC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream);
int lolb_tensorGInvPowC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream);
int lolb_tensorGInvDecC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream);
int lolb_tensorCRTC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream);
int lolb_tensorCRTInvC(const lolb_plan* plan, lolb_complex* y, int64_t batch, void* stream);
int lolb_mulC(const lolb_plan* plan, lolb_complex* a, const lolb_complex* b, int64_t batch, int64_t b_batch, void* stream);
/* y: batch elements of totm*tupSize doubles, i.i.d. Gaussians in, decoding-basis coefficients out (random.cpp:61-64) */
int lolb_tensorGaussianDec(const lolb_plan* plan, double* y, int64_t batch, void* stream);
/*
 * The Gaussian source on the device.  In the reference the inputs of tensorGaussianDec are drawn on the host by
 * `realGaussians` (lol/Crypto/Lol/GaussRandom.hs:34-59, polar Box-Muller over a MonadRandom) and handed to C
 * (lol-cpp/.../CPP.hs:376-389).  Here the same transform runs over a counter-based generator (Philox4x32-10): a value is a
 * pure function of (seed, first_element + b, position), so a batch can be produced in pieces and does not depend on
 * the launch shape.  Parity with the reference is distributional (another uniform source).
 *   lolb_realGaussians  y[batch][n] (device) <- i.i.d. Gaussians of scaled variance svar, i.e. true variance svar / (2 pi)
 *   lolb_tGaussianDec   `tGaussianDec v` of class Tensor (Tensor.hs:143): draws of scaled variance v * m / rad(m), then the
 *                       transform of tensorGaussianDec, in one pass over y (8 n bytes per element) where the streaming kernel
 *                       serves the index
 */
int lolb_realGaussians(double svar, uint64_t seed, uint64_t first_element, double* y, int64_t n, int64_t batch, void* stream);
int lolb_tGaussianDec(const lolb_plan* plan, double v, uint64_t seed, uint64_t first_element, double* y, int64_t batch, void* stream);
/* out[b*tupSize + t] (device) <- the value the drop-in symbol leaves in y[t]; y is not modified */
int lolb_tensorNormSqR(const lolb_plan* plan, const hInt_t* y, hInt_t* out, int64_t batch, void* stream);
int lolb_tensorNormSqD(const lolb_plan* plan, const double* y, double* out, int64_t batch, void* stream);

/*
 * Ring extensions O_m'/O_m, m | m': the two-index methods of `class Tensor` (embedPow, embedDec, embedCRT, twacePowDec,
 * twaceCRT, coeffs; lol/Crypto/Lol/Cyclotomic/Tensor.hs:160-190).  The reference runs them on the host in Haskell over
 * index vectors (lol-cpp/Crypto/Lol/Cyclotomic/Tensor/CPP/Extension.hs:54-129; tables Tensor.hs:380-510); here the tables
 * of one (m, m') pair live on the device and each operator is one gather pass over a device-resident batch.
 * An extension is created from the plans of O_m (`lo`) and O_m' (`hi`), which must be over the same ring (both Rq with
 * equal tupSize and moduli, or both modulus-free with equal tupSize) and must outlive it.  `ring` selects the coefficient
 * type of the call: LOLB_RING_RQ for Rq plans; LOLB_RING_R (int64), _DOUBLE or _C (complex) for modulus-free plans.
 * Operands are DEVICE pointers, distinct, `x` with `batch` elements of the source ring and `y` of the target ring:
 *   lolb_twacePowDec  O_m' -> O_m   y[i]  = x[extIndicesPowDec[i]]                       Extension.hs:99-103
 *   lolb_embedPow     O_m  -> O_m'  y[i'] = j0(i') == 0 ? x[j1(i')] : 0                  Extension.hs:60-70
 *   lolb_embedDec     O_m  -> O_m'  y[i'] = 0 | x[sh] | -x[sh] per baseIndicesDec        Extension.hs:71-77
 *   lolb_embedCRT     O_m  -> O_m'  y[i'] = x[baseIndicesCRT[i']]                        Extension.hs:81-85
 *   lolb_coeffsPowDec O_m' -> (O_m)^(phi'/phi), y laid out [batch][phi'/phi][phi][k]     Extension.hs:90-93
 *   lolb_twaceCRT     O_m' -> O_m   y[i]  = sum of the phi'/phi entries of tweak . x lying above i, tweak =
 *                                   m'hat^-1 mhat embedCRT(gInvCRT_m) gCRT_m'            Extension.hs:110-129
 *   lolb_powBasisPow  -> (O_m')^(phi'/phi), y laid out [phi'/phi][phi'][k]: vector r is `one` where baseIndicesPow = (r, 0),
 *                                   the O_m-basis of O_m' in the powerful basis (no input, no batch)   Extension.hs:133-143
 * embedCRT / twaceCRT return LOLB_ERR_NO_CRT where the reference's `CRTrans` yields Nothing (no CRT of index m' over the
 * moduli) and for the R / Double rings.
 */
typedef struct lolb_ext lolb_ext;
#define LOLB_RING_RQ     0
#define LOLB_RING_R      1
#define LOLB_RING_DOUBLE 2
#define LOLB_RING_C      3
int lolb_ext_create(lolb_ext** out, const lolb_plan* lo, const lolb_plan* hi);
void lolb_ext_destroy(lolb_ext* ext);
int32_t lolb_ext_totient(const lolb_ext* ext, int upper);   /* phi(m), or phi(m') when upper != 0 */
/* copy an index table out (host int32 buffer), for cross-checking against Tensor.hs:429-478 */
#define LOLB_EXT_INDICES_POWDEC 0   /* extIndicesPowDec  [phi]  */
#define LOLB_EXT_INDICES_CRT    1   /* extIndicesCRT     [phi'] */
#define LOLB_EXT_BASE_POW_J0    2   /* fst <$> baseIndicesPow [phi'] */
#define LOLB_EXT_BASE_POW_J1    3   /* snd <$> baseIndicesPow = baseIndicesCRT [phi'] */
#define LOLB_EXT_BASE_DEC       4   /* baseIndicesDec [phi']: -1 = Nothing, else 2 * index + (negate ? 1 : 0) */
#define LOLB_EXT_INDICES_COEFFS 5   /* extIndicesCoeffs [phi'/phi][phi] */
#define LOLB_EXT_TABLES         6
int lolb_ext_get_table(const lolb_ext* ext, int which, int32_t* out);
/* the same tables computed on the host from the two prime-power lists alone (no device, no plans): returns the entry
 * count and fills `out` when it is not NULL; -1 on a bad argument or when m does not divide m' */
int64_t lolb_ext_index_table(const PrimeExponent* pe, hShort_t nPE, const PrimeExponent* pe2, hShort_t nPE2, int which, int32_t* out);
int lolb_twacePowDec(const lolb_ext* ext, int ring, const void* x, void* y, int64_t batch, void* stream);
int lolb_embedPow(const lolb_ext* ext, int ring, const void* x, void* y, int64_t batch, void* stream);
int lolb_embedDec(const lolb_ext* ext, int ring, const void* x, void* y, int64_t batch, void* stream);
int lolb_embedCRT(const lolb_ext* ext, int ring, const void* x, void* y, int64_t batch, void* stream);
int lolb_coeffsPowDec(const lolb_ext* ext, int ring, const void* x, void* y, int64_t batch, void* stream);
int lolb_twaceCRT(const lolb_ext* ext, int ring, const void* x, void* y, int64_t batch, void* stream);
int lolb_powBasisPow(const lolb_ext* ext, int ring, void* y, void* stream);

/*
 * Coefficient-wise maps either side of the transforms when Lol switches moduli or rounds an error term; host `fmapT`
 * closures in the reference (lol/Crypto/Lol/Cyclotomic/UCyc.hs:267-300, 427-445), one streaming pass each here.
 * Device pointers, [batch][totm][tupSize] layout, Rq plans.
 *   lolb_liftRq        y = lift x per limb: representative in [-q/2, q/2) (decode', ZqBasic.hs:92-94; UCyc.hs:288-296)
 *   lolb_reduceRq      y[.][t] = z mod q_t; z has z_tupsize = 1 (one integer per coefficient, reduced into every limb) or
 *                      tupSize int64 per coefficient (reduce', ZqBasic.hs:88-90; UCyc.hs:267-275)
 *   lolb_rescaleDropRq removes limb `drop` of the product ring: y_t = q_d^-1 (x_t - reduce(lift x_d)), t != d, y is
 *                      [batch][totm][tupSize-1] (Prelude.hs:226-232 drop = 0, :259-265 drop = tupSize-1; rescalePow,
 *                      UCyc.hs:298-300; Cyc.hs:529-541).  LOLB_ERR_NOT_INVERTIBLE when q_d is not a unit mod some q_t.
 *   lolb_rescaleModRq  y = fst (divModCent (q'_t * lift x) q_t) mod q'_t per limb (rescaleMod, Prelude.hs:143-153)
 *   lolb_roundCosetRq  y = rep + p_t * round((e - rep) / p_t), rep = lift zp, p = the plan's moduli (roundCoset,
 *                      Prelude.hs:155-162; errorCoset, UCyc.hs:436-445); zp == NULL: y = round e (errorRounded,
 *                      UCyc.hs:427-434).  e is double, y int64, round is half-to-even.
 */
int lolb_liftRq(const lolb_plan* plan, const hInt_t* x, hInt_t* y, int64_t batch, void* stream);
int lolb_reduceRq(const lolb_plan* plan, const hInt_t* z, int z_tupsize, hInt_t* y, int64_t batch, void* stream);
int lolb_rescaleDropRq(const lolb_plan* plan, int drop, const hInt_t* x, hInt_t* y, int64_t batch, void* stream);
int lolb_rescaleModRq(const lolb_plan* plan, const hInt_t* qs_new, const hInt_t* x, hInt_t* y, int64_t batch, void* stream);
int lolb_roundCosetRq(const lolb_plan* plan, const double* e, const hInt_t* zp, hInt_t* y, int64_t batch, void* stream);

/*
 * Host-buffer batched calls (what an FFI caller with Haskell-owned vectors
 * uses): `y` is a HOST pointer to batch elements; the call pipelines
 * host->device copy, kernel(s) and device->host copy over chunks on internal
 * streams and returns when the result is back in `y`.  `ops` is a
 * NUL-terminated list of operator names separated by ',' applied in order,
 * e.g. "CRT", "CRT,CRTInv", "L,GPow" (names as in the drop-in symbols, without
 * the `tensor` prefix and ring suffix).
 */
int lolb_rq_apply_host(const lolb_plan* plan, const char* ops, hInt_t* y, int64_t batch);
/*
 * The same pipeline over a narrow wire format: `y` holds the residues as uint32 (every modulus the plan accepts is below
 * 2^32, and the reference itself needs q^2 to fit an int64, types.h:79-84), in the element layout y[(b*totm + j)*tupSize + t].
 * Half the bytes cross PCIe; the device widens each chunk to the int64 layout, runs the same kernels and narrows the
 * result.  Replaces one `SV.thaw` + one FFI call per element (CPP.hs:325-337) for callers that keep Z_q vectors as Word32.
 * An empty `ops` runs the copies alone: the pipeline's ceiling on this host.
 */
int lolb_rq_apply_host_u32(const lolb_plan* plan, const char* ops, uint32_t* y, int64_t batch);
/*
 * Device memory for callers without a CUDA binding of their own (haskell/.../Tensor/CUDA.hs keeps a ring element in GPU
 * memory behind a ForeignPtr whose finalizer is lolb_dev_free, so the `Tensor` methods chain without crossing PCIe):
 * cudaMalloc / cudaFree / cudaMemcpyAsync on the current device.  lolb_dev_download returns when the bytes are in dst.
 */
void* lolb_dev_alloc(uint64_t bytes);
void lolb_dev_free(void* p);
int lolb_dev_upload(void* dst_dev, const void* src_host, uint64_t bytes, void* stream);
int lolb_dev_download(void* dst_host, const void* src_dev, uint64_t bytes, void* stream);
int lolb_dev_copy(void* dst_dev, const void* src_dev, uint64_t bytes, void* stream);
/* pinned host memory for the above (cudaHostAlloc / cudaFreeHost) */
void* lolb_host_alloc(uint64_t bytes);
void lolb_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* LOL_B200_H_ */
