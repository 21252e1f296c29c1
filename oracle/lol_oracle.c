/*
 * lol_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See lol_oracle.h for scope, pinning and the rule on who may call this.
 *
 * Differences of form (not of result) from the reference:
 *  - Zq values are kept canonical in [0,q) after every operation instead of in
 *    (-q,q) with one fix-up at exit (types.h:52-94, zq.cpp:57-67); the residue
 *    returned is the same.  Inputs are reduced into [0,q) on entry.
 *  - one dense formula per operator for every odd p (the reference unrolls
 *    p = 3, 5, 7 and is dense from 11 on).
 *  - the current modulus is a file-static, like the reference's Zq::q
 *    (common.cpp:14): single-threaded by construction.
 */
#include "lol_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---------- small integer helpers ---------------------------------------- */

/* reference: common.cpp:16-27 (ipow) */
static int64_t lo_ipow(int64_t base, int e)
{
  int64_t r = 1;
  for (int i = 0; i < e; i++) r *= base;
  return r;
}

/* reference: crt.cpp:21-33 (bitrev): reverse the `digits` base-p digits of j */
static int64_t lo_digit_rev(int p, int digits, int64_t j)
{
  int64_t out = 0;
  for (int d = 0; d < digits; d++) { out = out * p + j % p; j /= p; }
  return out;
}

/* reference: g.cpp:157-167 (oddRad) */
static int64_t lo_odd_radical(const lo_pe_t* pe, int npe)
{
  int64_t r = 1;
  for (int i = 0; i < npe; i++) if (pe[i].prime != 2) r *= pe[i].prime;
  return r;
}

/* reference: zq.cpp:20-54 (reciprocal): b^{-1} mod a in [0,a), or 0 */
static int64_t lo_mod_inverse(int64_t a, int64_t b)
{
  int64_t r0 = a, r1 = ((b % a) + a) % a, t0 = 0, t1 = 1;
  while (r1 != 0) {
    int64_t qt = r0 / r1, rr = r0 - qt * r1, tt = t0 - qt * t1;
    r0 = r1; r1 = rr; t0 = t1; t1 = tt;
  }
  if (r0 != 1) return 0;
  return ((t0 % a) + a) % a;
}

/* ---------- rings ---------------------------------------------------------- */

static int64_t g_q = 1;                       /* reference: common.cpp:14 */
static void lo_set_modulus(int64_t q) { g_q = q; }

/* products of two canonical residues fit int64 for every q the reference
 * supports (q*q < 2^63, types.h:79-84) */
static inline int64_t zq_add(int64_t a, int64_t b) { int64_t s = a + b; return s >= g_q ? s - g_q : s; }
static inline int64_t zq_sub(int64_t a, int64_t b) { int64_t s = a - b; return s < 0 ? s + g_q : s; }
static inline int64_t zq_mul(int64_t a, int64_t b) { return (int64_t)(((unsigned __int128)(uint64_t)a * (uint64_t)b) % (uint64_t)g_q); }
static inline int64_t zq_int(int64_t i) { int64_t r = i % g_q; return r < 0 ? r + g_q : r; }

/* wrapping int64 (the reference's hInt_t arithmetic, on two's complement) */
static inline int64_t zz_add(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
static inline int64_t zz_sub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }
static inline int64_t zz_mul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }

/* reference: types.h:122-164 (class Complex) */
static inline lo_cplx_t cx_add(lo_cplx_t a, lo_cplx_t b) { lo_cplx_t r = { a.re + b.re, a.im + b.im }; return r; }
static inline lo_cplx_t cx_sub(lo_cplx_t a, lo_cplx_t b) { lo_cplx_t r = { a.re - b.re, a.im - b.im }; return r; }
static inline lo_cplx_t cx_mul(lo_cplx_t a, lo_cplx_t b) { lo_cplx_t r = { a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re }; return r; }
static inline lo_cplx_t cx_int(int64_t i) { lo_cplx_t r = { (double)i, 0.0 }; return r; }

#define RT int64_t
#define RN(n) n##_zq
#define R_ADD(a, b) zq_add((a), (b))
#define R_SUB(a, b) zq_sub((a), (b))
#define R_MUL(a, b) zq_mul((a), (b))
#define R_INT(i) zq_int((i))
#include "prime_ops.inc"
#include "crt_ops.inc"
#undef RT
#undef RN
#undef R_ADD
#undef R_SUB
#undef R_MUL
#undef R_INT

#define RT int64_t
#define RN(n) n##_zz
#define R_ADD(a, b) zz_add((a), (b))
#define R_SUB(a, b) zz_sub((a), (b))
#define R_MUL(a, b) zz_mul((a), (b))
#define R_INT(i) ((int64_t)(i))
#include "prime_ops.inc"
#undef RT
#undef RN
#undef R_ADD
#undef R_SUB
#undef R_MUL
#undef R_INT

#define RT double
#define RN(n) n##_dd
#define R_ADD(a, b) ((a) + (b))
#define R_SUB(a, b) ((a) - (b))
#define R_MUL(a, b) ((a) * (b))
#define R_INT(i) ((double)(i))
#include "prime_ops.inc"
#undef RT
#undef RN
#undef R_ADD
#undef R_SUB
#undef R_MUL
#undef R_INT

#define RT lo_cplx_t
#define RN(n) n##_cx
#define R_ADD(a, b) cx_add((a), (b))
#define R_SUB(a, b) cx_sub((a), (b))
#define R_MUL(a, b) cx_mul((a), (b))
#define R_INT(i) cx_int((i))
#include "prime_ops.inc"
#include "crt_ops.inc"
#undef RT
#undef RN
#undef R_ADD
#undef R_SUB
#undef R_MUL
#undef R_INT

/* ---------- drivers -------------------------------------------------------- */

/* reference: tensor.h:39-74 (tensorFuserPrime): for prime power i the prime
 * operator runs at (lts * p^{e-1}, rts), then rts *= phi(p^e). */
#define DEFINE_FUSE_PRIME(SUF, T)                                                        \
  typedef void (*prime_fn_##SUF)(T*, int, int64_t, int64_t, int);                        \
  static void fuse_prime_##SUF(T* y, int k, prime_fn_##SUF f, int64_t totm,              \
                               const lo_pe_t* pe, int npe, const int64_t* qs)            \
  {                                                                                      \
    int64_t lts = totm, rts = 1;                                                         \
    for (int i = 0; i < npe; i++) {                                                      \
      int64_t pem1 = lo_ipow(pe[i].prime, pe[i].exponent - 1);                           \
      int64_t phi = (pe[i].prime - 1) * pem1;                                            \
      lts /= phi;                                                                        \
      for (int limb = 0; limb < k; limb++) {                                             \
        if (qs) lo_set_modulus(qs[limb]);                                                \
        f(y + limb, k, lts * pem1, rts, pe[i].prime);                                    \
      }                                                                                  \
      rts *= phi;                                                                        \
    }                                                                                    \
  }
DEFINE_FUSE_PRIME(zq, int64_t)
DEFINE_FUSE_PRIME(zz, int64_t)
DEFINE_FUSE_PRIME(dd, double)
DEFINE_FUSE_PRIME(cx, lo_cplx_t)

/* bring every limb into [0,q) (the reference expects canonical input from
 * Haskell; zq.cpp:57-67 is the exit-side counterpart) */
static void zq_reduce_all(int64_t* y, int k, int64_t totm, const int64_t* qs)
{
  for (int limb = 0; limb < k; limb++) {
    lo_set_modulus(qs[limb]);
    for (int64_t j = 0; j < totm; j++) y[j * k + limb] = zq_int(y[j * k + limb]);
  }
}

/* ---------- CRT ------------------------------------------------------------ */

void lo_tensorCRTRq(int16_t k, int64_t* y, int32_t totm, const lo_pe_t* pe, int16_t npe, int64_t** ru, const int64_t* qs)
{
  zq_reduce_all(y, k, totm, qs);
  fuse_crt_zq(y, k, totm, pe, npe, ru, qs, 0);
}

void lo_tensorCRTInvRq(int16_t k, int64_t* y, int32_t totm, const lo_pe_t* pe, int16_t npe, int64_t** ruinv, const int64_t* mhatInv, const int64_t* qs)
{
  zq_reduce_all(y, k, totm, qs);
  fuse_crt_zq(y, k, totm, pe, npe, ruinv, qs, 1);
  for (int limb = 0; limb < k; limb++) {          /* crt.cpp:573-579 */
    lo_set_modulus(qs[limb]);
    int64_t s = zq_int(mhatInv[limb]);
    for (int64_t j = 0; j < totm; j++) y[j * k + limb] = zq_mul(y[j * k + limb], s);
  }
}

void lo_tensorCRTC(int16_t k, lo_cplx_t* y, int32_t totm, const lo_pe_t* pe, int16_t npe, lo_cplx_t** ru)
{
  fuse_crt_cx(y, k, totm, pe, npe, ru, 0, 0);
}

void lo_tensorCRTInvC(int16_t k, lo_cplx_t* y, int32_t totm, const lo_pe_t* pe, int16_t npe, lo_cplx_t** ruinv, const lo_cplx_t* mhatInv)
{
  fuse_crt_cx(y, k, totm, pe, npe, ruinv, 0, 1);
  for (int limb = 0; limb < k; limb++)            /* crt.cpp:592-597 */
    for (int64_t j = 0; j < totm; j++) y[j * k + limb] = cx_mul(y[j * k + limb], mhatInv[limb]);
}

/* ---------- L, G ----------------------------------------------------------- */

#define DEFINE_RQ(NAME, OP)                                                              \
  void lo_tensor##NAME##Rq(int16_t k, int64_t* y, int32_t totm, const lo_pe_t* pe, int16_t npe, const int64_t* qs) \
  { zq_reduce_all(y, k, totm, qs); fuse_prime_zq(y, k, OP##_zq, totm, pe, npe, qs); }
#define DEFINE_PLAIN(NAME, TAG, SUF, T, OP)                                              \
  void lo_tensor##NAME##TAG(int16_t k, T* y, int32_t totm, const lo_pe_t* pe, int16_t npe) \
  { fuse_prime_##SUF(y, k, OP##_##SUF, totm, pe, npe, 0); }

DEFINE_RQ(L, op_L)          /* l.cpp:109-115 */
DEFINE_RQ(LInv, op_LInv)    /* l.cpp:150-156 */
DEFINE_RQ(GPow, op_GPow)    /* g.cpp:130-134 */
DEFINE_RQ(GDec, op_GDec)    /* g.cpp:146-150 */
DEFINE_PLAIN(L, R, zz, int64_t, op_L)             /* l.cpp:125-129 */
DEFINE_PLAIN(LInv, R, zz, int64_t, op_LInv)       /* l.cpp:166-170 */
DEFINE_PLAIN(L, Double, dd, double, op_L)         /* l.cpp:131-134 */
DEFINE_PLAIN(LInv, Double, dd, double, op_LInv)   /* l.cpp:172-175 */
DEFINE_PLAIN(L, C, cx, lo_cplx_t, op_L)           /* l.cpp:136-139 */
DEFINE_PLAIN(LInv, C, cx, lo_cplx_t, op_LInv)     /* l.cpp:177-180 */
DEFINE_PLAIN(GPow, R, zz, int64_t, op_GPow)       /* g.cpp:125-128 */
DEFINE_PLAIN(GDec, R, zz, int64_t, op_GDec)       /* g.cpp:141-144 */
DEFINE_PLAIN(GPow, C, cx, lo_cplx_t, op_GPow)     /* g.cpp:136-139 */
DEFINE_PLAIN(GDec, C, cx, lo_cplx_t, op_GDec)     /* g.cpp:152-155 */

/* reference: g.cpp:186-207, 239-260: prime transform (scaled by p per odd
 * prime), then times rad_odd(m)^{-1} mod q per limb; 0 when not invertible. */
static int16_t ginv_rq(int k, int64_t* y, int64_t totm, const lo_pe_t* pe, int npe, const int64_t* qs, prime_fn_zq f)
{
  zq_reduce_all(y, k, totm, qs);
  fuse_prime_zq(y, k, f, totm, pe, npe, qs);
  int64_t rad = lo_odd_radical(pe, npe);
  for (int limb = 0; limb < k; limb++) {
    int64_t inv = lo_mod_inverse(qs[limb], rad);
    if (inv == 0) return 0;
    lo_set_modulus(qs[limb]);
    for (int64_t j = 0; j < totm; j++) y[j * k + limb] = zq_mul(y[j * k + limb], inv);
  }
  return 1;
}
int16_t lo_tensorGInvPowRq(int16_t k, int64_t* y, int32_t totm, const lo_pe_t* pe, int16_t npe, const int64_t* qs) { return ginv_rq(k, y, totm, pe, npe, qs, op_GInvPow_zq); }
int16_t lo_tensorGInvDecRq(int16_t k, int64_t* y, int32_t totm, const lo_pe_t* pe, int16_t npe, const int64_t* qs) { return ginv_rq(k, y, totm, pe, npe, qs, op_GInvDec_zq); }

/* intended semantics of g.cpp:169-184, 222-237 (see header): every entry must
 * be an exact multiple of rad_odd(m); on failure y holds the undivided
 * transform and 0 is returned. */
static int16_t ginv_r(int k, int64_t* y, int64_t totm, const lo_pe_t* pe, int npe, prime_fn_zz f)
{
  fuse_prime_zz(y, k, f, totm, pe, npe, 0);
  int64_t rad = lo_odd_radical(pe, npe);
  for (int64_t i = 0; i < totm * k; i++) if (y[i] % rad != 0) return 0;
  for (int64_t i = 0; i < totm * k; i++) y[i] /= rad;
  return 1;
}
int16_t lo_tensorGInvPowR(int16_t k, int64_t* y, int32_t totm, const lo_pe_t* pe, int16_t npe) { return ginv_r(k, y, totm, pe, npe, op_GInvPow_zz); }
int16_t lo_tensorGInvDecR(int16_t k, int64_t* y, int32_t totm, const lo_pe_t* pe, int16_t npe) { return ginv_r(k, y, totm, pe, npe, op_GInvDec_zz); }

/* intended semantics of g.cpp:209-220, 262-273: real division by rad_odd(m) */
static int16_t ginv_c(int k, lo_cplx_t* y, int64_t totm, const lo_pe_t* pe, int npe, prime_fn_cx f)
{
  fuse_prime_cx(y, k, f, totm, pe, npe, 0);
  double inv = 1.0 / (double)lo_odd_radical(pe, npe);
  for (int64_t i = 0; i < totm * k; i++) { y[i].re *= inv; y[i].im *= inv; }
  return 1;
}
int16_t lo_tensorGInvPowC(int16_t k, lo_cplx_t* y, int32_t totm, const lo_pe_t* pe, int16_t npe) { return ginv_c(k, y, totm, pe, npe, op_GInvPow_cx); }
int16_t lo_tensorGInvDecC(int16_t k, lo_cplx_t* y, int32_t totm, const lo_pe_t* pe, int16_t npe) { return ginv_c(k, y, totm, pe, npe, op_GInvDec_cx); }

/* ---------- norm ----------------------------------------------------------- */

/* reference: norm.cpp:39-59: y[limb] <- < y, ((x)(I+J)) y > */
void lo_tensorNormSqR(int16_t k, int64_t* y, int32_t totm, const lo_pe_t* pe, int16_t npe)
{
  size_t cnt = (size_t)totm * (size_t)k;
  int64_t* orig = (int64_t*)malloc(cnt * sizeof(int64_t));
  memcpy(orig, y, cnt * sizeof(int64_t));
  fuse_prime_zz(y, k, op_NormSq_zz, totm, pe, npe, 0);
  for (int limb = 0; limb < k; limb++) {
    int64_t dot = 0;
    for (int64_t j = 0; j < totm; j++) dot = zz_add(dot, zz_mul(orig[j * k + limb], y[j * k + limb]));
    y[limb] = dot;
  }
  free(orig);
}

/* reference: norm.cpp:61-80 */
void lo_tensorNormSqD(int16_t k, double* y, int32_t totm, const lo_pe_t* pe, int16_t npe)
{
  size_t cnt = (size_t)totm * (size_t)k;
  double* orig = (double*)malloc(cnt * sizeof(double));
  memcpy(orig, y, cnt * sizeof(double));
  fuse_prime_dd(y, k, op_NormSq_dd, totm, pe, npe, 0);
  for (int limb = 0; limb < k; limb++) {
    double dot = 0;
    for (int64_t j = 0; j < totm; j++) dot += orig[j * k + limb] * y[j * k + limb];
    y[limb] = dot;
  }
  free(orig);
}

/* ---------- Gaussian ------------------------------------------------------- */

/* reference: random.cpp:19-50 (primeD): along each line of length p-1,
 * out[row] = (1/sqrt2) * sum_col 2 * c(row, col) * in[col-1], col = 1..p-1,
 * c = Re ru[(row*col mod p) * rustride] for col <= p/2, Im ... otherwise. */
static void gauss_prime(double* y, int k, int64_t lts, int64_t rts, int p, int64_t rustride, const lo_cplx_t* ru)
{
  if (p == 2) return;
  double* tmp = (double*)malloc(sizeof(double) * (size_t)p);
  for (int64_t blk = 0; blk < lts; blk++)
    for (int64_t r = 0; r < rts; r++) {
      for (int row = 0; row < p - 1; row++) {
        double acc = 0;
        for (int col = 1; col <= p - 1; col++) {
          const lo_cplx_t w = ru[(((int64_t)row * col) % p) * rustride * k];
          double c = (col <= (p >> 1)) ? w.re : w.im;
          acc += 2 * c * y[((blk * (p - 1) + (col - 1)) * rts + r) * k];
        }
        tmp[row] = acc / sqrt(2);
      }
      for (int row = 0; row < p - 1; row++) y[((blk * (p - 1) + row) * rts + r) * k] = tmp[row];
    }
  free(tmp);
}

/* reference: random.cpp:52-64 (ppD, tensorGaussianDec) on tensor.h:76-95 */
void lo_tensorGaussianDec(int16_t k, double* y, int32_t totm, const lo_pe_t* pe, int16_t npe, lo_cplx_t** ru)
{
  int64_t lts = totm, rts = 1;
  for (int i = 0; i < npe; i++) {
    int64_t pem1 = lo_ipow(pe[i].prime, pe[i].exponent - 1);
    int64_t phi = (pe[i].prime - 1) * pem1;
    lts /= phi;
    for (int limb = 0; limb < k; limb++)
      gauss_prime(y + limb, k, lts * pem1, rts, pe[i].prime, pem1, ru[i] + limb);
    rts *= phi;
  }
}

/* ---------- pointwise ------------------------------------------------------ */

/* reference: mul.cpp:14-30 */
void lo_mulRq(int16_t k, int64_t* a, const int64_t* b, int32_t totm, const int64_t* qs)
{
  for (int limb = 0; limb < k; limb++) {
    lo_set_modulus(qs[limb]);
    for (int64_t j = 0; j < totm; j++)
      a[j * k + limb] = zq_mul(zq_int(a[j * k + limb]), zq_int(b[j * k + limb]));
  }
}

/* reference: mul.cpp:32-35 */
void lo_mulC(int16_t k, lo_cplx_t* a, const lo_cplx_t* b, int32_t totm)
{
  for (int64_t i = 0; i < (int64_t)totm * k; i++) a[i] = cx_mul(a[i], b[i]);
}
