// coeff_stream.cu -- the coefficient-wise maps Lol applies either side of the tensor transforms when it switches moduli
// or rounds an error term.  In the reference each is a host `fmapT` closure over the
// Storable vector (lol/Crypto/Lol/Cyclotomic/UCyc.hs:267-300, 427-445); here each is one streaming pass over the
// device-resident batch, [batch][n][k] layout, one 8-byte word per thread and four independent words in flight:
//
//   lift        y = decode'(x) per limb: the representative in [-q/2, q/2)           ZqBasic.hs:92-94; UCyc.hs:288-296
//   reduce      y[t] = z mod q_t (non-negative), z one int64 per coefficient or per limb   ZqBasic.hs:88-90; UCyc.hs:267-275
//   rescaleDrop (x_1..x_k) -> q_d^-1 (x_t - reduce(lift x_d)) for t != d: removes limb d of the product ring
//               (Prelude.hs:226-232, 259-265: `Rescale (a,b) b` and `Rescale (a,b) a`; rescalePow, UCyc.hs:298-300;
//               the fast path of rescaleCyc, Cyc.hs:529-541)
//   rescaleMod  y = fst (divModCent (q' * lift x) q) mod q'  per limb                Prelude.hs:143-153; Numeric.hs:227-234
//   roundCoset  y = rep + p * round((e - rep) / p), rep = lift zp; without zp: y = round e (roundMult 1)
//               (Prelude.hs:155-162; Numeric.hs:207-210; errorRounded / errorCoset, UCyc.hs:427-445)
//
// All integer results are exact and equal the host formulas bit for bit; roundCoset performs the same IEEE double
// operations in the same order (one subtraction, one division, round-half-even).  Bytes per coefficient word: 16
// (lift, rescaleMod), 8 + 8 k'/k (reduce), 8 k/(k-1) + 8 (rescaleDrop), 24 (roundCoset with zp), 16 (without).
#include "fused.cuh"
#include "numtheory.h"

using namespace lolb;

namespace {

// x mod q for any uint64 x, q < 2^32, mu = floor(2^64 / q)
__device__ __forceinline__ int64_t barrett_u64(uint64_t x, uint32_t q, uint64_t mu)
{
  uint64_t r = x - __umul64hi(x, mu) * q;                          // [0, 3q)
  if (r >= q) r -= q;
  if (r >= q) r -= q;
  return (int64_t)r;
}

// Haskell `mod` (non-negative) of any int64; canonical inputs take the first branch
__device__ __forceinline__ int64_t canon64(int64_t x, uint32_t q, uint64_t mu)
{
  if ((uint64_t)x < (uint64_t)q) return x;
  if (x >= 0) return barrett_u64((uint64_t)x, q, mu);
  const int64_t r = barrett_u64(0ull - (uint64_t)x, q, mu);        // |x| mod q
  return r == 0 ? 0 : (int64_t)q - r;
}

__device__ __forceinline__ int64_t lift64(int64_t c, int64_t q) { return 2 * c < q ? c : c - q; }      // ZqBasic.hs:92-94

// floor(a / q) for any int64 a (Haskell `div`), by a Barrett quotient with at most two corrections
__device__ __forceinline__ int64_t floor_div(int64_t a, uint32_t q, uint64_t mu)
{
  const bool neg = a < 0;
  const uint64_t x = neg ? 0ull - (uint64_t)a - 1 : (uint64_t)a;  // a < 0: floor(a / q) = -(floor((|a| - 1) / q) + 1)
  uint64_t qh = __umul64hi(x, mu);
  uint64_t r = x - qh * q;
  if (r >= q) { r -= q; qh++; }
  if (r >= q) { r -= q; qh++; }
  return neg ? -(int64_t)qh - 1 : (int64_t)qh;
}

// Every operator sees an output word as (g, u): coefficient index g (over batch * n) and position u inside the
// coefficient's tuple of `period` words.  The launch makes the grid stride a multiple of the period, so u (the limb) and
// everything derived from it are loop invariants of a thread and no division runs inside the loop.
struct OpLift {
  const long long* x; long long* y; int k; ZqConsts Z;
  typedef long long In;
  __device__ int period() const { return k; }
  __device__ In load(int64_t g, int u) const { return __ldcs(x + g * k + u); }
  __device__ void apply(int64_t g, int u, In v) const
  {
    const uint32_t q = Z.q[u];
    __stcs(y + g * k + u, (long long)lift64(canon64(v, q, Z.mu[u]), q));
  }
};

struct OpReduce {
  const long long* z; long long* y; int k; int kz; ZqConsts Z;
  typedef long long In;
  __device__ int period() const { return k; }
  __device__ In load(int64_t g, int u) const { return kz == k ? __ldcs(z + g * k + u) : __ldg(z + g); }
  __device__ void apply(int64_t g, int u, In v) const { __stcs(y + g * k + u, (long long)canon64(v, Z.q[u], Z.mu[u])); }
};

struct OpRescaleDrop {
  const long long* x; long long* y; int k; int d; ZqConsts Z;      // Z.scale[t] = q_d^-1 mod q_t
  typedef longlong2 In;                                            // (x_t, x_d)
  __device__ int period() const { return k - 1; }
  __device__ int limb(int u) const { return u < d ? u : u + 1; }
  __device__ In load(int64_t g, int u) const { return make_longlong2(__ldg(x + g * k + limb(u)), __ldg(x + g * k + d)); }
  __device__ void apply(int64_t g, int u, In v) const
  {
    const int t = limb(u);
    const uint32_t qt = Z.q[t], qd = Z.q[d];
    const int64_t xt = canon64(v.x, qt, Z.mu[t]);
    const int64_t z = lift64(canon64(v.y, qd, Z.mu[d]), qd);     // lift x_d
    int64_t diff = xt - canon64(z, qt, Z.mu[t]);                 // x_t - reduce z
    if (diff < 0) diff += qt;
    __stcs(y + g * (k - 1) + u, barrett_u64((uint64_t)diff * Z.scale[t], qt, Z.mu[t]));
  }
};

// The limb drop in 32-bit arithmetic for canonical inputs (measured on B200, m = 14400, k = 2: 88 % of the HBM peak against
// 72 % for the three 64-bit Barrett steps of OpRescaleDrop, which -DLOLB_COEFF_FAST=0 still builds).  reduce(lift x_d) = (x_d mod q_t) - [2 x_d >= q_d] (q_d mod q_t)  (mod q_t), with
// x_d mod q_t from a 32-bit Barrett quotient (mu >> 32), and the product with the constant q_d^-1 by Shoup's method
// (w' = floor(w 2^32 / q_t)): 4 multiply-adds per word instead of three 64-bit Barrett steps.
#ifndef LOLB_COEFF_FAST
#define LOLB_COEFF_FAST 1
#endif
struct OpRescaleDropFast {
  const long long* x; long long* y; int k; int d; ZqConsts Z;      // Z.scale[t] = w = q_d^-1 mod q_t
  uint32_t dm[kMaxLimbs];                                           // q_d mod q_t
  uint32_t shoup[kMaxLimbs];                                        // floor(w 2^32 / q_t)
  typedef longlong2 In;
  __device__ int period() const { return k - 1; }
  __device__ int limb(int u) const { return u < d ? u : u + 1; }
  __device__ In load(int64_t g, int u) const { return make_longlong2(__ldg(x + g * k + limb(u)), __ldg(x + g * k + d)); }
  __device__ void apply(int64_t g, int u, In v) const
  {
    const int t = limb(u);
    const uint32_t qt = Z.q[t], qd = Z.q[d];
    const uint32_t xt = (uint32_t)canon64(v.x, qt, Z.mu[t]);
    const uint32_t c = (uint32_t)canon64(v.y, qd, Z.mu[d]);
    // r = c mod q_t: quotient estimate from the top half of mu, at most two corrections
    const uint32_t qh = (uint32_t)(((uint64_t)c * (uint32_t)(Z.mu[t] >> 32)) >> 32);
    uint64_t r = (uint64_t)c - (uint64_t)qh * qt;
    if (r >= qt) r -= qt;
    if (r >= qt) r -= qt;
    if (r >= qt) r -= qt;
    uint32_t zr = (uint32_t)r;                                      // reduce(lift x_d)
    if (2 * (uint64_t)c >= qd) zr = zr >= dm[t] ? zr - dm[t] : zr + (qt - dm[t]);
    const uint32_t diff = xt >= zr ? xt - zr : xt + (qt - zr);      // x_t - reduce z
    // diff * w mod q_t, w constant: Shoup
    const uint32_t sh = (uint32_t)(((uint64_t)diff * shoup[t]) >> 32);
    uint64_t o = (uint64_t)diff * Z.scale[t] - (uint64_t)sh * qt;   // in [0, 2 q_t)
    if (o >= qt) o -= qt;
    __stcs(y + g * (k - 1) + u, (long long)o);
  }
};

struct OpRescaleMod {
  const long long* x; long long* y; int k; ZqConsts Z; uint32_t q2[kMaxLimbs];
  typedef long long In;
  __device__ int period() const { return k; }
  __device__ In load(int64_t g, int u) const { return __ldcs(x + g * k + u); }
  __device__ void apply(int64_t g, int u, In v) const
  {
    const uint32_t q = Z.q[u];
    const int64_t qn = q2[u];
    const int64_t a = qn * lift64(canon64(v, q, Z.mu[u]), q) + q / 2;      // |q' lift x| < 2^63: q' < 2^32, |lift x| <= 2^31
    int64_t quot = floor_div(a, q, Z.mu[u]);                                // -ceil(q'/2) <= quot <= floor((q'+1)/2)
    if (quot < 0) quot += qn;                                               // fromIntegral into Z_q': one correction is enough
    if (quot >= qn) quot -= qn;                                             // (a `while` here compiles to a 64-bit division call)
    __stcs(y + g * k + u, (long long)quot);
  }
};

struct OpRoundCoset {
  const double* e; const long long* zp; long long* y; int k; ZqConsts Z;
  struct In { double e; long long z; };
  __device__ int period() const { return k; }
  __device__ In load(int64_t g, int u) const { return In{__ldcs(e + g * k + u), zp ? __ldcs(zp + g * k + u) : 0}; }
  __device__ void apply(int64_t g, int u, In v) const
  {
    if (!zp) { __stcs(y + g * k + u, (long long)rint(v.e)); return; }      // roundMult 1 = round (half to even)
    const uint32_t p = Z.q[u];
    const int64_t rep = lift64(canon64(v.z, p, Z.mu[u]), p);
    const double r = __ddiv_rn(__dsub_rn(v.e, (double)rep), (double)p);
    __stcs(y + g * k + u, (long long)(rep + (int64_t)p * (int64_t)rint(r)));
  }
};

template <class OP>
__global__ void __launch_bounds__(256) k_coeff_stream(const __grid_constant__ OP op, int64_t groups)
{
  const int period = op.period();
  const int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t gstep = (int64_t)gridDim.x * blockDim.x / period;            // the launch makes this exact
  const int u = (int)(i0 % period);
  int64_t g = i0 / period;
  for (; g + 3 * gstep < groups; g += 4 * gstep) {
    typename OP::In v[4];
#pragma unroll
    for (int a = 0; a < 4; a++) v[a] = op.load(g + a * gstep, u);
#pragma unroll
    for (int a = 0; a < 4; a++) op.apply(g + a * gstep, u, v[a]);
  }
  for (; g < groups; g += gstep) op.apply(g, u, op.load(g, u));
}

template <class OP>
int launch(const lolb_plan* pl, const OP& op, int64_t groups, int period, void* stream, const char* what)
{
  if (groups <= 0) return LOLB_OK;
  const int64_t total = groups * period;
  int64_t blocks = (total + 1023) / 1024;                         // four words per thread
  const int64_t cap = (int64_t)pl->num_sms * 16;
  if (blocks > cap) blocks = cap;
  blocks = (blocks + period - 1) / period * period;               // grid stride a multiple of the period: a thread keeps its limb
  k_coeff_stream<OP><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(op, groups);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  count_launch();
  return LOLB_OK;
}

int check(const lolb_plan* pl, const void* a, const void* b, int64_t batch, const char* fn)
{
  if (!pl || pl->kind != PLAN_RQ) { set_error(std::string(fn) + ": needs an Rq plan"); return LOLB_ERR_ARG; }
  if (batch < 0 || (batch > 0 && (!a || !b))) { set_error(std::string(fn) + ": bad batch or NULL operand"); return LOLB_ERR_ARG; }
  return LOLB_OK;
}

}  // namespace

extern "C" int lolb_liftRq(const lolb_plan* plan, const hInt_t* x, hInt_t* y, int64_t batch, void* stream)
{
  int rc = check(plan, x, y, batch, __func__);
  if (rc) return rc;
  OpLift op{(const long long*)x, (long long*)y, plan->k, plan->zq_plain};
  return launch(plan, op, batch * plan->n, plan->k, stream, "k_coeff_stream<lift>");
}

extern "C" int lolb_reduceRq(const lolb_plan* plan, const hInt_t* z, int z_tupsize, hInt_t* y, int64_t batch, void* stream)
{
  int rc = check(plan, z, y, batch, __func__);
  if (rc) return rc;
  if (z_tupsize != 1 && z_tupsize != plan->k) { set_error("lolb_reduceRq: z_tupsize must be 1 or the plan's tupSize"); return LOLB_ERR_ARG; }
  OpReduce op{(const long long*)z, (long long*)y, plan->k, z_tupsize, plan->zq_plain};
  return launch(plan, op, batch * plan->n, plan->k, stream, "k_coeff_stream<reduce>");
}

extern "C" int lolb_rescaleDropRq(const lolb_plan* plan, int drop, const hInt_t* x, hInt_t* y, int64_t batch, void* stream)
{
  int rc = check(plan, x, y, batch, __func__);
  if (rc) return rc;
  if (plan->k < 2 || drop < 0 || drop >= plan->k) { set_error("lolb_rescaleDropRq: needs tupSize >= 2 and 0 <= drop < tupSize"); return LOLB_ERR_ARG; }
  if ((const void*)x == (const void*)y && batch > 0) { set_error("lolb_rescaleDropRq: operands must not alias"); return LOLB_ERR_ARG; }
#if LOLB_COEFF_FAST
  OpRescaleDropFast op{(const long long*)x, (long long*)y, plan->k, drop, plan->zq_plain, {}, {}};
#else
  OpRescaleDrop op{(const long long*)x, (long long*)y, plan->k, drop, plan->zq_plain};
#endif
  for (int t = 0; t < plan->k; t++) {
    if (t == drop) continue;
    const int64_t inv = mod_inverse(plan->qs[t], plan->qs[drop] % plan->qs[t]);      // recip (reduce q_d): `Field b`
    if (inv == 0) { set_error("lolb_rescaleDropRq: the dropped modulus is not invertible modulo another limb"); return LOLB_ERR_NOT_INVERTIBLE; }
    op.Z.scale[t] = (uint32_t)inv;
#if LOLB_COEFF_FAST
    op.dm[t] = (uint32_t)(plan->qs[drop] % plan->qs[t]);
    op.shoup[t] = (uint32_t)((((u128)(uint64_t)inv) << 32) / (uint64_t)plan->qs[t]);
#endif
  }
  return launch(plan, op, batch * plan->n, plan->k - 1, stream, "k_coeff_stream<rescaleDrop>");
}

extern "C" int lolb_rescaleModRq(const lolb_plan* plan, const hInt_t* qs_new, const hInt_t* x, hInt_t* y, int64_t batch, void* stream)
{
  int rc = check(plan, x, y, batch, __func__);
  if (rc) return rc;
  if (!qs_new) { set_error("lolb_rescaleModRq: NULL target moduli"); return LOLB_ERR_ARG; }
  OpRescaleMod op{(const long long*)x, (long long*)y, plan->k, plan->zq_plain, {}};
  for (int t = 0; t < plan->k; t++) {
    if (qs_new[t] < 1 || qs_new[t] >= ((int64_t)1 << 32)) { set_error("lolb_rescaleModRq: target modulus out of range [1, 2^32)"); return LOLB_ERR_ARG; }
    op.q2[t] = (uint32_t)qs_new[t];
  }
  return launch(plan, op, batch * plan->n, plan->k, stream, "k_coeff_stream<rescaleMod>");
}

extern "C" int lolb_roundCosetRq(const lolb_plan* plan, const double* e, const hInt_t* zp, hInt_t* y, int64_t batch, void* stream)
{
  int rc = check(plan, e, y, batch, __func__);
  if (rc) return rc;
  OpRoundCoset op{e, (const long long*)zp, (long long*)y, plan->k, plan->zq_plain};
  return launch(plan, op, batch * plan->n, plan->k, stream, "k_coeff_stream<roundCoset>");
}
