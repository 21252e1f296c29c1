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

__device__ __forceinline__ int64_t canon64(int64_t x, int64_t q)
{
  if ((uint64_t)x < (uint64_t)q) return x;
  const int64_t r = x % q;
  return r < 0 ? r + q : r;
}

__device__ __forceinline__ int64_t lift64(int64_t c, int64_t q) { return 2 * c < q ? c : c - q; }      // ZqBasic.hs:92-94

__device__ __forceinline__ uint32_t barrett_mul(uint32_t a, uint32_t b, uint32_t q, uint64_t mu)
{
  const uint64_t x = (uint64_t)a * b;
  uint64_t r = x - __umul64hi(x, mu) * q;
  if (r >= q) r -= q;
  if (r >= q) r -= q;
  return (uint32_t)r;
}

struct OpLift {
  const long long* x; long long* y; int k; ZqConsts Z;
  typedef long long In;
  __device__ In load(int64_t i) const { return __ldcs(x + i); }
  __device__ void apply(int64_t i, In v) const
  {
    const int64_t q = Z.q[(int)(i % k)];
    __stcs(y + i, (long long)lift64(canon64(v, q), q));
  }
};

struct OpReduce {
  const long long* z; long long* y; int k; int kz; ZqConsts Z;
  typedef long long In;
  __device__ In load(int64_t i) const { return kz == k ? __ldcs(z + i) : __ldg(z + i / k); }
  __device__ void apply(int64_t i, In v) const { __stcs(y + i, (long long)canon64(v, (int64_t)Z.q[(int)(i % k)])); }
};

struct OpRescaleDrop {
  const long long* x; long long* y; int k; int d; ZqConsts Z;      // Z.scale[t] = q_d^-1 mod q_t
  typedef longlong2 In;                                            // (x_t, x_d)
  __device__ int limb(int64_t i, int64_t* c) const
  {
    *c = i / (k - 1);
    const int u = (int)(i - *c * (k - 1));
    return u < d ? u : u + 1;
  }
  __device__ In load(int64_t i) const
  {
    int64_t c;
    const int t = limb(i, &c);
    return make_longlong2(__ldg(x + c * k + t), __ldg(x + c * k + d));
  }
  __device__ void apply(int64_t i, In v) const
  {
    int64_t c;
    const int t = limb(i, &c);
    const int64_t qt = Z.q[t], qd = Z.q[d];
    const int64_t xt = canon64(v.x, qt);
    const int64_t z = lift64(canon64(v.y, qd), qd);              // lift x_d
    int64_t diff = xt - canon64(z, qt);                          // x_t - reduce z
    if (diff < 0) diff += qt;
    __stcs(y + i, (long long)barrett_mul((uint32_t)diff, Z.scale[t], (uint32_t)qt, Z.mu[t]));
  }
};

struct OpRescaleMod {
  const long long* x; long long* y; int k; ZqConsts Z; uint32_t q2[kMaxLimbs];
  typedef long long In;
  __device__ In load(int64_t i) const { return __ldcs(x + i); }
  __device__ void apply(int64_t i, In v) const
  {
    const int t = (int)(i % k);
    const int64_t q = Z.q[t], qn = q2[t];
    const int64_t a = qn * lift64(canon64(v, q), q) + q / 2;     // |q' lift x| < 2^63: q' < 2^32, |lift x| <= 2^31
    int64_t quot = a / q;                                        // divMod: floor
    if (a % q < 0) quot -= 1;
    __stcs(y + i, (long long)canon64(quot, qn));                 // fromIntegral into Z_q'
  }
};

struct OpRoundCoset {
  const double* e; const long long* zp; long long* y; int k; ZqConsts Z;
  struct In { double e; long long z; };
  __device__ In load(int64_t i) const { return In{__ldcs(e + i), zp ? __ldcs(zp + i) : 0}; }
  __device__ void apply(int64_t i, In v) const
  {
    if (!zp) { __stcs(y + i, (long long)rint(v.e)); return; }    // roundMult 1 = round (half to even)
    const int64_t p = Z.q[(int)(i % k)];
    const int64_t rep = lift64(canon64(v.z, p), p);
    const double r = __ddiv_rn(__dsub_rn(v.e, (double)rep), (double)p);
    __stcs(y + i, (long long)(rep + (p == 1 ? (int64_t)rint(__dsub_rn(v.e, (double)rep)) : p * (int64_t)rint(r))));
  }
};

template <class OP>
__global__ void __launch_bounds__(256) k_coeff_stream(const __grid_constant__ OP op, int64_t total)
{
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < total; i += 4 * stride) {
    typename OP::In v[4];
#pragma unroll
    for (int a = 0; a < 4; a++) v[a] = op.load(i + a * stride);
#pragma unroll
    for (int a = 0; a < 4; a++) op.apply(i + a * stride, v[a]);
  }
  for (; i < total; i += stride) op.apply(i, op.load(i));
}

template <class OP>
int launch(const lolb_plan* pl, const OP& op, int64_t total, void* stream, const char* what)
{
  if (total <= 0) return LOLB_OK;
  int64_t blocks = (total + 1023) / 1024;                         // four words per thread
  const int64_t cap = (int64_t)pl->num_sms * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_coeff_stream<OP><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(op, total);
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
  return launch(plan, op, batch * plan->n * plan->k, stream, "k_coeff_stream<lift>");
}

extern "C" int lolb_reduceRq(const lolb_plan* plan, const hInt_t* z, int z_tupsize, hInt_t* y, int64_t batch, void* stream)
{
  int rc = check(plan, z, y, batch, __func__);
  if (rc) return rc;
  if (z_tupsize != 1 && z_tupsize != plan->k) { set_error("lolb_reduceRq: z_tupsize must be 1 or the plan's tupSize"); return LOLB_ERR_ARG; }
  OpReduce op{(const long long*)z, (long long*)y, plan->k, z_tupsize, plan->zq_plain};
  return launch(plan, op, batch * plan->n * plan->k, stream, "k_coeff_stream<reduce>");
}

extern "C" int lolb_rescaleDropRq(const lolb_plan* plan, int drop, const hInt_t* x, hInt_t* y, int64_t batch, void* stream)
{
  int rc = check(plan, x, y, batch, __func__);
  if (rc) return rc;
  if (plan->k < 2 || drop < 0 || drop >= plan->k) { set_error("lolb_rescaleDropRq: needs tupSize >= 2 and 0 <= drop < tupSize"); return LOLB_ERR_ARG; }
  if ((const void*)x == (const void*)y && batch > 0) { set_error("lolb_rescaleDropRq: operands must not alias"); return LOLB_ERR_ARG; }
  OpRescaleDrop op{(const long long*)x, (long long*)y, plan->k, drop, plan->zq_plain};
  for (int t = 0; t < plan->k; t++) {
    if (t == drop) continue;
    const int64_t inv = mod_inverse(plan->qs[t], plan->qs[drop] % plan->qs[t]);      // recip (reduce q_d): `Field b`
    if (inv == 0) { set_error("lolb_rescaleDropRq: the dropped modulus is not invertible modulo another limb"); return LOLB_ERR_NOT_INVERTIBLE; }
    op.Z.scale[t] = (uint32_t)inv;
  }
  return launch(plan, op, batch * plan->n * (plan->k - 1), stream, "k_coeff_stream<rescaleDrop>");
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
  return launch(plan, op, batch * plan->n * plan->k, stream, "k_coeff_stream<rescaleMod>");
}

extern "C" int lolb_roundCosetRq(const lolb_plan* plan, const double* e, const hInt_t* zp, hInt_t* y, int64_t batch, void* stream)
{
  int rc = check(plan, e, y, batch, __func__);
  if (rc) return rc;
  OpRoundCoset op{e, (const long long*)zp, (long long*)y, plan->k, plan->zq_plain};
  return launch(plan, op, batch * plan->n * plan->k, stream, "k_coeff_stream<roundCoset>");
}
