// ext_stream.cu -- the ring-extension operators of Lol's `Tensor` class for O_m'/O_m, m | m' (SURVEY.md section 8f
// rank 2).  In the reference these are host-side Haskell over boxed index vectors
// (lol-cpp/Crypto/Lol/Cyclotomic/Tensor/CPP/Extension.hs:54-129, index tables lol/Crypto/Lol/Cyclotomic/Tensor.hs:380-510);
// here the index tables are built once per (m, m') pair, kept on the device, and every operator is one streaming
// gather pass over the device-resident batch in the [batch][phi][k] layout of every other entry point:
//
//   twacePowDec  y[i]  = x[extIndicesPowDec[i]]                                   Extension.hs:99-103
//   embedPow     y[i'] = j0(i') == 0 ? x[j1(i')] : 0                              Extension.hs:60-70  (baseIndicesPow)
//   embedDec     y[i'] = Nothing -> 0 | (sh, neg) -> +-x[sh]                      Extension.hs:71-77  (baseIndicesDec)
//   embedCRT     y[i'] = x[baseIndicesCRT[i']]                                    Extension.hs:81-85
//   coeffs       y[i1][i0] = x[extIndicesCoeffs[i1][i0]]                          Extension.hs:90-93
//   twaceCRT     y[i]  = sum_{r < phi'/phi} (tweak . x)[extIndicesCRT[i phi'/phi + r]]
//                tweak = m'hat^-1 mhat embedCRT(gInvCRT_m) gCRT_m'                Extension.hs:110-129
//
// The five gathers share one kernel driven by a per-output code table (-1 = zero, else 2 source + negate).  A thread
// owns one position of the output element and walks the batch, so the code lookup, the limb and its modulus are loop
// invariants; consecutive threads write consecutive words.  Bytes per ring element: 8 k (phi_out + phi_read) for the
// gathers (phi_read = number of non-zero outputs), 8 k (phi' + phi) for twaceCRT, tables excluded (L1/L2 resident).
// Z_q arithmetic is exact on canonical residues (64-bit Barrett), so results equal the host formulas bit for bit.
#include <algorithm>
#include <cstring>

#include "fused.cuh"
#include "numtheory.h"
#include "rings.cuh"

using namespace lolb;

struct lolb_ext {
  const lolb_plan* lo = nullptr;   // O_m   (not owned)
  const lolb_plan* hi = nullptr;   // O_m'  (not owned)
  int kind = 0, k = 1;
  int32_t phi = 1, phi2 = 1, rel = 1;
  std::vector<int32_t> h_tab[LOLB_EXT_TABLES];
  int32_t* d_code[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // twace, embedPow, embedDec, embedCRT, coeffs
  int32_t* d_base_j = nullptr;     // [2][phi']: baseIndicesPow as (j0, j1) planes, for powBasisPow
  int32_t* d_crt_idx = nullptr;    // [rel][phi]: extIndicesCRT transposed so that consecutive threads read consecutive entries
  uint32_t* d_tweak = nullptr;     // Rq: [rel][phi][k] canonical residues
  double2* d_ctweak = nullptr;     // complex: [rel][phi][k]
};

namespace {

enum { CODE_TWACE = 0, CODE_EMBED_POW, CODE_EMBED_DEC, CODE_EMBED_CRT, CODE_COEFFS };

// ------------------------------------------------------------------ element words
// One "word" is what a thread moves per batch element: one coefficient of one limb (8 bytes: int64 / double),
// two adjacent limbs of one coefficient (16 bytes, int64 with an even tupSize) or one complex coefficient (16 bytes).
struct WordZq1 {
  typedef long long V;
  uint32_t q;
  __device__ WordZq1(const ZqConsts& Z, int u) : q(Z.q[u]) {}
  __device__ static V zero() { return 0; }
  __device__ V neg(V v) const
  {
    if ((unsigned long long)v >= q) { long long r = v % (long long)q; v = r < 0 ? r + q : r; }
    return v == 0 ? 0 : (long long)q - v;
  }
};
struct WordZq2 {
  typedef longlong2 V;
  WordZq1 a, b;
  __device__ WordZq2(const ZqConsts& Z, int u) : a(Z, 2 * u), b(Z, 2 * u + 1) {}
  __device__ static V zero() { return make_longlong2(0, 0); }
  __device__ V neg(V v) const { return make_longlong2(a.neg(v.x), b.neg(v.y)); }
};
struct WordI64 {
  typedef long long V;
  __device__ WordI64(const ZqConsts&, int) {}
  __device__ static V zero() { return 0; }
  __device__ V neg(V v) const { return (long long)(0ull - (unsigned long long)v); }
};
struct WordI64x2 {
  typedef longlong2 V;
  __device__ WordI64x2(const ZqConsts&, int) {}
  __device__ static V zero() { return make_longlong2(0, 0); }
  __device__ V neg(V v) const { return make_longlong2((long long)(0ull - (unsigned long long)v.x), (long long)(0ull - (unsigned long long)v.y)); }
};
struct WordF64 {
  typedef double V;
  __device__ WordF64(const ZqConsts&, int) {}
  __device__ static V zero() { return 0.0; }
  __device__ V neg(V v) const { return -v; }
};
struct WordC64 {
  typedef double2 V;
  __device__ WordC64(const ZqConsts&, int) {}
  __device__ static V zero() { return make_double2(0.0, 0.0); }
  __device__ V neg(V v) const { return make_double2(-v.x, -v.y); }
};

// y[b][j][u] <- code[j] < 0 ? 0 : (+-) x[b][code[j] >> 1][u];  kw words per coefficient
template <class W>
__global__ void __launch_bounds__(256)
k_ext_gather(const typename W::V* __restrict__ x, typename W::V* __restrict__ y, const int32_t* __restrict__ code,
             int32_t n_out, int32_t n_in, int32_t kw, int64_t batch, const __grid_constant__ ZqConsts Z)
{
  typedef typename W::V V;
  const int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= (int64_t)n_out * kw) return;
  const int32_t j = w / kw, u = w - j * kw;
  const int32_t c = __ldg(code + j);
  const int64_t so = (int64_t)n_out * kw, si = (int64_t)n_in * kw;
  const int64_t step = gridDim.y;
  V* yo = y + w;
  if (c < 0) {
    for (int64_t b = blockIdx.y; b < batch; b += step) __stcs(yo + b * so, W::zero());
    return;
  }
  const V* xi = x + (int64_t)(c >> 1) * kw + u;
  const W ops(Z, u);
  const bool neg = (c & 1) != 0;
  int64_t b = blockIdx.y;
  for (; b + 3 * step < batch; b += 4 * step) {          // four independent loads in flight
    V v0 = __ldcs(xi + b * si), v1 = __ldcs(xi + (b + step) * si), v2 = __ldcs(xi + (b + 2 * step) * si), v3 = __ldcs(xi + (b + 3 * step) * si);
    if (neg) { v0 = ops.neg(v0); v1 = ops.neg(v1); v2 = ops.neg(v2); v3 = ops.neg(v3); }
    __stcs(yo + b * so, v0);
    __stcs(yo + (b + step) * so, v1);
    __stcs(yo + (b + 2 * step) * so, v2);
    __stcs(yo + (b + 3 * step) * so, v3);
  }
  for (; b < batch; b += step) {
    V v = __ldcs(xi + b * si);
    if (neg) v = ops.neg(v);
    __stcs(yo + b * so, v);
  }
}

// y[b][i][t] <- sum_r tweak[r][i][t] * x[b][idx[r][i]][t]  over Z_q_t.
// A thread owns output position (i, t) and takes NB batch elements at a time, so one index / tweak lookup and one
// address computation serve NB loads (the kernel is instruction-bound otherwise: ~40 instructions per 8-byte word with a
// Barrett reduction per product).  Products are accumulated exactly in 64 bits and reduced once per `chunk` terms,
// chunk = floor((2^64 - q) / (q-1)^2) >= 1 over the largest modulus: any order of exact reductions gives the same
// residue, so the result is the host formula's bit for bit.
template <int NB>
__device__ __forceinline__ void twace_crt_group(const long long* __restrict__ xb, long long* __restrict__ yb, const int32_t* __restrict__ idx,
                                                const uint32_t* __restrict__ tw, int32_t phi, int32_t rel, int32_t k, int32_t chunk,
                                                int64_t so, int64_t si, int32_t i, int32_t w, const ZqRing& R)
{
  uint64_t acc[NB];
#pragma unroll
  for (int a = 0; a < NB; a++) acc[a] = 0;
  for (int32_t r0 = 0; r0 < rel; r0 += chunk) {
    const int32_t r1 = min(rel, r0 + chunk);
#pragma unroll 2
    for (int32_t r = r0; r < r1; r++) {
      const int64_t off = (int64_t)__ldg(idx + (int64_t)r * phi + i) * k;
      const uint32_t tv = __ldg(tw + (int64_t)r * so + w);
      long long xv[NB];
#pragma unroll
      for (int a = 0; a < NB; a++) xv[a] = __ldg(xb + a * si + off);
#pragma unroll
      for (int a = 0; a < NB; a++) acc[a] += (uint64_t)R.load(xv[a]) * tv;
    }
#pragma unroll
    for (int a = 0; a < NB; a++) acc[a] = R.reduce64(acc[a]);
  }
#pragma unroll
  for (int a = 0; a < NB; a++) __stcs(yb + a * so, (long long)acc[a]);
}

__global__ void __launch_bounds__(256)
k_ext_twace_crt_zq(const long long* __restrict__ x, long long* __restrict__ y, const int32_t* __restrict__ idx,
                   const uint32_t* __restrict__ tw, int32_t phi, int32_t phi2, int32_t rel, int32_t k, int32_t chunk, int64_t batch,
                   const __grid_constant__ ZqConsts Z)
{
  constexpr int NB = 4;
  const int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= (int64_t)phi * k) return;
  const int32_t i = w / k, t = w - i * k;
  const ZqRing R = ZqRing::make(Z, t);
  const int64_t so = (int64_t)phi * k, si = (int64_t)phi2 * k;
  for (int64_t b = (int64_t)blockIdx.y * NB; b < batch; b += (int64_t)gridDim.y * NB) {
    const long long* xb = x + b * si + t;
    long long* yb = y + b * so + w;
    if (b + NB <= batch) {
      twace_crt_group<NB>(xb, yb, idx, tw, phi, rel, k, chunk, so, si, i, w, R);
    } else {
      for (int64_t a = 0; b + a < batch; a++) twace_crt_group<1>(xb + a * si, yb + a * so, idx, tw, phi, rel, k, chunk, so, si, i, w, R);
    }
  }
}

// the same over the complex numbers; sum in the order r = 0, 1, ... (foldl1' (+), Extension.hs:129)
__global__ void __launch_bounds__(256)
k_ext_twace_crt_c(const double2* __restrict__ x, double2* __restrict__ y, const int32_t* __restrict__ idx,
                  const double2* __restrict__ tw, int32_t phi, int32_t phi2, int32_t rel, int32_t k, int64_t batch)
{
  const int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= (int64_t)phi * k) return;
  const int32_t i = w / k, t = w - i * k;
  const C64Ring R;
  const int64_t so = (int64_t)phi * k, si = (int64_t)phi2 * k;
  for (int64_t b = blockIdx.y; b < batch; b += gridDim.y) {
    const double2* xb = x + b * si + t;
    double2 acc = make_double2(0.0, 0.0);
    for (int32_t r = 0; r < rel; r++) {
      const int32_t s = __ldg(idx + (int64_t)r * phi + i);
      const double2 p = R.mul(__ldg(tw + (int64_t)r * so + w), __ldg(xb + (int64_t)s * k));
      acc = r == 0 ? p : R.add(acc, p);
    }
    __stcs(y + b * so + w, acc);
  }
}

// tweak[r][i][t] = ratio_t * gInvCRT_m[i][t] * gCRT_m'[idx[r][i]][t]   (embedCRT(gInv)[i'] = gInv[baseIndicesCRT[i']] and
// baseIndicesCRT[extIndicesCRT[i rel + r]] == i)
__global__ void k_ext_tweak_zq(uint32_t* __restrict__ tw, const int32_t* __restrict__ idx, const long long* __restrict__ ginv_lo,
                               const long long* __restrict__ g_hi, int32_t phi, int32_t rel, int32_t k,
                               const __grid_constant__ ZqConsts Z /* scale = m'hat^-1 mhat */)
{
  const int64_t total = (int64_t)rel * phi * k;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int32_t t = (int32_t)(e % k);
    const int64_t ri = e / k;
    const int32_t i = (int32_t)(ri % phi);
    const ZqRing R = ZqRing::make(Z, t);
    const uint32_t a = R.load(ginv_lo[(int64_t)i * k + t]);
    const uint32_t g = R.load(g_hi[(int64_t)idx[ri] * k + t]);
    tw[e] = R.mul(R.mul(a, g), Z.scale[t]);
  }
}

// complex: tweak = (mhat / m'hat) * gCRT_m'[src] / gCRT_m[i]
__global__ void k_ext_tweak_c(double2* __restrict__ tw, const int32_t* __restrict__ idx, const double2* __restrict__ g_lo,
                              const double2* __restrict__ g_hi, int32_t phi, int32_t rel, int32_t k, double ratio)
{
  const int64_t total = (int64_t)rel * phi * k;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int32_t t = (int32_t)(e % k);
    const int64_t ri = e / k;
    const int32_t i = (int32_t)(ri % phi);
    const double2 d = g_lo[(int64_t)i * k + t], n = g_hi[(int64_t)idx[ri] * k + t];
    const double den = d.x * d.x + d.y * d.y;
    tw[e] = make_double2(ratio * (n.x * d.x + n.y * d.y) / den, ratio * (n.y * d.x - n.x * d.y) / den);
  }
}

__global__ void k_ext_unit(double2* __restrict__ y, int32_t k)      // scalarPow 1: coefficient 0 of every limb
{
  if ((int)threadIdx.x < k) y[threadIdx.x] = make_double2(1.0, 0.0);
}

// ------------------------------------------------------------------ host side: index tables (Tensor.hs:391-498)
struct MergedPP { int p, e, e2; int32_t phi, phi2; };

int32_t tot_pp(int p, int e) { return e == 0 ? 1 : (int32_t)((p - 1) * ipow64(p, e - 1)); }

// fromIndexPair (Tensor.hs:401-406), iteratively: digit l of the result is i0r_l + i1r_l * phi_l in radix phi'_l
int32_t from_index_pair(const std::vector<MergedPP>& mp, int64_t i1, int64_t i0)
{
  int64_t out = 0, stride = 1;
  for (const MergedPP& f : mp) {
    const int64_t relf = f.phi2 / f.phi;
    out += ((i0 % f.phi) + (i1 % relf) * f.phi) * stride;
    i0 /= f.phi; i1 /= relf; stride *= f.phi2;
  }
  return (int32_t)out;
}

// toIndexPair (Tensor.hs:393-399)
void to_index_pair(const std::vector<MergedPP>& mp, int64_t j, int64_t* i1, int64_t* i0)
{
  int64_t a1 = 0, a0 = 0, s1 = 1, s0 = 1;
  for (const MergedPP& f : mp) {
    const int64_t d = j % f.phi2;
    j /= f.phi2;
    a1 += (d / f.phi) * s1; a0 += (d % f.phi) * s0;
    s1 *= f.phi2 / f.phi; s0 *= f.phi;
  }
  *i1 = a1; *i0 = a0;
}

// baseIndexDec (Tensor.hs:483-498): -1 = Nothing, else 2 index + negate
int32_t base_index_dec(const std::vector<MergedPP>& mp, int64_t j)
{
  int64_t idx = 0, s0 = 1;
  int neg = 0;
  for (const MergedPP& f : mp) {
    const int64_t d = j % f.phi2;
    j /= f.phi2;
    int64_t cur;
    if (f.p > 2 && f.e == 0 && f.e2 > 0) {
      if (d == 0) cur = 0;
      else if (d == 1) { cur = 0; neg ^= 1; }
      else return -1;
    } else {
      if (d >= f.phi) return -1;
      cur = d;
    }
    idx += cur * s0;
    s0 *= f.phi;
  }
  return (int32_t)(idx * 2 + neg);
}

int upload_i32(int32_t** dst, const std::vector<int32_t>& v)
{
  LOLB_CUDA(cudaMalloc((void**)dst, v.size() * sizeof(int32_t)));
  LOLB_CUDA(cudaMemcpy(*dst, v.data(), v.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  return LOLB_OK;
}

// every table of Tensor.hs:429-478 for one merged prime-power list (host only, no device involved)
void compute_tables(const std::vector<MergedPP>& mp, int32_t phi, int32_t phi2, std::vector<int32_t> (&tab)[LOLB_EXT_TABLES])
{
  const int32_t rel = phi2 / phi;
  std::vector<int32_t>& powdec = tab[LOLB_EXT_INDICES_POWDEC];
  std::vector<int32_t>& crt = tab[LOLB_EXT_INDICES_CRT];
  std::vector<int32_t>& j0 = tab[LOLB_EXT_BASE_POW_J0];
  std::vector<int32_t>& j1 = tab[LOLB_EXT_BASE_POW_J1];
  std::vector<int32_t>& dec = tab[LOLB_EXT_BASE_DEC];
  std::vector<int32_t>& cof = tab[LOLB_EXT_INDICES_COEFFS];
  powdec.resize(phi); crt.resize(phi2); j0.resize(phi2); j1.resize(phi2); dec.resize(phi2); cof.resize(phi2);
  for (int32_t i = 0; i < phi; i++) powdec[i] = from_index_pair(mp, 0, i);
  for (int32_t j = 0; j < phi2; j++) {
    crt[j] = from_index_pair(mp, j % rel, j / rel);          // swap . (`divMod` rel)
    cof[j] = from_index_pair(mp, j / phi, j % phi);          // vector i1 = j / phi, entry i0 = j % phi
    int64_t a1, a0;
    to_index_pair(mp, j, &a1, &a0);
    j0[j] = (int32_t)a1; j1[j] = (int32_t)a0;
    dec[j] = base_index_dec(mp, j);
  }
}

// mergePPs (Tensor.hs:502-507): every prime power of m' with the exponent of the same prime in m (0 if absent)
int merge_pps(const PrimeExponent* lo, int nlo, const PrimeExponent* hi, int nhi, std::vector<MergedPP>* mp, int64_t* phi, int64_t* phi2)
{
  int used = 0;
  *phi = *phi2 = 1;
  mp->clear();
  for (int a = 0; a < nhi; a++) {
    int e = 0;
    if (hi[a].prime < 2 || hi[a].exponent < 1 || (a > 0 && hi[a].prime <= hi[a - 1].prime)) return LOLB_ERR_ARG;
    for (int b = 0; b < nlo; b++) if (lo[b].prime == hi[a].prime) { e = lo[b].exponent; used++; }
    if (e < 0 || e > hi[a].exponent) return LOLB_ERR_ARG;
    mp->push_back(MergedPP{hi[a].prime, e, hi[a].exponent, tot_pp(hi[a].prime, e), tot_pp(hi[a].prime, hi[a].exponent)});
    *phi *= mp->back().phi; *phi2 *= mp->back().phi2;
    if (*phi2 > ((int64_t)1 << 30)) return LOLB_ERR_ARG;
  }
  return used == nlo ? LOLB_OK : LOLB_ERR_ARG;
}

int build_tables(lolb_ext* x, const std::vector<MergedPP>& mp)
{
  const int32_t phi = x->phi, phi2 = x->phi2, rel = x->rel;
  compute_tables(mp, phi, phi2, x->h_tab);
  const std::vector<int32_t>& powdec = x->h_tab[LOLB_EXT_INDICES_POWDEC];
  const std::vector<int32_t>& crt = x->h_tab[LOLB_EXT_INDICES_CRT];
  const std::vector<int32_t>& j0 = x->h_tab[LOLB_EXT_BASE_POW_J0];
  const std::vector<int32_t>& j1 = x->h_tab[LOLB_EXT_BASE_POW_J1];
  const std::vector<int32_t>& dec = x->h_tab[LOLB_EXT_BASE_DEC];
  const std::vector<int32_t>& cof = x->h_tab[LOLB_EXT_INDICES_COEFFS];
  std::vector<int32_t> code(phi);
  for (int32_t i = 0; i < phi; i++) code[i] = powdec[i] * 2;
  int rc = upload_i32(&x->d_code[CODE_TWACE], code);
  code.resize(phi2);
  for (int32_t j = 0; j < phi2 && !rc; j++) code[j] = j0[j] == 0 ? j1[j] * 2 : -1;
  if (!rc) rc = upload_i32(&x->d_code[CODE_EMBED_POW], code);
  if (!rc) rc = upload_i32(&x->d_code[CODE_EMBED_DEC], dec);
  for (int32_t j = 0; j < phi2; j++) code[j] = j1[j] * 2;
  if (!rc) rc = upload_i32(&x->d_code[CODE_EMBED_CRT], code);
  for (int32_t j = 0; j < phi2; j++) code[j] = cof[j] * 2;
  if (!rc) rc = upload_i32(&x->d_code[CODE_COEFFS], code);
  for (int32_t j = 0; j < phi2; j++) code[(int64_t)(j % rel) * phi + j / rel] = crt[j];      // [rel][phi]
  if (!rc) rc = upload_i32(&x->d_crt_idx, code);
  code.assign(j0.begin(), j0.end());
  code.insert(code.end(), j1.begin(), j1.end());
  if (!rc) rc = upload_i32(&x->d_base_j, code);
  return rc;
}

int build_tweak_zq(lolb_ext* x)
{
  const lolb_plan *lo = x->lo, *hi = x->hi;
  if (!lo->d_gcrtinv || !hi->d_gcrt || (int)hi->mhatinv.size() != x->k) return LOLB_OK;   // no CRT: twaceCRT reports it
  ZqConsts Z = hi->zq_plain;
  const int64_t mhat = (lo->m % 2 == 0) ? lo->m / 2 : lo->m;                               // valueHat, FactoredDefs.hs:444-445
  for (int t = 0; t < x->k; t++) {
    const uint64_t q = (uint64_t)hi->qs[t];
    Z.scale[t] = (uint32_t)((u128)((uint64_t)hi->mhatinv[t] % q) * ((uint64_t)mhat % q) % q);
  }
  const int64_t total = (int64_t)x->rel * x->phi * x->k;
  LOLB_CUDA(cudaMalloc((void**)&x->d_tweak, total * sizeof(uint32_t)));
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  k_ext_tweak_zq<<<blocks, 256>>>(x->d_tweak, x->d_crt_idx, (const long long*)lo->d_gcrtinv, (const long long*)hi->d_gcrt,
                                  x->phi, x->rel, x->k, Z);
  LOLB_CUDA(cudaGetLastError());
  count_launch();
  LOLB_CUDA(cudaDeviceSynchronize());
  return LOLB_OK;
}

// complex gCRT vectors are not stored in the plans: g = crt(mulGPow(scalarPow 1)) (Tensor.hs:319-337 is its closed form)
int complex_g(const lolb_plan* pl, double2** out)
{
  const size_t bytes = (size_t)pl->n * pl->k * sizeof(double2);
  LOLB_CUDA(cudaMalloc((void**)out, bytes));
  LOLB_CUDA(cudaMemset(*out, 0, bytes));
  k_ext_unit<<<1, 32>>>(*out, pl->k);
  LOLB_CUDA(cudaGetLastError());
  count_launch();
  int rc = lolb_tensorGPowC(pl, (lolb_complex*)*out, 1, nullptr);
  if (!rc) rc = lolb_tensorCRTC(pl, (lolb_complex*)*out, 1, nullptr);
  return rc;
}

int build_tweak_c(lolb_ext* x)
{
  if (x->k > 32) return LOLB_OK;
  double2 *g_lo = nullptr, *g_hi = nullptr;
  int rc = complex_g(x->lo, &g_lo);
  if (!rc) rc = complex_g(x->hi, &g_hi);
  if (!rc) {
    const int64_t total = (int64_t)x->rel * x->phi * x->k;
    cudaError_t e = cudaMalloc((void**)&x->d_ctweak, total * sizeof(double2));
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc tweak");
    if (!rc) {
      const int64_t mhat = (x->lo->m % 2 == 0) ? x->lo->m / 2 : x->lo->m;
      const int64_t mhat2 = (x->hi->m % 2 == 0) ? x->hi->m / 2 : x->hi->m;
      const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
      k_ext_tweak_c<<<blocks, 256>>>(x->d_ctweak, x->d_crt_idx, g_lo, g_hi, x->phi, x->rel, x->k, (double)mhat / (double)mhat2);
      e = cudaGetLastError();
      if (e == cudaSuccess) { count_launch(); e = cudaDeviceSynchronize(); }
      if (e != cudaSuccess) rc = cuda_fail(e, "k_ext_tweak_c");
    }
  }
  if (g_lo) cudaFree(g_lo);
  if (g_hi) cudaFree(g_hi);
  if (rc == LOLB_ERR_NO_CRT) { if (x->d_ctweak) { cudaFree(x->d_ctweak); x->d_ctweak = nullptr; } rc = LOLB_OK; }   // gathers still work
  return rc;
}

// ------------------------------------------------------------------ launches
// block size: a multiple of 32 in [128, 256] that divides the words of one element when there is one (no ragged last block)
int pick_threads(int64_t words)
{
  if (words < 256) return (int)((words + 31) / 32) * 32;
  for (int c = 256; c >= 128; c -= 32) if (words % c == 0) return c;
  return 256;
}

dim3 batch_grid(const lolb_plan* pl, int64_t words, int threads, int64_t batch)
{
  dim3 grid((unsigned)((words + threads - 1) / threads), 1, 1);
  int64_t gy = ((int64_t)pl->num_sms * 2048 / threads + grid.x - 1) / grid.x * 2;     // ~2 waves of resident threads
  if (gy > batch) gy = batch;
  if (gy > 65535) gy = 65535;
  if (gy < 1) gy = 1;
  grid.y = (unsigned)gy;
  return grid;
}

template <class W>
int launch_gather(const lolb_ext* x, const int32_t* code, int32_t n_out, int32_t n_in, int32_t kw, const void* src, void* dst,
                  int64_t batch, cudaStream_t st)
{
  const int64_t words = (int64_t)n_out * kw;
  const int threads = pick_threads(words);
  const dim3 grid = batch_grid(x->hi, words, threads, batch);
  k_ext_gather<W><<<grid, threads, 0, st>>>((const typename W::V*)src, (typename W::V*)dst, code, n_out, n_in, kw, batch, x->hi->zq_plain);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_ext_gather");
  count_launch();
  return LOLB_OK;
}

int check_ring(const lolb_ext* x, int ring, const char* fn)
{
  if (!x) { set_error(std::string(fn) + ": NULL extension"); return LOLB_ERR_ARG; }
  const bool ok = x->kind == PLAN_RQ ? ring == LOLB_RING_RQ : (ring == LOLB_RING_R || ring == LOLB_RING_DOUBLE || ring == LOLB_RING_C);
  if (!ok) { set_error(std::string(fn) + ": ring does not match the plans this extension was created from"); return LOLB_ERR_ARG; }
  return LOLB_OK;
}

int gather(const lolb_ext* x, int ring, int which, bool up, const void* src, void* dst, int64_t batch, void* stream, const char* fn)
{
  int rc = check_ring(x, ring, fn);
  if (rc) return rc;
  if (batch < 0 || (batch > 0 && (!src || !dst))) { set_error(std::string(fn) + ": bad batch or NULL operand"); return LOLB_ERR_ARG; }
  if (src == dst && batch > 0) { set_error(std::string(fn) + ": operands must not alias"); return LOLB_ERR_ARG; }
  if (batch == 0) return LOLB_OK;
  const int32_t n_out = (up || which == CODE_COEFFS) ? x->phi2 : x->phi;
  const int32_t n_in = (up && which != CODE_COEFFS) ? x->phi : x->phi2;
  const int32_t* code = x->d_code[which];
  cudaStream_t st = (cudaStream_t)stream;
  const int k = x->k;
  const bool al16 = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
  switch (ring) {
    case LOLB_RING_RQ:
      if (k % 2 == 0 && al16) return launch_gather<WordZq2>(x, code, n_out, n_in, k / 2, src, dst, batch, st);
      return launch_gather<WordZq1>(x, code, n_out, n_in, k, src, dst, batch, st);
    case LOLB_RING_R:
      if (k % 2 == 0 && al16) return launch_gather<WordI64x2>(x, code, n_out, n_in, k / 2, src, dst, batch, st);
      return launch_gather<WordI64>(x, code, n_out, n_in, k, src, dst, batch, st);
    case LOLB_RING_DOUBLE:
      return launch_gather<WordF64>(x, code, n_out, n_in, k, src, dst, batch, st);
    default:
      return launch_gather<WordC64>(x, code, n_out, n_in, k, src, dst, batch, st);
  }
}

}  // namespace

// ------------------------------------------------------------------ C ABI
extern "C" int lolb_ext_create(lolb_ext** out, const lolb_plan* lo, const lolb_plan* hi)
{
  if (!out || !lo || !hi) { set_error("lolb_ext_create: NULL argument"); return LOLB_ERR_ARG; }
  *out = nullptr;
  if (lo->kind != hi->kind || lo->k != hi->k || (lo->kind == PLAN_RQ && lo->qs != hi->qs)) {
    set_error("lolb_ext_create: the two plans must be over the same ring (kind, tupSize, moduli)");
    return LOLB_ERR_ARG;
  }
  std::vector<MergedPP> mp;
  int64_t phi = 1, phi2 = 1;
  if (merge_pps(lo->pe.data(), (int)lo->pe.size(), hi->pe.data(), (int)hi->pe.size(), &mp, &phi, &phi2) || phi != lo->n || phi2 != hi->n) {
    set_error("lolb_ext_create: m does not divide m'");
    return LOLB_ERR_ARG;
  }
  lolb_ext* x = new lolb_ext();
  x->lo = lo; x->hi = hi; x->kind = lo->kind; x->k = lo->k;
  x->phi = lo->n; x->phi2 = hi->n; x->rel = hi->n / lo->n;
  cudaError_t e = cudaSetDevice(hi->device);
  int rc = e == cudaSuccess ? LOLB_OK : cuda_fail(e, "cudaSetDevice");
  if (!rc) rc = build_tables(x, mp);
  if (!rc) rc = x->kind == PLAN_RQ ? build_tweak_zq(x) : build_tweak_c(x);
  if (rc) { lolb_ext_destroy(x); return rc; }
  *out = x;
  return LOLB_OK;
}

extern "C" void lolb_ext_destroy(lolb_ext* x)
{
  if (!x) return;
  for (int32_t* p : x->d_code) if (p) cudaFree(p);
  if (x->d_crt_idx) cudaFree(x->d_crt_idx);
  if (x->d_base_j) cudaFree(x->d_base_j);
  if (x->d_tweak) cudaFree(x->d_tweak);
  if (x->d_ctweak) cudaFree(x->d_ctweak);
  delete x;
}

extern "C" int32_t lolb_ext_totient(const lolb_ext* x, int upper) { return !x ? 0 : upper ? x->phi2 : x->phi; }

extern "C" int64_t lolb_ext_index_table(const PrimeExponent* pe, hShort_t nPE, const PrimeExponent* pe2, hShort_t nPE2, int which, int32_t* out)
{
  std::vector<MergedPP> mp;
  int64_t phi = 1, phi2 = 1;
  if (nPE < 0 || nPE2 < 0 || (nPE > 0 && !pe) || (nPE2 > 0 && !pe2) || which < 0 || which >= LOLB_EXT_TABLES ||
      merge_pps(pe, nPE, pe2, nPE2, &mp, &phi, &phi2)) {
    set_error("lolb_ext_index_table: bad argument or m does not divide m'");
    return -1;
  }
  const int64_t count = which == LOLB_EXT_INDICES_POWDEC ? phi : phi2;
  if (!out) return count;
  std::vector<int32_t> tab[LOLB_EXT_TABLES];
  compute_tables(mp, (int32_t)phi, (int32_t)phi2, tab);
  memcpy(out, tab[which].data(), (size_t)count * sizeof(int32_t));
  return count;
}

extern "C" int lolb_ext_get_table(const lolb_ext* x, int which, int32_t* out)
{
  if (!x || !out || which < 0 || which >= LOLB_EXT_TABLES) { set_error("lolb_ext_get_table: bad argument"); return LOLB_ERR_ARG; }
  memcpy(out, x->h_tab[which].data(), x->h_tab[which].size() * sizeof(int32_t));
  return LOLB_OK;
}

// powBasisPow' (Extension.hs:133-143): the phi'/phi vectors of O_m' that form an O_m-basis of O_m', in the powerful basis:
// vector r has `one` where baseIndicesPow = (r, 0) and zero elsewhere.  One thread per 8-byte (or 16-byte complex) word.
template <class T>
__global__ void k_ext_pow_basis(T* __restrict__ y, const int32_t* __restrict__ base_j, int32_t phi2, int32_t k, int64_t total, T one, T zero)
{
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t w = idx / k;
    const int32_t j = (int32_t)(w % phi2), r = (int32_t)(w / phi2);
    y[idx] = (base_j[j] == r && base_j[phi2 + j] == 0) ? one : zero;
  }
}

extern "C" int lolb_powBasisPow(const lolb_ext* x, int ring, void* y, void* stream)
{
  int rc = check_ring(x, ring, __func__);
  if (rc) return rc;
  if (!y) { set_error("lolb_powBasisPow: NULL output"); return LOLB_ERR_ARG; }
  const int64_t total = (int64_t)x->rel * x->phi2 * x->k;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)x->hi->num_sms * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (ring == LOLB_RING_C) k_ext_pow_basis<double2><<<blocks, 256, 0, st>>>((double2*)y, x->d_base_j, x->phi2, x->k, total, make_double2(1.0, 0.0), make_double2(0.0, 0.0));
  else if (ring == LOLB_RING_DOUBLE) k_ext_pow_basis<double><<<blocks, 256, 0, st>>>((double*)y, x->d_base_j, x->phi2, x->k, total, 1.0, 0.0);
  else k_ext_pow_basis<long long><<<blocks, 256, 0, st>>>((long long*)y, x->d_base_j, x->phi2, x->k, total, 1LL, 0LL);      // q >= 2: one = 1 in every limb
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_ext_pow_basis");
  count_launch();
  return LOLB_OK;
}

extern "C" int lolb_twacePowDec(const lolb_ext* x, int ring, const void* src, void* dst, int64_t batch, void* stream)
{ return gather(x, ring, CODE_TWACE, false, src, dst, batch, stream, __func__); }

extern "C" int lolb_embedPow(const lolb_ext* x, int ring, const void* src, void* dst, int64_t batch, void* stream)
{ return gather(x, ring, CODE_EMBED_POW, true, src, dst, batch, stream, __func__); }

extern "C" int lolb_embedDec(const lolb_ext* x, int ring, const void* src, void* dst, int64_t batch, void* stream)
{ return gather(x, ring, CODE_EMBED_DEC, true, src, dst, batch, stream, __func__); }

extern "C" int lolb_coeffsPowDec(const lolb_ext* x, int ring, const void* src, void* dst, int64_t batch, void* stream)
{ return gather(x, ring, CODE_COEFFS, false, src, dst, batch, stream, __func__); }

// embedCRT' and twaceCRT' exist only where O_m' has a CRT over the ring (`CRTrans mon r`, Extension.hs:81-85, 110-116):
// Z_q with m' | q - 1, or the complex numbers.
static int require_crt(const lolb_ext* x, int ring, const char* fn)
{
  int rc = check_ring(x, ring, fn);
  if (rc) return rc;
  if (ring == LOLB_RING_R || ring == LOLB_RING_DOUBLE) { set_error(std::string(fn) + ": no CRT basis over this ring"); return LOLB_ERR_NO_CRT; }
  if (ring == LOLB_RING_RQ && !x->hi->has_fwd) { set_error(std::string(fn) + ": no CRT over this modulus / index (ZqBasic.hs:159-165)"); return LOLB_ERR_NO_CRT; }
  return LOLB_OK;
}

extern "C" int lolb_embedCRT(const lolb_ext* x, int ring, const void* src, void* dst, int64_t batch, void* stream)
{
  int rc = require_crt(x, ring, __func__);
  return rc ? rc : gather(x, ring, CODE_EMBED_CRT, true, src, dst, batch, stream, __func__);
}

extern "C" int lolb_twaceCRT(const lolb_ext* x, int ring, const void* src, void* dst, int64_t batch, void* stream)
{
  int rc = require_crt(x, ring, __func__);
  if (rc) return rc;
  if (batch < 0 || (batch > 0 && (!src || !dst))) { set_error("lolb_twaceCRT: bad batch or NULL operand"); return LOLB_ERR_ARG; }
  if (src == dst && batch > 0) { set_error("lolb_twaceCRT: operands must not alias"); return LOLB_ERR_ARG; }
  if (batch == 0) return LOLB_OK;
  const int64_t words = (int64_t)x->phi * x->k;
  const int threads = pick_threads(words);
  cudaStream_t st = (cudaStream_t)stream;
  if (ring == LOLB_RING_RQ) {
    if (!x->d_tweak) { set_error("lolb_twaceCRT: no gCRT / mhat^-1 tables in the plans (no CRT over this modulus / index)"); return LOLB_ERR_NO_CRT; }
    // exact 64-bit accumulation: reduce once per `chunk` products (see the kernel)
    int64_t chunk = x->rel;
    for (int t = 0; t < x->k; t++) {
      const u128 q = (u128)(uint64_t)x->hi->qs[t];
      const u128 c = q <= 2 ? (u128)chunk : ((((u128)1) << 64) - q) / ((q - 1) * (q - 1));
      if (c < (u128)chunk) chunk = (int64_t)c;
    }
    if (chunk < 1) chunk = 1;
    const dim3 grid = batch_grid(x->hi, words, threads, (batch + 3) / 4);      // four batch elements per thread iteration
    k_ext_twace_crt_zq<<<grid, threads, 0, st>>>((const long long*)src, (long long*)dst, x->d_crt_idx, x->d_tweak, x->phi, x->phi2,
                                                 x->rel, x->k, (int32_t)chunk, batch, x->hi->zq_plain);
  } else {
    const dim3 grid = batch_grid(x->hi, words, threads, batch);
    if (!x->d_ctweak) { set_error("lolb_twaceCRT: no complex CRT tables in the plans"); return LOLB_ERR_NO_CRT; }
    k_ext_twace_crt_c<<<grid, threads, 0, st>>>((const double2*)src, (double2*)dst, x->d_crt_idx, x->d_ctweak, x->phi, x->phi2, x->rel,
                                                x->k, batch);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_ext_twace_crt");
  count_launch();
  return LOLB_OK;
}
