// she_stream.cu -- the coefficient-wise steps either side of the CRT in SymmSHE's ciphertext multiply and quadratic
// key switch (BASELINE.json configs[3]).  The reference runs them on the host between its
// FFI calls; here they are single streaming passes over the device-resident batch, in the same [batch][n][k] layout:
//
//   ct_mul     (c1 * c2) for two linear ciphertexts, then mulG on every coefficient of the product polynomial
//              (SymmSHE.hs:443-449: `CT d2 (k1+k2+1) (l1*l2) (mulG <$> c1 * c2)`; the Cyc product is coefficient-wise
//              in the CRT basis, UCyc.hs:232, and mulG there is the product with the gCRT vector, CPP.hs:230)
//   decompose  gadget decomposition of a Pow-basis element followed by `reduce` of every digit into the product ring
//              (SymmSHE.hs:314 `fmap reduce <$> decompose c`; Cyc.hs:603; tuples concatenate the per-limb digits,
//              Gadget.hs:97-101; TrivGad = [lift x], ZqBasic.hs:230-232; BaseBGad = decomp radices . lift,
//              ZqBasic.hs:257-264 with decomp / divModCent of Numeric.hs:202-205, 227-234; lift = decode',
//              ZqBasic.hs:92-94: the representative in [-q/2, q/2))
//   knapsack   c_j += sum_i digit_i * hint_i[j] for the two coefficients of the linear hint polynomials
//              (SymmSHE.hs:302-305 `sum $ zipWith (*>>) (adviseCRT <$> xs) hint`, added to [c0,c1] at :372)
//
// All arithmetic is exact in Z_q per limb (canonical residues in and out), so results equal the host formulas bit for
// bit.  Bytes per ring element: ct_mul 7 x 8nk (+ the cached gCRT vector), decompose (1 + l) x 8nk, knapsack
// (l + 4) x 8nk (+ the cached hints), l = number of gadget digits.
#include "fused.cuh"

namespace lolb {

namespace {

__device__ __forceinline__ uint32_t she_barrett(uint64_t x, uint32_t q, uint64_t mu)
{
  uint64_t r = x - __umul64hi(x, mu) * q;       // [0, 2q) ... [0, 3q)
  if (r >= q) r -= q;
  if (r >= q) r -= q;
  return (uint32_t)r;
}

// any int64 -> canonical residue (inputs are canonical by contract; anything else is reduced like `c % q`, types.h:62-66)
__device__ __forceinline__ uint32_t she_canon(int64_t x, uint32_t q)
{
  if ((uint64_t)x < (uint64_t)q) return (uint32_t)x;
  int64_t r = x % (int64_t)q;
  return (uint32_t)(r < 0 ? r + (int64_t)q : r);
}

__device__ __forceinline__ uint32_t she_addmod(uint32_t a, uint32_t b, uint32_t q)
{
  const uint64_t s = (uint64_t)a + b;
  return (uint32_t)(s >= q ? s - q : s);
}

struct Limb { uint32_t q; uint64_t mu; };

// one coefficient of the tensor product of two linear polynomials, times g
struct Prod3 { int64_t d0, d1, d2; };
__device__ __forceinline__ Prod3 ct_coeff(int64_t a0, int64_t a1, int64_t b0, int64_t b1, int64_t g, bool has_g, Limb L)
{
  const uint32_t x0 = she_canon(a0, L.q), x1 = she_canon(a1, L.q), y0 = she_canon(b0, L.q), y1 = she_canon(b1, L.q);
  uint32_t d0 = she_barrett((uint64_t)x0 * y0, L.q, L.mu);
  uint32_t d1 = she_addmod(she_barrett((uint64_t)x0 * y1, L.q, L.mu), she_barrett((uint64_t)x1 * y0, L.q, L.mu), L.q);
  uint32_t d2 = she_barrett((uint64_t)x1 * y1, L.q, L.mu);
  if (has_g) {
    const uint32_t gg = she_canon(g, L.q);
    d0 = she_barrett((uint64_t)d0 * gg, L.q, L.mu);
    d1 = she_barrett((uint64_t)d1 * gg, L.q, L.mu);
    d2 = she_barrett((uint64_t)d2 * gg, L.q, L.mu);
  }
  return Prod3{(int64_t)d0, (int64_t)d1, (int64_t)d2};
}

// two coefficients (16 bytes) per thread per operand; `pairs` = batch * n * k / 2, g_pairs = n * k / 2
__global__ void __launch_bounds__(256)
k_ct_mul(const longlong2* a0, const longlong2* a1, const longlong2* b0, const longlong2* b1, const longlong2* __restrict__ g,
         longlong2* d0, longlong2* d1, longlong2* d2, int64_t pairs, int64_t g_pairs, int k,      // d* may alias a* / b*
         const __grid_constant__ ZqConsts Z)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pairs; i += (int64_t)gridDim.x * blockDim.x) {
    const longlong2 x0 = __ldcs(a0 + i), x1 = __ldcs(a1 + i), y0 = __ldcs(b0 + i), y1 = __ldcs(b1 + i);
    const longlong2 gg = g ? __ldg(g + (i % g_pairs)) : make_longlong2(1, 1);
    const int l0 = (int)((2 * i) % k), l1 = (int)((2 * i + 1) % k);
    const Prod3 p = ct_coeff(x0.x, x1.x, y0.x, y1.x, gg.x, g != nullptr, Limb{Z.q[l0], Z.mu[l0]});
    const Prod3 r = ct_coeff(x0.y, x1.y, y0.y, y1.y, gg.y, g != nullptr, Limb{Z.q[l1], Z.mu[l1]});
    __stcs(d0 + i, make_longlong2(p.d0, r.d0));
    __stcs(d1 + i, make_longlong2(p.d1, r.d1));
    __stcs(d2 + i, make_longlong2(p.d2, r.d2));
  }
}

// scalar form for an odd n*k (m = 1, 2 with an odd tupSize)
__global__ void __launch_bounds__(256)
k_ct_mul_1(const int64_t* a0, const int64_t* a1, const int64_t* b0, const int64_t* b1, const int64_t* __restrict__ g,
           int64_t* d0, int64_t* d1, int64_t* d2, int64_t count, int64_t g_count, int k, const __grid_constant__ ZqConsts Z)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i % k);
    const int64_t x0 = a0[i], x1 = a1[i], y0 = b0[i], y1 = b1[i];
    const Prod3 p = ct_coeff(x0, x1, y0, y1, g ? g[i % g_count] : 1, g != nullptr, Limb{Z.q[l], Z.mu[l]});
    d0[i] = p.d0; d1[i] = p.d1; d2[i] = p.d2;
  }
}

struct GadgetGeom {
  int k, ell;
  int64_t base;                  // 0: TrivGad; b >= 2: BaseBGad b
  int digits[kMaxLimbs];         // digits of limb l (gadlen, ZqBasic.hs:241-243)
  int first[kMaxLimbs];          // index of its first digit in the concatenation (Gadget.hs:101)
};

// integer digit -> canonical residue mod q (`reduce`, fromIntegral into ZqBasic)
__device__ __forceinline__ int64_t she_reduce_digit(int64_t d, uint32_t q)
{
  int64_t r = d % (int64_t)q;
  return r < 0 ? r + (int64_t)q : r;
}

// One thread per coefficient tuple (all k limbs of one (element, j)): reads 8k contiguous bytes, writes 8k contiguous
// bytes into each of the ell digit arrays.  digits[d] is a full [tuples][k] array at offset d * tuples * k.
__global__ void __launch_bounds__(256)
k_decompose(const int64_t* __restrict__ x, int64_t* __restrict__ digits, int64_t tuples, const __grid_constant__ GadgetGeom G,
            const __grid_constant__ ZqConsts Z)
{
  const int k = G.k;
  const int64_t plane = tuples * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < tuples; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t* src = x + i * k;
    for (int l = 0; l < k; l++) {
      const uint32_t q = Z.q[l];
      const uint32_t c = she_canon(__ldcs(src + l), q);
      int64_t v = (2 * (uint64_t)c < (uint64_t)q) ? (int64_t)c : (int64_t)c - (int64_t)q;        // lift
      int64_t* dst = digits + (int64_t)G.first[l] * plane + i * k;
      const int nd = G.digits[l];
      for (int d = 0; d < nd; d++) {
        int64_t digit;
        if (d == nd - 1) {
          digit = v;                                        // decomp [] v = [v]
        } else {                                            // (quo, r) = v `divModCent` b
          const int64_t shift = G.base / 2;
          const int64_t t = v + shift;
          int64_t quo = t / G.base;
          if (t % G.base < 0) quo -= 1;                     // floor division (Haskell divMod)
          digit = t - quo * G.base - shift;
          v = quo;
        }
        if (k == 2) {
          __stcs(reinterpret_cast<longlong2*>(dst), make_longlong2(she_reduce_digit(digit, Z.q[0]), she_reduce_digit(digit, Z.q[1])));
        } else {
          for (int t2 = 0; t2 < k; t2++) __stcs(dst + t2, she_reduce_digit(digit, Z.q[t2]));
        }
        dst += plane;
      }
    }
  }
}

// c0 += sum_i digit_i * h[i][0], c1 += sum_i digit_i * h[i][1]; hints are [ell][2][n][k], one ring element each
template <typename V>
struct VecIO;
template <>
struct VecIO<longlong2> {
  static constexpr int W = 2;
  __device__ static void get(const longlong2& v, int64_t (&o)[2]) { o[0] = v.x; o[1] = v.y; }
  __device__ static longlong2 put(const int64_t (&o)[2]) { return make_longlong2(o[0], o[1]); }
};
template <>
struct VecIO<long long> {
  static constexpr int W = 1;
  __device__ static void get(const long long& v, int64_t (&o)[1]) { o[0] = v; }
  __device__ static long long put(const int64_t (&o)[1]) { return o[0]; }
};

template <typename V>
__global__ void __launch_bounds__(256)
k_knapsack(const V* __restrict__ digits, int ell, const V* __restrict__ hints, V* __restrict__ c0, V* __restrict__ c1,
           int64_t units, int64_t elem_units, int k, const __grid_constant__ ZqConsts Z)
{
  constexpr int W = VecIO<V>::W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < units; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i % elem_units;
    int64_t a0[W], a1[W];
    VecIO<V>::get(__ldcs(c0 + i), a0);
    VecIO<V>::get(__ldcs(c1 + i), a1);
    uint32_t s0[W], s1[W], q[W];
    uint64_t mu[W];
#pragma unroll
    for (int w = 0; w < W; w++) {
      const int l = (int)((W * i + w) % k);
      q[w] = Z.q[l]; mu[w] = Z.mu[l];
      s0[w] = she_canon(a0[w], q[w]);
      s1[w] = she_canon(a1[w], q[w]);
    }
    for (int d = 0; d < ell; d++) {
      int64_t x[W], h0[W], h1[W];
      VecIO<V>::get(__ldcs(digits + (int64_t)d * units + i), x);
      VecIO<V>::get(__ldg(hints + ((int64_t)d * 2 + 0) * elem_units + j), h0);
      VecIO<V>::get(__ldg(hints + ((int64_t)d * 2 + 1) * elem_units + j), h1);
#pragma unroll
      for (int w = 0; w < W; w++) {
        const uint32_t xv = she_canon(x[w], q[w]);
        s0[w] = she_addmod(s0[w], she_barrett((uint64_t)xv * she_canon(h0[w], q[w]), q[w], mu[w]), q[w]);
        s1[w] = she_addmod(s1[w], she_barrett((uint64_t)xv * she_canon(h1[w], q[w]), q[w], mu[w]), q[w]);
      }
    }
#pragma unroll
    for (int w = 0; w < W; w++) { a0[w] = (int64_t)s0[w]; a1[w] = (int64_t)s1[w]; }
    __stcs(c0 + i, VecIO<V>::put(a0));
    __stcs(c1 + i, VecIO<V>::put(a1));
  }
}

int grid_for(const lolb_plan* pl, int64_t items)
{
  int64_t blocks = (items + 255) / 256;
  const int64_t cap = (int64_t)pl->num_sms * 32;
  if (blocks > cap) blocks = cap;
  return (int)(blocks < 1 ? 1 : blocks);
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

int gadget_geom(const lolb_plan* pl, int64_t base, GadgetGeom* G)
{
  if (base != 0 && base < 2) { set_error("gadget base must be 0 (TrivGad) or >= 2 (BaseBGad)"); return LOLB_ERR_ARG; }
  G->k = pl->k; G->base = base; G->ell = 0;
  for (int l = 0; l < pl->k; l++) {
    int nd = 1;
    if (base) {                                   // gadlen b q = 1 + gadlen b (q `div` b), gadlen _ 0 = 0
      nd = 0;
      for (int64_t q = pl->qs[l]; q != 0; q /= base) nd++;
    }
    G->digits[l] = nd; G->first[l] = G->ell; G->ell += nd;
  }
  return LOLB_OK;
}

}  // namespace

int she_gadget_length(const lolb_plan* pl, int64_t base)
{
  GadgetGeom G;
  if (gadget_geom(pl, base, &G)) return -1;
  return G.ell;
}

int she_ct_mul(const lolb_plan* pl, const int64_t* a0, const int64_t* a1, const int64_t* b0, const int64_t* b1,
               const int64_t* g, int64_t* d0, int64_t* d1, int64_t* d2, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  const int64_t nk = (int64_t)pl->n * pl->k, count = batch * nk;
  const bool vec = !(nk & 1) && aligned16(a0) && aligned16(a1) && aligned16(b0) && aligned16(b1) && aligned16(d0) && aligned16(d1) &&
                   aligned16(d2) && (!g || aligned16(g));
  if (vec)
    k_ct_mul<<<grid_for(pl, count / 2), 256, 0, st>>>((const longlong2*)a0, (const longlong2*)a1, (const longlong2*)b0, (const longlong2*)b1,
                                                     (const longlong2*)g, (longlong2*)d0, (longlong2*)d1, (longlong2*)d2, count / 2, nk / 2,
                                                     pl->k, pl->zq_plain);
  else
    k_ct_mul_1<<<grid_for(pl, count), 256, 0, st>>>(a0, a1, b0, b1, g, d0, d1, d2, count, nk, pl->k, pl->zq_plain);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_ct_mul");
  count_launch();
  return LOLB_OK;
}

int she_decompose(const lolb_plan* pl, const int64_t* x, int64_t* digits, int64_t batch, int64_t base, cudaStream_t st)
{
  GadgetGeom G;
  int rc = gadget_geom(pl, base, &G);
  if (rc) return rc;
  if (batch <= 0) return LOLB_OK;
  if (pl->k == 2 && !aligned16(digits)) { set_error("decompose: digits must be 16-byte aligned"); return LOLB_ERR_ARG; }
  const int64_t tuples = batch * pl->n;
  k_decompose<<<grid_for(pl, tuples), 256, 0, st>>>(x, digits, tuples, G, pl->zq_plain);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_decompose");
  count_launch();
  return LOLB_OK;
}

int she_knapsack(const lolb_plan* pl, const int64_t* digits, int ell, const int64_t* hints, int64_t* c0, int64_t* c1,
                 int64_t batch, cudaStream_t st)
{
  if (ell < 0) { set_error("knapsack: negative digit count"); return LOLB_ERR_ARG; }
  if (batch <= 0 || ell == 0) return LOLB_OK;
  const int64_t nk = (int64_t)pl->n * pl->k, count = batch * nk;
  const bool vec = !(nk & 1) && aligned16(digits) && aligned16(hints) && aligned16(c0) && aligned16(c1);
  if (vec)
    k_knapsack<longlong2><<<grid_for(pl, count / 2), 256, 0, st>>>((const longlong2*)digits, ell, (const longlong2*)hints, (longlong2*)c0,
                                                                  (longlong2*)c1, count / 2, nk / 2, pl->k, pl->zq_plain);
  else
    k_knapsack<long long><<<grid_for(pl, count), 256, 0, st>>>((const long long*)digits, ell, (const long long*)hints, (long long*)c0,
                                                              (long long*)c1, count, nk, pl->k, pl->zq_plain);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_knapsack");
  count_launch();
  return LOLB_OK;
}

}  // namespace lolb
