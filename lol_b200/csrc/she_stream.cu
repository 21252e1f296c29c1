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

// Per-limb constants of a thread.  A thread keeps its position inside the ring element and walks the batch, so the
// limb, its modulus and every broadcast operand (gCRT, hints) are loop invariants held in registers.
// MONT (odd q < 2^31): 32-bit Montgomery reduction, REDC(T) = T 2^-32 mod q in [0, 2q) for T < q 2^32; a product of two
// data words is repaired by a second REDC against a broadcast operand kept as w 2^64 mod q.  Otherwise 64-bit Barrett.
struct Limb {
  uint32_t q, qinv, r2;      // qinv = -q^-1 mod 2^32, r2 = 2^64 mod q (MONT only)
  uint64_t mu;
};

template <bool MONT>
__device__ __forceinline__ Limb make_limb(const ZqConsts& Z, int l)
{
  Limb L;
  L.q = Z.q[l]; L.mu = Z.mu[l]; L.qinv = 0; L.r2 = 0;
  if (MONT) {
    uint32_t inv = L.q;                                  // Newton: q * inv == 1 mod 2^3, doubling each step
#pragma unroll
    for (int i = 0; i < 5; i++) inv *= 2u - L.q * inv;
    L.qinv = 0u - inv;
    const uint32_t r1 = she_barrett(1ull << 32, L.q, L.mu);
    L.r2 = she_barrett((uint64_t)r1 * r1, L.q, L.mu);
  }
  return L;
}

__device__ __forceinline__ uint32_t redc(uint64_t t, const Limb& L)
{
  const uint32_t m = (uint32_t)t * L.qinv;
  return (uint32_t)((t + (uint64_t)m * L.q) >> 32);      // [0, 2q)
}

__device__ __forceinline__ uint32_t csub(uint32_t x, uint32_t q) { return x >= q ? x - q : x; }

// broadcast operand w -> the form the second reduction wants: w 2^64 mod q (MONT) or w itself
template <bool MONT>
__device__ __forceinline__ uint32_t prep_scale(int64_t w, const Limb& L)
{
  const uint32_t c = she_canon(w, L.q);
  return MONT ? she_barrett((uint64_t)c * L.r2, L.q, L.mu) : c;
}

// one coefficient of the product of two linear polynomials, times the prepared g: (a0 b0, a0 b1 + a1 b0, a1 b1) g
struct Prod3 { int64_t d0, d1, d2; };
template <bool MONT>
__device__ __forceinline__ Prod3 ct_coeff(int64_t a0, int64_t a1, int64_t b0, int64_t b1, uint32_t gs, const Limb& L)
{
  const uint32_t x0 = she_canon(a0, L.q), x1 = she_canon(a1, L.q), y0 = she_canon(b0, L.q), y1 = she_canon(b1, L.q);
  uint32_t d0, d1, d2;
  if (MONT) {                                            // q < 2^31: x0 y1 + x1 y0 < 2 q^2 < q 2^32
    d0 = csub(redc((uint64_t)redc((uint64_t)x0 * y0, L) * gs, L), L.q);
    d1 = csub(redc((uint64_t)redc((uint64_t)x0 * y1 + (uint64_t)x1 * y0, L) * gs, L), L.q);
    d2 = csub(redc((uint64_t)redc((uint64_t)x1 * y1, L) * gs, L), L.q);
  } else {
    d0 = she_barrett((uint64_t)she_barrett((uint64_t)x0 * y0, L.q, L.mu) * gs, L.q, L.mu);
    d1 = she_barrett((uint64_t)she_addmod(she_barrett((uint64_t)x0 * y1, L.q, L.mu), she_barrett((uint64_t)x1 * y0, L.q, L.mu), L.q) * gs,
                     L.q, L.mu);
    d2 = she_barrett((uint64_t)she_barrett((uint64_t)x1 * y1, L.q, L.mu) * gs, L.q, L.mu);
  }
  return Prod3{(int64_t)d0, (int64_t)d1, (int64_t)d2};
}

// W coefficients (8 W bytes) per thread per operand: V = longlong2 (W = 2) or long long when n k is odd
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

// thread <-> unit `u` of the ring element (W coefficients), blockIdx.y strides over the batch; d* may alias a* / b*
template <typename V, bool MONT>
__global__ void __launch_bounds__(256)
k_ct_mul(const V* a0, const V* a1, const V* b0, const V* b1, const V* __restrict__ g, V* d0, V* d1, V* d2, int64_t batch,
         int elem_units, int k, const __grid_constant__ ZqConsts Z)
{
  constexpr int W = VecIO<V>::W;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= elem_units) return;
  Limb L[W];
  uint32_t gs[W];
  int64_t gv[W];
  if (g) VecIO<V>::get(__ldg(g + u), gv);
#pragma unroll
  for (int w = 0; w < W; w++) {
    L[w] = make_limb<MONT>(Z, (W * u + w) % k);
    gs[w] = prep_scale<MONT>(g ? gv[w] : 1, L[w]);
  }
  for (int64_t b = blockIdx.y; b < batch; b += gridDim.y) {
    const int64_t i = b * elem_units + u;
    int64_t x0[W], x1[W], y0[W], y1[W], o0[W], o1[W], o2[W];
    VecIO<V>::get(__ldcs(a0 + i), x0);
    VecIO<V>::get(__ldcs(a1 + i), x1);
    VecIO<V>::get(__ldcs(b0 + i), y0);
    VecIO<V>::get(__ldcs(b1 + i), y1);
#pragma unroll
    for (int w = 0; w < W; w++) {
      const Prod3 p = ct_coeff<MONT>(x0[w], x1[w], y0[w], y1[w], gs[w], L[w]);
      o0[w] = p.d0; o1[w] = p.d1; o2[w] = p.d2;
    }
    __stcs(d0 + i, VecIO<V>::put(o0));
    __stcs(d1 + i, VecIO<V>::put(o1));
    __stcs(d2 + i, VecIO<V>::put(o2));
  }
}

struct GadgetGeom {
  int k, ell;
  int64_t base;                  // 0: TrivGad; b >= 2: BaseBGad b
  int shift;                     // log2(base) when base is a power of two, else -1
  int digits[kMaxLimbs];         // digits of limb l (gadlen, ZqBasic.hs:241-243)
  int first[kMaxLimbs];          // index of its first digit in the concatenation (Gadget.hs:101)
};

// integer digit -> canonical residue mod q (`reduce`, fromIntegral into ZqBasic)
__device__ __forceinline__ int64_t she_reduce_digit(int64_t d, uint32_t q, uint64_t mu)
{
  const uint64_t a = d < 0 ? (uint64_t)(-d) : (uint64_t)d;
  const uint32_t r = a < q ? (uint32_t)a : she_barrett(a, q, mu);
  return (int64_t)((d < 0 && r) ? q - r : r);
}

// (quo, r) = v `divModCent` b  (Numeric.hs:227-234): floor division of v + b/2, remainder moved to [-b/2, b/2)
__device__ __forceinline__ int64_t div_mod_cent(int64_t& v, const GadgetGeom& G)
{
  const int64_t half = G.base / 2, t = v + half;
  int64_t quo;
  if (G.shift >= 0) quo = t >> G.shift;                   // arithmetic shift floors
  else { quo = t / G.base; if (t % G.base < 0) quo -= 1; }
  const int64_t r = t - quo * G.base - half;
  v = quo;
  return r;
}

// thread <-> coefficient tuple j of the ring element (all k limbs, 8k contiguous bytes), blockIdx.y strides over the
// batch.  digits[d] is a full [batch][n][k] array at offset d * batch * n * k.
template <int K>      // K = 2: 128-bit accesses; K = 0: any tupSize
__global__ void __launch_bounds__(256)
k_decompose(const int64_t* __restrict__ x, int64_t* __restrict__ digits, int64_t batch, int n, const __grid_constant__ GadgetGeom G,
            const __grid_constant__ ZqConsts Z)
{
  const int k = K ? K : G.k;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t plane = batch * (int64_t)n * k;
  for (int64_t b = blockIdx.y; b < batch; b += gridDim.y) {
    const int64_t at = (b * n + j) * k;
    int64_t raw[K ? K : 1];
    if (K == 2) { const longlong2 r = __ldcs(reinterpret_cast<const longlong2*>(x + at)); raw[0] = r.x; raw[1] = r.y; }
    for (int l = 0; l < k; l++) {
      const uint32_t q = Z.q[l];
      const uint32_t c = she_canon(K == 2 ? raw[l] : __ldcs(x + at + l), q);
      int64_t v = (2 * (uint64_t)c < (uint64_t)q) ? (int64_t)c : (int64_t)c - (int64_t)q;        // lift (ZqBasic.hs:92-94)
      int64_t* dst = digits + (int64_t)G.first[l] * plane + at;
      const int nd = G.digits[l];
      for (int d = 0; d < nd; d++) {
        const int64_t digit = d == nd - 1 ? v : div_mod_cent(v, G);      // decomp (Numeric.hs:202-205)
        if (K == 2) {
          __stcs(reinterpret_cast<longlong2*>(dst),
                 make_longlong2(she_reduce_digit(digit, Z.q[0], Z.mu[0]), she_reduce_digit(digit, Z.q[1], Z.mu[1])));
        } else {
          for (int t2 = 0; t2 < k; t2++) __stcs(dst + t2, she_reduce_digit(digit, Z.q[t2], Z.mu[t2]));
        }
        dst += plane;
      }
    }
  }
}

// c0 += sum_i digit_i * h[i][0], c1 += sum_i digit_i * h[i][1]; hints are [ell][2][n][k], one ring element each.
// ELL > 0: the 2 ELL hint words of the thread's position stay in registers across the batch; ELL = 0: any digit
// count, hints re-read (L1/L2) per ring element.  MONT: hints kept as h 2^32 mod q, so REDC(x h~) = x h; `lazy`
// (ell * q < 2^32) sums the ell products in 64 bits and reduces once.
template <typename V, int ELL, bool MONT>
__global__ void __launch_bounds__(256)
k_knapsack(const V* __restrict__ digits, int ell_rt, const V* __restrict__ hints, V* __restrict__ c0, V* __restrict__ c1, int64_t batch,
           int elem_units, int k, int lazy, const __grid_constant__ ZqConsts Z)
{
  constexpr int W = VecIO<V>::W;
  constexpr int HR = ELL ? ELL : 1;
  const int ell = ELL ? ELL : ell_rt;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= elem_units) return;
  const int64_t units = batch * elem_units;
  Limb L[W];
#pragma unroll
  for (int w = 0; w < W; w++) L[w] = make_limb<MONT>(Z, (W * u + w) % k);
  auto prep_hint = [&](int64_t h, int w) -> uint32_t {
    const uint32_t c = she_canon(h, L[w].q);
    return MONT ? she_barrett((uint64_t)c << 32, L[w].q, L[w].mu) : c;
  };
  uint32_t h0[HR][W], h1[HR][W];
  if (ELL) {
#pragma unroll
    for (int d = 0; d < HR; d++) {
      int64_t t0[W], t1[W];
      VecIO<V>::get(__ldg(hints + ((int64_t)d * 2 + 0) * elem_units + u), t0);
      VecIO<V>::get(__ldg(hints + ((int64_t)d * 2 + 1) * elem_units + u), t1);
#pragma unroll
      for (int w = 0; w < W; w++) { h0[d][w] = prep_hint(t0[w], w); h1[d][w] = prep_hint(t1[w], w); }
    }
  }
  // two ring elements per iteration: every load of both is issued before the first use
  constexpr int U = 2;
  for (int64_t b = blockIdx.y; b < batch; b += (int64_t)U * gridDim.y) {
    int64_t idx[U];
    bool live[U];
    V ra0[U], ra1[U], rx[U][ELL ? ELL : 1];
#pragma unroll
    for (int e = 0; e < U; e++) {
      const int64_t be = b + (int64_t)e * gridDim.y;
      live[e] = be < batch;
      idx[e] = (live[e] ? be : b) * elem_units + u;
      ra0[e] = __ldcs(c0 + idx[e]);
      ra1[e] = __ldcs(c1 + idx[e]);
      if (ELL) {
#pragma unroll
        for (int d = 0; d < ELL; d++) rx[e][d] = __ldcs(digits + (int64_t)d * units + idx[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < U; e++) {
      int64_t a0[W], a1[W];
      VecIO<V>::get(ra0[e], a0);
      VecIO<V>::get(ra1[e], a1);
      uint32_t s0[W], s1[W];
      uint64_t acc0[W], acc1[W];
#pragma unroll
      for (int w = 0; w < W; w++) { s0[w] = she_canon(a0[w], L[w].q); s1[w] = she_canon(a1[w], L[w].q); acc0[w] = 0; acc1[w] = 0; }
#pragma unroll
      for (int d = 0; d < (ELL ? ELL : ell); d++) {
        int64_t x[W];
        if (ELL) {
          VecIO<V>::get(rx[e][d], x);
        } else {
          VecIO<V>::get(__ldcs(digits + (int64_t)d * units + idx[e]), x);
          int64_t t0[W], t1[W];
          VecIO<V>::get(__ldg(hints + ((int64_t)d * 2 + 0) * elem_units + u), t0);
          VecIO<V>::get(__ldg(hints + ((int64_t)d * 2 + 1) * elem_units + u), t1);
#pragma unroll
          for (int w = 0; w < W; w++) { h0[0][w] = prep_hint(t0[w], w); h1[0][w] = prep_hint(t1[w], w); }
        }
        const int hd = ELL ? d : 0;
#pragma unroll
        for (int w = 0; w < W; w++) {
          const uint32_t xv = she_canon(x[w], L[w].q);
          if (MONT && lazy) {
            acc0[w] += (uint64_t)xv * h0[hd][w];
            acc1[w] += (uint64_t)xv * h1[hd][w];
          } else if (MONT) {
            s0[w] = csub(s0[w] + csub(redc((uint64_t)xv * h0[hd][w], L[w]), L[w].q), L[w].q);
            s1[w] = csub(s1[w] + csub(redc((uint64_t)xv * h1[hd][w], L[w]), L[w].q), L[w].q);
          } else {
            s0[w] = she_addmod(s0[w], she_barrett((uint64_t)xv * h0[hd][w], L[w].q, L[w].mu), L[w].q);
            s1[w] = she_addmod(s1[w], she_barrett((uint64_t)xv * h1[hd][w], L[w].q, L[w].mu), L[w].q);
          }
        }
      }
#pragma unroll
      for (int w = 0; w < W; w++) {
        if (MONT && lazy) {      // REDC < 2q: canonical first, so that the sum stays below 2q < 2^32
          s0[w] = csub(s0[w] + csub(redc(acc0[w], L[w]), L[w].q), L[w].q);
          s1[w] = csub(s1[w] + csub(redc(acc1[w], L[w]), L[w].q), L[w].q);
        }
        a0[w] = (int64_t)s0[w]; a1[w] = (int64_t)s1[w];
      }
      if (live[e]) {
        __stcs(c0 + idx[e], VecIO<V>::put(a0));
        __stcs(c1 + idx[e], VecIO<V>::put(a1));
      }
    }
  }
}

// grid: x covers the units of one ring element, y strides over the batch with about two waves of resident threads
dim3 grid_for(const lolb_plan* pl, int64_t elem_units, int64_t batch)
{
  dim3 g((unsigned)((elem_units + 255) / 256), 1, 1);
  int64_t gy = ((int64_t)pl->num_sms * 8 + g.x - 1) / g.x * 2;
  if (gy > batch) gy = batch;
  if (gy > 65535) gy = 65535;
  g.y = (unsigned)(gy < 1 ? 1 : gy);
  return g;
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// Montgomery class: every modulus odd and below 2^31
bool mont_ok(const lolb_plan* pl)
{
  for (int64_t q : pl->qs) if (!(q & 1) || q >= ((int64_t)1 << 31)) return false;
  return true;
}

int gadget_geom(const lolb_plan* pl, int64_t base, GadgetGeom* G)
{
  if (base != 0 && base < 2) { set_error("gadget base must be 0 (TrivGad) or >= 2 (BaseBGad)"); return LOLB_ERR_ARG; }
  G->k = pl->k; G->base = base; G->ell = 0; G->shift = -1;
  if (base && !(base & (base - 1))) { G->shift = 0; while (((int64_t)1 << G->shift) < base) G->shift++; }
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

int she_gadget_digits(const lolb_plan* pl, int64_t base, int* nd, int* shift)
{
  GadgetGeom G;
  if (gadget_geom(pl, base, &G)) return -1;
  for (int l = 0; l < pl->k; l++) nd[l] = G.digits[l];
  *shift = G.shift;
  return G.ell;
}

int she_ct_mul(const lolb_plan* pl, const int64_t* a0, const int64_t* a1, const int64_t* b0, const int64_t* b1,
               const int64_t* g, int64_t* d0, int64_t* d1, int64_t* d2, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  const int64_t nk = (int64_t)pl->n * pl->k;
  const bool vec = !(nk & 1) && aligned16(a0) && aligned16(a1) && aligned16(b0) && aligned16(b1) && aligned16(d0) && aligned16(d1) &&
                   aligned16(d2) && (!g || aligned16(g));
  const bool mont = mont_ok(pl);
#define CT(V, M, UNITS)                                                                                                       \
  k_ct_mul<V, M><<<grid_for(pl, UNITS, batch), 256, 0, st>>>((const V*)a0, (const V*)a1, (const V*)b0, (const V*)b1, (const V*)g,  \
                                                             (V*)d0, (V*)d1, (V*)d2, batch, (int)(UNITS), pl->k, pl->zq_plain)
  if (vec) { if (mont) CT(longlong2, true, nk / 2); else CT(longlong2, false, nk / 2); }
  else { if (mont) CT(long long, true, nk); else CT(long long, false, nk); }
#undef CT
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
  if (pl->k == 2 && aligned16(x) && aligned16(digits))
    k_decompose<2><<<grid_for(pl, pl->n, batch), 256, 0, st>>>(x, digits, batch, pl->n, G, pl->zq_plain);
  else
    k_decompose<0><<<grid_for(pl, pl->n, batch), 256, 0, st>>>(x, digits, batch, pl->n, G, pl->zq_plain);
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
  const int64_t nk = (int64_t)pl->n * pl->k;
  const bool vec = !(nk & 1) && aligned16(digits) && aligned16(hints) && aligned16(c0) && aligned16(c1);
  const bool mont = mont_ok(pl);
  int64_t qmax = 0;
  for (int64_t q : pl->qs) qmax = q > qmax ? q : qmax;
  const int lazy = (int64_t)ell * qmax < ((int64_t)1 << 32);
#define KS(V, E, M, UNITS)                                                                                                    \
  k_knapsack<V, E, M><<<grid_for(pl, UNITS, batch), 256, 0, st>>>((const V*)digits, ell, (const V*)hints, (V*)c0, (V*)c1, batch, \
                                                                  (int)(UNITS), pl->k, lazy, pl->zq_plain)
#define KS_ELL(V, M, UNITS)                                                                                                   \
  switch (ell) {                                                                                                               \
    case 1: KS(V, 1, M, UNITS); break;                                                                                         \
    case 2: KS(V, 2, M, UNITS); break;                                                                                         \
    case 3: KS(V, 3, M, UNITS); break;                                                                                         \
    case 4: KS(V, 4, M, UNITS); break;                                                                                         \
    case 5: KS(V, 5, M, UNITS); break;                                                                                         \
    case 6: KS(V, 6, M, UNITS); break;                                                                                         \
    default: KS(V, 0, M, UNITS); break;                                                                                        \
  }
  if (vec) { if (mont) { KS_ELL(longlong2, true, nk / 2) } else { KS_ELL(longlong2, false, nk / 2) } }
  else { if (mont) { KS_ELL(long long, true, nk) } else { KS_ELL(long long, false, nk) } }
#undef KS_ELL
#undef KS
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_knapsack");
  count_launch();
  return LOLB_OK;
}

}  // namespace lolb
