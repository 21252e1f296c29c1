// fused_stream.cu -- streaming (register-only, no shared memory, no barrier) kernels for the cheap linear operators
// of the path over Z_q:  L, L^-1, *g in Pow/Dec, /g in Pow/Dec (l.cpp:28-98, g.cpp:16-123 lifted by tensor.h:39-74)
// and the coefficient-wise product (mul.cpp:14-30).
//
// Line operators.  For an index with one or two odd prime factors every operator is a Kronecker product of small
// integer matrices along the (p-1)-long lines of each odd-prime axis.  A thread owns one TILE: all (pA-1)*(pB-1)
// coefficients that share every other tensor digit, loads them (consecutive threads <-> consecutive fastest
// digit, so each warp instruction touches contiguous memory), applies axis A then axis B exactly over the
// integers in int64 (|entries| <= p, <= p terms: no overflow for q < 2^32), reduces once per axis modulo q,
// and stores.  Algorithmic traffic: 16 bytes per coefficient, one read and one write.
#include <cstdlib>
#include <type_traits>

#include "fused.cuh"

namespace lolb {

namespace {

__device__ __forceinline__ uint32_t barrett64(uint64_t x, uint32_t q, uint64_t mu)
{
  uint64_t r = x - __umul64hi(x, mu) * q;       // [0, 2q)
  if (r >= q) r -= q;
  if (r >= q) r -= q;
  return (uint32_t)r;
}

// exact integer form of the prime-index operators on one line v[0..D), D = P-1
template <int KIND, int P, typename I>
__device__ __forceinline__ void line_op(I (&v)[P - 1])
{
  constexpr int D = P - 1;
  if (KIND == PASS_L) {                                   // l.cpp:28-57
#pragma unroll
    for (int a = 1; a < D; a++) v[a] += v[a - 1];
  } else if (KIND == PASS_LINV) {                         // l.cpp:67-98
#pragma unroll
    for (int a = D - 1; a >= 1; a--) v[a] -= v[a - 1];
  } else if (KIND == PASS_GPOW) {                         // g.cpp:16-35
    const I last = v[D - 1];
#pragma unroll
    for (int a = D - 1; a >= 1; a--) v[a] += last - v[a - 1];
    v[0] += last;
  } else if (KIND == PASS_GDEC) {                         // g.cpp:37-58
    I acc = v[0];
#pragma unroll
    for (int a = D - 1; a >= 1; a--) { acc += v[a]; v[a] -= v[a - 1]; }
    v[0] += acc;
  } else if (KIND == PASS_GINVPOW) {                      // g.cpp:60-90
    I lo = 0, hi = 0;
#pragma unroll
    for (int a = 0; a < D; a++) lo += v[a];
#pragma unroll
    for (int a = D - 1; a >= 0; a--) {
      const I z = v[a];
      v[a] = (I)(P - 1 - a) * lo - (I)(a + 1) * hi;
      lo -= z; hi += z;
    }
  } else if (KIND == PASS_GINVDEC) {                      // g.cpp:92-123
    I s = 0;
#pragma unroll
    for (int a = 0; a < D; a++) s += (I)(a + 1) * v[a];
    I acc = s;
#pragma unroll
    for (int a = D - 1; a >= 1; a--) { const I keep = acc; acc -= v[a] * (I)P; v[a] = keep; }
    v[0] = acc;
  }
}

struct LineGeom {
  int32_t n, k;
  int32_t RA, RB;        // strides (rts) of the two axes; RB = n when there is no second axis
  int32_t M;             // RB / (RA * dA)
  int32_t tiles;         // n / (dA * dB)
};

__device__ __forceinline__ uint32_t barrett32(uint32_t x, uint32_t q, uint32_t mu32)
{
  uint32_t r = x - __umulhi(x, mu32) * q;       // [0, 2q)
  return min(r, r - q);
}

__device__ __forceinline__ int32_t reduce_biased(int32_t x, int64_t bias, uint32_t q, uint64_t mu)
{ return (int32_t)barrett32((uint32_t)x + (uint32_t)bias, q, (uint32_t)(mu >> 32)); }
__device__ __forceinline__ int64_t reduce_biased(int64_t x, int64_t bias, uint32_t q, uint64_t mu)
{ return (int64_t)barrett64((uint64_t)(x + bias), q, mu); }

// NARROW: P*P*q < 2^31 on both axes, so the exact integer intermediates fit int32 and one 32-bit Barrett step reduces
template <int KIND, int PA, int PB, bool NARROW>
__global__ void __launch_bounds__(256)
k_line_stream(int64_t* __restrict__ y, int64_t batch, const __grid_constant__ LineGeom G, const __grid_constant__ ZqConsts Z, int scale)
{
  constexpr int DA = PA - 1, DB = PB > 1 ? PB - 1 : 1;
  const int k = G.k;
  // the tile / limb of a thread is fixed; blockIdx.y strides over ring elements
  const int rem = blockIdx.x * blockDim.x + threadIdx.x;
  if (rem >= G.tiles * k) return;
  const int limb = rem % k;
  int t = rem / k;
  const int lo = t % G.RA; t /= G.RA;
  const int mid = t % G.M;
  const int hi = t / G.M;
  const uint32_t q = Z.q[limb];
  const uint64_t mu = Z.mu[limb];
  const uint32_t s = Z.scale[limb];
  // bias: a multiple of q above the largest negative intermediate (|x| <= P*P*q)
  const int64_t biasA = (int64_t)q * (PA * PA), biasB = (int64_t)q * (PB * PB);
  const size_t off = ((size_t)lo + (size_t)G.RA * DA * mid + (size_t)G.RB * DB * hi) * k + limb;
  const size_t sa = (size_t)G.RA * k, sb = (size_t)G.RB * k;
  typedef typename std::conditional<NARROW, int32_t, int64_t>::type I;
  // U ring elements per iteration, so that a thread with a small tile (2 values for a single axis of p = 3) still has 8 loads in flight
  constexpr int U = DA * DB <= 2 ? 4 : DA * DB <= 4 ? 2 : 1;
  for (int64_t e0 = (int64_t)blockIdx.y * U; e0 < batch; e0 += (int64_t)gridDim.y * U) {
    int64_t raw[U][DB][DA];      // every load is issued before the first use
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int64_t* base = y + (size_t)(e0 + u < batch ? e0 + u : e0) * G.n * k + off;
#pragma unroll
      for (int b = 0; b < DB; b++)
#pragma unroll
        for (int a = 0; a < DA; a++) raw[u][b][a] = __ldcs(base + sa * a + sb * b);
    }
    // canonical input is the contract; anything else is reduced first like `c % q` (types.h:62-66)
    bool odd_input = false;
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int b = 0; b < DB; b++)
#pragma unroll
        for (int a = 0; a < DA; a++) odd_input |= (uint64_t)raw[u][b][a] >= (uint64_t)q;
    if (odd_input) {
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int b = 0; b < DB; b++)
#pragma unroll
          for (int a = 0; a < DA; a++) { int64_t r = raw[u][b][a] % (int64_t)q; raw[u][b][a] = r < 0 ? r + q : r; }
    }
    I v[U][DB][DA];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
      for (int b = 0; b < DB; b++)
#pragma unroll
        for (int a = 0; a < DA; a++) v[u][b][a] = (I)raw[u][b][a];
#pragma unroll
    for (int u = 0; u < U; u++) {
#pragma unroll
      for (int b = 0; b < DB; b++) {
        line_op<KIND, PA, I>(v[u][b]);
#pragma unroll
        for (int a = 0; a < DA; a++) v[u][b][a] = reduce_biased(v[u][b][a], biasA, q, mu);
      }
      if constexpr (PB > 1) {
#pragma unroll
        for (int a = 0; a < DA; a++) {
          I w[DB];
#pragma unroll
          for (int b = 0; b < DB; b++) w[b] = v[u][b][a];
          line_op<KIND, PB, I>(w);
#pragma unroll
          for (int b = 0; b < DB; b++) v[u][b][a] = reduce_biased(w[b], biasB, q, mu);
        }
      }
      if (scale) {
#pragma unroll
        for (int b = 0; b < DB; b++)
#pragma unroll
          for (int a = 0; a < DA; a++) v[u][b][a] = (I)barrett64((uint64_t)(uint32_t)v[u][b][a] * s, q, mu);
      }
      if (e0 + u < batch) {
        int64_t* base = y + (size_t)(e0 + u) * G.n * k + off;
#pragma unroll
        for (int b = 0; b < DB; b++)
#pragma unroll
          for (int a = 0; a < DA; a++) __stcs(base + sa * a + sb * b, (int64_t)(uint32_t)v[u][b][a]);
      }
    }
  }
}

// coefficient-wise product, two coefficients (16 bytes) per thread per operand
__global__ void __launch_bounds__(256)
k_mul_stream(longlong2* __restrict__ a, const longlong2* __restrict__ b, int64_t pairs, int64_t b_pairs, int k,
             const __grid_constant__ ZqConsts Z)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pairs; i += (int64_t)gridDim.x * blockDim.x) {
    const longlong2 x = __ldcs(a + i);
    const longlong2 w = (b_pairs == pairs) ? __ldcs(b + i) : __ldg(b + (i % b_pairs));
    const int l0 = (int)((2 * i) % k), l1 = (int)((2 * i + 1) % k);
    const uint32_t q0 = Z.q[l0], q1 = Z.q[l1];
    uint64_t x0 = (uint64_t)x.x, x1 = (uint64_t)x.y, w0 = (uint64_t)w.x, w1 = (uint64_t)w.y;
    if (x0 >= q0 || w0 >= q0) { int64_t r = x.x % (int64_t)q0; x0 = r < 0 ? r + q0 : r; r = w.x % (int64_t)q0; w0 = r < 0 ? r + q0 : r; }
    if (x1 >= q1 || w1 >= q1) { int64_t r = x.y % (int64_t)q1; x1 = r < 0 ? r + q1 : r; r = w.y % (int64_t)q1; w1 = r < 0 ? r + q1 : r; }
    longlong2 o;
    o.x = (int64_t)barrett64(x0 * w0, q0, Z.mu[l0]);
    o.y = (int64_t)barrett64(x1 * w1, q1, Z.mu[l1]);
    __stcs(a + i, o);
  }
}

template <int KIND, int PA, int PB, bool NARROW>
int launch_line_n(const lolb_plan* pl, const LineGeom& G, const ZqConsts& zc, bool scale, int64_t* y, int64_t batch, cudaStream_t st)
{
  const int per_elem = G.tiles * G.k;
  int threads = per_elem >= 256 ? 256 : ((per_elem + 31) / 32) * 32;
  for (int c = 256; c >= 128; c -= 32) if (per_elem % c == 0) { threads = c; break; }   // avoid a ragged last block
  dim3 grid((per_elem + threads - 1) / threads, 1, 1);
  constexpr int DT = (PA - 1) * (PB > 1 ? PB - 1 : 1), U = DT <= 2 ? 4 : DT <= 4 ? 2 : 1;      // elements per iteration (k_line_stream)
  int64_t gy = ((int64_t)pl->num_sms * 2048 / threads + grid.x - 1) / grid.x * 2;     // ~2 waves of resident threads
  if (gy > (batch + U - 1) / U) gy = (batch + U - 1) / U;
  if (gy > 65535) gy = 65535;
  grid.y = (unsigned)gy;
  k_line_stream<KIND, PA, PB, NARROW><<<grid, threads, 0, st>>>(y, batch, G, zc, scale ? 1 : 0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_line_stream");
  count_launch();
  return LOLB_OK;
}

template <int KIND, int PA, int PB>
int launch_line(const lolb_plan* pl, const LineGeom& G, const ZqConsts& zc, bool scale, int64_t* y, int64_t batch, cudaStream_t st)
{
  constexpr int64_t PM = (PA > PB ? PA : PB);
  bool narrow = true;
  for (int t = 0; t < pl->k; t++) narrow = narrow && (int64_t)zc.q[t] * PM * PM < ((int64_t)1 << 31);
  return narrow ? launch_line_n<KIND, PA, PB, true>(pl, G, zc, scale, y, batch, st)
                : launch_line_n<KIND, PA, PB, false>(pl, G, zc, scale, y, batch, st);
}

template <int PA, int PB>
int dispatch_kind(const lolb_plan* pl, int kind, const LineGeom& G, const ZqConsts& zc, bool scale, int64_t* y, int64_t batch, cudaStream_t st)
{
  switch (kind) {
    case PASS_L: return launch_line<PASS_L, PA, PB>(pl, G, zc, scale, y, batch, st);
    case PASS_LINV: return launch_line<PASS_LINV, PA, PB>(pl, G, zc, scale, y, batch, st);
    case PASS_GPOW: return launch_line<PASS_GPOW, PA, PB>(pl, G, zc, scale, y, batch, st);
    case PASS_GDEC: return launch_line<PASS_GDEC, PA, PB>(pl, G, zc, scale, y, batch, st);
    case PASS_GINVPOW: return launch_line<PASS_GINVPOW, PA, PB>(pl, G, zc, scale, y, batch, st);
    case PASS_GINVDEC: return launch_line<PASS_GINVDEC, PA, PB>(pl, G, zc, scale, y, batch, st);
    default: return LOLB_FUSED_UNAVAILABLE;
  }
}

}  // namespace

// odd prime axes of the plan as (p, rts); returns how many
static int odd_axes(const lolb_plan* pl, int (&p)[4], int64_t (&rts)[4])
{
  int cnt = 0;
  int64_t r = 1;
  for (const PrimeExponent& pe : pl->pe) {
    int64_t phi = (int64_t)(pe.prime - 1);
    for (int i = 1; i < pe.exponent; i++) phi *= pe.prime;
    if (pe.prime != 2) { if (cnt < 4) { p[cnt] = pe.prime; rts[cnt] = r; } cnt++; }
    r *= phi;
  }
  return cnt;
}

// The operators of different prime axes commute (a Kronecker product), so an index with more odd primes than one kernel's tile
// can hold is served by a few launches, each applying one or two axes: every launch is one read and one write of the batch.
// Kernels exist for the single primes 3, 5, 7, 13 and the pairs below; preference = fewest launches for the reference's rings
// ({7,13}, {5,7,13}, {3,5,7,13}: two launches each).
struct LineStep { int pa, pb; int64_t ra, rb; };

static int line_steps(const lolb_plan* pl, LineStep (&steps)[4])
{
  int p[4]; int64_t r[4];
  const int cnt = odd_axes(pl, p, r);
  if (cnt > 4) return -1;
  for (int i = 0; i < cnt; i++) if (p[i] != 3 && p[i] != 5 && p[i] != 7 && p[i] != 13) return -1;
  // (a 6 x 12 tile for {7,13} in one launch was measured: 255 registers + 0.8 KB of spills, 32 % of the roofline at m = 2912 and 10 % at
  // m = 728 against 48 % / 46 % for the two single-axis launches -- not instantiated)
  static const int pref[][2] = {{5, 7}, {3, 13}, {3, 5}, {3, 7}};
  bool used[4] = {false, false, false, false};
  int ns = 0;
  for (const auto& pr : pref) {
    int ia = -1, ib = -1;
    for (int i = 0; i < cnt; i++) { if (!used[i] && p[i] == pr[0]) ia = i; if (!used[i] && p[i] == pr[1]) ib = i; }
    if (ia >= 0 && ib >= 0) { used[ia] = used[ib] = true; steps[ns++] = LineStep{p[ia], p[ib], r[ia], r[ib]}; }
  }
  for (int i = 0; i < cnt; i++) if (!used[i]) steps[ns++] = LineStep{p[i], 1, r[i], 0};
  return ns;
}

const char* fused_stream_line_name(const lolb_plan* pl)
{
  LineStep st[4];
  const int ns = line_steps(pl, st);
  return ns == 0 ? "identity" : ns > 0 ? "line_stream" : "generic";
}

int fused_stream_line(const lolb_plan* pl, int kind, const ZqConsts& zc, bool scale, int64_t* y, int64_t batch, cudaStream_t st)
{
  LineStep steps[4];
  const int ns = line_steps(pl, steps);
  if (batch <= 0) return LOLB_OK;
  if (ns < 0) return LOLB_FUSED_UNAVAILABLE;
  // ns == 0: every prime-index operator is the identity for p = 2 (l.cpp:35, g.cpp:18,39,62,94) and rad_odd = 1:
  // canonical input is already the result, no launch
  for (int i = 0; i < ns; i++) {
    const LineStep& S = steps[i];
    const bool sc = scale && i == ns - 1;      // the rad_odd^-1 factor of the divisions by g rides on the last launch
    LineGeom G{};
    G.n = pl->n; G.k = pl->k;
    G.RA = (int32_t)S.ra;
    int rc;
    if (S.pb == 1) {
      // a single axis: "hi" is always 0, everything above the axis folds into `mid`
      G.RB = pl->n; G.M = (int32_t)(pl->n / (S.ra * (S.pa - 1))); G.tiles = pl->n / (S.pa - 1);
      switch (S.pa) {
        case 3: rc = dispatch_kind<3, 1>(pl, kind, G, zc, sc, y, batch, st); break;
        case 5: rc = dispatch_kind<5, 1>(pl, kind, G, zc, sc, y, batch, st); break;
        case 7: rc = dispatch_kind<7, 1>(pl, kind, G, zc, sc, y, batch, st); break;
        default: rc = dispatch_kind<13, 1>(pl, kind, G, zc, sc, y, batch, st); break;
      }
    } else {
      G.RB = (int32_t)S.rb;
      G.M = (int32_t)(S.rb / (S.ra * (S.pa - 1)));
      G.tiles = pl->n / ((S.pa - 1) * (S.pb - 1));
      if (S.pa == 5) rc = dispatch_kind<5, 7>(pl, kind, G, zc, sc, y, batch, st);
      else if (S.pb == 13) rc = dispatch_kind<3, 13>(pl, kind, G, zc, sc, y, batch, st);
      else if (S.pb == 5) rc = dispatch_kind<3, 5>(pl, kind, G, zc, sc, y, batch, st);
      else rc = dispatch_kind<3, 7>(pl, kind, G, zc, sc, y, batch, st);
    }
    if (rc) return rc;
  }
  return LOLB_OK;
}

int fused_stream_mul(const lolb_plan* pl, int64_t* a, const int64_t* b, int64_t batch, int64_t b_batch, cudaStream_t st)
{
  const int64_t nk = (int64_t)pl->n * pl->k;
  if (batch <= 0) return LOLB_OK;
  // pairs of coefficients must not straddle elements of b when broadcasting, and pointers must be 16-byte aligned
  if ((nk & 1) || ((uintptr_t)a & 15) || ((uintptr_t)b & 15)) return LOLB_FUSED_UNAVAILABLE;
  const int64_t pairs = batch * nk / 2, b_pairs = b_batch * nk / 2;
  int64_t blocks = (pairs + 255) / 256;
  const int64_t cap = (int64_t)pl->num_sms * 32;
  if (blocks > cap) blocks = cap;
  k_mul_stream<<<(int)blocks, 256, 0, st>>>((longlong2*)a, (const longlong2*)b, pairs, b_pairs, pl->k, pl->zq_plain);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_mul_stream");
  count_launch();
  return LOLB_OK;
}

}  // namespace lolb
