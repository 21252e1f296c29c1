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

// |intermediates| of line_op<KIND, P> for inputs in [0, q), as a multiple of q: the divisions by g multiply by up to P twice; L, L^-1
// and the multiplications by g only add up to P terms (and go below zero by less than q).  The bias added before the reduction and
// the 32-bit mode (everything below 2^31) are sized from it.
template <int KIND>
__host__ __device__ constexpr int line_mult(int P) { return (KIND == PASS_GINVPOW || KIND == PASS_GINVDEC) ? P * P : P + 2; }

// Arithmetic mode of the tile kernel for an operator and a set of moduli: 1 = int32 intermediates (line_mult q < 2^31), 2 = the
// divisions by g with running sums kept reduced (lineop_ginv_red: q max(2 P, P (P - 1) / 2) < 2^32 at P = 13, i.e. q < 55 M -- the
// 25-bit moduli of the HomomPRF chains), 0 = int64 intermediates
inline int line_tile_mode(int kind, const ZqConsts& zc, int k)
{
  const bool ginv = kind == PASS_GINVPOW || kind == PASS_GINVDEC;
  bool narrow = true, red = ginv;
  for (int t = 0; t < k; t++) {
    narrow = narrow && (int64_t)zc.q[t] * (ginv ? 13 * 13 : 13 + 2) < ((int64_t)1 << 31);
    red = red && (int64_t)zc.q[t] * 78 < ((int64_t)1 << 32);
  }
  return narrow ? 1 : red ? 2 : 0;
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

// x s mod q for a fixed s < q < 2^31 with sp = floor(s 2^32 / q) (Shoup): three 32-bit multiplies instead of a 64-bit Barrett step
__device__ __forceinline__ uint32_t mul_fixed(uint32_t x, uint32_t s, uint32_t sp, uint32_t q)
{
  const uint32_t r = x * s - __umulhi(x, sp) * q;      // [0, 2q)
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
  const uint32_t sp = (NARROW && scale) ? (uint32_t)(((uint64_t)s << 32) / q) : 0u;      // NARROW implies q < 2^31
  // bias: a multiple of q above the largest negative intermediate
  const int64_t biasA = (int64_t)q * line_mult<KIND>(PA), biasB = (int64_t)q * line_mult<KIND>(PB);
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
          for (int a = 0; a < DA; a++)
            v[u][b][a] = NARROW ? (I)mul_fixed((uint32_t)v[u][b][a], s, sp, q) : (I)barrett64((uint64_t)(uint32_t)v[u][b][a] * s, q, mu);
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

// ------------------------------------------------------------------ the same operators with the ring elements staged in shared memory
// For indices with more odd-prime axes than a register tile can hold ({7,13}: 72 values) or with short runs below the axes (m = 5460,
// 4095: 16- / 32-byte runs per lane group), a CTA loads whole ring elements (tupSize 1, contiguous 16-byte loads) into a u32 tile,
// runs one in-place pass per odd-prime axis (a thread per line of p - 1 words, consecutive threads on consecutive lines) and
// writes the elements back the same way: one HBM round trip for any number of axes.
struct TileGeom {
  int32_t n, naxes, epb, threads;  // n: words per element (tupSize folded in: word w belongs to limb w mod k)
  int32_t p[4], rts[4], lines[4];
  uint32_t m_rts[4];              // ceil(2^32 / d): exact floor(x / d) by __umulhi for x d < 2^32
  int32_t k;
  uint32_t m_k;                   // ceil(2^32 / k)
};

__device__ __forceinline__ int fdiv(uint32_t x, int d, uint32_t magic) { return d == 1 ? (int)x : (int)__umulhi(x, magic); }

// The divisions by g on one line with every running value kept in [0, q): the same integers as line_op modulo q (g.cpp:60-123), products
// bounded by P q instead of P^2 q, so moduli up to 2^32 / 78 stay in 32-bit arithmetic.  v in [0, q) on entry and on exit.
template <int KIND, int P>
__device__ __forceinline__ void lineop_ginv_red(uint32_t (&v)[P - 1], const uint32_t q, const uint32_t mu32)
{
  constexpr int D = P - 1;
  if (KIND == PASS_GINVPOW) {                             // v[a] = (P-1-a) lo - (a+1) hi, lo = sum_{i<=a} v, hi = sum_{i>a} v
    uint32_t sum = 0;
#pragma unroll
    for (int a = 0; a < D; a++) sum += v[a];              // <= (P-1) q
    uint32_t lo = barrett32(sum, q, mu32), hi = 0;
#pragma unroll
    for (int a = D - 1; a >= 0; a--) {
      const uint32_t z = v[a];
      v[a] = barrett32((uint32_t)(P - 1 - a) * lo + (uint32_t)P * q - (uint32_t)(a + 1) * hi, q, mu32);      // in (0, 2 P q)
      lo = lo >= z ? lo - z : lo + q - z;
      hi += z; hi = hi >= q ? hi - q : hi;
    }
  } else {                                                // v[D-1] = s = sum (a+1) v[a]; v[a-1] = v[a] - P v_in[a]
    uint32_t s = 0;
#pragma unroll
    for (int a = 0; a < D; a++) s += (uint32_t)(a + 1) * v[a];      // <= P (P-1) / 2 q
    uint32_t acc = barrett32(s, q, mu32);
#pragma unroll
    for (int a = D - 1; a >= 1; a--) {
      const uint32_t keep = acc;
      acc = barrett32(acc + (uint32_t)P * q - (uint32_t)P * v[a], q, mu32);      // in (0, (P+1) q)
      v[a] = keep;
    }
    v[0] = acc;
  }
}

// MULTI (tupSize > 1): the interleaved element is the tensor with one more innermost axis of length k -- strides and n carry the
// factor k, the limb of a line is its index mod k, and the per-limb constants (Shoup factors in `sps`) are read per line.
template <int KIND, int P, typename I, bool MULTI, bool RED>
__device__ __forceinline__ void tile_axis(uint32_t* tile, const TileGeom& G, const int ax, const int units, const ZqConsts& Z, const bool scale,
                                          const uint32_t* sps)
{
  constexpr int D = P - 1;
  const int rts = G.rts[ax], lines = G.lines[ax];
  const int total = units * lines;
  for (int L = threadIdx.x; L < total; L += blockDim.x) {
    // line L of the CTA's elements: element e, block hi, offset lo -> e n + hi rts D + lo = (L / rts) rts D + lo: one division per line
    const int uh = fdiv((uint32_t)L, rts, G.m_rts[ax]), lo = L - uh * rts;
    uint32_t* base = tile + (size_t)uh * rts * D + lo;
    const int limb = MULTI ? L - fdiv((uint32_t)L, G.k, G.m_k) * G.k : 0;      // rts is a multiple of k
    const uint32_t q = Z.q[limb];
    const uint64_t mu = Z.mu[limb];
    if constexpr (RED) {
      uint32_t w[D];
#pragma unroll
      for (int a = 0; a < D; a++) w[a] = base[a * rts];
      lineop_ginv_red<KIND, P>(w, q, (uint32_t)(mu >> 32));
#pragma unroll
      for (int a = 0; a < D; a++) base[a * rts] = scale ? mul_fixed(w[a], Z.scale[limb], sps[limb], q) : w[a];
    } else {
      I v[D];
#pragma unroll
      for (int a = 0; a < D; a++) v[a] = (I)base[a * rts];
      line_op<KIND, P, I>(v);
      const int64_t bias = (int64_t)q * line_mult<KIND>(P);
#pragma unroll
      for (int a = 0; a < D; a++) {
        uint32_t r = (uint32_t)reduce_biased(v[a], bias, q, mu);
        if (scale) r = sizeof(I) == 4 ? mul_fixed(r, Z.scale[limb], sps[limb], q) : barrett64((uint64_t)r * Z.scale[limb], q, mu);
        base[a * rts] = r;
      }
    }
  }
}

template <int KIND, int MODE, bool MULTI>
__global__ void __launch_bounds__(512)
k_line_tile(int64_t* __restrict__ y, int64_t batch, const __grid_constant__ TileGeom G, const __grid_constant__ ZqConsts Z, int scale)
{
  constexpr bool NARROW = MODE != 0, RED = MODE == 2 && (KIND == PASS_GINVPOW || KIND == PASS_GINVDEC);
  typedef typename std::conditional<NARROW, int32_t, int64_t>::type I;
  // shared memory: the u32 tile of the current group, then a raw int64 staging buffer the NEXT group arrives in by cp.async
  // (16-byte, L2-only) while the passes run on the current one: a CTA overlaps its own HBM reads with its arithmetic.
  extern __shared__ __align__(16) uint32_t line_tile[];
  const int tile_words = G.epb * G.n;
  longlong2* raw = reinterpret_cast<longlong2*>(line_tile + ((tile_words + 3) & ~3));      // 16-byte aligned
  __shared__ uint32_t sps[kMaxLimbs];                                                            // Shoup factors of rad_odd^-1 per limb
  if (threadIdx.x < G.k) sps[threadIdx.x] = (NARROW && scale) ? (uint32_t)(((uint64_t)Z.scale[threadIdx.x] << 32) / Z.q[threadIdx.x]) : 0u;      // NARROW implies q < 2^31
  const int64_t ngroups = (batch + G.epb - 1) / G.epb;
  auto prefetch = [&](int64_t g) {
    const int64_t e0 = g * G.epb;
    const int pairs = (int)(batch - e0 < G.epb ? batch - e0 : G.epb) * G.n / 2;      // n is even (checked by the host)
    const longlong2* src = reinterpret_cast<const longlong2*>(y + (size_t)e0 * G.n);
    for (int i = threadIdx.x; i < pairs; i += blockDim.x) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(raw + i);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int64_t)blockIdx.x < ngroups) prefetch(blockIdx.x);
  for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const int64_t e0 = g * G.epb;
    const int cnt = (int)(batch - e0 < G.epb ? batch - e0 : G.epb);
    longlong2* dst = reinterpret_cast<longlong2*>(y + (size_t)e0 * G.n);
    const int pairs = cnt * G.n / 2;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();      // the staged group is complete, and every thread is past the previous iteration's reads of the tile
    // ---- in: word w of the piece -> tile[w]
    for (int i = threadIdx.x; i < pairs; i += blockDim.x) {
      const longlong2 r = raw[i];
      uint32_t c0 = (uint32_t)r.x, c1 = (uint32_t)r.y;
      const int l0 = MULTI ? 2 * i - fdiv((uint32_t)(2 * i), G.k, G.m_k) * G.k : 0, l1 = MULTI ? (l0 + 1 == G.k ? 0 : l0 + 1) : 0;
      const uint32_t q0 = Z.q[l0], q1 = Z.q[l1];
      if ((uint64_t)r.x >= (uint64_t)q0) { int64_t t = r.x % (int64_t)q0; c0 = (uint32_t)(t < 0 ? t + q0 : t); }      // like the reference's c % q
      if ((uint64_t)r.y >= (uint64_t)q1) { int64_t t = r.y % (int64_t)q1; c1 = (uint32_t)(t < 0 ? t + q1 : t); }
      *reinterpret_cast<uint2*>(line_tile + 2 * i) = make_uint2(c0, c1);
    }
    __syncthreads();
    if (g + gridDim.x < ngroups) prefetch(g + gridDim.x);
    // ---- one pass per odd-prime axis, in place
    for (int ax = 0; ax < G.naxes; ax++) {
      const bool sc = scale && ax == G.naxes - 1;
      switch (G.p[ax]) {
        case 3: tile_axis<KIND, 3, I, MULTI, RED>(line_tile, G, ax, cnt, Z, sc, sps); break;
        case 5: tile_axis<KIND, 5, I, MULTI, RED>(line_tile, G, ax, cnt, Z, sc, sps); break;
        case 7: tile_axis<KIND, 7, I, MULTI, RED>(line_tile, G, ax, cnt, Z, sc, sps); break;
        case 11: tile_axis<KIND, 11, I, MULTI, RED>(line_tile, G, ax, cnt, Z, sc, sps); break;
        default: tile_axis<KIND, 13, I, MULTI, RED>(line_tile, G, ax, cnt, Z, sc, sps); break;
      }
      __syncthreads();
    }
    // ---- out
    for (int i = threadIdx.x; i < pairs; i += blockDim.x) {
      const uint2 v = *reinterpret_cast<const uint2*>(line_tile + 2 * i);
      __stcs(dst + i, make_longlong2((int64_t)v.x, (int64_t)v.y));
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
  for (int t = 0; t < pl->k; t++) narrow = narrow && (int64_t)zc.q[t] * line_mult<KIND>((int)PM) < ((int64_t)1 << 31);
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

static uint32_t magic_div(uint32_t d) { return (uint32_t)((((uint64_t)1 << 32) + d - 1) / d); }      // exact floor(x / d) by umulhi while x d < 2^32

// the shared-memory kernel: odd primes from {3, 5, 7, 11, 13}, at most four odd axes, n k even, one element within the opt-in shared memory
static bool line_tile_geom(const lolb_plan* pl, TileGeom* G)
{
  int p[4]; int64_t r[4];
  const int cnt = odd_axes(pl, p, r);
  if (pl->k > kMaxLimbs || cnt < 1 || cnt > 4) return false;
  for (int i = 0; i < cnt; i++) if (p[i] != 3 && p[i] != 5 && p[i] != 7 && p[i] != 11 && p[i] != 13) return false;
  const int64_t nk = (int64_t)pl->n * pl->k;
  if ((nk & 1) || nk > 16384) return false;      // 12 bytes of shared memory per word
  G->n = (int32_t)nk;
  G->k = pl->k;
  G->m_k = magic_div((uint32_t)pl->k);
  G->naxes = cnt;
  // measured 2048 / 4096 / 8192 / 16384 words per CTA: m = 2912 L 50 / 60 / 55 / 58 %, m = 5460 38 / 49 / 47 / 46 % of HBM
  // the CTA shape (threads, elements) that leaves no idle threads in the last round of an axis pass
  // (lolb_internal.cuh::choose_tile_shape) within `tile_words` (LOLB_LINE_TILE_WORDS: tuning runs; 4096 measured best)
  static const int tile_words = [] { const char* e = getenv("LOLB_LINE_TILE_WORDS"); return e ? atoi(e) : 4096; }();
  const TileShape sh = choose_tile_shape(nk, p, cnt, sizeof(uint32_t), (size_t)tile_words * sizeof(uint32_t), false);
  const int64_t epb = sh.epb;
  G->threads = sh.threads;
  G->epb = (int32_t)epb;
  // an element too large for two CTAs per SM (m = 5824, tupSize 4: 110 KB with the staging buffer): twice the threads in the one CTA
  if ((size_t)epb * nk * 12 > 72 * 1024 && G->threads <= 256) G->threads *= 2;
  for (int i = 0; i < cnt; i++) {
    G->p[i] = p[i]; G->rts[i] = (int32_t)(r[i] * pl->k); G->lines[i] = (int32_t)(nk / (p[i] - 1));
    G->m_rts[i] = magic_div((uint32_t)(r[i] * pl->k));
  }
  return true;
}

template <int KIND>
static int launch_line_tile(const lolb_plan* pl, const TileGeom& G, const ZqConsts& zc, bool scale, int64_t* y, int64_t batch, cudaStream_t st)
{
  const int mode = line_tile_mode(KIND, zc, pl->k);
  const size_t smem = (size_t)G.epb * G.n * (sizeof(uint32_t) + sizeof(int64_t)) + 16;      // u32 tile + the raw staging buffer of the next group
  const int64_t groups = (batch + G.epb - 1) / G.epb;
  int per_sm = (int)(200 * 1024 / (smem + 1024));
  if (per_sm > 2048 / G.threads) per_sm = 2048 / G.threads;
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)pl->num_sms * per_sm;
  if (grid > groups) grid = groups;
  cudaError_t e = cudaSuccess;
  auto go = [&](auto kern) {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) kern<<<(int)grid, G.threads, smem, st>>>(y, batch, G, zc, scale ? 1 : 0);
  };
  constexpr bool GINV = KIND == PASS_GINVPOW || KIND == PASS_GINVDEC;
  if (pl->k == 1) {
    if (mode == 1) go(k_line_tile<KIND, 1, false>);
    else if (mode == 2) { if constexpr (GINV) go(k_line_tile<KIND, 2, false>); }
    else go(k_line_tile<KIND, 0, false>);
  } else if (mode == 1) go(k_line_tile<KIND, 1, true>);
  else if (mode == 2) { if constexpr (GINV) go(k_line_tile<KIND, 2, true>); }
  else return LOLB_FUSED_UNAVAILABLE;      // several limbs in the 64-bit mode: the register-tile launches (fused_stream_line routes them there)
  if (e != cudaSuccess) return cuda_fail(e, "k_line_tile shared memory");
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_line_tile");
  count_launch();
  return LOLB_OK;
}

// which kernel serves the plan: 0 identity, 1 register tiles (one launch), 2 shared-memory tile, 3 register tiles in several launches, -1 none
static int line_route(const lolb_plan* pl, LineStep (&st)[4], int* nsteps, TileGeom* G)
{
  const int ns = line_steps(pl, st);
  *nsteps = ns;
  if (ns == 0) return 0;
  if (ns == 1) return 1;
  static const bool use_tile = [] { const char* e = getenv("LOLB_LINE_TILE"); return !e || atoi(e) != 0; }();
  if (use_tile && line_tile_geom(pl, G)) return 2;      // any tupSize (folded into the strides) since the staged, double-buffered version
  return ns > 1 ? 3 : -1;
}

const char* fused_stream_line_name(const lolb_plan* pl, bool ginv)
{
  LineStep st[4];
  TileGeom G;
  int ns;
  int route = line_route(pl, st, &ns, &G);
  if (route == 2 && pl->k > 1) {      // several limbs: the tile kernel in its 32-bit mode only (fused_stream_line)
    if (line_tile_mode(ginv ? PASS_GINVPOW : PASS_L, ginv ? pl->zq_radinv : pl->zq_plain, pl->k) == 0) route = 3;
  }
  return route == 0 ? "identity" : route == 2 ? "line_tile" : route > 0 ? "line_stream" : "generic";
}

int fused_stream_line(const lolb_plan* pl, int kind, const ZqConsts& zc, bool scale, int64_t* y, int64_t batch, cudaStream_t st)
{
  LineStep steps[4];
  TileGeom TG;
  int ns;
  const int route = line_route(pl, steps, &ns, &TG);
  if (batch <= 0) return LOLB_OK;
  if (route < 0) return LOLB_FUSED_UNAVAILABLE;
  // several limbs in the tile kernel pay off in its 32-bit mode (m = 5824, tupSize 2 / 4: L 77 % / 52 % against 47 % for the register
  // tiles); in the 64-bit mode (25-bit moduli under /g) the register-tile launches are as fast or faster (35 % / 29 % against 35 % / 37 %)
  bool tile_ok = route == 2 && !((uintptr_t)y & 15);
  if (tile_ok && pl->k > 1 && ns > 1) tile_ok = line_tile_mode(kind, zc, pl->k) != 0;
  if (tile_ok) {
    switch (kind) {
      case PASS_L: return launch_line_tile<PASS_L>(pl, TG, zc, scale, y, batch, st);
      case PASS_LINV: return launch_line_tile<PASS_LINV>(pl, TG, zc, scale, y, batch, st);
      case PASS_GPOW: return launch_line_tile<PASS_GPOW>(pl, TG, zc, scale, y, batch, st);
      case PASS_GDEC: return launch_line_tile<PASS_GDEC>(pl, TG, zc, scale, y, batch, st);
      case PASS_GINVPOW: return launch_line_tile<PASS_GINVPOW>(pl, TG, zc, scale, y, batch, st);
      case PASS_GINVDEC: return launch_line_tile<PASS_GINVDEC>(pl, TG, zc, scale, y, batch, st);
      default: return LOLB_FUSED_UNAVAILABLE;
    }
  }
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
