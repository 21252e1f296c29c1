// fused_plain.cu -- streaming kernels for the modulus-free rings (wrapping int64 "R", double, complex double):
// L, L^-1, *g (l.cpp:28-98, g.cpp:16-58), the Gaussian E_m transform (random.cpp:19-64) and g-norm (norm.cpp:15-80)
// for indices with one or two small odd primes.  Same tile scheme as fused_stream.cu: a thread owns the
// (pA-1)*(pB-1) coefficients that share all other tensor digits, applies axis A then axis B in registers, one HBM
// read and one HBM write per coefficient, no shared memory.  tupSize > 1 is folded into the strides for the
// element-wise operators (I (x) A (x) I_{R*k}); Gaussian and norm are per-limb and take tupSize = 1 here.
#include <algorithm>

#include "fused.cuh"
#include "rings.cuh"

namespace lolb {

namespace {

struct PlainGeom {
  int32_t n;             // coefficients per element (tupSize folded in)
  int32_t RA, RB;        // strides of the two axes (tupSize folded in); RB = n without a second axis
  int32_t M;             // RB / (RA * dA)
  int32_t tiles;         // n / (dA * dB)
};

// ring-generic prime-index operators on one line, operation order of the reference
template <int KIND, int P, class R>
__device__ __forceinline__ void line_ring(const R& ring, typename R::T (&v)[P - 1])
{
  typedef typename R::T T;
  constexpr int D = P - 1;
  if (KIND == PASS_L) {
#pragma unroll
    for (int a = 1; a < D; a++) v[a] = ring.add(v[a], v[a - 1]);
  } else if (KIND == PASS_LINV) {
#pragma unroll
    for (int a = D - 1; a >= 1; a--) v[a] = ring.sub(v[a], v[a - 1]);
  } else if (KIND == PASS_GPOW) {
    const T last = v[D - 1];
#pragma unroll
    for (int a = D - 1; a >= 1; a--) v[a] = ring.add(v[a], ring.sub(last, v[a - 1]));
    v[0] = ring.add(v[0], last);
  } else if (KIND == PASS_GDEC) {
    T acc = v[0];
#pragma unroll
    for (int a = D - 1; a >= 1; a--) { acc = ring.add(acc, v[a]); v[a] = ring.sub(v[a], v[a - 1]); }
    v[0] = ring.add(v[0], acc);
  } else if (KIND == PASS_GINVPOW) {
    T lo = ring.zero(), hi = ring.zero();
#pragma unroll
    for (int a = 0; a < D; a++) lo = ring.add(lo, v[a]);
#pragma unroll
    for (int a = D - 1; a >= 0; a--) {
      const T z = v[a];
      v[a] = ring.sub(ring.mul(ring.from_int(P - 1 - a), lo), ring.mul(ring.from_int(a + 1), hi));
      lo = ring.sub(lo, z); hi = ring.add(hi, z);
    }
  } else if (KIND == PASS_GINVDEC) {
    T s = ring.zero();
#pragma unroll
    for (int a = 0; a < D; a++) s = ring.add(s, ring.mul(ring.from_int(a + 1), v[a]));
    T acc = s;
    const T pp = ring.from_int(P);
#pragma unroll
    for (int a = D - 1; a >= 1; a--) { const T keep = acc; acc = ring.sub(acc, ring.mul(v[a], pp)); v[a] = keep; }
    v[0] = acc;
  } else if (KIND == PASS_NORMSQ) {
    T s = ring.zero();
#pragma unroll
    for (int a = 0; a < D; a++) s = ring.add(s, v[a]);
#pragma unroll
    for (int a = 0; a < D; a++) v[a] = ring.add(v[a], s);
  }
}

struct TileIndex {
  size_t off, sa, sb;
  __device__ __forceinline__ TileIndex(const PlainGeom& G, int t, int DA, int DB)
  {
    const int lo = t % G.RA; t /= G.RA;
    const int mid = t % G.M;
    const int hi = t / G.M;
    off = (size_t)lo + (size_t)G.RA * DA * mid + (size_t)G.RB * DB * hi;
    sa = (size_t)G.RA; sb = (size_t)G.RB;
  }
};

template <class R, int KIND, int PA, int PB>
__global__ void __launch_bounds__(256)
k_line_plain(typename R::IO* __restrict__ y, int64_t batch, const __grid_constant__ PlainGeom G, double rscale)
{
  typedef typename R::T T;
  constexpr int DA = PA - 1, DB = PB > 1 ? PB - 1 : 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G.tiles) return;
  const TileIndex ix(G, t, DA, DB);
  const R ring{};
  for (int64_t e = blockIdx.y; e < batch; e += gridDim.y) {
    typename R::IO* base = y + (size_t)e * G.n + ix.off;
    T v[DB][DA];
#pragma unroll
    for (int b = 0; b < DB; b++)
#pragma unroll
      for (int a = 0; a < DA; a++) v[b][a] = __ldcs(base + ix.sa * a + ix.sb * b);
#pragma unroll
    for (int b = 0; b < DB; b++) line_ring<KIND, PA, R>(ring, v[b]);
    if constexpr (PB > 1) {
#pragma unroll
      for (int a = 0; a < DA; a++) {
        T w[DB];
#pragma unroll
        for (int b = 0; b < DB; b++) w[b] = v[b][a];
        line_ring<KIND, PB, R>(ring, w);
#pragma unroll
        for (int b = 0; b < DB; b++) v[b][a] = w[b];
      }
    }
    if constexpr (sizeof(T) == 16) {          // complex: optional real scale (g.cpp:209-220 intent)
      if (rscale != 0.0) {
#pragma unroll
        for (int b = 0; b < DB; b++)
#pragma unroll
          for (int a = 0; a < DA; a++) v[b][a] = make_double2(__dmul_rn(v[b][a].x, rscale), __dmul_rn(v[b][a].y, rscale));
      }
    }
#pragma unroll
    for (int b = 0; b < DB; b++)
#pragma unroll
      for (int a = 0; a < DA; a++) __stcs(base + ix.sa * a + ix.sb * b, v[b][a]);
  }
}

// 2 * E_p as a dense (p-1) x (p-1) real matrix per odd prime, row-major; built on the host from the complex roots
struct GaussMats {
  double a[6][6];
  double b[6][6];
};

// out[row] = (sum_col (2 * E[row][col]) * in[col]) / sqrt(2), accumulation order of random.cpp:33-40.  The matrices
// arrive as 2*E (exact), and the final division is a multiplication by the correctly rounded 1/sqrt(2): at most one
// ulp away from the reference's quotient (the Gaussian path is specified to 1e-9 relative; the generic engine keeps
// the division and is the cross-check), and it removes a ~25-instruction double division per coefficient and axis.
template <int P>
__device__ __forceinline__ void gauss_line(double (&v)[P - 1], const double (&E)[6][6])
{
  constexpr int D = P - 1;
  double o[D];
  const double inv_sqrt2 = 0.70710678118654752440;
#pragma unroll
  for (int row = 0; row < D; row++) {
    double acc = 0.0;
#pragma unroll
    for (int col = 0; col < D; col++) acc = __dadd_rn(acc, __dmul_rn(E[row][col], v[col]));
    o[row] = __dmul_rn(acc, inv_sqrt2);
  }
#pragma unroll
  for (int row = 0; row < D; row++) v[row] = o[row];
}

// ------------------------------------------------------------------ on-device Gaussian source (GaussRandom.hs:34-59)
// The reference draws its continuous Gaussians on the host: `realGaussian` is the polar form of Box-Muller over a
// `MonadRandom` -- (u, v) uniform in (-1, 1)^2 until 0 < t = u^2 + v^2 < 1, then (u, v) * sqrt(-var ln t / t) with
// var = svar / pi (svar = 2 pi x the true variance).  Here the same transform runs per thread over a counter-based
// generator, Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11): a draw is a pure function of (seed, ring element, pair
// index, attempt), so results do not depend on the launch shape and batches can be generated in pieces.  A different
// uniform source than any Haskell `RandomGen`, so parity with the reference is distributional (tests: moments, KS, the
// reference's gSqNorm bound).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

// 53 random bits -> uniform on the open interval (-1, 1), symmetric about 0
__device__ __forceinline__ double uniform_pm1(uint32_t hi, uint32_t lo)
{
  const uint64_t bits = ((uint64_t)hi << 21) | (lo >> 11);                   // 53 bits
  return ((double)bits + 0.5) * (2.0 / 9007199254740992.0) - 1.0;
}

// one accepted pair of independent N(0, var2 / 2) values
__device__ __forceinline__ void gauss_pair(uint64_t seed, uint64_t element, uint32_t pair, double var2, double& g0, double& g1)
{
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  for (uint32_t attempt = 0;; attempt++) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)element, (uint32_t)(element >> 32), pair, attempt), key);
    const double u = uniform_pm1(r.x, r.y), v = uniform_pm1(r.z, r.w);
    const double t = u * u + v * v;
    if (t < 1.0 && t > 0.0) {                                                // uvGuard, GaussRandom.hs:47
      const double com = sqrt(-var2 * log(t) / t);
      g0 = u * com;
      g1 = v * com;
      return;
    }
  }
}

// realGaussians (GaussRandom.hs:52-59): y[e][0 .. n) i.i.d. N(0, var2 / 2); one thread per pair, coefficients (2p, 2p+1)
__global__ void __launch_bounds__(256)
k_real_gaussians(double* __restrict__ y, int64_t n, int64_t batch, uint64_t seed, uint64_t first, double var2)
{
  const int64_t pairs = (n + 1) / 2, total = pairs * batch;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / pairs, p = idx - e * pairs;
    double g0, g1;
    gauss_pair(seed, first + (uint64_t)e, (uint32_t)p, var2, g0, g1);
    double* dst = y + e * n + 2 * p;
    if (2 * p + 1 < n) {
      if ((n & 1) == 0) __stcs(reinterpret_cast<double2*>(dst), make_double2(g0, g1));
      else { dst[0] = g0; dst[1] = g1; }
    } else dst[0] = g0;
  }
}

template <int PA, int PB, bool GEN = false>
__global__ void __launch_bounds__(256)
k_gauss_stream(double* __restrict__ y, int64_t batch, const __grid_constant__ PlainGeom G, const __grid_constant__ GaussMats E,
               uint64_t seed = 0, uint64_t first = 0, double var2 = 0.0)
{
  constexpr int DA = PA - 1, DB = PB > 1 ? PB - 1 : 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G.tiles) return;
  const TileIndex ix(G, t, DA, DB);
  for (int64_t e = blockIdx.y; e < batch; e += gridDim.y) {
    double* base = y + (size_t)e * G.n + ix.off;
    double v[DB][DA];
    if (GEN) {
      // tGaussianDec in one pass (CPP.hs:376-389): the tile's inputs are drawn here instead of read -- 8 n bytes per
      // element move instead of 24 n (fill + in-place transform); pair (b, i) of tile t is stream (e, (t DB + b) DA/2 + i)
#pragma unroll
      for (int b = 0; b < DB; b++)
#pragma unroll
        for (int i = 0; i < DA / 2; i++)
          gauss_pair(seed, first + (uint64_t)e, (uint32_t)((t * DB + b) * (DA / 2) + i), var2, v[b][2 * i], v[b][2 * i + 1]);
    } else {
#pragma unroll
      for (int b = 0; b < DB; b++)
#pragma unroll
        for (int a = 0; a < DA; a++) v[b][a] = __ldcs(base + ix.sa * a + ix.sb * b);
    }
#pragma unroll
    for (int b = 0; b < DB; b++) gauss_line<PA>(v[b], E.a);
    if constexpr (PB > 1) {
#pragma unroll
      for (int a = 0; a < DA; a++) {
        double w[DB];
#pragma unroll
        for (int b = 0; b < DB; b++) w[b] = v[b][a];
        gauss_line<PB>(w, E.b);
#pragma unroll
        for (int b = 0; b < DB; b++) v[b][a] = w[b];
      }
    }
#pragma unroll
    for (int b = 0; b < DB; b++)
#pragma unroll
      for (int a = 0; a < DA; a++) __stcs(base + ix.sa * a + ix.sb * b, v[b][a]);
  }
}

// ------------------------------------------------------------------ the same operators with the ring elements staged in shared memory
// Indices whose odd primes do not fit a register tile ({7,13}: 72 values, {5,7,13}, {3,5,7,13}: every other ring of the reference's
// benchmark lists, which time `error` = tGaussianDec and the line operators on each of them): a CTA loads whole ring elements with
// contiguous 16-byte accesses into a shared-memory tile of ring values, runs one in-place pass per odd-prime axis (a thread per
// line of p - 1 values) and stores the same way -- one HBM round trip for any number of axes, like k_line_tile over Z_q
// (fused_stream.cu).  KIND = PASS_GAUSS applies 2 E_p / sqrt(2) per axis (random.cpp:19-50) with the matrices read as constant-bank
// operands; GEN draws the inputs into the tile instead of loading them (pair p of an element -> coefficients 2p, 2p + 1, the layout
// of k_real_gaussians, so the one-pass result equals realGaussians followed by the transform bit for bit).
struct PTileGeom {
  int32_t n, naxes, epb, threads;  // n: values per element (tupSize folded in); threads: CTA size chosen with epb
  int32_t p[4], rts[4];
  uint32_t m_rts[4];               // ceil(2^32 / rts)
};

struct GaussAll {                  // 2 E_p row-major, one slot per supported prime so that every index is a compile-time constant
  double m3[4], m5[16], m7[36], m11[100], m13[144];
};

template <int P>
__device__ __forceinline__ const double* gauss_slot(const GaussAll& E)
{
  if constexpr (P == 3) return E.m3;
  else if constexpr (P == 5) return E.m5;
  else if constexpr (P == 7) return E.m7;
  else if constexpr (P == 11) return E.m11;
  else return E.m13;
}

// LPT: lines per thread of the Gaussian pass -- 2 when the inputs are loaded (every constant fetched once per two lines, eight chains in
// flight: 45 % -> 56 % of HBM at m = 2912), 1 when they are drawn in the kernel (the draw wants the registers: 0.80 against 0.97 ms)
template <class R, int KIND, int P, int LPT>
__device__ __forceinline__ void plain_tile_axis(typename R::T* tile, const PTileGeom& G, const int ax, const int units, const GaussAll& E)
{
  typedef typename R::T T;
  constexpr int D = P - 1;
  const R ring{};
  const int rts = G.rts[ax];
  const int total = units * (G.n / D);
  if constexpr (KIND == PASS_GAUSS) {
    const double* M = gauss_slot<P>(E);
    const double inv_sqrt2 = 0.70710678118654752440;
    constexpr int RG = D % 4 == 0 ? 4 : 2;             // rows in flight per line at a bounded register count
    for (int L0 = threadIdx.x; L0 < total; L0 += LPT * blockDim.x) {
      T* base[LPT];
      bool live[LPT];
      T v[LPT][D];
#pragma unroll
      for (int t = 0; t < LPT; t++) {
        const int L = L0 + t * blockDim.x;
        live[t] = L < total;
        const int Lc = live[t] ? L : L0;
        const int uh = rts == 1 ? Lc : (int)__umulhi((uint32_t)Lc, G.m_rts[ax]), lo = Lc - uh * rts;
        base[t] = tile + (size_t)uh * rts * D + lo;
#pragma unroll
        for (int a = 0; a < D; a++) v[t][a] = base[t][a * rts];
      }
#pragma unroll 1
      for (int r0 = 0; r0 < D; r0 += RG) {             // rolled: the row group is a warp-uniform offset into the constant bank
        double acc[LPT][RG];
#pragma unroll
        for (int t = 0; t < LPT; t++)
#pragma unroll
          for (int r = 0; r < RG; r++) acc[t][r] = 0.0;
#pragma unroll
        for (int col = 0; col < D; col++)
#pragma unroll
          for (int r = 0; r < RG; r++) {
            const double m = M[(r0 + r) * D + col];
#pragma unroll
            for (int t = 0; t < LPT; t++) acc[t][r] = __dadd_rn(acc[t][r], __dmul_rn(m, v[t][col]));      // order of random.cpp:33-40
          }
#pragma unroll
        for (int t = 0; t < LPT; t++)
          if (live[t]) {
#pragma unroll
            for (int r = 0; r < RG; r++) base[t][(r0 + r) * rts] = __dmul_rn(acc[t][r], inv_sqrt2);
          }
      }
    }
  } else {
    for (int L = threadIdx.x; L < total; L += blockDim.x) {
      const int uh = rts == 1 ? L : (int)__umulhi((uint32_t)L, G.m_rts[ax]), lo = L - uh * rts;
      T* base = tile + (size_t)uh * rts * D + lo;
      T v[D];
#pragma unroll
      for (int a = 0; a < D; a++) v[a] = base[a * rts];
      line_ring<KIND, P, R>(ring, v);
#pragma unroll
      for (int a = 0; a < D; a++) base[a * rts] = v[a];
    }
  }
}

__device__ __forceinline__ int64_t wsum(int64_t v);
__device__ __forceinline__ double wsum(double v);

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Two tile buffers (NBUF = 2, everything but GEN): the next group's elements arrive by cp.async (16-byte, L2-only) while the
// passes run on the current one, so one CTA overlaps its own HBM reads with its arithmetic instead of relying on co-resident CTAs.
template <class R, int KIND, bool GEN>
__global__ void __launch_bounds__((KIND == PASS_GAUSS && !GEN) ? 192 : 256, KIND == PASS_GAUSS ? 3 : 1)      // Gaussian transform: 192-thread CTAs, three per SM (<= 113 registers)
k_plain_tile(typename R::IO* __restrict__ y, int64_t batch, const __grid_constant__ PTileGeom G, const __grid_constant__ GaussAll E,
             double rscale, uint64_t seed, uint64_t first, double var2, typename R::IO* __restrict__ out)
{
  typedef typename R::T T;
  static_assert(sizeof(T) == sizeof(typename R::IO), "the tile holds the values as stored");
  extern __shared__ __align__(16) unsigned char plain_tile_raw[];
  constexpr int PER16 = 16 / (int)sizeof(T);          // values per 16-byte word: 2 (int64, double) or 1 (complex)
  const int buf_words = G.epb * G.n / PER16;           // 16-byte words per buffer; n is even for the 8-byte rings (checked by the host)
  const int64_t ngroups = (batch + G.epb - 1) / G.epb;
  auto group_words = [&](int64_t g) { const int64_t e0 = g * G.epb; return (int)(batch - e0 < G.epb ? batch - e0 : G.epb) * (G.n / PER16); };
  auto prefetch = [&](int64_t g, int buf) {
    const double2* src = reinterpret_cast<const double2*>(y + (size_t)g * G.epb * G.n);
    double2* dst = reinterpret_cast<double2*>(plain_tile_raw) + (size_t)buf * buf_words;
    const int words = group_words(g);
    for (int i = threadIdx.x; i < words; i += blockDim.x) cp_async16(dst + i, src + i);
    cp_async_commit();
  };
  int cur = 0;
  if constexpr (!GEN) { if ((int64_t)blockIdx.x < ngroups) prefetch(blockIdx.x, 0); }
  for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const int64_t e0 = g * G.epb;
    const int cnt = (int)(batch - e0 < G.epb ? batch - e0 : G.epb);
    const int words = cnt * (G.n / PER16);
    double2* tile16 = reinterpret_cast<double2*>(plain_tile_raw) + (size_t)cur * buf_words;
    T* tile = reinterpret_cast<T*>(tile16);
    double2* dst = reinterpret_cast<double2*>(y + (size_t)e0 * G.n);
    if constexpr (GEN) {
      const int half = G.n / 2;
      for (int i = threadIdx.x; i < words; i += blockDim.x) {
        const int u = i / half, p = i - u * half;
        double g0, g1;
        gauss_pair(seed, first + (uint64_t)(e0 + u), (uint32_t)p, var2, g0, g1);
        tile16[i] = make_double2(g0, g1);
      }
      __syncthreads();
    } else {
      cp_async_wait_all();
      __syncthreads();            // the current tile is complete, and every thread is past the previous iteration's reads of the other buffer
      if (g + gridDim.x < ngroups) prefetch(g + gridDim.x, cur ^ 1);
    }
    for (int ax = 0; ax < G.naxes; ax++) {
      switch (G.p[ax]) {
        case 3: plain_tile_axis<R, KIND, 3, GEN ? 1 : 2>(tile, G, ax, cnt, E); break;
        case 5: plain_tile_axis<R, KIND, 5, GEN ? 1 : 2>(tile, G, ax, cnt, E); break;
        case 7: plain_tile_axis<R, KIND, 7, GEN ? 1 : 2>(tile, G, ax, cnt, E); break;
        case 11: plain_tile_axis<R, KIND, 11, GEN ? 1 : 2>(tile, G, ax, cnt, E); break;
        default: plain_tile_axis<R, KIND, 13, GEN ? 1 : 2>(tile, G, ax, cnt, E); break;
      }
      __syncthreads();
    }
    if constexpr (KIND == PASS_NORMSQ) {
      // norm.cpp:15-80: out[e] = sum_j y[j] ((x)(I + J) y)[j] -- the tile holds the second factor, the first is re-read (an L2 hit: the
      // element was fetched by this CTA a few microseconds ago); y is not written
      if constexpr (sizeof(T) == 8) {
        __shared__ T part[64 * 8];
        const R ring{};
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
        for (int u = 0; u < cnt; u++) {
          const T* orig = reinterpret_cast<const T*>(y) + (size_t)(e0 + u) * G.n;
          const T* second = tile + (size_t)u * G.n;
          T sacc = ring.zero();
          for (int j0 = threadIdx.x; j0 < G.n; j0 += 4 * blockDim.x) {      // four re-reads in flight per thread
            T o[4];
#pragma unroll
            for (int w = 0; w < 4; w++) { const int j = j0 + w * blockDim.x; o[w] = j < G.n ? __ldg(orig + j) : ring.zero(); }
#pragma unroll
            for (int w = 0; w < 4; w++) { const int j = j0 + w * blockDim.x; if (j < G.n) sacc = ring.add(sacc, ring.mul(o[w], second[j])); }
          }
          sacc = wsum(sacc);
          if (lane == 0) part[u * 8 + warp] = sacc;
        }
        __syncthreads();
        if (threadIdx.x < cnt) {
          T sacc = part[threadIdx.x * 8];
          for (int w = 1; w < nw; w++) sacc = ring.add(sacc, part[threadIdx.x * 8 + w]);
          out[e0 + threadIdx.x] = sacc;
        }
      }
    } else {
      for (int i = threadIdx.x; i < words; i += blockDim.x) {
        double2 v = tile16[i];
        if constexpr (sizeof(T) == 16) {                 // complex: optional real scale (g.cpp:209-220 intent), as k_line_plain
          if (rscale != 0.0) v = make_double2(__dmul_rn(v.x, rscale), __dmul_rn(v.y, rscale));
        }
        __stcs(dst + i, v);
      }
    }
    if constexpr (GEN) __syncthreads(); else cur ^= 1;
  }
}

__device__ __forceinline__ int64_t wsum(int64_t v)
{
  for (int o = 16; o > 0; o >>= 1) v = (int64_t)((uint64_t)v + (uint64_t)__shfl_down_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double wsum(double v)
{
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// one CTA per ring element (grid-stride): out[e] = sum_j y[j] * ((x)(I+J) y)[j]; fixed reduction tree => deterministic
template <class R, int PA, int PB>
__global__ void __launch_bounds__(512)
k_normsq_stream(const typename R::IO* __restrict__ y, typename R::IO* __restrict__ out, int64_t batch, const __grid_constant__ PlainGeom G)
{
  typedef typename R::T T;
  constexpr int DA = PA - 1, DB = PB > 1 ? PB - 1 : 1;
  __shared__ T red[32];
  const R ring{};
  // a thread's first tile is the same for every element: its index arithmetic (two divisions) is done once
  const TileIndex ix0(G, threadIdx.x < G.tiles ? threadIdx.x : 0, DA, DB);
  for (int64_t e = blockIdx.x; e < batch; e += gridDim.x) {
    T acc = ring.zero();
    for (int t = threadIdx.x; t < G.tiles; t += blockDim.x) {
      const TileIndex ix = t == (int)threadIdx.x ? ix0 : TileIndex(G, t, DA, DB);
      const typename R::IO* base = y + (size_t)e * G.n + ix.off;
      T v[DB][DA], o[DB][DA];
#pragma unroll
      for (int b = 0; b < DB; b++)
#pragma unroll
        for (int a = 0; a < DA; a++) { v[b][a] = __ldcs(base + ix.sa * a + ix.sb * b); o[b][a] = v[b][a]; }
#pragma unroll
      for (int b = 0; b < DB; b++) line_ring<PASS_NORMSQ, PA, R>(ring, v[b]);
      if constexpr (PB > 1) {
#pragma unroll
        for (int a = 0; a < DA; a++) {
          T w[DB];
#pragma unroll
          for (int b = 0; b < DB; b++) w[b] = v[b][a];
          line_ring<PASS_NORMSQ, PB, R>(ring, w);
#pragma unroll
          for (int b = 0; b < DB; b++) v[b][a] = w[b];
        }
      }
#pragma unroll
      for (int b = 0; b < DB; b++)
#pragma unroll
        for (int a = 0; a < DA; a++) acc = ring.add(acc, ring.mul(o[b][a], v[b][a]));
    }
    acc = wsum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      T v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : ring.zero();
      v = wsum(v);
      if (threadIdx.x == 0) out[e] = v;
    }
    __syncthreads();
  }
}

// odd prime axes of the plan as (p, rts, p^(e-1)); returns how many
int odd_axes(const lolb_plan* pl, int (&p)[4], int64_t (&rts)[4], int (&ppi)[4], int64_t (&mprime)[4])
{
  int cnt = 0;
  int64_t r = 1;
  for (size_t i = 0; i < pl->pe.size(); i++) {
    const PrimeExponent& pe = pl->pe[i];
    int64_t mp = 1;
    for (int j = 1; j < pe.exponent; j++) mp *= pe.prime;
    if (pe.prime != 2) { if (cnt < 4) { p[cnt] = pe.prime; rts[cnt] = r; ppi[cnt] = (int)i; mprime[cnt] = mp; } cnt++; }
    r *= (int64_t)(pe.prime - 1) * mp;
  }
  return cnt;
}

// supported (pA, pB) combinations; 1 = no second axis
bool combo_ok(int cnt, const int (&p)[4])
{
  if (cnt == 1) return p[0] == 3 || p[0] == 5 || p[0] == 7;
  if (cnt == 2) return p[0] == 3 && (p[1] == 5 || p[1] == 7);
  return false;
}

bool make_geom(const lolb_plan* pl, int fold_k, PlainGeom* G, int (&p)[4], int (&ppi)[4], int64_t (&mprime)[4], int* cnt_out)
{
  int64_t r[4];
  const int cnt = odd_axes(pl, p, r, ppi, mprime);
  *cnt_out = cnt;
  if (!combo_ok(cnt, p)) return false;
  const int64_t n = (int64_t)pl->n * fold_k;
  G->n = (int32_t)n;
  G->RA = (int32_t)(r[0] * fold_k);
  if (cnt == 1) {
    G->RB = (int32_t)n;
    G->M = (int32_t)(n / ((int64_t)G->RA * (p[0] - 1)));
    G->tiles = (int32_t)(n / (p[0] - 1));
  } else {
    G->RB = (int32_t)(r[1] * fold_k);
    G->M = (int32_t)((int64_t)G->RB / ((int64_t)G->RA * (p[0] - 1)));
    G->tiles = (int32_t)(n / ((p[0] - 1) * (p[1] - 1)));
  }
  return true;
}

dim3 tile_grid(const lolb_plan* pl, const PlainGeom& G, int64_t batch, int* threads_out)
{
  int threads = G.tiles >= 256 ? 256 : ((G.tiles + 31) / 32) * 32;
  for (int c = 256; c >= 128; c -= 32) if (G.tiles % c == 0) { threads = c; break; }
  dim3 grid((G.tiles + threads - 1) / threads, 1, 1);
  int64_t gy = ((int64_t)pl->num_sms * 2048 / threads + grid.x - 1) / grid.x * 2;
  if (gy > batch) gy = batch;
  if (gy > 65535) gy = 65535;
  grid.y = (unsigned)gy;
  *threads_out = threads;
  return grid;
}

int launched(const char* what)
{
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, what);
  count_launch();
  return LOLB_OK;
}

// shared-memory tile kernel: odd primes from {3, 5, 7, 11, 13}, at most four odd axes, one element within the opt-in shared memory
bool ptile_geom(const lolb_plan* pl, int fold_k, size_t tsize, int quad /* 0 line operators, 1 Gaussian drawn in the kernel, 2 Gaussian loaded */, PTileGeom* G, int (&p)[4], int (&ppi)[4], int64_t (&mp)[4])
{
  int64_t r[4];
  const int cnt = odd_axes(pl, p, r, ppi, mp);
  if (cnt < 1 || cnt > 4) return false;
  for (int i = 0; i < cnt; i++) if (p[i] != 3 && p[i] != 5 && p[i] != 7 && p[i] != 11 && p[i] != 13) return false;
  const int64_t n = (int64_t)pl->n * fold_k;
  if ((tsize == 8 && (n & 1)) || n * (int64_t)tsize > 100 * 1024) return false;
  // CTA shape: lolb_internal.cuh::choose_tile_shape; LOLB_PLAIN_TILE_BYTES / _EPB / _THREADS override for tuning runs
  static const int tile_bytes = [] { const char* e = getenv("LOLB_PLAIN_TILE_BYTES"); return e ? atoi(e) : 40960; }();
  static const int epb_env = [] { const char* e = getenv("LOLB_PLAIN_TILE_EPB"); return e ? atoi(e) : 0; }();
  static const int thr_env = [] { const char* e = getenv("LOLB_PLAIN_TILE_THREADS"); return e ? atoi(e) : 0; }();
  TileShape sh = choose_tile_shape(n, p, cnt, tsize, (size_t)tile_bytes, quad != 0, quad == 2 ? 2 : 1, quad == 2 ? 192 : 256);
  if (epb_env > 0) sh.epb = epb_env < 64 ? epb_env : 64;      // the norm's per-element partial sums are sized for 64 elements
  if (thr_env >= 32 && thr_env <= (quad == 2 ? 192 : 256)) sh.threads = thr_env / 32 * 32;
  const int64_t epb = sh.epb;
  G->threads = sh.threads;
  G->n = (int32_t)n;
  G->naxes = cnt;
  G->epb = (int32_t)epb;
  for (int i = 0; i < cnt; i++) {
    G->p[i] = p[i];
    G->rts[i] = (int32_t)(r[i] * fold_k);
    G->m_rts[i] = (uint32_t)((((uint64_t)1 << 32) + (uint64_t)G->rts[i] - 1) / (uint64_t)G->rts[i]);
  }
  return true;
}

template <class R, int KIND, bool GEN>
int launch_plain_tile(const lolb_plan* pl, const PTileGeom& G, const GaussAll& E, typename R::IO* y, int64_t batch, double rscale,
                      uint64_t seed, uint64_t first, double var2, cudaStream_t st, typename R::IO* out = nullptr)
{
  if ((uintptr_t)y & 15) return LOLB_FUSED_UNAVAILABLE;      // 16-byte accesses (cp.async, vector stores): an odd word offset goes to the generic engine
  const size_t smem = (size_t)G.epb * G.n * sizeof(typename R::T) * (GEN ? 1 : 2);      // two tile buffers: k_plain_tile prefetches the next group
  const int64_t groups = (batch + G.epb - 1) / G.epb;
  // CTAs per SM: the Gaussian transform takes everything that fits (228 KB per SM, 1 KB reserved per CTA: three CTAs of two 36 KB
  // buffers, 56 -> 60 %); the light operators measured faster with one CTA less (L at m = 2912: 86 % against 79 %)
  int per_sm = (int)((KIND == PASS_GAUSS ? 227 : 200) * 1024 / (smem + 1024));
  if (per_sm > 2048 / G.threads) per_sm = 2048 / G.threads;
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)pl->num_sms * per_sm;
  if (grid > groups) grid = groups;
  auto kern = k_plain_tile<R, KIND, GEN>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "k_plain_tile shared memory");
  }
  kern<<<(int)grid, G.threads, smem, st>>>(y, batch, G, E, rscale, seed, first, var2, out);
  return launched("k_plain_tile");
}

template <class R>
int plain_tile_kind(const lolb_plan* pl, int kind, const PTileGeom& G, typename R::IO* y, int64_t batch, double rscale, cudaStream_t st)
{
  static const GaussAll none{};
  switch (kind) {
    case PASS_L: return launch_plain_tile<R, PASS_L, false>(pl, G, none, y, batch, rscale, 0, 0, 0.0, st);
    case PASS_LINV: return launch_plain_tile<R, PASS_LINV, false>(pl, G, none, y, batch, rscale, 0, 0, 0.0, st);
    case PASS_GPOW: return launch_plain_tile<R, PASS_GPOW, false>(pl, G, none, y, batch, rscale, 0, 0, 0.0, st);
    case PASS_GDEC: return launch_plain_tile<R, PASS_GDEC, false>(pl, G, none, y, batch, rscale, 0, 0, 0.0, st);
    case PASS_GINVPOW: return launch_plain_tile<R, PASS_GINVPOW, false>(pl, G, none, y, batch, rscale, 0, 0, 0.0, st);
    case PASS_GINVDEC: return launch_plain_tile<R, PASS_GINVDEC, false>(pl, G, none, y, batch, rscale, 0, 0, 0.0, st);
    default: return LOLB_FUSED_UNAVAILABLE;
  }
}

bool plain_tile_enabled()
{
  static const bool on = [] { const char* e = getenv("LOLB_PLAIN_TILE"); return !e || atoi(e) != 0; }();
  return on;
}

template <class R, int PA, int PB>
int line_kind(const lolb_plan* pl, int kind, const PlainGeom& G, typename R::IO* y, int64_t batch, double rscale, cudaStream_t st)
{
  int threads;
  const dim3 grid = tile_grid(pl, G, batch, &threads);
  switch (kind) {
    case PASS_L: k_line_plain<R, PASS_L, PA, PB><<<grid, threads, 0, st>>>(y, batch, G, rscale); break;
    case PASS_LINV: k_line_plain<R, PASS_LINV, PA, PB><<<grid, threads, 0, st>>>(y, batch, G, rscale); break;
    case PASS_GPOW: k_line_plain<R, PASS_GPOW, PA, PB><<<grid, threads, 0, st>>>(y, batch, G, rscale); break;
    case PASS_GDEC: k_line_plain<R, PASS_GDEC, PA, PB><<<grid, threads, 0, st>>>(y, batch, G, rscale); break;
    case PASS_GINVPOW: k_line_plain<R, PASS_GINVPOW, PA, PB><<<grid, threads, 0, st>>>(y, batch, G, rscale); break;
    case PASS_GINVDEC: k_line_plain<R, PASS_GINVDEC, PA, PB><<<grid, threads, 0, st>>>(y, batch, G, rscale); break;
    default: return LOLB_FUSED_UNAVAILABLE;
  }
  return launched("k_line_plain");
}

template <class R>
int line_combo(const lolb_plan* pl, int kind, const PlainGeom& G, int cnt, const int (&p)[4], typename R::IO* y, int64_t batch, double rscale, cudaStream_t st)
{
  if (cnt == 1) {
    if (p[0] == 3) return line_kind<R, 3, 1>(pl, kind, G, y, batch, rscale, st);
    if (p[0] == 5) return line_kind<R, 5, 1>(pl, kind, G, y, batch, rscale, st);
    return line_kind<R, 7, 1>(pl, kind, G, y, batch, rscale, st);
  }
  if (p[1] == 5) return line_kind<R, 3, 5>(pl, kind, G, y, batch, rscale, st);
  return line_kind<R, 3, 7>(pl, kind, G, y, batch, rscale, st);
}

}  // namespace

// which kernel family serves the line operators (gauss = false; tupSize folded) or the Gaussian transform of a plan
const char* fused_plain_name(const lolb_plan* pl, bool gauss, bool cplx)
{
  PlainGeom G{};
  PTileGeom TG{};
  int p[4], ppi[4], cnt; int64_t mp[4];
  if (gauss && pl->k != 1) return "generic";
  if (make_geom(pl, gauss ? 1 : pl->k, &G, p, ppi, mp, &cnt)) return "plain_stream";
  if (cnt == 0) return "identity";
  if (plain_tile_enabled() && ptile_geom(pl, gauss ? 1 : pl->k, cplx ? 16 : 8, gauss ? 2 : 0, &TG, p, ppi, mp)) return "plain_tile";
  return "generic";
}

// ring: RING_I64 / RING_F64 / RING_C64.  GInv on int64 needs the per-element divisibility verdict: generic engine.
int fused_plain_line(const lolb_plan* pl, int ring, int kind, void* y, int64_t batch, double rscale, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  PlainGeom G{};
  int p[4], ppi[4], cnt; int64_t mp[4];
  if (!make_geom(pl, pl->k, &G, p, ppi, mp, &cnt)) {
    if (cnt == 0 && !(ring == RING_C64 && rscale != 0.0 && rscale != 1.0)) return LOLB_OK;     // identity for p = 2
    PTileGeom TG{};
    if (cnt == 0 || !plain_tile_enabled()) return LOLB_FUSED_UNAVAILABLE;
    if (ring == RING_I64) {
      if (kind == PASS_GINVPOW || kind == PASS_GINVDEC || !ptile_geom(pl, pl->k, 8, 0, &TG, p, ppi, mp)) return LOLB_FUSED_UNAVAILABLE;
      return plain_tile_kind<I64Ring>(pl, kind, TG, (int64_t*)y, batch, 0.0, st);
    }
    if (ring == RING_F64) {
      if (!ptile_geom(pl, pl->k, 8, 0, &TG, p, ppi, mp)) return LOLB_FUSED_UNAVAILABLE;
      return plain_tile_kind<F64Ring>(pl, kind, TG, (double*)y, batch, 0.0, st);
    }
    if (!ptile_geom(pl, pl->k, 16, 0, &TG, p, ppi, mp)) return LOLB_FUSED_UNAVAILABLE;
    return plain_tile_kind<C64Ring>(pl, kind, TG, (double2*)y, batch, rscale, st);
  }
  if (ring == RING_I64) {
    if (kind == PASS_GINVPOW || kind == PASS_GINVDEC) return LOLB_FUSED_UNAVAILABLE;
    return line_combo<I64Ring>(pl, kind, G, cnt, p, (int64_t*)y, batch, 0.0, st);
  }
  if (ring == RING_F64) return line_combo<F64Ring>(pl, kind, G, cnt, p, (double*)y, batch, 0.0, st);
  return line_combo<C64Ring>(pl, kind, G, cnt, p, (double2*)y, batch, rscale, st);
}

int fused_plain_gauss_gen(const lolb_plan* pl, double* y, int64_t batch, cudaStream_t st, bool gen, uint64_t seed, uint64_t first, double var2);

int fused_plain_gauss(const lolb_plan* pl, double* y, int64_t batch, cudaStream_t st)
{ return fused_plain_gauss_gen(pl, y, batch, st, false, 0, 0, 0.0); }

int fused_plain_real_gaussians(const lolb_plan* pl, double* y, int64_t n, int64_t batch, uint64_t seed, uint64_t first, double var2, cudaStream_t st)
{
  if (batch <= 0 || n <= 0) return LOLB_OK;
  const int64_t total = ((n + 1) / 2) * batch;
  const int sms = pl ? pl->num_sms : 148;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sms * 16);
  k_real_gaussians<<<blocks, 256, 0, st>>>(y, n, batch, seed, first, var2);
  return launched("k_real_gaussians");
}

int fused_plain_gauss_gen(const lolb_plan* pl, double* y, int64_t batch, cudaStream_t st, bool gen, uint64_t seed, uint64_t first, double var2)
{
  if (batch <= 0) return LOLB_OK;
  if (pl->k != 1) return LOLB_FUSED_UNAVAILABLE;
  PlainGeom G{};
  int p[4], ppi[4], cnt; int64_t mp[4];
  if (!make_geom(pl, 1, &G, p, ppi, mp, &cnt)) {
    if (cnt == 0) return LOLB_OK;
    PTileGeom TG{};
    if (!plain_tile_enabled() || !ptile_geom(pl, 1, 8, gen ? 1 : 2, &TG, p, ppi, mp)) return LOLB_FUSED_UNAVAILABLE;
    GaussAll A{};
    for (int ax = 0; ax < cnt; ax++) {
      const int P = p[ax];
      double* M = P == 3 ? A.m3 : P == 5 ? A.m5 : P == 7 ? A.m7 : P == 11 ? A.m11 : A.m13;
      const std::vector<lolb_complex>& T = pl->cru[ppi[ax]];
      for (int row = 0; row < P - 1; row++)
        for (int col = 1; col <= P - 1; col++) {
          const lolb_complex w = T[(size_t)(((int64_t)row * col) % P) * mp[ax]];
          M[row * (P - 1) + col - 1] = 2.0 * (col <= (P >> 1) ? w.real : w.imag);
        }
    }
    return gen ? launch_plain_tile<F64Ring, PASS_GAUSS, true>(pl, TG, A, y, batch, 0.0, seed, first, var2, st)
               : launch_plain_tile<F64Ring, PASS_GAUSS, false>(pl, TG, A, y, batch, 0.0, 0, 0, 0.0, st);
  }
  // E_p[row][col-1] = Re or Im of ru[(row*col mod p) * p^(e-1)]  (random.cpp:33-40)
  GaussMats E{};
  for (int ax = 0; ax < cnt; ax++) {
    const int P = p[ax];
    double (*M)[6] = ax == 0 ? E.a : E.b;
    const std::vector<lolb_complex>& T = pl->cru[ppi[ax]];
    for (int row = 0; row < P - 1; row++)
      for (int col = 1; col <= P - 1; col++) {
        const lolb_complex w = T[(size_t)(((int64_t)row * col) % P) * mp[ax]];
        M[row][col - 1] = 2.0 * (col <= (P >> 1) ? w.real : w.imag);
      }
  }
  int threads;
  const dim3 grid = tile_grid(pl, G, batch, &threads);
#define GS(PA, PB)                                                                                               \
  do {                                                                                                            \
    if (gen) k_gauss_stream<PA, PB, true><<<grid, threads, 0, st>>>(y, batch, G, E, seed, first, var2);           \
    else k_gauss_stream<PA, PB, false><<<grid, threads, 0, st>>>(y, batch, G, E);                                 \
  } while (0)
  if (cnt == 1) {
    if (p[0] == 3) GS(3, 1);
    else if (p[0] == 5) GS(5, 1);
    else GS(7, 1);
  } else if (p[1] == 5) GS(3, 5);
  else GS(3, 7);
#undef GS
  return launched(gen ? "k_gauss_stream<gen>" : "k_gauss_stream");
}

template <class R>
static int normsq_combo(const lolb_plan* pl, const typename R::IO* y, typename R::IO* out, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  if (pl->k != 1) return LOLB_FUSED_UNAVAILABLE;
  PlainGeom G{};
  int p[4], ppi[4], cnt; int64_t mp[4];
  if (!make_geom(pl, 1, &G, p, ppi, mp, &cnt)) {
    PTileGeom TG{};
    static const GaussAll none{};
    if (cnt == 0 || !plain_tile_enabled() || !ptile_geom(pl, 1, 8, 0, &TG, p, ppi, mp)) return LOLB_FUSED_UNAVAILABLE;
    return launch_plain_tile<R, PASS_NORMSQ, false>(pl, TG, none, const_cast<typename R::IO*>(y), batch, 0.0, 0, 0, 0.0, st, out);
  }
  int threads = G.tiles >= 512 ? 512 : ((G.tiles + 31) / 32) * 32;
  for (int c = 512; c >= 128; c -= 32) if (G.tiles % c == 0) { threads = c; break; }
  int64_t grid = (int64_t)pl->num_sms * (2048 / threads);
  if (grid > batch) grid = batch;
  if (cnt == 1) {
    if (p[0] == 3) k_normsq_stream<R, 3, 1><<<(int)grid, threads, 0, st>>>(y, out, batch, G);
    else if (p[0] == 5) k_normsq_stream<R, 5, 1><<<(int)grid, threads, 0, st>>>(y, out, batch, G);
    else k_normsq_stream<R, 7, 1><<<(int)grid, threads, 0, st>>>(y, out, batch, G);
  } else if (p[1] == 5) k_normsq_stream<R, 3, 5><<<(int)grid, threads, 0, st>>>(y, out, batch, G);
  else k_normsq_stream<R, 3, 7><<<(int)grid, threads, 0, st>>>(y, out, batch, G);
  return launched("k_normsq_stream");
}

int fused_plain_normsq_i64(const lolb_plan* pl, const int64_t* y, int64_t* out, int64_t batch, cudaStream_t st)
{ return normsq_combo<I64Ring>(pl, y, out, batch, st); }
int fused_plain_normsq_f64(const lolb_plan* pl, const double* y, double* out, int64_t batch, cudaStream_t st)
{ return normsq_combo<F64Ring>(pl, y, out, batch, st); }

}  // namespace lolb
