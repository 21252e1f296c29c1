// engine_axis.cu -- the generic CRT / CRT^-1 engine, second generation: any index m whose prime powers are small
// (phi(p^e) <= 54), Z_q and complex double, with the reference's stage list executed ONE PRIME POWER AT A TIME in
// registers.
//
// engine.cu runs every stage of the reference (crt.cpp:459-560: crtp, crtTwiddle, dftp, dftTwiddle ...) as its own pass
// over the element in shared memory: 17 passes and barriers at m = 1728, one thread per output coefficient, index
// arithmetic with run-time divisions -- 3 % of the HBM roofline.  All stages of one prime power p^e act inside axis
// lines of phi(p^e) coefficients (tensor.h:76-95), so here a thread loads one line into registers, runs ALL stages of
// that prime power on it with compile-time indices (the stage structure is a function of (p, e) only; table offsets
// come from the plan's pass list), and writes it back: one shared-memory round trip and one barrier per prime power.
// Several ring elements share a CTA so that small indices still fill it.  Same operator, same tables, exact arithmetic
// => bit-identical residues over Z_q; over complex double the summation order inside a stage is the reference's.
#include "lolb_internal.cuh"
#include "rings.cuh"

namespace lolb {

namespace {

constexpr int kAxThreads = 128;
constexpr int kAxMaxPP = 8;

__host__ __device__ constexpr int cpow(int b, int e) { return e <= 0 ? 1 : b * cpow(b, e - 1); }

struct AxisList {
  int32_t count;
  int32_t p[kAxMaxPP], e[kAxMaxPP], phi[kAxMaxPP], rts[kAxMaxPP];
  int32_t first[kAxMaxPP];    // first pass of the group in the PassList
  int32_t ruoff[kAxMaxPP];    // root table of the prime power inside the per-limb table block
};

template <class R, class W>
struct AxParams {
  typename R::IO* y;
  int64_t batch;
  int32_t n, k, epc;          // ring elements per CTA iteration
  int32_t pad;                // 1: one padding word per 32 coefficients (first axis 2^6: its stride-32 lines would otherwise share a bank)
  uint32_t magic_k, magic_n;  // ceil(2^32 / k), ceil(2^32 / n): exact quotients for operands < 2^15
  const W* tab;
  int32_t tab_stride;
  int32_t finish;
  ZqConsts zc;
  uint32_t qinv[kMaxLimbs];   // Zq: -q^-1 mod 2^32
  uint32_t mscale[kMaxLimbs]; // Zq: final scalar (mhat^-1) in Montgomery form
  double2 cscale[kMaxLimbs];
};

// Z_q for this engine: odd q < 2^28, canonical residues, CONSTANTS (roots, twiddles, final scalar) in Montgomery form
// c * 2^32 mod q, so  value x constant  is one REDC (IMAD.WIDE, IMAD, IMAD.HI) and a row of <= 13 products one REDC
struct ZqMontRing {
  typedef uint32_t T;
  typedef int64_t IO;
  uint32_t q, qinv;
  __device__ __forceinline__ T zero() const { return 0u; }
  __device__ __forceinline__ T add(T a, T b) const { const uint32_t s = a + b; return min(s, s - q); }
  __device__ __forceinline__ T sub(T a, T b) const { const uint32_t d = a - b; return min(d, d + q); }
  __device__ __forceinline__ T reduce64(uint64_t x) const      // x < q 2^32  ->  x 2^-32 mod q, canonical
  {
    const uint32_t m = (uint32_t)x * qinv;
    const uint32_t t = (uint32_t)((x + (uint64_t)m * q) >> 32);
    return min(t, t - q);
  }
  __device__ __forceinline__ T mul(T a, T w_mont) const { return reduce64((uint64_t)a * w_mont); }
  __device__ __forceinline__ T load(IO x) const
  {
    if ((uint64_t)x < (uint64_t)q) return (T)x;
    int64_t r = x % (int64_t)q;
    return (T)(r < 0 ? r + (int64_t)q : r);
  }
  __device__ __forceinline__ IO store(T v) const { return (IO)v; }
};

template <class W> __device__ __forceinline__ ZqMontRing ax_ring(const ZqMontRing*, const AxParams<ZqMontRing, W>& P, int limb) { return ZqMontRing{P.zc.q[limb], P.qinv[limb]}; }
template <class W> __device__ __forceinline__ C64Ring ax_ring(const C64Ring*, const AxParams<C64Ring, W>&, int) { return C64Ring{}; }

// I_{PHI/(D*C)} (x) A_D (x) I_C on the registers of a line; KIND: PASS_DFT / PASS_CRT / PASS_CRTINV (engine.cu pass_dense)
template <class R, int P_, int PHI, int D, int C, int KIND>
__device__ __forceinline__ void dense_regs(typename R::T (&v)[PHI], const R& ring, const typename R::T (&w)[P_])
{
  typedef typename R::T T;
#pragma unroll
  for (int blk = 0; blk < PHI / (D * C); blk++) {
#pragma unroll
    for (int rr = 0; rr < C; rr++) {
      T in[D], out[D];
#pragma unroll
      for (int col = 0; col < D; col++) in[col] = v[blk * D * C + col * C + rr];
      if (KIND == PASS_DFT && P_ == 2) {                    // crt.cpp:137-149
        out[0] = ring.add(in[0], in[1]);
        out[1] = ring.sub(in[0], in[1]);
      } else {
        if constexpr (sizeof(T) == 4) {
          // Z_q: a row is accumulated as exact 64-bit products (at most 13 terms, 13 q^2 < q 2^32 checked on the host)
          // and reduced ONCE; same residue as the reference's term-by-term reduction
#pragma unroll
          for (int row = 0; row < D; row++) {
            uint64_t acc = 0, shift = 0;
#pragma unroll
            for (int col = 0; col < D; col++) {
              const int wi = KIND == PASS_DFT ? (row * col) % P_ : KIND == PASS_CRT ? ((row + 1) * col) % P_ : (row * (col + 1)) % P_;
              acc += (uint64_t)in[col] * w[wi];
              if (KIND == PASS_CRTINV) shift += (uint64_t)in[col] * w[P_ - col - 1];
            }
            out[row] = KIND == PASS_CRTINV ? ring.sub(ring.reduce64(acc), ring.reduce64(shift)) : ring.reduce64(acc);
          }
        } else {
#pragma unroll
        for (int row = 0; row < D; row++) {
          T acc = ring.zero();
          if (KIND == PASS_DFT) {                           // sum_col in[col] * w[(row*col) % p]
#pragma unroll
            for (int col = 0; col < D; col++) acc = ring.add(acc, ring.mul(in[col], w[(row * col) % P_]));
          } else if (KIND == PASS_CRT) {                    // sum_col in[col] * w[((row+1)*col) % p]
#pragma unroll
            for (int col = 0; col < D; col++) acc = ring.add(acc, ring.mul(in[col], w[((row + 1) * col) % P_]));
          } else {                                          // crtpinv: sum_col in[col] * w[(row*(col+1)) % p] - shift
            T shift = ring.zero();
#pragma unroll
            for (int col = 0; col < D; col++) {
              acc = ring.add(acc, ring.mul(in[col], w[(row * (col + 1)) % P_]));
              shift = ring.add(shift, ring.mul(in[col], w[P_ - col - 1]));
            }
            acc = ring.sub(acc, shift);
          }
          out[row] = acc;
        }
        }
      }
#pragma unroll
      for (int row = 0; row < D; row++) v[blk * D * C + row * C + rr] = out[row];
    }
  }
}

// diagonal stage: v[a] *= table[(a / C) % DIM]; CRT = crtTwiddle (entries with i0 = 0 are 1), else dftTwiddle
// (entries with i0 = 0 or i1 = 0 are 1): those multiplications are skipped at compile time
template <class R, class W, int P_, int PHI, int DIM, int C, bool CRT>
__device__ __forceinline__ void diag_regs(typename R::T (&v)[PHI], const R& ring, const W* tw)
{
#pragma unroll
  for (int a = 0; a < PHI; a++) {
    const int pos = (a / C) % DIM;
    const bool one = CRT ? (pos / (P_ - 1) == 0) : (pos / P_ == 0 || pos % P_ == 0);
    if (!one) v[a] = ring.mul(v[a], tw[pos]);
  }
}

// one round of ppDFT / ppDFTInv on the registers (crt.cpp:476-485, 505-515)
template <class R, class W, int P_, int E_, bool INV, int RD>
__device__ __forceinline__ void axis_round(typename R::T (&v)[(P_ - 1) * cpow(P_, E_ - 1)], const R& ring, const W* tab,
                                           const Pass* ps, int& t, const typename R::T (&w)[P_])
{
  constexpr int PHI = (P_ - 1) * cpow(P_, E_ - 1);
  if constexpr (RD < E_ - 1) {
    if constexpr (!INV) {
      // dftp at local stride (p-1) p^round, then dftTwiddle of dimension p^(e-1-round)
      constexpr int C = (P_ - 1) * cpow(P_, RD), DIM = cpow(P_, E_ - 1 - RD);
      dense_regs<R, P_, PHI, P_, C, PASS_DFT>(v, ring, w); t++;
      if constexpr (DIM / P_ > 1) { diag_regs<R, W, P_, PHI, DIM, C, false>(v, ring, tab + ps[t].tab); t++; }
    } else {
      // dftTwiddle of dimension p^(round+1), then dftp at local stride (p-1) p^(e-2-round)
      constexpr int C = (P_ - 1) * cpow(P_, E_ - 2 - RD), DIM = cpow(P_, RD + 1);
      if constexpr (DIM / P_ > 1) { diag_regs<R, W, P_, PHI, DIM, C, false>(v, ring, tab + ps[t].tab); t++; }
      dense_regs<R, P_, PHI, P_, C, PASS_DFT>(v, ring, w); t++;
    }
  }
}

// all stages of prime power P_^E_ on one axis line (build_crt_dir in plan.cu emits exactly this sequence of passes;
// `ps` points at the first pass of the group, `ru` at the root table of the prime power)
template <class R, class W, int P_, int E_, bool INV>
__device__ __forceinline__ void axis_line(typename R::T* line, int stride, const R& ring, const W* tab, const W* ru, const Pass* ps)
{
  typedef typename R::T T;
  constexpr int MP = cpow(P_, E_ - 1), PHI = (P_ - 1) * MP;
  T v[PHI];
#pragma unroll
  for (int a = 0; a < PHI; a++) v[a] = line[a * stride];
  T w[P_];                                                  // the p-th roots of unity: ru[t * p^(e-1)]
#pragma unroll
  for (int i = 0; i < P_; i++) w[i] = ru[i * MP];
  int t = 0;                                                // pass cursor inside the group
  if constexpr (!INV) {
    if constexpr (P_ != 2) { dense_regs<R, P_, PHI, P_ - 1, 1, PASS_CRT>(v, ring, w); t++; }              // crtp
    if constexpr (MP > 1) { diag_regs<R, W, P_, PHI, PHI, 1, true>(v, ring, tab + ps[t].tab); t++; }       // crtTwiddle
    axis_round<R, W, P_, E_, INV, 0>(v, ring, tab, ps, t, w);
    axis_round<R, W, P_, E_, INV, 1>(v, ring, tab, ps, t, w);
    axis_round<R, W, P_, E_, INV, 2>(v, ring, tab, ps, t, w);
    axis_round<R, W, P_, E_, INV, 3>(v, ring, tab, ps, t, w);
    axis_round<R, W, P_, E_, INV, 4>(v, ring, tab, ps, t, w);
  } else {
    axis_round<R, W, P_, E_, INV, 0>(v, ring, tab, ps, t, w);
    axis_round<R, W, P_, E_, INV, 1>(v, ring, tab, ps, t, w);
    axis_round<R, W, P_, E_, INV, 2>(v, ring, tab, ps, t, w);
    axis_round<R, W, P_, E_, INV, 3>(v, ring, tab, ps, t, w);
    axis_round<R, W, P_, E_, INV, 4>(v, ring, tab, ps, t, w);
    if constexpr (MP > 1) { diag_regs<R, W, P_, PHI, PHI, 1, true>(v, ring, tab + ps[t].tab); t++; }
    if constexpr (P_ != 2) { dense_regs<R, P_, PHI, P_ - 1, 1, PASS_CRTINV>(v, ring, w); t++; }           // crtpinv
  }
#pragma unroll
  for (int a = 0; a < PHI; a++) line[a * stride] = v[a];
}

__device__ __forceinline__ int ax_pad(int j, int pad) { return pad ? j + (j >> 5) : j; }
// x / d for x < 2^15 with magic = ceil(2^32 / d) (d = 1 has no 32-bit magic)
__device__ __forceinline__ int ax_div(int x, int d, uint32_t magic) { return d == 1 ? x : (int)__umulhi((uint32_t)x, magic); }

// TIER bounds the registers: 0 = axis lines of at most 20 coefficients, 1 = adds 2^6 (32), 2 = adds 3^4 (54)
template <class R, class W, bool INV, int TIER>
__global__ void __launch_bounds__(kAxThreads)
k_engine_axis(const __grid_constant__ AxParams<R, W> P, const __grid_constant__ PassList PL, const __grid_constant__ AxisList AX)
{
  typedef typename R::T T;
  extern __shared__ __align__(16) unsigned char ax_smem[];
  T* buf = reinterpret_cast<T*>(ax_smem);                 // [slots][pad(n)], slot = element * k + limb
  const int n = P.n, k = P.k, pad = P.pad, np = ax_pad(n - 1, pad) + 1;
  const int64_t ngroups = (P.batch + P.epc - 1) / P.epc;
  for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const int64_t e0 = g * P.epc;
    const int cnt = (int)(P.batch - e0 < P.epc ? P.batch - e0 : P.epc);
    const int slots = cnt * k, total = cnt * n * k;
    typename R::IO* base = P.y + (size_t)e0 * n * k;
    // contiguous load, limbs de-interleaved: idx = (el * n + j) * k + limb
    for (int idx = threadIdx.x; idx < total; idx += kAxThreads) {
      const int tq = ax_div(idx, k, P.magic_k), limb = idx - tq * k;
      const int el = ax_div(tq, n, P.magic_n), j = tq - el * n;
      const R ring = ax_ring((const R*)nullptr, P, limb);
      buf[(el * k + limb) * np + ax_pad(j, pad)] = ring.load(base[idx]);
    }
    __syncthreads();
    for (int ax = 0; ax < AX.count; ax++) {
      const int phi = AX.phi[ax], rts = AX.rts[ax], lps = n / phi;        // lines per slot
      const Pass* ps = PL.pass + AX.first[ax];
      const int pe = AX.p[ax] * 16 + AX.e[ax];
      for (int ln = threadIdx.x; ln < slots * lps; ln += kAxThreads) {
        const int slot = ln / lps, l = ln - slot * lps;
        const int hi = l / rts, r = l - hi * rts;
        const int limb = slot % k;
        const R ring = ax_ring((const R*)nullptr, P, limb);
        const W* tab = P.tab + (size_t)limb * P.tab_stride;
        const int j0 = r + rts * phi * hi;
        // with padding every line stays linear: the host pads only when the first axis is 2^6 (lines of 32 at stride 1)
        // and every later stride is a multiple of 32
        T* line = buf + slot * np + ax_pad(j0, pad);
        const int stride = ax_pad(rts, pad);
#define AX_CASE(PP, EE) case PP * 16 + EE: axis_line<R, W, PP, EE, INV>(line, stride, ring, tab, tab + AX.ruoff[ax], ps); break;
        switch (pe) {
          AX_CASE(2, 1) AX_CASE(2, 2) AX_CASE(2, 3) AX_CASE(2, 4) AX_CASE(2, 5)
          AX_CASE(3, 1) AX_CASE(3, 2) AX_CASE(3, 3)
          AX_CASE(5, 1) AX_CASE(5, 2) AX_CASE(7, 1) AX_CASE(11, 1) AX_CASE(13, 1)
          default:
            if constexpr (TIER >= 1) { if (pe == 2 * 16 + 6) axis_line<R, W, 2, 6, INV>(line, stride, ring, tab, tab + AX.ruoff[ax], ps); }
            if constexpr (TIER >= 2) { if (pe == 3 * 16 + 4) axis_line<R, W, 3, 4, INV>(line, stride, ring, tab, tab + AX.ruoff[ax], ps); }
            break;
        }
#undef AX_CASE
      }
      __syncthreads();
    }
    for (int idx = threadIdx.x; idx < total; idx += kAxThreads) {
      const int tq = ax_div(idx, k, P.magic_k), limb = idx - tq * k;
      const int el = ax_div(tq, n, P.magic_n), j = tq - el * n;
      const R ring = ax_ring((const R*)nullptr, P, limb);
      T x = buf[(el * k + limb) * np + ax_pad(j, pad)];
      if (P.finish == FIN_SCALE) {
        if constexpr (sizeof(T) == 4) x = ring.mul(x, (T)P.mscale[limb]);
        else x = ring.mul(x, P.cscale[limb]);
      }
      base[idx] = ring.store(x);
    }
    __syncthreads();
  }
}

bool pp_supported(int p, int e)
{
  switch (p) {
    case 2: return e >= 1 && e <= 6;
    case 3: return e >= 1 && e <= 4;
    case 5: return e >= 1 && e <= 2;
    case 7: case 11: case 13: return e == 1;
  }
  return false;
}

// the pass groups of the plan's list, checked against what axis_line<> executes
bool build_axes(const lolb_plan* pl, const PassList& PL, bool inverse, AxisList* AX)
{
  const int npe = (int)pl->pe.size();
  if (npe < 1 || npe > kAxMaxPP) return false;
  int cursor = 0;
  int64_t rts = 1, ruoff = 0;
  AX->count = 0;
  for (int i = 0; i < npe; i++) {
    const int p = pl->pe[i].prime, e = pl->pe[i].exponent;
    if (!pp_supported(p, e)) return false;
    int mp = 1;
    for (int j = 1; j < e; j++) mp *= p;
    const int64_t my_ruoff = ruoff;
    ruoff += (int64_t)mp * p;
    const int phi = (p - 1) * mp;
    // number of passes build_crt_dir emits for this prime power
    int cnt = 0;
    if (p != 2) cnt++;                       // crtp / crtpinv
    if (mp > 1) cnt++;                       // crtTwiddle
    for (int round = 0; round < e - 1; round++) {
      cnt++;                                 // dftp
      int dim = 1;
      for (int j = 0; j < (inverse ? round + 1 : e - 1 - round); j++) dim *= p;
      if (dim / p > 1) cnt++;                // dftTwiddle
    }
    if (phi == 1 && cnt == 0) { rts *= phi; continue; }     // p^e = 2: identity, no passes
    if (cursor + cnt > PL.count) return false;
    const int a = AX->count++;
    AX->p[a] = p; AX->e[a] = e; AX->phi[a] = phi; AX->rts[a] = (int32_t)rts; AX->first[a] = cursor; AX->ruoff[a] = (int32_t)my_ruoff;
    cursor += cnt;
    rts *= phi;
  }
  return cursor == PL.count;
}

template <class R, class W>
int launch_axis(const lolb_plan* pl, bool inverse, typename R::IO* y, int64_t batch, const W* tab, int32_t tab_stride,
                AxParams<R, W>& P, cudaStream_t st)
{
  const PassList& PL = inverse ? pl->crt_inv : pl->crt_fwd;
  AxisList AX{};
  if (!build_axes(pl, PL, inverse, &AX)) return -1;
  const int n = pl->n, k = pl->k;
  int pad = AX.count > 0 && AX.phi[0] == 32 && AX.rts[0] == 1;
  int tier = 0;
  for (int a = 0; a < AX.count; a++) {
    if (a > 0 && (AX.rts[a] & 31) != 0) pad = 0;
    if (AX.phi[a] > 32) tier = 2; else if (AX.phi[a] > 20 && tier < 1) tier = 1;
  }
  const int np = pad ? (n - 1) + ((n - 1) >> 5) + 1 : n;
  const size_t slot_bytes = (size_t)np * sizeof(typename R::T);
  // elements per CTA: about 8192 coefficients of work, at most 40 KB of shared memory and 2^15 values (exact magic division)
  int epc = (int)(8192 / ((int64_t)n * k));
  if (epc < 1) epc = 1;
  while (epc > 1 && ((size_t)epc * k * slot_bytes > 40 * 1024 || (int64_t)epc * n * k >= 32768)) epc--;
  if ((size_t)epc * k * slot_bytes > 96 * 1024 || (int64_t)epc * n * k >= 32768) return -1;
  const size_t smem = (size_t)epc * k * slot_bytes;
  P.y = y; P.batch = batch; P.n = n; P.k = k; P.epc = epc; P.pad = pad;
  P.magic_k = (uint32_t)((((uint64_t)1 << 32) + k - 1) / k);
  P.magic_n = (uint32_t)((((uint64_t)1 << 32) + n - 1) / n);
  P.tab = tab; P.tab_stride = tab_stride;
  P.finish = inverse ? FIN_SCALE : FIN_NONE;
  auto kern = inverse ? (tier == 0 ? k_engine_axis<R, W, true, 0> : tier == 1 ? k_engine_axis<R, W, true, 1> : k_engine_axis<R, W, true, 2>)
                      : (tier == 0 ? k_engine_axis<R, W, false, 0> : tier == 1 ? k_engine_axis<R, W, false, 1> : k_engine_axis<R, W, false, 2>);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(k_engine_axis)");
  }
  int per_sm = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kAxThreads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const int64_t ngroups = (batch + epc - 1) / epc;
  int64_t grid = (int64_t)pl->num_sms * per_sm;
  if (grid > ngroups) grid = ngroups;
  kern<<<(int)grid, kAxThreads, smem, st>>>(P, PL, AX);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_engine_axis");
  count_launch();
  return LOLB_OK;
}

}  // namespace

// -1: shape not supported (caller falls back to engine.cu); otherwise a LOLB status
int engine_axis_crt_zq(const lolb_plan* pl, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  const uint32_t* tabm = inverse ? pl->d_tab_inv_m : pl->d_tab_fwd_m;
  if (!tabm) return -1;                                  // an even modulus or one >= 2^28: engine.cu
  AxParams<ZqMontRing, uint32_t> P{};
  P.zc = inverse ? pl->zq_mhat : pl->zq_plain;
  for (int t = 0; t < pl->k; t++) {
    const uint32_t q = (uint32_t)pl->qs[t];
    uint32_t inv = q;
    for (int i = 0; i < 5; i++) inv *= 2u - q * inv;
    P.qinv[t] = 0u - inv;
    P.mscale[t] = (uint32_t)((((uint64_t)P.zc.scale[t]) << 32) % q);
  }
  return launch_axis<ZqMontRing, uint32_t>(pl, inverse, y, batch, tabm, inverse ? pl->tab_stride_inv : pl->tab_stride_fwd, P, st);
}

int engine_axis_crt_c(const lolb_plan* pl, bool inverse, double2* y, int64_t batch, cudaStream_t st)
{
  if (batch <= 0) return LOLB_OK;
  AxParams<C64Ring, double2> P{};
  for (int i = 0; i < pl->k; i++) P.cscale[i] = pl->c_mhatinv[i];
  return launch_axis<C64Ring, double2>(pl, inverse, y, batch, inverse ? pl->d_ctab_inv : pl->d_ctab_fwd,
                                       inverse ? pl->ctab_stride_inv : pl->ctab_stride_fwd, P, st);
}

bool engine_axis_supported(const lolb_plan* pl, bool inverse)
{
  AxisList AX{};
  const PassList& PL = inverse ? pl->crt_inv : pl->crt_fwd;
  if (!(inverse ? pl->has_inv : pl->has_fwd)) return false;
  if (!build_axes(pl, PL, inverse, &AX)) return false;
  if (pl->kind == PLAN_RQ && !(inverse ? pl->d_tab_inv_m : pl->d_tab_fwd_m)) return false;
  const int64_t nk = (int64_t)pl->n * pl->k;
  const size_t bytes = (size_t)((pl->n - 1) + ((pl->n - 1) >> 5) + 1) * pl->k * (pl->kind == PLAN_C ? 16 : 4);
  return nk < 32768 && bytes <= 96 * 1024;
}

}  // namespace lolb
