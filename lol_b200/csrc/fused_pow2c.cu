// fused_pow2c.cu -- tensorCRTC / tensorCRTInvC (crt.cpp:583-598 over Complex{double,double}, types.h:122-164) for m = 2^e, tupSize 1.
//
// Operator (the same map as the Z_q kernels, pow2_common.cuh; pinned on the CPU by test_pow2_crt_is_negacyclic_evaluation): the
// reference's  crtTwiddle ; {dftp ; dftTwiddle} x (e-1)  evaluates  f(x) = sum_i y[i] x^rev(i)  at  psi^(2 pos + 1), psi = ru[0][1].
// Twist-free Cooley-Tukey form, rounds r = 0 .. e-2 on pairs (pos, pos + 2^r), p = pos mod 2^r:
//     forward   (u, t) -> (u + T t, u - T t),   T = ru[(2p+1) n / 2^(r+1)]        rounds ascending
//     inverse   (u, t) -> (u + t, (u - t) T'),  T' = ruinv[(2p+1) n / 2^(r+1)]    rounds descending, then * mhat^-1
// Over C a different evaluation order changes the rounding only: parity with the reference is to 1e-9 relative (observed 1e-15).
//
// Schedule: the elements of a group live in a (padded) shared-memory tile of complex doubles; up to three rounds per pass (a thread owns the
// 2^S values pos + j 2^r: the rounds in registers, one barrier per pass; passes of three rounds, the remainder in twos); the next
// group arrives by cp.async while the passes run (two buffers) when two buffers fit, else one buffer with plain loads.  The n - 1
// twiddles of a direction are one table in global memory (entry 2^r - 1 + p), one read per thread and pass.  32 n bytes of HBM
// traffic per element, one read and one write.  Larger elements than the shared memory holds (n > 8192) and tupSize > 1 stay on the
// generic engines.
#include <vector>

#include "fused.cuh"

namespace lolb {

namespace {

struct Pow2C {
  bool ok_fwd = false, ok_inv = false;
  int e = 0;                       // m = 2^e, n = 2^(e-1)
  double2* d_tw = nullptr;         // [2][n]: forward table then inverse table, entry (2^r - 1) + p
  double2 rot1[2], rot2[2];        // per direction: table[n/2] and table[n/4] (root^(n/2) = +-i, root^(n/4))
};

struct Pow2CGeom {
  int32_t n, rounds, epb, nbuf;
  int32_t npass;
  int8_t pr[8], ps[8];             // pass i of the forward transform: rounds pr[i] .. pr[i] + ps[i] - 1 (ps = 1, 2 or 3); the inverse runs them backwards
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// tile index of value i: one 16-byte slot of padding after every eight values, so that the stride-4 quads of the first pass (a thread
// owns four consecutive values = 64 bytes) fall in different bank groups for the eight threads of a quarter warp (4-way conflicts before)
__device__ __host__ __forceinline__ int pad8(int i) { return i + (i >> 3); }

// one round on the pair (u, t)
template <bool INV>
__device__ __forceinline__ void bfly(double2& u, double2& t, const double2 T)
{
  if (!INV) {
    const double2 w = cmul(t, T);
    t = csub(u, w);
    u = cadd(u, w);
  } else {
    const double2 d = csub(u, t);
    u = cadd(u, t);
    t = cmul(d, T);
  }
}

// S rounds (r .. r + S - 1) on the 2^S values pos + j 2^r of a thread, bits r .. r + S - 1 of pos clear; value j sits at position
// pos + j 2^r, so in round r + s its twiddle index is (pos + j 2^r) mod 2^(r+s) = p + (j mod 2^s) 2^r.
// One table read per thread and pass: with W = T_{r+S-1}[p] the other twiddles follow from the table's own structure,
//   T_{rho-1}[p] = T_rho[p]^2   and   T_{r+s}[p + j 2^r] = T_{r+s}[p] * root^(j n / 2^s),
// where root^(n/2) (`rot1`) and root^(n/4) (`rot2`) are read from the caller's table on the host -- the L1 / shared-memory pipe
// is this kernel's limiter (ncu: 78 % busy with seven 16-byte table reads per eight values), the FP64 pipe is not (25 %).
template <bool INV, int S>
__device__ __forceinline__ void pow2c_pass(double2* x, const int vals, const int r, const double2* __restrict__ tw, const double2 rot1,
                                           const double2 rot2)
{
  constexpr int V = 1 << S;
  const int st = 1 << r;
  const int groups = vals >> S;
  for (int i = threadIdx.x; i < groups; i += blockDim.x) {
    const int p = i & (st - 1);
    const int pos = ((i >> r) << (r + S)) | p;
    int ix[V];
    double2 v[V];
#pragma unroll
    for (int j = 0; j < V; j++) { ix[j] = pad8(pos + j * st); v[j] = x[ix[j]]; }
    // W[s][j] = T_{r+s}[p + j 2^r], j < 2^s
    double2 W[S][V / 2];
    W[S - 1][0] = __ldg(tw + ((st << (S - 1)) - 1) + p);
#pragma unroll
    for (int s = S - 2; s >= 0; s--) W[s][0] = cmul(W[s + 1][0], W[s + 1][0]);
    if constexpr (S >= 2) W[1][1] = cmul(W[1][0], rot1);
    if constexpr (S >= 3) {
      W[2][1] = cmul(W[2][0], rot2);
      W[2][2] = cmul(W[2][0], rot1);
      W[2][3] = cmul(W[2][1], rot1);
    }
#pragma unroll
    for (int ss = 0; ss < S; ss++) {
      const int s = INV ? S - 1 - ss : ss;
      const int half = 1 << s;
#pragma unroll
      for (int a = 0; a < V; a++)
        if (!(a & half)) bfly<INV>(v[a], v[a + half], W[s][a & (half - 1)]);
    }
#pragma unroll
    for (int j = 0; j < V; j++) x[ix[j]] = v[j];
  }
}

template <bool INV>
__global__ void __launch_bounds__(512)
k_pow2c(double2* __restrict__ y, int64_t batch, const __grid_constant__ Pow2CGeom G, const double2* __restrict__ tw, double2 scale,
        double2 rot1, double2 rot2)
{
  extern __shared__ __align__(16) unsigned char pow2c_raw[];
  const int n = G.n;
  const int buf_vals = pad8(G.epb * n) + 1;      // padded values per buffer
  const int64_t ngroups = (batch + G.epb - 1) / G.epb;
  auto prefetch = [&](int64_t g, int buf) {
    const int64_t e0 = g * G.epb;
    const int vals = (int)(batch - e0 < G.epb ? batch - e0 : G.epb) * n;
    const double2* src = y + (size_t)e0 * n;
    double2* dst = reinterpret_cast<double2*>(pow2c_raw) + (size_t)buf * buf_vals;
    for (int i = threadIdx.x; i < vals; i += blockDim.x) {
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst + pad8(i));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + i) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int cur = 0;
  if (G.nbuf == 2 && (int64_t)blockIdx.x < ngroups) prefetch(blockIdx.x, 0);
  for (int64_t g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const int64_t e0 = g * G.epb;
    const int cnt = (int)(batch - e0 < G.epb ? batch - e0 : G.epb);
    const int vals = cnt * n;
    double2* x = reinterpret_cast<double2*>(pow2c_raw) + (size_t)cur * buf_vals;
    double2* dst = y + (size_t)e0 * n;
    if (G.nbuf == 2) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();            // the current tile is complete, and every thread is past the previous iteration's reads of the other buffer
      if (g + gridDim.x < ngroups) prefetch(g + gridDim.x, cur ^ 1);
    } else {
      for (int i0 = threadIdx.x; i0 < vals; i0 += 4 * blockDim.x) {      // four 16-byte loads in flight per thread
        double2 r[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { const int i = i0 + u * blockDim.x; r[u] = i < vals ? __ldcs(dst + i) : make_double2(0.0, 0.0); }
#pragma unroll
        for (int u = 0; u < 4; u++) { const int i = i0 + u * blockDim.x; if (i < vals) x[pad8(i)] = r[u]; }
      }
      __syncthreads();
    }
    // ---- passes of up to three rounds; forward ascending, inverse descending
    for (int q = 0; q < G.npass; q++) {
      const int idx = INV ? G.npass - 1 - q : q;
      const int r = G.pr[idx];
      switch (G.ps[idx]) {
        case 3: pow2c_pass<INV, 3>(x, vals, r, tw, rot1, rot2); break;
        case 2: pow2c_pass<INV, 2>(x, vals, r, tw, rot1, rot2); break;
        default: pow2c_pass<INV, 1>(x, vals, r, tw, rot1, rot2); break;
      }
      __syncthreads();
    }
    for (int i = threadIdx.x; i < vals; i += blockDim.x) {
      double2 v = x[pad8(i)];
      if (INV) v = cmul(v, scale);
      __stcs(dst + i, v);
    }
    if (G.nbuf == 2) cur ^= 1; else __syncthreads();
  }
}

bool shape_ok(const lolb_plan* pl)
{
  return pl->kind == PLAN_C && pl->k == 1 && pl->pe.size() == 1 && pl->pe[0].prime == 2 && pl->pe[0].exponent >= 3 && pl->pe[0].exponent <= 14;
}

}  // namespace

int fused_pow2c_select(lolb_plan* pl, void** slot)
{
  if (!shape_ok(pl)) return LOLB_OK;
  Pow2C* F = (Pow2C*)*slot;
  if (!F) { F = new Pow2C(); *slot = F; }
  const int e = pl->pe[0].exponent, n = 1 << (e - 1);
  const size_t m = (size_t)1 << e;
  F->e = e;
  F->ok_fwd = pl->has_fwd && pl->cru.size() == 1 && pl->cru[0].size() == m;
  F->ok_inv = pl->has_inv && pl->cruinv.size() == 1 && pl->cruinv[0].size() == m;
  std::vector<double2> tw((size_t)2 * n, make_double2(1.0, 0.0));
  for (int dir = 0; dir < 2; dir++) {
    if (!(dir ? F->ok_inv : F->ok_fwd)) continue;
    const std::vector<lolb_complex>& T = dir ? pl->cruinv[0] : pl->cru[0];
    F->rot1[dir] = make_double2(T[(size_t)n / 2].real, T[(size_t)n / 2].imag);
    F->rot2[dir] = make_double2(T[(size_t)n / 4].real, T[(size_t)n / 4].imag);
    for (int r = 0; r < e - 1; r++)
      for (int p = 0; p < (1 << r); p++) {
        const lolb_complex w = T[(size_t)(2 * p + 1) * (size_t)(n >> (r + 1))];
        tw[(size_t)dir * n + ((size_t)1 << r) - 1 + p] = make_double2(w.real, w.imag);
      }
  }
  if (F->d_tw) { cudaFree(F->d_tw); F->d_tw = nullptr; }
  LOLB_CUDA(cudaMalloc((void**)&F->d_tw, tw.size() * sizeof(double2)));
  LOLB_CUDA(cudaMemcpy(F->d_tw, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice));
  return LOLB_OK;
}

void fused_pow2c_release(void* slot)
{
  Pow2C* F = (Pow2C*)slot;
  if (!F) return;
  if (F->d_tw) cudaFree(F->d_tw);
  delete F;
}

bool fused_pow2c_available(const void* slot, bool inverse)
{
  const Pow2C* F = (const Pow2C*)slot;
  return F && (inverse ? F->ok_inv : F->ok_fwd);
}

int fused_pow2c_crt(const lolb_plan* pl, const void* slot, bool inverse, double2* y, int64_t batch, cudaStream_t st)
{
  const Pow2C* F = (const Pow2C*)slot;
  if (!fused_pow2c_available(slot, inverse) || ((uintptr_t)y & 15)) return LOLB_FUSED_UNAVAILABLE;
  if (batch <= 0) return LOLB_OK;
  Pow2CGeom G{};
  G.n = 1 << (F->e - 1);
  G.rounds = F->e - 1;
  {      // rounds in passes of three, the remainder as passes of two (R = 3a: 3..3; 3a + 2: 3..3 2; 3a + 1: 3..3 2 2; R = 2: 2; R = 4: 2 2)
    int R = G.rounds, r = 0, np = 0;
    int threes = R / 3;
    const int rem = R % 3;
    if (rem == 1 && threes > 0) threes--;
    for (int i = 0; i < threes; i++) { G.pr[np] = (int8_t)r; G.ps[np++] = 3; r += 3; }
    while (R - r >= 2) { G.pr[np] = (int8_t)r; G.ps[np++] = 2; r += 2; }
    if (R - r == 1) { G.pr[np] = (int8_t)r; G.ps[np++] = 1; r += 1; }
    G.npass = np;
  }
  const size_t el = (size_t)G.n * sizeof(double2);
  G.nbuf = 2 * el <= 80 * 1024 ? 2 : 1;                         // n <= 2048: the next group prefetched into a second buffer
  int64_t epb = (int64_t)(16 * 1024 / el);                      // small rings: several elements per CTA (at least 1024 values)
  if (epb < 1) epb = 1;
  G.epb = (int32_t)epb;
  const size_t smem = (size_t)(pad8((int)(G.epb * G.n)) + 1) * sizeof(double2) * G.nbuf;
  const int64_t groups = (batch + G.epb - 1) / G.epb;
  const int threads = F->e >= 13 ? 512 : 256;      // one or two resident CTAs for the large elements: more threads in each
  int per_sm = (int)(200 * 1024 / (smem + 1024));
  if (per_sm > 2048 / threads) per_sm = 2048 / threads;
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)pl->num_sms * per_sm;
  if (grid > groups) grid = groups;
  const double2 scale = inverse ? pl->c_mhatinv[0] : make_double2(1.0, 0.0);
  const double2* tw = F->d_tw + (inverse ? G.n : 0);
  cudaError_t e = cudaSuccess;
  auto go = [&](auto kern) {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) kern<<<(int)grid, threads, smem, st>>>(y, batch, G, tw, scale, F->rot1[inverse ? 1 : 0], F->rot2[inverse ? 1 : 0]);
  };
  if (inverse) go(k_pow2c<true>); else go(k_pow2c<false>);
  if (e != cudaSuccess) return cuda_fail(e, "k_pow2c shared memory");
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_pow2c");
  count_launch();
  return LOLB_OK;
}

}  // namespace lolb
