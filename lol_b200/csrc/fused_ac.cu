// fused_ac.cu -- fused complex-double CRT / CRT^-1 (tensorCRTC / tensorCRTInvC, crt.cpp:583-598) for
// m = 2^6 * 3^2 * 5^2 = 14400: the schedule of fused_a.cu (one HBM read and one HBM write per ring element, 5^2 axis in
// registers, 3^2 axis in registers, 2^6 axis across the 32 lanes with the register-exchange network) over
// Complex{double,double} (types.h:122-164) instead of Z_q.
//
// Same operator as the reference: CRT_m = CRT_64 (x) CRT_9 (x) CRT_25, CRT_{p^e} = (DFT_{p^(e-1)} (x) I_{p-1}) . That .
// (I_{p^(e-1)} (x) CRT_p), roots cis(2 pi j / p^e) from the plan (CRTrans.hs:88-95 or the caller's tables), inverse
// scaled by mhat^-1 (crt.cpp:592-597).  The evaluation order differs from the reference's stage order and uses fused
// multiply-adds, so results agree to rounding (tests: 1e-9 relative, observed ~1e-15), not bit for bit; the generic
// pass engine keeps the reference's order and is the cross-check.
//
// Bound: FP64.  ~60 double-precision instructions per coefficient against 32 bytes of HBM traffic.
#include <cstdlib>

#include "fused.cuh"
#include "numtheory.h"

namespace lolb {

namespace {

constexpr int kN = 3840, kD3 = 20;
constexpr int kLaneRows = 20;      // rows 0-7: 2^6-axis twiddles; rows 8-19: forward m3[i0][r][c] * crtTwiddle_64(lane)

struct FusedACConsts {
  double2 m5[5][4][4];      // fwd: (twiddle . CRT_5) per block i0;   inv: (CRT_5^-1' . twiddle) * mhat^-1
  double2 d5[5][5];         // DFT_5 over the block index
  double2 m3[3][2][2];
  double2 d3[3][3];
  const double2* lane_tw;   // device [kLaneRows][32]
};

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cmad(double2 acc, double2 a, double2 b)
{
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, acc.x)), fma(a.x, b.y, fma(a.y, b.x, acc.y)));
}
__device__ __forceinline__ double2 shfl_xor2(double2 v, int mask)
{
  return make_double2(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}
__device__ __forceinline__ double2 sel(bool p, double2 a, double2 b) { return make_double2(p ? a.x : b.x, p ? a.y : b.y); }

// CRT_25 / CRT_25^-1 on the 20 values of one (i1,i2) column, v[4*i0 + c]
template <bool INV>
__device__ __forceinline__ void blocks5(double2 (&v)[20], const FusedACConsts& C)
{
#pragma unroll
  for (int i0 = 0; i0 < 5; i0++) {
    double2 o[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      double2 acc = cmul(C.m5[i0][r][0], v[4 * i0]);
#pragma unroll
      for (int c = 1; c < 4; c++) acc = cmad(acc, C.m5[i0][r][c], v[4 * i0 + c]);
      o[r] = acc;
    }
#pragma unroll
    for (int r = 0; r < 4; r++) v[4 * i0 + r] = o[r];
  }
}

template <bool INV>
__device__ __forceinline__ void axis5(double2 (&v)[20], const FusedACConsts& C)
{
  if (!INV) blocks5<INV>(v, C);
#pragma unroll
  for (int c = 0; c < 4; c++) {            // DFT_5 across the block index for residue column c
    double2 o[5];
    o[0] = cadd(cadd(cadd(v[c], v[4 + c]), cadd(v[8 + c], v[12 + c])), v[16 + c]);
#pragma unroll
    for (int row = 1; row < 5; row++) {
      double2 acc = v[c];
#pragma unroll
      for (int col = 1; col < 5; col++) acc = cmad(acc, C.d5[row][col], v[4 * col + c]);
      o[row] = acc;
    }
#pragma unroll
    for (int row = 0; row < 5; row++) v[4 * row + c] = o[row];
  }
  if (INV) blocks5<INV>(v, C);
}

// CRT_9 / CRT_9^-1 on the 6 values x[2*i0 + c] of one (i3, i1).  Forward: the 2x2 blocks use the PER-LANE constants
// m3l = m3 * crtTwiddle_64(column = lane) (the diagonal twiddle of the 2^6 axis commutes with the 3^2 axis).
template <bool INV>
__device__ __forceinline__ void axis3(double2 (&x)[6], const FusedACConsts& C, const double2* m3l /* shared, stride 32 */)
{
  if (!INV) {
#pragma unroll
    for (int i0 = 0; i0 < 3; i0++) {
      const double2 a = x[2 * i0], b = x[2 * i0 + 1];
      x[2 * i0] = cmad(cmul(m3l[(4 * i0) * 32], a), m3l[(4 * i0 + 1) * 32], b);
      x[2 * i0 + 1] = cmad(cmul(m3l[(4 * i0 + 2) * 32], a), m3l[(4 * i0 + 3) * 32], b);
    }
  }
#pragma unroll
  for (int c = 0; c < 2; c++) {
    const double2 a = x[c], b = x[2 + c], d = x[4 + c];
    x[c] = cadd(cadd(a, b), d);
    x[2 + c] = cmad(cmad(a, C.d3[1][1], b), C.d3[1][2], d);
    x[4 + c] = cmad(cmad(a, C.d3[2][1], b), C.d3[2][2], d);
  }
  if (INV) {
#pragma unroll
    for (int i0 = 0; i0 < 3; i0++) {
      const double2 a = x[2 * i0], b = x[2 * i0 + 1];
      x[2 * i0] = cmad(cmul(C.m3[i0][0][0], a), C.m3[i0][0][1], b);
      x[2 * i0 + 1] = cmad(cmul(C.m3[i0][1][0], a), C.m3[i0][1][1], b);
    }
  }
}

// One radix-2 round of the exchange network on lane bit `bit` (see fused_a.cu).  Forward: (u,t) -> (u+t, (u-t)*tw), the
// sign of (u-t) for lanes that hold (t,u) lives in the lane's twiddle.  Inverse: (u,t) -> (u+t*tw, u-t*tw).
template <bool INV, bool TRIVIAL>
__device__ __forceinline__ void exchange_round(double2 (&c0)[3], double2 (&c1)[3], int lane, int bit, double2 tw)
{
  const bool hi = (lane >> bit) & 1;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const double2 send = sel(hi, c0[j], c1[j]);
    const double2 keep = sel(hi, c1[j], c0[j]);
    const double2 recv = shfl_xor2(send, 1 << bit);
    if (TRIVIAL) {
      const double2 u = sel(hi, recv, keep), t = sel(hi, keep, recv);
      c0[j] = cadd(u, t);
      c1[j] = csub(u, t);
    } else if (!INV) {
      c0[j] = cadd(keep, recv);
      c1[j] = cmul(tw, csub(keep, recv));
    } else {
      const double2 t = cmul(tw, sel(hi, keep, recv));
      const double2 u = sel(hi, recv, keep);
      c0[j] = cadd(u, t);
      c1[j] = csub(u, t);
    }
  }
}

// K = compile-time tupSize (1: immediate address offsets) or 0 for a run-time k
template <bool INV, int K, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
k_fused_ac(double2* __restrict__ y, int64_t batch, int k_rt, int limb, const __grid_constant__ FusedACConsts C)
{
  const int k = K ? K : k_rt;
  extern __shared__ __align__(16) double2 tile[];          // [kN] element tile, then [kLaneRows][32] per-lane constants
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the per-lane constants live in shared memory (19 complex values per lane would otherwise be pinned in registers
  // across phase 1, whose 20-value column already needs 80)
  double2* ltab = tile + kN;
  for (int i = threadIdx.x; i < kLaneRows * 32; i += blockDim.x) ltab[i] = C.lane_tw[i];
  __syncthreads();

  for (int64_t e = blockIdx.x; e < batch; e += gridDim.x) {
    double2* base = y + ((size_t)e * kN) * k + limb;
    // ---------------- phase 1: 5^2 axis; warp-task = i2, lane = i1
    for (int t = warp; t < 6; t += WARPS) {
      const int col = t * 32 + lane;
      double2 v[20];
#pragma unroll
      for (int a = 0; a < 20; a++) v[a] = __ldcs(base + (size_t)(col + a * 192) * k);
      axis5<INV>(v, C);
#pragma unroll
      for (int a = 0; a < 20; a++) tile[col + a * 192] = v[a];
    }
    __syncthreads();
    // ---------------- phase 2: 3^2 axis in the thread, 2^6 axis across the warp; warp-task = i3
    for (int i3 = warp; i3 < kD3; i3 += WARPS) {
      double2 x[6], c0[3], c1[3];
      const double2* ltw = ltab + lane;                    // row i of the per-lane constants: ltw[i * 32]
#pragma unroll
      for (int i2 = 0; i2 < 6; i2++) x[i2] = tile[i3 * 192 + i2 * 32 + lane];
      axis3<INV>(x, C, ltw + 8 * 32);
#pragma unroll
      for (int j = 0; j < 3; j++) { c0[j] = x[2 * j]; c1[j] = x[2 * j + 1]; }
      if (!INV) {
#pragma unroll
        for (int r = 0; r < 4; r++) exchange_round<false, false>(c0, c1, lane, r, ltw[(1 + r) * 32]);
        exchange_round<false, true>(c0, c1, lane, 4, make_double2(1.0, 0.0));
        // lane owns rows 2j + (lane&1), columns (lane>>1) and (lane>>1)+16
        double2* out = base + (size_t)(i3 * 192 + (lane & 1) * 32 + (lane >> 1)) * k;
#pragma unroll
        for (int j = 0; j < 3; j++) {
          __stcs(out + (size_t)(j * 64) * k, c0[j]);
          __stcs(out + (size_t)(j * 64 + 16) * k, c1[j]);
        }
      } else {
        exchange_round<true, true>(c0, c1, lane, 4, make_double2(1.0, 0.0));
#pragma unroll
        for (int r = 3; r >= 0; r--) exchange_round<true, false>(c0, c1, lane, r, ltw[r * 32]);
        // lane owns rows 2j + (lane>>4), columns 2*(lane&15) and 2*(lane&15)+1; crtTwiddle with inverse roots last
        double2* out = base + (size_t)(i3 * 192 + (lane >> 4) * 32 + 2 * (lane & 15)) * k;
#pragma unroll
        for (int j = 0; j < 3; j++) {
          __stcs(out + (size_t)(j * 64) * k, cmul(ltw[5 * 32], c0[j]));
          __stcs(out + (size_t)(j * 64 + 1) * k, cmul(ltw[6 * 32], c1[j]));
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ host: constants from the plan's root tables

typedef struct { double re, im; } cplx;
inline cplx cx(double re, double im) { cplx z; z.re = re; z.im = im; return z; }
inline cplx operator*(cplx a, cplx b) { return cx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
inline cplx operator-(cplx a, cplx b) { return cx(a.re - b.re, a.im - b.im); }
inline double2 d2(cplx a) { return make_double2(a.re, a.im); }

struct FusedAC {
  bool ok_fwd = false, ok_inv = false;
  std::vector<FusedACConsts> fwd, inv;     // per limb
  double2* d_lane_tw = nullptr;            // [k][2][kLaneRows][32]
};

void build_consts(const lolb_plan* pl, bool inverse, int limb, FusedACConsts* C, double2* lane_tw)
{
  const int k = pl->k;
  const auto& T = inverse ? pl->cruinv : pl->cru;
  auto rdc = [&](int i, int64_t j) { const lolb_complex c = T[i][(size_t)j * k + limb]; return cx(c.real, c.imag); };
  auto r64 = [&](int64_t j) { return rdc(0, j % 64); };
  auto r9 = [&](int64_t j) { return rdc(1, j % 9); };
  auto r25 = [&](int64_t j) { return rdc(2, j % 25); };
  const cplx one = cx(1.0, 0.0);
  const cplx scale = inverse ? cx(pl->c_mhatinv[limb].x, pl->c_mhatinv[limb].y) : one;
  for (int row = 0; row < 5; row++) for (int col = 0; col < 5; col++) C->d5[row][col] = d2(r25(5 * ((row * col) % 5)));
  for (int row = 0; row < 3; row++) for (int col = 0; col < 3; col++) C->d3[row][col] = d2(r9(3 * ((row * col) % 3)));
  cplx m3[3][2][2];
  for (int i0 = 0; i0 < 5; i0++)
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) {
        cplx v;
        if (!inverse) {   // crtTwiddle(i0, r) * CRT_5[r][c]   (crt.cpp:60-79, 272-295)
          const cplx tw = i0 ? r25((int64_t)i0 * (r + 1)) : one;
          v = tw * r25(5 * (((r + 1) * c) % 5));
        } else {          // (w^-r(c+1) - w^(c+1)) * crtTwiddle(i0, c) * mhat^-1   (crt.cpp:376-399)
          const cplx tw = i0 ? r25((int64_t)i0 * (c + 1)) : one;
          const cplx mat = r25(5 * ((r * (c + 1)) % 5)) - r25(5 * (5 - c - 1));
          v = tw * mat * scale;
        }
        C->m5[i0][r][c] = d2(v);
      }
  for (int i0 = 0; i0 < 3; i0++)
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 2; c++) {
        cplx v;
        if (!inverse) {
          const cplx tw = i0 ? r9((int64_t)i0 * (r + 1)) : one;
          v = tw * r9(3 * (((r + 1) * c) % 3));
        } else {
          const cplx tw = i0 ? r9((int64_t)i0 * (c + 1)) : one;
          const cplx mat = r9(3 * ((r * (c + 1)) % 3)) - r9(3 * (3 - c - 1));
          v = tw * mat;
        }
        m3[i0][r][c] = v;
        C->m3[i0][r][c] = d2(v);
      }
  // 2^6 axis, rows as in fused_a.cu.  Forward: [0] unused (trivial round), [1+r] round r, [8..19] m3 * crtTwiddle(lane).
  // Inverse: [r] round r, [5],[6] crtTwiddle of columns 2*(lane&15) + {0,1}.
  for (int i = 0; i < kLaneRows * 32; i++) lane_tw[i] = make_double2(1.0, 0.0);
  for (int lane = 0; lane < 32; lane++) {
    if (!inverse) {
      const cplx tw0 = lane ? r64(digit_rev(2, 5, lane)) : one;           // crtTwiddle of 2^6 (crt.cpp:43-58), column = lane
      for (int i0 = 0; i0 < 3; i0++)
        for (int r = 0; r < 2; r++)
          for (int c = 0; c < 2; c++) lane_tw[(8 + 4 * i0 + 2 * r + c) * 32 + lane] = d2(m3[i0][r][c] * tw0);
      for (int r = 0; r < 5; r++) {
        const int i0 = lane >> (r + 1);
        const cplx tw = i0 ? r64(digit_rev(2, 4 - r, i0) * (2 << r)) : one;
        // lanes whose bit r is set hold (t, u) instead of (u, t): they multiply (t - u) by -tw
        lane_tw[(1 + r) * 32 + lane] = ((lane >> r) & 1) ? make_double2(-tw.re, -tw.im) : d2(tw);
      }
    } else {
      for (int r = 0; r < 5; r++) {
        const int i0 = (lane >> r) & ((1 << (4 - r)) - 1);
        lane_tw[r * 32 + lane] = i0 ? d2(r64(digit_rev(2, 4 - r, i0) * (2 << r))) : make_double2(1.0, 0.0);
      }
      for (int s = 0; s < 2; s++) {
        const int col = 2 * (lane & 15) + s;
        lane_tw[(5 + s) * 32 + lane] = col ? d2(r64(digit_rev(2, 5, col))) : make_double2(1.0, 0.0);
      }
    }
  }
}

bool shape_is_a(const lolb_plan* pl)
{
  if (pl->kind != PLAN_C || pl->pe.size() != 3) return false;
  const PrimeExponent want[3] = {{2, 6}, {3, 2}, {5, 2}};
  for (int i = 0; i < 3; i++) if (pl->pe[i].prime != want[i].prime || pl->pe[i].exponent != want[i].exponent) return false;
  return true;
}

}  // namespace

int fused_ac_select(lolb_plan* pl, void** slot)
{
  if (!shape_is_a(pl)) return LOLB_OK;
  FusedAC* F = (FusedAC*)*slot;
  if (!F) { F = new FusedAC(); *slot = F; }
  const int k = pl->k;
  auto complete = [&](const std::vector<std::vector<lolb_complex>>& T) {
    if (T.size() != 3) return false;
    const size_t want[3] = {64, 9, 25};
    for (int i = 0; i < 3; i++) if (T[i].size() != want[i] * (size_t)k) return false;
    return true;
  };
  F->ok_fwd = pl->has_fwd && complete(pl->cru);
  F->ok_inv = pl->has_inv && complete(pl->cruinv);
  std::vector<double2> lt((size_t)k * 2 * kLaneRows * 32, make_double2(1.0, 0.0));
  F->fwd.assign(k, FusedACConsts{});
  F->inv.assign(k, FusedACConsts{});
  for (int t = 0; t < k; t++) {
    if (F->ok_fwd) build_consts(pl, false, t, &F->fwd[t], lt.data() + ((size_t)t * 2 + 0) * kLaneRows * 32);
    if (F->ok_inv) build_consts(pl, true, t, &F->inv[t], lt.data() + ((size_t)t * 2 + 1) * kLaneRows * 32);
  }
  if (F->d_lane_tw) { cudaFree(F->d_lane_tw); F->d_lane_tw = nullptr; }
  LOLB_CUDA(cudaMalloc((void**)&F->d_lane_tw, lt.size() * sizeof(double2)));
  LOLB_CUDA(cudaMemcpy(F->d_lane_tw, lt.data(), lt.size() * sizeof(double2), cudaMemcpyHostToDevice));
  for (int t = 0; t < k; t++) {
    F->fwd[t].lane_tw = F->d_lane_tw + ((size_t)t * 2 + 0) * kLaneRows * 32;
    F->inv[t].lane_tw = F->d_lane_tw + ((size_t)t * 2 + 1) * kLaneRows * 32;
  }
  return LOLB_OK;
}

void fused_ac_release(void* slot)
{
  FusedAC* F = (FusedAC*)slot;
  if (!F) return;
  if (F->d_lane_tw) cudaFree(F->d_lane_tw);
  delete F;
}

bool fused_ac_available(const void* slot, bool inverse)
{
  const FusedAC* F = (const FusedAC*)slot;
  return F && (inverse ? F->ok_inv : F->ok_fwd);
}

template <int W, int MB>
static int launch_ac(const lolb_plan* pl, const FusedAC* F, bool inverse, double2* y, int64_t batch, cudaStream_t st)
{
  const size_t smem = (size_t)(kN + kLaneRows * 32) * sizeof(double2);
  static PerDeviceOnce once;
  if (once.first()) {
    LOLB_CUDA(cudaFuncSetAttribute(k_fused_ac<true, 1, W, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LOLB_CUDA(cudaFuncSetAttribute(k_fused_ac<false, 1, W, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LOLB_CUDA(cudaFuncSetAttribute(k_fused_ac<true, 0, W, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LOLB_CUDA(cudaFuncSetAttribute(k_fused_ac<false, 0, W, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  int64_t grid = (int64_t)pl->num_sms * MB;
  if (grid > batch) grid = batch;
  for (int t = 0; t < pl->k; t++) {
    if (pl->k == 1) {
      if (inverse) k_fused_ac<true, 1, W, MB><<<(int)grid, W * 32, smem, st>>>(y, batch, 1, t, F->inv[t]);
      else k_fused_ac<false, 1, W, MB><<<(int)grid, W * 32, smem, st>>>(y, batch, 1, t, F->fwd[t]);
    } else {
      if (inverse) k_fused_ac<true, 0, W, MB><<<(int)grid, W * 32, smem, st>>>(y, batch, pl->k, t, F->inv[t]);
      else k_fused_ac<false, 0, W, MB><<<(int)grid, W * 32, smem, st>>>(y, batch, pl->k, t, F->fwd[t]);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_fused_ac");
    count_launch();
  }
  return LOLB_OK;
}

int fused_ac_crt(const lolb_plan* pl, const void* slot, bool inverse, double2* y, int64_t batch, cudaStream_t st)
{
  const FusedAC* F = (const FusedAC*)slot;
  if (!fused_ac_available(slot, inverse)) return LOLB_FUSED_UNAVAILABLE;
  if (batch <= 0) return LOLB_OK;
  // measured at m = 14400, batch 32768 (% of HBM peak, forward / inverse): 3 warps x 3 CTAs/SM (168 registers, spills)
  // 47 / 42; 3 x 2 (255 registers) 47 / 49; 4 x 3 49 / 49; 2 x 3 47 / 39; 6 warps x 2 CTAs/SM (158 registers, no
  // spill, phase 1 = one column task per warp) 61 / 61 -- the only configuration kept
  return launch_ac<6, 2>(pl, F, inverse, y, batch, st);
}

}  // namespace lolb
