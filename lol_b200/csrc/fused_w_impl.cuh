// fused_w_impl.cuh -- fused Z_q CRT / CRT^-1 for indices m = 2^a * (one to three odd prime powers): the reference's other
// benchmark rings (lol/Crypto/Lol/Benchmarks/Default.hs:41-48: F64*F27, F64*F81 and the Twace-Embed rings
// F32*F7*F13, F8*F7*F13, F8*F5*F7*F13; lol-apps tunnel ring F64*F7*F13), one HBM read and one HBM write per ring element like fused_a.cu does for m = 14400.
//
// The operator is the reference's (crt.cpp:518-581 on tensor.h:76-95): CRT_m = (x)_i CRT_{p_i^e_i}, first factor fastest,
// CRT_{p^e} = (DFT_{p^(e-1)} (x) I_{p-1}) . That . (I_{p^(e-1)} (x) CRT_p), DFT_{p^(e-1)} as e-1 radix-p rounds with diagonal
// twiddles between them (ppDFT, crt.cpp:459-486).  Factors on different axes commute and the arithmetic is exact, so the
// schedule below yields the reference's residues bit for bit.
//
//   element X[ic][ib][ia][i1]:  i1 < L = 2^(a-1) (the 2^a axis, fastest), ia, ib the "middle" odd prime powers, ic the last
//
//   line     all stages of ONE odd prime power on the phi(p^e) values a thread holds in registers, compile-time indices,
//            every stage a set of small dense rows  sum_j c_j x_j  accumulated lazily and reduced once; the diagonal
//            twiddles (crtTwiddle, dftTwiddle) and mhat^-1 are folded into the row constants on the host (kernel
//            parameter bank: every lane of a warp uses the same constant)
//   network  the 2^a axis across L lanes: the register-exchange butterfly network of fused_a.cu generalised to any
//            L = 2 .. 32 and any number of value pairs per lane (sub-warp groups when L < 32)
//
//   k_fused_w1  (no last axis: m = 2^a p^e, e.g. 1728, 5184) a group of L lanes owns one ring element: loads its D2 values
//               per lane straight from HBM, line(s), network, stores.  No shared memory, no barrier.
//   k_fused_w2  (m = 2^a .. p_c^e_c) phase 1: thread <- one column along ic from HBM, line, u32 tile in shared memory;
//               phase 2: group of L lanes <- one ic-row of the tile, middle line(s), network, stores.
#ifndef LOLB_W_PART
#error "fused_w_impl.cuh is compiled through fused_w_p0.cu .. fused_w_p4.cu (LOLB_W_PART selects the kernels a translation unit instantiates)"
#endif
#include <complex>
#include <cstdlib>
#include <mutex>

#include "fused.cuh"
#include "numtheory.h"

namespace lolb {

namespace {

#define WHD __host__ __device__ __forceinline__

constexpr int ipw(int b, int e) { return e <= 0 ? 1 : b * ipw(b, e - 1); }

WHD uint32_t w_min(uint32_t a, uint32_t b) { return a < b ? a : b; }
WHD uint32_t w_mulhi(uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// ------------------------------------------------------------------ arithmetic (the two policies of fused_a.cu, callable
// from the host as well so that the device-free emulation below runs the very same line code)
struct WMod {
  uint32_t q, q2;
  uint32_t r0;     // WS: floor(2^32 / q);  WM: -q^-1 mod 2^32
  uint32_t one;    // WS: 1;                WM: 2^32 mod q
  uint32_t r2;     // WM: 2^64 mod q
};

// WS: a row of T terms needs 2 T q^2 < 2^32; residues lazily in [0, 2q), Barrett reduction
struct WS {
  typedef uint32_t T;       // value in registers / shared memory
  typedef int64_t IO;       // value in HBM (the ABI)
  typedef uint32_t Acc;
  static constexpr bool kZq = true;
  static constexpr bool kSym = false;
  static constexpr int MAXT = 1 << 20;      // terms per reduction: unbounded (the host checks 2 T q^2 < 2^32 for the longest row)
  uint32_t q, q2, mu, nq;
  WHD uint32_t add(uint32_t a, uint32_t b) const { return a + b; }                  // lazy: the caller folds or reduces
  WHD uint32_t sub(uint32_t u, uint32_t t) const { return u + q2 - t; }             // u - t + 2q, t < 2q
  WHD WS(const WMod& M) : q(M.q), q2(M.q2), mu(M.r0), nq(0u - M.q) {}
  WHD Acc mul(uint32_t c, uint32_t v) const { return c * v; }
  WHD Acc mad(Acc a, uint32_t c, uint32_t v) const { return a + c * v; }
  WHD Acc unit(uint32_t v) const { return v; }
  WHD uint32_t red(Acc x) const { return w_mulhi(x, mu) * nq + x; }
  WHD uint32_t fold(uint32_t x) const { return w_min(x, x - q2); }
  WHD uint32_t canon(uint32_t x) const { return w_min(x, x - q); }
};

// WS6: the same 32-bit arithmetic for moduli whose longest rows (12 or 13 terms: p = 13) would overflow: a row is cut into
// pieces of at most 6 terms, each reduced on its own (2 * 6 * q^2 < 2^32), the pieces folded together
struct WS6 : WS {
  static constexpr bool kSym = true;
  static constexpr int MAXT = 6;
  WHD WS6(const WMod& M) : WS(M) {}
};

// WM: odd q, 2 T q < 2^32; 64-bit accumulation, one Montgomery reduction per row, constants in Montgomery form
struct WM {
  typedef uint32_t T;
  typedef int64_t IO;
  typedef uint64_t Acc;
  static constexpr bool kZq = true;
  static constexpr bool kSym = true;
  static constexpr int MAXT = 1 << 20;
  uint32_t q, q2, qinv, one;
  WHD uint32_t add(uint32_t a, uint32_t b) const { return a + b; }
  WHD uint32_t sub(uint32_t u, uint32_t t) const { return u + q2 - t; }
  WHD WM(const WMod& M) : q(M.q), q2(M.q2), qinv(M.r0), one(M.one) {}
  WHD Acc mul(uint32_t c, uint32_t v) const { return (uint64_t)c * v; }
  WHD Acc mad(Acc a, uint32_t c, uint32_t v) const { return a + (uint64_t)c * v; }
  WHD Acc unit(uint32_t v) const { return (uint64_t)one * v; }
  WHD uint32_t red(Acc x) const
  {
    const uint32_t m = (uint32_t)x * qinv;
    return (uint32_t)((x + (uint64_t)m * q) >> 32);
  }
  WHD uint32_t fold(uint32_t x) const { return w_min(x, x - q2); }
  WHD uint32_t canon(uint32_t x) const { return w_min(x, x - q); }
};

// WC: complex double (tensorCRTC / tensorCRTInvC, crt.cpp:583-598; class Complex, types.h:122-164).  The same schedule; no
// reductions, no lazy ranges.  Rounding differs from the reference's evaluation order (FMA contraction, folded twiddles):
// parity is to 1e-9 relative like every floating-point path.
struct WC {
  typedef double2 T;
  typedef double2 IO;
  typedef double2 Acc;
  static constexpr bool kZq = false;
  static constexpr bool kSym = true;
  static constexpr int MAXT = 1 << 20;
  WHD WC(const WMod&) {}
  WHD Acc mul(double2 c, double2 v) const { return make_double2(c.x * v.x - c.y * v.y, c.x * v.y + c.y * v.x); }
  WHD Acc mad(Acc a, double2 c, double2 v) const { return make_double2(a.x + (c.x * v.x - c.y * v.y), a.y + (c.x * v.y + c.y * v.x)); }
  WHD Acc unit(double2 v) const { return v; }
  WHD double2 red(Acc x) const { return x; }
  WHD double2 fold(double2 x) const { return x; }
  WHD double2 canon(double2 x) const { return x; }
  WHD double2 add(double2 a, double2 b) const { return make_double2(a.x + b.x, a.y + b.y); }
  WHD double2 sub(double2 a, double2 b) const { return make_double2(a.x - b.x, a.y - b.y); }
};

// ------------------------------------------------------------------ one odd prime power p^e: constant layout
//   M1[i0][r][c]          i0 < p^(e-1); forward  crtTwiddle(i0, r) . CRT_p[r][c]   (crt.cpp:60-79, 248-346)
//                                       inverse  CRT_p^-1'[r][c] . crtTwiddle(i0, c) (. mhat^-1 on one axis) (crt.cpp:349-457)
//   W_d[hi][r][a]         round on base-p digit d of i0, hi = the digits above d: forward dftTwiddle(hi, r) . DFT_p[r][a],
//                         inverse DFT_p[r][a] . dftTwiddle(hi, a)   (crt.cpp:84-126, 131-246, 459-516)
template <int P, int E>
struct PPT {
  static constexpr int p = P, e = E, R = E - 1, mp = ipw(P, E - 1), d = P - 1, phi = (P - 1) * ipw(P, E - 1);
  static constexpr int n_m1 = mp * d * d;
  static constexpr int w_off(int dig)
  {
    int o = n_m1;
    for (int t = 0; t < dig; t++) o += ipw(P, R - 1 - t) * P * P;
    return o;
  }
  static constexpr int n_consts = w_off(R);
};
template <>
struct PPT<1, 1> {      // "no prime power here"
  static constexpr int p = 1, e = 1, R = 0, mp = 1, d = 0, phi = 1, n_m1 = 0, n_consts = 0;
  static constexpr int w_off(int) { return 0; }
};
typedef PPT<1, 1> PPNone;

// one radix-p round on digit DIG of the block index, values of the line at v[base + (i0 * (p-1) + cc) * STRIDE]
template <class PPx, bool INV, int DIG, int STRIDE, int COFF, class AR, class CT, class TV, int NV>
WHD void pp_round(TV (&v)[NV], const int base, const CT& C, const AR& A)
{
  constexpr int P = PPx::p, D = PPx::d, R = PPx::R;
  constexpr int NHI = ipw(P, R - 1 - DIG), NLO = ipw(P, DIG), WOFF = COFF + PPx::w_off(DIG);
#pragma unroll
  for (int hi = 0; hi < NHI; hi++) {
#pragma unroll
    for (int lo = 0; lo < NLO; lo++) {
#pragma unroll
      for (int cc = 0; cc < D; cc++) {
        TV x[P], o[P];
#pragma unroll
        for (int a = 0; a < P; a++) x[a] = v[base + (((hi * P + a) * NLO + lo) * D + cc) * STRIDE];
#pragma unroll
        for (int r = 0; r < P; r++) {
          const bool all_ones = r == 0 && (!INV || hi == 0);      // DFT row 0 carries no twiddle in the forward direction
          const bool col0_one = INV || hi == 0;                   // the inverse twiddles its inputs: input 0 is never scaled
          if (all_ones) {
            TV s = x[0];
#pragma unroll
            for (int a = 1; a < P; a++) s = A.add(s, x[a]);
            o[r] = A.red(A.unit(s));
          } else {
            typename AR::Acc acc = col0_one ? A.unit(x[0]) : A.mul(C.c[WOFF + (hi * P + r) * P], x[0]);
            TV part = TV();
#pragma unroll
            for (int a = 1; a < P; a++) {
              if (a % AR::MAXT == 0) { part = a == AR::MAXT ? A.red(acc) : A.fold(A.add(part, A.red(acc))); acc = A.mul(C.c[WOFF + (hi * P + r) * P + a], x[a]); }
              else acc = A.mad(acc, C.c[WOFF + (hi * P + r) * P + a], x[a]);
            }
            o[r] = P > AR::MAXT ? A.fold(A.add(part, A.red(acc))) : A.red(acc);
          }
        }
#pragma unroll
        for (int r = 0; r < P; r++) v[base + (((hi * P + r) * NLO + lo) * D + cc) * STRIDE] = o[r];
      }
    }
  }
}

template <class PPx, bool INV, int STRIDE, int COFF, class AR, class CT, class TV, int NV>
WHD void pp_blocks(TV (&v)[NV], const int base, const CT& C, const AR& A)
{
  constexpr int D = PPx::d, MP = PPx::mp;
#pragma unroll
  for (int i0 = 0; i0 < MP; i0++) {
    TV o[D];
#pragma unroll
    for (int r = 0; r < D; r++) {
      typename AR::Acc acc = A.mul(C.c[COFF + (i0 * D + r) * D], v[base + (i0 * D) * STRIDE]);
      TV part = TV();
#pragma unroll
      for (int cc = 1; cc < D; cc++) {
        if (cc % AR::MAXT == 0) { part = cc == AR::MAXT ? A.red(acc) : A.fold(A.add(part, A.red(acc))); acc = A.mul(C.c[COFF + (i0 * D + r) * D + cc], v[base + (i0 * D + cc) * STRIDE]); }
        else acc = A.mad(acc, C.c[COFF + (i0 * D + r) * D + cc], v[base + (i0 * D + cc) * STRIDE]);
      }
      o[r] = D > AR::MAXT ? A.fold(A.add(part, A.red(acc))) : A.red(acc);
    }
#pragma unroll
    for (int r = 0; r < D; r++) v[base + (i0 * D + r) * STRIDE] = o[r];
  }
}

#ifndef LOLB_W_SYM
#define LOLB_W_SYM 1      // 0: dense (p-1) x (p-1) blocks for the primes too (the first version; A/B builds)
#endif
// Used where a product costs more than an addition: the 64-bit-accumulate class (IMAD.WIDE, half rate), the class whose rows are
// cut into 6-term pieces, and complex doubles.  For the plain 32-bit class (one IMAD per product) the dense block measured faster
// (m = 2912, q = 8737: 71 % / 72 % of HBM dense, 69 % / 67 % with the split), so WS keeps it.
template <class PPx, class AR> struct PPSym { static constexpr bool on = LOLB_W_SYM && AR::kSym && PPx::e == 1 && PPx::p >= 5; };

// CRT_p / CRT_p^-1 of a PRIME p >= 5 (no twiddles: crtTwiddle and ppDFT are empty for e = 1) with half the multiplications.
// With k = row + 1 and w = w_p:   forward  out_k = sum_{c=0}^{p-2} w^(kc) x_c,  k = 1 .. p-1   (crtp, crt.cpp:248-346)
//                                 inverse  out_r = s (sum_{k=1}^{p-1} w^(-rk) y_k - sum_k w^k y_k),  r = 0 .. p-2   (crtpinv, :349-457)
// Pair the inputs (c, p - c) -- e_c = x_c + x_{p-c}, d_c = x_c - x_{p-c} (x_{p-1} = 0 forward) -- and the outputs (k, p - k):
//   C_k = sum_c cos_kc e_c,  S_k = sum_c sin_kc d_c,  cos_kc = (w^kc + w^-kc) / 2,  sin_kc = (w^kc - w^-kc) / 2,  c, k = 1 .. h = (p-1)/2
//   forward  out_k = x_0 + C_k + S_k,  out_{p-k} = x_0 + C_k - S_k
//   inverse  A_r = C_r - S_r,  A_{p-r} = C_r + S_r,  A_0 = s sum_c e_c;  the common term is A_{p-1} = C_1 + S_1;  out_r = A_r - A_{p-1}
// 2 h^2 = (p-1)^2 / 2 products instead of (p-1)^2; exact arithmetic, so the residues are the reference's.
// Constants at COFF: cos[k-1][c-1] (h x h), then sin[k-1][c-1], then s (inverse; cos and sin carry it too).
template <class PPx, bool INV, int STRIDE, int COFF, class AR, class CT, class TV, int NV>
WHD void pp_prime_sym(TV (&v)[NV], const int base, const CT& C, const AR& A)
{
  constexpr int P = PPx::p, H = (P - 1) / 2;
  static_assert(H <= AR::MAXT, "a row of the half-size blocks must fit one lazy accumulation");      // h products (< 2 q^2 each) + x_0 < 2q: inside every class bound (w_class)
  TV e[H], d[H], cs[H], sn[H];
  if constexpr (!INV) {
    const TV x0 = v[base];
    e[0] = d[0] = v[base + STRIDE];
#pragma unroll
    for (int c = 2; c <= H; c++) {
      const TV a = v[base + c * STRIDE], b = v[base + (P - c) * STRIDE];
      e[c - 1] = A.fold(A.add(a, b));
      d[c - 1] = A.fold(A.sub(a, b));
    }
#pragma unroll
    for (int k = 1; k <= H; k++) {
      typename AR::Acc ac = A.unit(x0), as = A.mul(C.c[COFF + H * H + (k - 1) * H], d[0]);
#pragma unroll
      for (int c = 1; c <= H; c++) ac = A.mad(ac, C.c[COFF + (k - 1) * H + (c - 1)], e[c - 1]);
#pragma unroll
      for (int c = 2; c <= H; c++) as = A.mad(as, C.c[COFF + H * H + (k - 1) * H + (c - 1)], d[c - 1]);
      cs[k - 1] = A.red(ac);
      sn[k - 1] = A.red(as);
    }
#pragma unroll
    for (int k = 1; k <= H; k++) {
      v[base + (k - 1) * STRIDE] = A.fold(A.add(cs[k - 1], sn[k - 1]));
      v[base + (P - k - 1) * STRIDE] = A.fold(A.sub(cs[k - 1], sn[k - 1]));
    }
  } else {
#pragma unroll
    for (int k = 1; k <= H; k++) {
      const TV a = v[base + (k - 1) * STRIDE], b = v[base + (P - k - 1) * STRIDE];
      e[k - 1] = A.fold(A.add(a, b));
      d[k - 1] = A.fold(A.sub(a, b));
    }
    TV sum = e[0];
#pragma unroll
    for (int k = 2; k <= H; k++) sum = A.add(sum, e[k - 1]);
    const TV a0 = A.red(A.mul(C.c[COFF + 2 * H * H], sum));
#pragma unroll
    for (int r = 1; r <= H; r++) {
      typename AR::Acc ac = A.mul(C.c[COFF + (r - 1) * H], e[0]), as = A.mul(C.c[COFF + H * H + (r - 1) * H], d[0]);
#pragma unroll
      for (int k = 2; k <= H; k++) {
        ac = A.mad(ac, C.c[COFF + (r - 1) * H + (k - 1)], e[k - 1]);
        as = A.mad(as, C.c[COFF + H * H + (r - 1) * H + (k - 1)], d[k - 1]);
      }
      cs[r - 1] = A.red(ac);
      sn[r - 1] = A.red(as);
    }
    const TV shift = A.fold(A.add(cs[0], sn[0]));      // A_{p-1}
    v[base] = A.fold(A.sub(a0, shift));
#pragma unroll
    for (int r = 1; r <= H; r++) {
      v[base + r * STRIDE] = A.fold(A.sub(A.fold(A.sub(cs[r - 1], sn[r - 1])), shift));
      if (r >= 2) v[base + (P - r) * STRIDE] = A.fold(A.sub(A.fold(A.add(cs[r - 1], sn[r - 1])), shift));
    }
  }
}

// CRT_{p^e} / CRT_{p^e}^-1 on one line (ppcrt / ppcrtinv, crt.cpp:518-560)
template <class PPx, bool INV, int STRIDE, int COFF, class AR, class CT, class TV, int NV>
WHD void pp_line(TV (&v)[NV], const int base, const CT& C, const AR& A)
{
  if constexpr (PPSym<PPx, AR>::on) {
    pp_prime_sym<PPx, INV, STRIDE, COFF>(v, base, C, A);
  } else if constexpr (PPx::p > 1) {
    constexpr int R = PPx::R;
    static_assert(R <= 4, "prime-power exponent too large for the unrolled rounds");
    if constexpr (!INV) {
      pp_blocks<PPx, false, STRIDE, COFF>(v, base, C, A);
      if constexpr (R > 0) pp_round<PPx, false, 0, STRIDE, COFF>(v, base, C, A);
      if constexpr (R > 1) pp_round<PPx, false, 1, STRIDE, COFF>(v, base, C, A);
      if constexpr (R > 2) pp_round<PPx, false, 2, STRIDE, COFF>(v, base, C, A);
      if constexpr (R > 3) pp_round<PPx, false, 3, STRIDE, COFF>(v, base, C, A);
    } else {
      if constexpr (R > 3) pp_round<PPx, true, 3, STRIDE, COFF>(v, base, C, A);
      if constexpr (R > 2) pp_round<PPx, true, 2, STRIDE, COFF>(v, base, C, A);
      if constexpr (R > 1) pp_round<PPx, true, 1, STRIDE, COFF>(v, base, C, A);
      if constexpr (R > 0) pp_round<PPx, true, 0, STRIDE, COFF>(v, base, C, A);
      pp_blocks<PPx, true, STRIDE, COFF>(v, base, C, A);
    }
  }
}

// ------------------------------------------------------------------ shapes
// A: exponent of 2 (0 or 1: no 2^a axis), PA / PB: the odd prime powers handled together with the 2^a axis, PC: the
// last prime power (phase 1 of k_fused_w2; PPNone selects k_fused_w1), PD: an optional prime power between them that is
// transformed in place in the tile (phase 1b), for indices with four odd prime powers.  EPB: ring elements per CTA
// iteration (w2).  Tensor order (fastest first): 2^a, PA, PB, PD, PC.
template <int A_, class PA_, class PB_, class PD_, class PC_, int EPB_, int MINB_>
struct WShape {
  typedef PA_ PA;
  typedef PB_ PB;
  typedef PD_ PD;
  typedef PC_ PC;
  static constexpr int A = A_, LOG = A_ >= 2 ? A_ - 1 : 0, L = 1 << LOG;
  // the 2^a axis spans LL lane bits; beyond 32 columns (a = 7) a lane holds H = 2 column halves (col, col + 32) of every row
  static constexpr int LL = LOG > 5 ? 5 : LOG, LW = 1 << LL, H = L / LW, GPW = 32 / LW;
  static_assert(LOG <= 6, "2^a axis: a <= 7");
  static constexpr int DA = PA::phi, DB = PB::phi, D2 = DA * DB, NP = D2 / 2, DD = PD::phi;
  static constexpr int SUB = L * D2;                      // coefficients of one phase-2 block (fixed id, ic)
  static constexpr int COLS = SUB * DD, ROWS = PC::phi, N = COLS * ROWS;
  static constexpr int OFF_A = 0, OFF_B = PA::n_consts, OFF_D = OFF_B + PB::n_consts, OFF_C = OFF_D + PD::n_consts;
  static constexpr int NC = (OFF_C + PC::n_consts) > 0 ? (OFF_C + PC::n_consts) : 1;
  static constexpr int EPB = EPB_, MINB = MINB_;
  static constexpr bool TWO_PHASE = PC::p > 1;
  // tile strides: the blocks handled by the sub-warp groups of one warp must start in different banks
  // (block b of a warp's 32/L groups sits at b * SRS: with SRS = L * odd the groups cover all 32 banks exactly once)
  static constexpr int SRS = L >= 32 ? SUB : L * (D2 | 1);      // block stride
  static constexpr int RS = DD * SRS;                                                                      // stride of ic
  static_assert(DD == 1 || TWO_PHASE, "the in-tile axis needs the tile");
  static_assert(H == 1 || TWO_PHASE, "a = 7 is served by the two-phase kernel only");
  static_assert(D2 % 2 == 0, "an odd prime power is required next to the 2^a axis");
};

// complex doubles: CTAs of 128 threads per SM asked for.  Measured on B200 (% of the 32 n-byte roofline, CRT / CRT^-1, at 1 -> 3 -> 4):
// m = 1728  42 / 79 -> 58 / 66 -> 61 / 68;  m = 2912  44 / 42 -> 55 / 52 -> 61 / 55;  m = 3640  25 / 24 -> 32 / 30 -> 36 / 33;
// m = 11648  35 / 33 -> 41 / 39 -> 30 / 24;  m = 5184  27 / 34 -> 23 / 29 -> 17 / 21 (54 complex values per lane need every register)
template <class SH, bool INV>
constexpr int wc_minb()
{
  if (SH::D2 * SH::H >= 48) return 1;
  if (!SH::TWO_PHASE) return INV ? 1 : 4;
  return SH::H == 2 ? 3 : 4;
}
constexpr int kWLaneRows = 10;     // per-lane constants of the network, per column half (see build_lane_table)
constexpr int kWThreads = 128;

template <class T, int NC>
struct WConsts {
  WMod mod;
  const T* lane_tw;      // device [2 halves][kWLaneRows][32]
  T c[NC];
};

__device__ __forceinline__ uint32_t w_shfl_xor(uint32_t v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }
__device__ __forceinline__ double2 w_shfl_xor(double2 v, int mask)
{
  return make_double2(__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask));
}

// ------------------------------------------------------------------ the 2^a axis across L lanes
// State: the lane holds NP pairs (v[2j], v[2j+1]).  Round on lane bit b: the lane keeps one value of each pair, swaps the
// other with lane ^ 2^b and then owns both inputs of NP butterflies.  Forward (u,t) -> (u+t, (u-t) tw); inverse
// (u,t) -> (u + t tw, u - t tw).  Ownership afterwards (derived in DESIGN.md 4.9):
//   forward: pair j, slot s  =  odd-axis row 2j + (l & 1),          column (s << (LOG-1)) | (l >> 1)
//   inverse: pair j, slot s  =  odd-axis row 2j + (l >> (LOG-1)),   column 2 (l & (L/2 - 1)) + s
template <bool INV, bool TRIVIAL, int NP, class AR>
__device__ __forceinline__ void w_round(typename AR::T (&v)[2 * NP], const int l, const int bit, const typename AR::T tw, const AR& A)
{
  typedef typename AR::T T;
  const bool hi = (l >> bit) & 1;
#pragma unroll
  for (int j = 0; j < NP; j++) {
    const T send = hi ? v[2 * j] : v[2 * j + 1];
    const T keep = hi ? v[2 * j + 1] : v[2 * j];
    const T recv = w_shfl_xor(send, 1 << bit);
    if (TRIVIAL) {                 // every twiddle of the round is 1 (crt.cpp:92-106 skips i0 = 0)
      const T u = hi ? recv : keep, t = hi ? keep : recv;
      v[2 * j] = A.fold(A.add(u, t));
      v[2 * j + 1] = A.fold(A.sub(u, t));
    } else if (!INV) {             // u + t is symmetric; the sign of u - t lives in the lane's twiddle (host: -tw for hi lanes)
      v[2 * j] = A.fold(A.add(keep, recv));
      v[2 * j + 1] = A.red(A.mul(tw, A.sub(keep, recv)));
    } else {
      const T t = A.red(A.mul(tw, hi ? keep : recv));
      const T u = hi ? recv : keep;
      v[2 * j] = A.fold(A.add(u, t));
      v[2 * j + 1] = A.fold(A.sub(u, t));
    }
  }
}

// last inverse round (lane bit 0) merged with the inverse crtTwiddle of the 2^a axis: the lane ends with columns 2c, 2c+1
// whose twiddles a, b are per-lane constants:  a (u + tw t) = a u + (a tw) t,  b (u - tw t) = b u + (-b tw) t
template <int NP, class AR>
__device__ __forceinline__ void w_last_inv(typename AR::T (&v)[2 * NP], const int l, const typename AR::T a, const typename AR::T atw,
                                           const typename AR::T b, const typename AR::T nbtw, const AR& A)
{
  typedef typename AR::T T;
  const bool hi = l & 1;
#pragma unroll
  for (int j = 0; j < NP; j++) {
    const T send = hi ? v[2 * j] : v[2 * j + 1];
    const T keep = hi ? v[2 * j + 1] : v[2 * j];
    const T recv = w_shfl_xor(send, 1);
    const T t = hi ? keep : recv, u = hi ? recv : keep;
    v[2 * j] = A.red(A.mad(A.mul(a, u), atw, t));
    v[2 * j + 1] = A.red(A.mad(A.mul(b, u), nbtw, t));
  }
}

// lane-table rows (per column half): forward [0] crtTwiddle of the lane's column, [1 + r] round r (the sign of hi lanes folded
// in); inverse [r] round r (r >= 1), [6] a, [7] a tw_0, [8] b, [9] -b tw_0.  LL = lane bits of the axis; TOP_IN_LANES: the
// highest column bit is a lane bit (a <= 6), so its round carries no twiddle (crt.cpp:92-106 skips i0 = 0).
template <int LL, bool TOP_IN_LANES, bool INV, int NP, class AR>
__device__ __forceinline__ void w_network(typename AR::T (&v)[2 * NP], const int l, const typename AR::T (&lt)[kWLaneRows], const AR& A)
{
  if constexpr (LL >= 1) {
    constexpr int NT = TOP_IN_LANES ? LL - 1 : LL;      // rounds with twiddles
    if constexpr (!INV) {
#pragma unroll
      for (int i = 0; i < 2 * NP; i++) v[i] = A.red(A.mul(lt[0], v[i]));      // crtTwiddle (crt.cpp:43-58)
#pragma unroll
      for (int r = 0; r < NT; r++) w_round<false, false, NP>(v, l, r, lt[1 + r], A);
      if constexpr (TOP_IN_LANES) w_round<false, true, NP>(v, l, LL - 1, lt[0], A);
    } else {
      if constexpr (TOP_IN_LANES && LL >= 2) w_round<true, true, NP>(v, l, LL - 1, lt[0], A);
#pragma unroll
      for (int r = NT - 1; r >= 1; r--) w_round<true, false, NP>(v, l, r, lt[r], A);
      w_last_inv<NP>(v, l, lt[6], lt[7], lt[8], lt[9], A);
    }
  }
}

// the round on the column bit above the lanes (a = 7): both inputs sit in the same thread, the twiddle is 1
template <int N2, class AR>
__device__ __forceinline__ void w_top_round(typename AR::T (&v0)[N2], typename AR::T (&v1)[N2], const AR& A)
{
#pragma unroll
  for (int i = 0; i < N2; i++) {
    const typename AR::T u = v0[i], t = v1[i];
    v0[i] = A.fold(A.add(u, t));
    v1[i] = A.fold(A.sub(u, t));
  }
}

// position inside a block of D2 * L coefficients of value (pair j, slot s, column half h) after the network
template <int LL, int L, bool INV>
__device__ __forceinline__ int w_out_pos(const int l, const int j, const int s, const int h)
{
  constexpr int LW = 1 << LL;
  if constexpr (LL == 0) return 2 * j + s;
  else if constexpr (!INV) return (2 * j + (l & 1)) * L + ((s << (LL - 1)) | (l >> 1)) + LW * h;
  else return (2 * j + (l >> (LL - 1))) * L + 2 * (l & (LW / 2 - 1)) + s + LW * h;
}

__device__ __noinline__ uint32_t w_reduce_any(int64_t x, uint32_t q)      // non-canonical input, like `c % q` (types.h:62-66)
{
  int64_t r = x % (int64_t)q;
  return (uint32_t)(r < 0 ? r + q : r);
}

// middle line(s) + network + store of one group's block of D2 * L coefficients; v[h] = the lane's values of column half h
// the loads of one thread-task: NV values at `step` apart, all issued before the first use.  Zq: 64-bit words narrowed to the
// 32-bit working range; one out-of-range word sends the task through `c % q` like the reference's constructor.
// STREAM: the words are touched once (tupSize 1): evict-first.  With several limbs per tuple the sectors are shared with the CTAs
// of the other limbs and must stay in L2 until those have come by: default policy.
template <bool STREAM, class TP>
__device__ __forceinline__ TP w_ld(const TP* p) { if constexpr (STREAM) return __ldcs(p); else return __ldcg(p); }
template <bool STREAM, class TP>
__device__ __forceinline__ void w_st(TP* p, const TP v) { if constexpr (STREAM) __stcs(p, v); else __stcg(p, v); }

template <class AR, bool STREAM, int NV>
__device__ __forceinline__ void w_load(typename AR::T (&v)[NV], const typename AR::IO* __restrict__ src, const size_t step, const bool live,
                                       const uint32_t q)
{
  if constexpr (AR::kZq) {
    uint32_t hi_or = 0, lo_max = 0;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      const int64_t raw = live ? w_ld<STREAM>(src + (size_t)i * step) : 0;
      v[i] = (uint32_t)raw;
      hi_or |= (uint32_t)((uint64_t)raw >> 32);
      lo_max = max(lo_max, v[i]);
    }
    if (hi_or != 0 || lo_max >= q) {
#pragma unroll
      for (int i = 0; i < NV; i++) v[i] = w_reduce_any(src[(size_t)i * step], q);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = live ? w_ld<STREAM>(src + (size_t)i * step) : make_double2(0.0, 0.0);
  }
}

template <class SH, bool INV, class AR, int K>
__device__ __forceinline__ void w_finish(typename AR::T (&v)[SH::H][SH::D2], const int l, const typename AR::T (&lt)[SH::H][kWLaneRows],
                                         const WConsts<typename AR::T, SH::NC>& C, const AR& A,
                                         typename AR::IO* __restrict__ dst /* block base (+ limb) */, const int k, const bool live)
{
#pragma unroll
  for (int h = 0; h < SH::H; h++) {
    // axis a: stride 1, one line per ib;  axis b: stride DA, one line per ia
#pragma unroll
    for (int ib = 0; ib < SH::DB; ib++) pp_line<typename SH::PA, INV, 1, SH::OFF_A>(v[h], ib * SH::DA, C, A);
#pragma unroll
    for (int ia = 0; ia < SH::DA; ia++) pp_line<typename SH::PB, INV, SH::DA, SH::OFF_B>(v[h], ia, C, A);
  }
  if constexpr (SH::H == 2 && INV) w_top_round<SH::D2>(v[0], v[1], A);
#pragma unroll
  for (int h = 0; h < SH::H; h++) w_network<SH::LL, SH::H == 1, INV, SH::NP>(v[h], l, lt[h], A);
  if constexpr (SH::H == 2 && !INV) w_top_round<SH::D2>(v[0], v[1], A);
  if (live) {
#pragma unroll
    for (int h = 0; h < SH::H; h++) {
#pragma unroll
      for (int j = 0; j < SH::NP; j++) {
        if constexpr (AR::kZq) {
          const int64_t a = (int64_t)A.canon(v[h][2 * j]), b = (int64_t)A.canon(v[h][2 * j + 1]);
          // the pair is adjacent in memory (inverse: always; forward: 2^a axes of <= 2 lanes): one 16-byte store.  Measured at
          // q = 3144961: m = 4095 36 % / 35 % -> 51 % / 49 % of HBM, m = 5460 52 % -> 58 % forward.  Sending the blocks of the
          // narrow shapes (L <= 4) back through the tile for a coalesced third phase was measured too and is slower
          // (4095: 50 / 47, 5460: 50 / 48, 3640: 58 / 56): the extra barrier and tile pass cost more than the sectors save.
          if (K == 1 && (INV || SH::LL <= 1)) {
            w_st<true>(reinterpret_cast<longlong2*>(dst + w_out_pos<SH::LL, SH::L, INV>(l, j, 0, h)), make_longlong2(a, b));
          } else {
            w_st<K == 1>(dst + (size_t)w_out_pos<SH::LL, SH::L, INV>(l, j, 0, h) * k, a);
            w_st<K == 1>(dst + (size_t)w_out_pos<SH::LL, SH::L, INV>(l, j, 1, h) * k, b);
          }
        } else {
          w_st<K == 1>(dst + (size_t)w_out_pos<SH::LL, SH::L, INV>(l, j, 0, h) * k, v[h][2 * j]);
          w_st<K == 1>(dst + (size_t)w_out_pos<SH::LL, SH::L, INV>(l, j, 1, h) * k, v[h][2 * j + 1]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ m = 2^a p^e (p_b^e_b): a group of L lanes per element
template <class SH, bool INV, class AR, int K>
__device__ __forceinline__ void w1_body(typename AR::IO* __restrict__ y, const int64_t batch, const int k_rt, const int limb,
                                        const WConsts<typename AR::T, SH::NC>& C, const int bid, const int nblk)
{
  typedef typename AR::T T;
  const int k = K ? K : k_rt;
  const AR A(C.mod);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = lane & (SH::LW - 1), sub = lane >> SH::LL;
  T lt[1][kWLaneRows];
#pragma unroll
  for (int i = 0; i < kWLaneRows; i++) lt[0][i] = C.lane_tw[i * 32 + lane];
  const int64_t nwt = (batch + SH::GPW - 1) / SH::GPW;      // warp-tasks: GPW elements each
  for (int64_t wt = (int64_t)bid * (kWThreads / 32) + warp; wt < nwt; wt += (int64_t)nblk * (kWThreads / 32)) {
    const int64_t e = wt * SH::GPW + sub;
    const bool live = e < batch;
    typename AR::IO* ebase = y + (size_t)(live ? e : 0) * SH::N * k + limb;
    T v[1][SH::D2];
    w_load<AR, K == 1>(v[0], ebase + (size_t)l * k, (size_t)SH::L * k, live, C.mod.q);
    w_finish<SH, INV, AR, K>(v, l, lt, C, A, ebase, k, live);
  }
}

template <class SH, bool INV, class AR, int K>
__global__ void __launch_bounds__(kWThreads, AR::kZq ? SH::MINB : wc_minb<SH, INV>())
k_fused_w1(typename AR::IO* __restrict__ y, const int64_t batch, const int k_rt, const int limb,
           const __grid_constant__ WConsts<typename AR::T, SH::NC> C)
{
  w1_body<SH, INV, AR, K>(y, batch, k_rt, limb, C, (int)blockIdx.x, (int)gridDim.x);
}

// ------------------------------------------------------------------ m = 2^a (p_a^e_a) (p_b^e_b) p_c^e_c: two phases, on-chip tile
// (u32 words for Zq; complex doubles with a quarter of the elements per CTA, in dynamic shared memory)
template <class SH, class AR> struct WTile {
  static constexpr int EPB = AR::kZq ? SH::EPB : (SH::EPB >= 4 ? SH::EPB / 4 : 1);
  static constexpr size_t BYTES = (size_t)EPB * SH::ROWS * SH::RS * sizeof(typename AR::T);
};

template <class SH, bool INV, class AR, int K>
__device__ __forceinline__ void w2_body(typename AR::IO* __restrict__ y, const int64_t batch, const int k_rt, const int limb,
                                        const WConsts<typename AR::T, SH::NC>& C, const int bid, const int nblk)
{
  typedef typename SH::PC PC;
  typedef typename AR::T T;
  constexpr int EPB = WTile<SH, AR>::EPB, COLS = SH::COLS, ROWS = SH::ROWS, RS = SH::RS, N = SH::N;
  const int k = K ? K : k_rt;
  extern __shared__ __align__(16) unsigned char w2_smem[];
  T* tile = reinterpret_cast<T*>(w2_smem);
  const AR A(C.mod);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = lane & (SH::LW - 1), sub = lane >> SH::LL;
  T lt[SH::H][kWLaneRows];
  auto load_lane_table = [&]() {
#pragma unroll
    for (int h = 0; h < SH::H; h++)
#pragma unroll
      for (int i = 0; i < kWLaneRows; i++) lt[h][i] = C.lane_tw[(h * kWLaneRows + i) * 32 + lane];
  };
  if constexpr (AR::kZq) load_lane_table();      // complex: 40 registers per column half, loaded where phase 2 starts instead
  const int64_t ngroups = (batch + EPB - 1) / EPB;
  for (int64_t g = bid; g < ngroups; g += nblk) {
    const int64_t e0 = g * EPB;
    const int cnt = (int)(batch - e0 < EPB ? batch - e0 : EPB);
    // ---------------- phase 1: the last prime power; thread-task = (element slot, column), coefficients at stride COLS
    for (int t = threadIdx.x; t < cnt * COLS; t += kWThreads) {
      const int slot = t / COLS, col = t - slot * COLS;
      T v[ROWS];
      w_load<AR, K == 1>(v, y + ((size_t)(e0 + slot) * N + col) * k + limb, (size_t)COLS * k, true, C.mod.q);
      pp_line<PC, INV, 1, SH::OFF_C>(v, 0, C, A);
      const int cd = col / SH::SUB, cx = col - cd * SH::SUB;
      T* dst = tile + slot * ROWS * RS + cd * SH::SRS + cx;
#pragma unroll
      for (int i = 0; i < ROWS; i++) dst[i * RS] = v[i];
    }
    __syncthreads();
    // ---------------- phase 1b (four odd prime powers): the PD axis in place in the tile; thread-task = (slot, ic, x)
    if constexpr (SH::DD > 1) {
      for (int t = threadIdx.x; t < cnt * ROWS * SH::SUB; t += kWThreads) {
        const int sr = t / SH::SUB, x = t - sr * SH::SUB;      // sr = slot * ROWS + ic
        T* line = tile + sr * RS + x;
        T v[SH::DD];
#pragma unroll
        for (int i = 0; i < SH::DD; i++) v[i] = line[i * SH::SRS];
        pp_line<typename SH::PD, INV, 1, SH::OFF_D>(v, 0, C, A);
#pragma unroll
        for (int i = 0; i < SH::DD; i++) line[i * SH::SRS] = v[i];
      }
      __syncthreads();
    }
    // ---------------- phase 2: group-task = (element slot, ic, id); L lanes x D2 values, middle lines + network
    constexpr int BLOCKS = ROWS * SH::DD;      // phase-2 blocks per element
    const int ntask = cnt * BLOCKS;
    if constexpr (!AR::kZq) load_lane_table();
    for (int t0 = warp * SH::GPW; t0 < ntask; t0 += (kWThreads / 32) * SH::GPW) {      // warp-uniform trip count
      const int t = t0 + sub;
      const bool live = t < ntask;
      const int tt = live ? t : 0;
      const int slot = tt / BLOCKS, blk = tt - slot * BLOCKS;      // blk = ic * DD + id: the block's position in the element
      const int row = blk / SH::DD, bd = blk - row * SH::DD;
      const T* srow = tile + (slot * ROWS + row) * RS + bd * SH::SRS + l;
      T v[SH::H][SH::D2];
#pragma unroll
      for (int h = 0; h < SH::H; h++)
#pragma unroll
        for (int i = 0; i < SH::D2; i++) v[h][i] = srow[i * SH::L + h * SH::LW];
      typename AR::IO* dst = y + ((size_t)(e0 + slot) * N + (size_t)blk * SH::SUB) * k + limb;
      w_finish<SH, INV, AR, K>(v, l, lt, C, A, dst, k, live);
    }
    __syncthreads();
  }
}

template <class SH, bool INV, class AR, int K>
__global__ void __launch_bounds__(kWThreads, AR::kZq ? SH::MINB : wc_minb<SH, INV>())
k_fused_w2(typename AR::IO* __restrict__ y, const int64_t batch, const int k_rt, const int limb,
           const __grid_constant__ WConsts<typename AR::T, SH::NC> C)
{
  w2_body<SH, INV, AR, K>(y, batch, k_rt, limb, C, (int)blockIdx.x, (int)gridDim.x);
}

// ------------------------------------------------------------------ tupSize > 1: KL limbs of the same elements in ONE launch
// A limb's CTA uses 8 of every 8 k bytes, i.e. a part of every 32-byte sector of its elements; launched limb after limb, each
// launch moves the batch ~6x through HBM (sector fetch for the load, read-fill + write-back for the partial store; ncu: 5.9x).
// Here block b works on limb b % KL of element group b / KL, so the KL CTAs that share the sectors of a group are dispatched
// together: the first fetch serves all of them from L2 and their partial writes merge there before the sectors go back.  Each limb
// keeps its own compile-time constant offsets (one copy of the body per limb, selected once per CTA).
template <class T, int NC, int KL>
struct WConstsM { WConsts<T, NC> l[KL]; };

template <class SH, bool INV, class AR, int KL>
__global__ void __launch_bounds__(kWThreads, AR::kZq ? SH::MINB : wc_minb<SH, INV>())
k_fused_wm(typename AR::IO* __restrict__ y, const int64_t batch, const int k_rt, const int limb0,
           const __grid_constant__ WConstsM<typename AR::T, SH::NC, KL> CM)
{
  const int t = (int)blockIdx.x % KL, bid = (int)blockIdx.x / KL, nblk = (int)gridDim.x / KL;
#define LOLB_WM_CASE(I)                                                                                          \
  case I:                                                                                                        \
    if constexpr (I < KL) {                                                                                      \
      if constexpr (SH::TWO_PHASE) w2_body<SH, INV, AR, 0>(y, batch, k_rt, limb0 + I, CM.l[I < KL ? I : 0], bid, nblk);   \
      else w1_body<SH, INV, AR, 0>(y, batch, k_rt, limb0 + I, CM.l[I < KL ? I : 0], bid, nblk);                  \
    }                                                                                                            \
    break;
  switch (t) {
    LOLB_WM_CASE(0) LOLB_WM_CASE(1) LOLB_WM_CASE(2) LOLB_WM_CASE(3)
    default: break;
  }
#undef LOLB_WM_CASE
}

// ------------------------------------------------------------------ host: constants from the plan's root tables

enum WClass { WC_NONE = 0, WC_S, WC_M, WC_S6, WC_C };

inline WClass w_class(uint64_t q, int pmax)
{
  if (2 * (uint64_t)pmax * q * q < ((uint64_t)1 << 32)) return WC_S;
  if (pmax > 6 && 2 * (uint64_t)6 * q * q + 4 * q < ((uint64_t)1 << 32)) return WC_S6;      // rows cut into pieces of 6 terms
  if ((q & 1) && 2 * (uint64_t)pmax * q < ((uint64_t)1 << 32)) return WC_M;
  return WC_NONE;
}

// the field the constants are computed in on the host, and the type they are stored as
struct FieldZq {
  typedef uint64_t V;
  typedef uint32_t Out;
  uint64_t q;
  V one() const { return 1; }
  V mul(V a, V b) const { return mulmod64(a, b, q); }
  V add(V a, V b) const { return (a + b) % q; }
  V sub(V a, V b) const { return (a + q - b) % q; }
  V neg(V a) const { return (q - a) % q; }
  V half() const { return (q + 1) / 2; }      // q odd
  Out out(V a) const { return (uint32_t)a; }
};
struct FieldC {
  typedef std::complex<double> V;
  typedef double2 Out;
  V one() const { return V(1.0, 0.0); }
  V mul(V a, V b) const { return V(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real()); }
  V add(V a, V b) const { return a + b; }
  V sub(V a, V b) const { return a - b; }
  V neg(V a) const { return -a; }
  V half() const { return V(0.5, 0.0); }
  Out out(V a) const { return make_double2(a.real(), a.imag()); }
};

struct RootTab {      // root table of one prime power for one limb (forward or inverse roots), canonical
  const std::vector<int64_t>* tab;
  int k, limb;
  int64_t pp;
  uint64_t q;
  uint64_t operator()(int64_t j) const
  {
    int64_t v = (*tab)[(size_t)(((j % pp) + pp) % pp) * k + limb] % (int64_t)q;
    if (v < 0) v += q;
    return (uint64_t)v;
  }
};
struct RootTabC {     // the same over C: the plan's cis tables (CRTrans.hs:88-95)
  const std::vector<lolb_complex>* tab;
  int k, limb;
  int64_t pp;
  std::complex<double> operator()(int64_t j) const
  {
    const lolb_complex c = (*tab)[(size_t)(((j % pp) + pp) % pp) * k + limb];
    return std::complex<double>(c.real, c.imag);
  }
};

// constants of prime power (p, e) into out[0 .. n_consts): layout of PPT; `scale` multiplies the inverse block matrices
template <class FLD, class TAB>
void build_pp_consts(int p, int e, bool inverse, bool sym, const FLD& f, const TAB& T, typename FLD::V scale, typename FLD::Out* out)
{
  typedef typename FLD::V V;
  const int d = p - 1, R = e - 1;
  const int64_t mp = ipow64(p, e - 1);
  size_t o = 0;
  if (LOLB_W_SYM && sym && e == 1 && p >= 5) {      // a prime: the half-size cosine / sine blocks of pp_prime_sym (same storage: (p-1)^2 words)
    const int h = (p - 1) / 2;
    auto w = [&](int64_t j) { return inverse ? T(-j) : T(j); };      // w_p^j from either table
    const V s = inverse ? scale : f.one();
    for (int k = 1; k <= h; k++)
      for (int c = 1; c <= h; c++) {
        const V a = w((int64_t)k * c), b = w(-(int64_t)k * c);
        out[(k - 1) * h + (c - 1)] = f.out(f.mul(s, f.mul(f.half(), f.add(a, b))));
        out[h * h + (k - 1) * h + (c - 1)] = f.out(f.mul(s, f.mul(f.half(), f.sub(a, b))));
      }
    out[2 * h * h] = f.out(s);
    for (int i = 2 * h * h + 1; i < d * d; i++) out[i] = f.out(f.sub(f.one(), f.one()));
    return;
  }
  for (int64_t i0 = 0; i0 < mp; i0++)
    for (int r = 0; r < d; r++)
      for (int c = 0; c < d; c++) {
        V v;
        if (!inverse) {        // crtTwiddle(i0, r) . w_p^((r+1) c)
          const V tw = i0 ? T(digit_rev(p, R, i0) * (r + 1)) : f.one();
          v = f.mul(tw, T(mp * (((int64_t)(r + 1) * c) % p)));
        } else {               // (w^-(r (c+1)) - w^(c+1)) . crtTwiddle(i0, c) . scale, T = inverse roots (crt.cpp:369-398)
          const V tw = i0 ? T(digit_rev(p, R, i0) * (c + 1)) : f.one();
          const V mat = f.sub(T(mp * (((int64_t)r * (c + 1)) % p)), T(mp * (p - c - 1)));
          v = f.mul(f.mul(tw, mat), scale);
        }
        out[o++] = f.out(v);
      }
  for (int dig = 0; dig < R; dig++) {
    const int64_t nhi = ipow64(p, R - 1 - dig), stride = ipow64(p, dig + 1);
    for (int64_t hi = 0; hi < nhi; hi++)
      for (int r = 0; r < p; r++)
        for (int a = 0; a < p; a++) {
          const int x = inverse ? a : r;      // the twiddled index: outputs (forward, after dftp) / inputs (inverse, before dftp)
          const V tw = (hi && x) ? T(digit_rev(p, R - 1 - dig, hi) * x * stride) : f.one();
          out[o++] = f.out(f.mul(tw, T(mp * (((int64_t)r * a) % p))));
        }
  }
}

// per-lane constants of the network for 2^a (a >= 2): out[2 column halves][kWLaneRows][32], the group pattern repeated across
// the warp.  LOG = a - 1 column bits, LL = min(LOG, 5) of them lane bits; with a = 7 the top bit is the column half h.
template <class FLD, class TAB>
void build_lane_table(int a_exp, bool inverse, const FLD& f, const TAB& T, typename FLD::Out* out)
{
  typedef typename FLD::V V;
  const int LOG = a_exp >= 2 ? a_exp - 1 : 0, LL = LOG > 5 ? 5 : LOG, LW = 1 << LL, H = (1 << LOG) / LW;
  for (int i = 0; i < 2 * kWLaneRows * 32; i++) out[i] = f.out(f.one());
  if (LOG == 0) return;
  const int NT = H == 1 ? LL - 1 : LL;      // rounds with twiddles among the lane rounds (the top column bit carries none)
  for (int h = 0; h < H; h++) {
    typename FLD::Out* o = out + (size_t)h * kWLaneRows * 32;
    for (int lane = 0; lane < 32; lane++) {
      const int l = lane & (LW - 1);
      // dftTwiddle of the round on column bit r for the butterfly whose higher column bits are `hi_bits` (crt.cpp:92-106)
      auto round_tw = [&](int r, int hi_bits) -> V {
        return hi_bits ? T(digit_rev(2, LOG - 1 - r, hi_bits) * ((int64_t)2 << r)) : f.one();
      };
      if (!inverse) {
        const int col = h * LW + l;
        o[0 * 32 + lane] = f.out(col ? T(digit_rev(2, LOG, col)) : f.one());      // crtTwiddle of the lane's column
        for (int r = 0; r < NT; r++) {
          const V tw = round_tw(r, (h << (LL - 1 - r)) | (l >> (r + 1)));
          // lanes whose bit r is set hold (t, u) instead of (u, t): they multiply (t - u) by -tw
          o[(1 + r) * 32 + lane] = f.out(((l >> r) & 1) ? f.neg(tw) : tw);
        }
      } else {
        for (int r = 1; r < NT; r++)
          o[r * 32 + lane] = f.out(round_tw(r, (h << (LL - 1 - r)) | ((l >> r) & ((1 << (LL - 1 - r)) - 1))));
        const V tw0 = round_tw(0, (h << (LL - 1)) | (l & ((1 << (LL - 1)) - 1)));
        const int col = 2 * (l & (LW / 2 - 1)) + h * LW;
        const V ca = col ? T(digit_rev(2, LOG, col)) : f.one(), cb = T(digit_rev(2, LOG, col + 1));
        o[6 * 32 + lane] = f.out(ca);
        o[7 * 32 + lane] = f.out(f.mul(ca, tw0));
        o[8 * 32 + lane] = f.out(cb);
        o[9 * 32 + lane] = f.out(f.neg(f.mul(cb, tw0)));
      }
    }
  }
}

void w_mod_consts(uint64_t q, int cls, WMod* M)
{
  M->q = (uint32_t)q; M->q2 = (uint32_t)(2 * q); M->one = 1; M->r2 = 0;
  M->r0 = (uint32_t)((((uint64_t)1) << 32) / q);
  if (cls == WC_M) {
    uint32_t inv = (uint32_t)q;                       // q * inv == 1 mod 2^3 initially (q odd)
    for (int i = 0; i < 5; i++) inv *= 2u - (uint32_t)q * inv;
    M->r0 = 0u - inv;
    M->one = (uint32_t)((((uint64_t)1) << 32) % q);
    M->r2 = (uint32_t)((((uint64_t)M->one) << 32) % q);
  }
}

inline uint32_t w_mont(uint32_t c, uint64_t q) { return (uint32_t)((((uint64_t)c) << 32) % q); }

// the shape list: one entry per instantiated kernel family
struct WShapeId { int a; int pa, ea, pb, eb, pd, ed, pc, ec; };

template <class SH>
constexpr WShapeId shape_id() { return WShapeId{SH::A, SH::PA::p, SH::PA::e, SH::PB::p, SH::PB::e, SH::PD::p, SH::PD::e, SH::PC::p, SH::PC::e}; }

#ifndef LOLB_W27_MINB
#define LOLB_W27_MINB 4      // measured on B200, CRT / CRT^-1 of HBM peak: 4 -> 95 % / 94 %, 6 -> 87 % / 88 %, 8 -> 87 % / 87 %.  CTAs of 128 threads per SM the register allocation must allow (tuning: tools/build_variant.py)
#endif
#ifndef LOLB_W81_MINB
#define LOLB_W81_MINB 3      // 2 -> 72 % / 75 %, 3 -> 81 % / 83 %, 4 -> 71 % / 84 %
#endif
typedef WShape<6, PPT<3, 3>, PPNone, PPNone, PPNone, 1, LOLB_W27_MINB> SH_64_27;     // m = 1728  (n = 576)
typedef WShape<6, PPT<3, 4>, PPNone, PPNone, PPNone, 1, LOLB_W81_MINB> SH_64_81;     // m = 5184  (n = 1728)
// Two-phase shapes: CTAs of 128 threads per SM asked for (the last WShape argument).  Measured on B200 after the half-size prime
// blocks, % of HBM CRT / CRT^-1 at 4 -> 8 CTAs/SM: m = 2912 / 8737 71 / 72 -> 81 / 82 (10: 66 / 66), / 3144961 61 / 58 -> 66 / 64;
// m = 728 / 8737 77 / 79 -> 86 / 89; m = 3640 / 14561 71 / 74 -> 74 / 76; m = 5824 / 3144961 62 / 59 -> 68 / 65; m = 11648 63 / 59 ->
// 69 / 65; m = 5460 58 / 55 -> 61 / 58; m = 2016 / 2017 63 / 63 -> 82 / 84 (10: 86 / 88); m = 4095: 6 is best (56 / 49; 8: 47 / 46).
typedef WShape<5, PPT<7, 1>, PPNone, PPNone, PPT<13, 1>, 4, 8> SH_32_7_13;           // m = 2912  (n = 1152)
typedef WShape<3, PPT<7, 1>, PPNone, PPNone, PPT<13, 1>, 16, 8> SH_8_7_13;           // m = 728   (n = 288)
typedef WShape<3, PPT<5, 1>, PPT<7, 1>, PPNone, PPT<13, 1>, 8, 8> SH_8_5_7_13;       // m = 3640  (n = 1152)
typedef WShape<5, PPT<3, 2>, PPNone, PPNone, PPT<7, 1>, 4, 10> SH_32_9_7;            // m = 2016  (n = 576)
typedef WShape<6, PPT<7, 1>, PPNone, PPNone, PPT<13, 1>, 2, 8> SH_64_7_13;           // m = 5824  (n = 2304; lol-apps tunnel benchmark ring H1)
typedef WShape<7, PPT<7, 1>, PPNone, PPNone, PPT<13, 1>, 1, 8> SH_128_7_13;  // m = 11648 (n = 4608; Twace-Embed / tunnel H0): two column halves per lane
typedef WShape<2, PPT<3, 1>, PPT<5, 1>, PPT<7, 1>, PPT<13, 1>, 8, 8> SH_4_3_5_7_13;      // m = 5460 (n = 1152; tunnel H4): 7 in the tile
typedef WShape<0, PPT<3, 2>, PPT<5, 1>, PPT<7, 1>, PPT<13, 1>, 4, 6> SH_9_5_7_13;        // m = 4095 (n = 1728; tunnel H5): odd index, no network

typedef WShape<6, PPT<7, 1>, PPNone, PPNone, PPNone, 1, 8> SH_64_7;      // m = 448 (n = 192): the plaintext-side ring H1 = F64*F7 of the tunnel benchmark / HomomPRF example

constexpr int kNumShapes = 11;

template <class T>
struct FusedWT {
  int shape = -1;
  bool ok_fwd = false, ok_inv = false;
  std::vector<int> cls;                   // WClass per limb
  std::vector<std::vector<T>> cf, ci;     // per limb: flat constants (already in the limb's representation)
  std::vector<WMod> mod;
  T* d_lane = nullptr;                    // [k][2 directions][2 column halves][kWLaneRows][32]
  std::vector<T> h_lane;
};
typedef FusedWT<uint32_t> FusedW;
typedef FusedWT<double2> FusedWC;      // the complex plans (tensorCRTC / tensorCRTInvC)

bool shape_matches(const lolb_plan* pl, const WShapeId& id, int* pmax)
{
  std::vector<PrimeExponent> want;
  if (id.a > 0) want.push_back({2, (hShort_t)id.a});
  if (id.pa > 1) want.push_back({(hShort_t)id.pa, (hShort_t)id.ea});
  if (id.pb > 1) want.push_back({(hShort_t)id.pb, (hShort_t)id.eb});
  if (id.pd > 1) want.push_back({(hShort_t)id.pd, (hShort_t)id.ed});
  if (id.pc > 1) want.push_back({(hShort_t)id.pc, (hShort_t)id.ec});
  if (want.size() != pl->pe.size()) return false;
  for (size_t i = 0; i < want.size(); i++)
    if (want[i].prime != pl->pe[i].prime || want[i].exponent != pl->pe[i].exponent) return false;
  *pmax = id.pc > 1 ? id.pc : id.pd > 1 ? id.pd : id.pb > 1 ? id.pb : id.pa;
  return true;
}

const WShapeId kShapeIds[kNumShapes] = {shape_id<SH_64_27>(), shape_id<SH_64_81>(), shape_id<SH_32_7_13>(),
                                        shape_id<SH_8_7_13>(), shape_id<SH_8_5_7_13>(), shape_id<SH_32_9_7>(), shape_id<SH_64_7_13>(),
                                        shape_id<SH_128_7_13>(), shape_id<SH_4_3_5_7_13>(), shape_id<SH_9_5_7_13>(), shape_id<SH_64_7>()};

// host-side constants of one plan (no CUDA calls): shared by fused_w_select and the device-free emulation
template <class FW, class FLD, class MKTAB>
void fill_consts(const lolb_plan* pl, FW* F, int t, int dir, const FLD& f, typename FLD::V scale, const MKTAB& mktab)
{
  const bool sym = F->cls[t] != WC_S;      // = AR::kSym of the limb's arithmetic class
  const WShapeId& id = kShapeIds[F->shape];
  const int npe = (int)pl->pe.size(), first_odd = id.a > 0 ? 1 : 0;
  auto& out = dir ? F->ci[t] : F->cf[t];
  for (int i = first_odd; i < npe; i++) {      // order = (PA, PB, PD, PC) = the plan's odd prime powers in order
    const int p = pl->pe[i].prime, e = pl->pe[i].exponent;
    const size_t n_m1 = (size_t)ipow64(p, e - 1) * (p - 1) * (p - 1);
    size_t n_w = 0;
    for (int dig = 0; dig < e - 1; dig++) n_w += (size_t)ipow64(p, e - 2 - dig) * p * p;
    const size_t at = out.size();
    out.resize(at + n_m1 + n_w);
    build_pp_consts(p, e, dir != 0, sym, f, mktab(i, ipow64(p, e)), i == first_odd ? scale : f.one(), out.data() + at);      // mhat^-1 rides on the first odd axis
  }
  if (id.a >= 2)
    build_lane_table(id.a, dir != 0, f, mktab(0, ipow64(2, id.a)), F->h_lane.data() + ((size_t)t * 2 + dir) * 2 * kWLaneRows * 32);
}

int find_shape(const lolb_plan* pl, int* pmax)
{
  for (int s = 0; s < kNumShapes; s++)
    if (shape_matches(pl, kShapeIds[s], pmax)) return s;
  return -1;
}

int build_fused_w(const lolb_plan* pl, FusedW* F)
{
  F->shape = -1;
  if (pl->kind != PLAN_RQ) return LOLB_OK;
  int pmax = 0;
  F->shape = find_shape(pl, &pmax);
  if (F->shape < 0) return LOLB_OK;
  const int k = pl->k;
  F->cls.assign(k, WC_NONE);
  for (int t = 0; t < k; t++) {
    F->cls[t] = w_class((uint64_t)pl->qs[t], pmax);
    if (F->cls[t] == WC_NONE) { F->shape = -1; return LOLB_OK; }
  }
  if (k > 1) {      // one launch serves several limbs (k_fused_wm): all limbs in the widest class any of them needs (each class contains the narrower ones' moduli)
    auto rank = [](int c) { return c == WC_S ? 0 : c == WC_S6 ? 1 : 2; };
    int widest = F->cls[0];
    for (int t = 1; t < k; t++) if (rank(F->cls[t]) > rank(widest)) widest = F->cls[t];
    for (int t = 0; t < k; t++) F->cls[t] = widest;
  }
  const int npe = (int)pl->pe.size();
  F->ok_fwd = pl->ru.size() == (size_t)npe;
  F->ok_inv = pl->ruinv.size() == (size_t)npe && (int)pl->mhatinv.size() == k;
  F->cf.assign(k, {});
  F->ci.assign(k, {});
  F->mod.assign(k, WMod{});
  F->h_lane.assign((size_t)k * 2 * 2 * kWLaneRows * 32, 1u);
  for (int t = 0; t < k; t++) {
    const uint64_t q = (uint64_t)pl->qs[t];
    w_mod_consts(q, F->cls[t], &F->mod[t]);
    for (int dir = 0; dir < 2; dir++) {
      if (dir == 0 ? !F->ok_fwd : !F->ok_inv) continue;
      const auto& tabs = dir ? pl->ruinv : pl->ru;
      uint64_t scale = 1;
      if (dir) { int64_t s = pl->mhatinv[t] % (int64_t)q; if (s < 0) s += q; scale = (uint64_t)s; }
      fill_consts(pl, F, t, dir, FieldZq{q}, scale, [&](int i, int64_t pp) { return RootTab{&tabs[i], k, t, pp, q}; });
      if (F->cls[t] == WC_M) {
        uint32_t* lane = F->h_lane.data() + ((size_t)t * 2 + dir) * 2 * kWLaneRows * 32;
        for (auto& c : (dir ? F->ci[t] : F->cf[t])) c = w_mont(c, q);
        for (int i = 0; i < 2 * kWLaneRows * 32; i++) lane[i] = w_mont(lane[i], q);
      }
    }
  }
  return LOLB_OK;
}

int build_fused_wc(const lolb_plan* pl, FusedWC* F)
{
  F->shape = -1;
  if (pl->kind != PLAN_C) return LOLB_OK;
  int pmax = 0;
  F->shape = find_shape(pl, &pmax);
  if (F->shape < 0) return LOLB_OK;
  const int k = pl->k, npe = (int)pl->pe.size();
  F->cls.assign(k, WC_C);
  F->ok_fwd = pl->cru.size() == (size_t)npe;
  F->ok_inv = pl->cruinv.size() == (size_t)npe;
  F->cf.assign(k, {});
  F->ci.assign(k, {});
  F->mod.assign(k, WMod{});
  F->h_lane.assign((size_t)k * 2 * 2 * kWLaneRows * 32, make_double2(1.0, 0.0));
  for (int t = 0; t < k; t++)
    for (int dir = 0; dir < 2; dir++) {
      if (dir == 0 ? !F->ok_fwd : !F->ok_inv) continue;
      const auto& tabs = dir ? pl->cruinv : pl->cru;
      const std::complex<double> scale = dir ? std::complex<double>(pl->c_mhatinv[t].x, pl->c_mhatinv[t].y) : std::complex<double>(1.0, 0.0);
      fill_consts(pl, F, t, dir, FieldC{}, scale, [&](int i, int64_t pp) { return RootTabC{&tabs[i], k, t, pp}; });
    }
  return LOLB_OK;
}

template <class SH, bool INV, class AR, int K>
int launch_w(const lolb_plan* pl, const FusedWT<typename AR::T>* F, int limb, typename AR::IO* y, int64_t batch, cudaStream_t st)
{
  typedef typename AR::T T;
  WConsts<T, SH::NC> C;
  C.mod = F->mod[limb];
  C.lane_tw = F->d_lane + ((size_t)limb * 2 + (INV ? 1 : 0)) * 2 * kWLaneRows * 32;
  const std::vector<T>& src = INV ? F->ci[limb] : F->cf[limb];
  if ((int)src.size() != SH::OFF_C + SH::PC::n_consts) { set_error("fused_w: constant layout mismatch"); return LOLB_ERR_ARG; }
  for (size_t i = 0; i < src.size(); i++) C.c[i] = src[i];
  int64_t grid = (int64_t)pl->num_sms * (AR::kZq ? SH::MINB : (wc_minb<SH, INV>() < 2 ? 2 : wc_minb<SH, INV>()));
  if constexpr (SH::TWO_PHASE) {
    constexpr int EPB = WTile<SH, AR>::EPB;
    constexpr size_t BYTES = WTile<SH, AR>::BYTES;
    const int64_t groups = (batch + EPB - 1) / EPB;
    if (grid > groups) grid = groups;
    if (BYTES > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k_fused_w2<SH, INV, AR, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BYTES);
      if (e != cudaSuccess) return cuda_fail(e, "k_fused_w2 shared memory");
    }
    k_fused_w2<SH, INV, AR, K><<<(int)grid, kWThreads, BYTES, st>>>(y, batch, pl->k, limb, C);
  } else {
    const int64_t ctas = (batch + (kWThreads / 32) * SH::GPW - 1) / ((kWThreads / 32) * SH::GPW);
    if (grid > ctas) grid = ctas;
    k_fused_w1<SH, INV, AR, K><<<(int)grid, kWThreads, 0, st>>>(y, batch, pl->k, limb, C);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_fused_w");
  count_launch();
  return LOLB_OK;
}

template <class SH>
int launch_shape_c(const lolb_plan* pl, const FusedWC* F, bool inverse, double2* y, int64_t batch, cudaStream_t st)
{
  for (int t = 0; t < pl->k; t++) {
    const bool k1 = pl->k == 1;
    int rc = inverse ? (k1 ? launch_w<SH, true, WC, 1>(pl, F, t, y, batch, st) : launch_w<SH, true, WC, 0>(pl, F, t, y, batch, st))
                     : (k1 ? launch_w<SH, false, WC, 1>(pl, F, t, y, batch, st) : launch_w<SH, false, WC, 0>(pl, F, t, y, batch, st));
    if (rc) return rc;
  }
  return LOLB_OK;
}

// KL limbs (limb0 .. limb0 + KL - 1) of a Z_q plan in one launch
template <class SH, bool INV, class AR, int KL>
int launch_wm(const lolb_plan* pl, const FusedW* F, int limb0, int64_t* y, int64_t batch, cudaStream_t st)
{
  typedef typename AR::T T;
  static WConstsM<T, SH::NC, KL> CM;      // filled under the lock below: too large for the stack of a deep call chain
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < KL; i++) {
    const int limb = limb0 + i;
    CM.l[i].mod = F->mod[limb];
    CM.l[i].lane_tw = F->d_lane + ((size_t)limb * 2 + (INV ? 1 : 0)) * 2 * kWLaneRows * 32;
    const std::vector<T>& src = INV ? F->ci[limb] : F->cf[limb];
    if ((int)src.size() != SH::OFF_C + SH::PC::n_consts) { set_error("fused_w: constant layout mismatch"); return LOLB_ERR_ARG; }
    for (size_t j = 0; j < src.size(); j++) CM.l[i].c[j] = src[j];
  }
  int64_t groups_in_flight = (int64_t)pl->num_sms * SH::MINB / KL;
  if (groups_in_flight < 1) groups_in_flight = 1;
  size_t smem = 0;
  if constexpr (SH::TWO_PHASE) {
    constexpr int EPB = WTile<SH, AR>::EPB;
    smem = WTile<SH, AR>::BYTES;
    const int64_t groups = (batch + EPB - 1) / EPB;
    if (groups_in_flight > groups) groups_in_flight = groups;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(k_fused_wm<SH, INV, AR, KL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return cuda_fail(e, "k_fused_wm shared memory");
    }
  } else {
    const int64_t ctas = (batch + (kWThreads / 32) * SH::GPW - 1) / ((kWThreads / 32) * SH::GPW);
    if (groups_in_flight > ctas) groups_in_flight = ctas;
  }
  k_fused_wm<SH, INV, AR, KL><<<(int)(groups_in_flight * KL), kWThreads, smem, st>>>(y, batch, pl->k, limb0, CM);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_fused_wm");
  count_launch();
  return LOLB_OK;
}

template <class SH, class AR>
int launch_limbs(const lolb_plan* pl, const FusedW* F, bool inverse, int limb0, int kl, int64_t* y, int64_t batch, cudaStream_t st)
{
  const bool k1 = pl->k == 1;
  switch (kl) {
    case 1: return inverse ? (k1 ? launch_w<SH, true, AR, 1>(pl, F, limb0, y, batch, st) : launch_w<SH, true, AR, 0>(pl, F, limb0, y, batch, st))
                           : (k1 ? launch_w<SH, false, AR, 1>(pl, F, limb0, y, batch, st) : launch_w<SH, false, AR, 0>(pl, F, limb0, y, batch, st));
    case 2: return inverse ? launch_wm<SH, true, AR, 2>(pl, F, limb0, y, batch, st) : launch_wm<SH, false, AR, 2>(pl, F, limb0, y, batch, st);
    case 3: return inverse ? launch_wm<SH, true, AR, 3>(pl, F, limb0, y, batch, st) : launch_wm<SH, false, AR, 3>(pl, F, limb0, y, batch, st);
    case 4: return inverse ? launch_wm<SH, true, AR, 4>(pl, F, limb0, y, batch, st) : launch_wm<SH, false, AR, 4>(pl, F, limb0, y, batch, st);
  }
  return LOLB_ERR_ARG;
}

template <class SH>
int launch_shape(const lolb_plan* pl, const FusedW* F, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  const char* me = getenv("LOLB_W_MULTI");      // 0: one launch per limb (tests, A/B)
  const bool multi = !me || atoi(me) != 0;
  const int k = pl->k, step = multi ? 4 : 1;
  for (int t = 0; t < k; t += step) {
    const int kl = k - t < step ? k - t : step;
    int rc;
    if (F->cls[t] == WC_S6) {
      if constexpr (SH::PC::p > 6 || SH::PD::p > 6 || SH::PB::p > 6 || SH::PA::p > 6) rc = launch_limbs<SH, WS6>(pl, F, inverse, t, kl, y, batch, st);
      else rc = LOLB_FUSED_UNAVAILABLE;
    } else if (F->cls[t] == WC_M) rc = launch_limbs<SH, WM>(pl, F, inverse, t, kl, y, batch, st);
    else rc = launch_limbs<SH, WS>(pl, F, inverse, t, kl, y, batch, st);
    if (rc) return rc;
  }
  return LOLB_OK;
}

// ------------------------------------------------------------------ device-free emulation (CPU tests of the constants, the
// line code above compiled for the host, and a lane-by-lane replica of the network with the same ownership rules)
template <class SH, bool INV, class AR>
void emulate_shape(const FusedWT<typename AR::T>* F, int limb, int k, typename AR::IO* y)
{
  typedef typename AR::T T;
  WConsts<T, SH::NC> C;
  C.mod = F->mod[limb];
  C.lane_tw = nullptr;
  const std::vector<T>& src = INV ? F->ci[limb] : F->cf[limb];
  for (size_t i = 0; i < src.size(); i++) C.c[i] = src[i];
  const T* lane_tab = F->h_lane.data() + ((size_t)limb * 2 + (INV ? 1 : 0)) * 2 * kWLaneRows * 32;
  const AR A(C.mod);
  constexpr int L = SH::L, D2 = SH::D2, NP = SH::NP, COLS = SH::COLS, ROWS = SH::ROWS;
  std::vector<T> tile((size_t)SH::N);
  auto ld = [&](int j) -> T {
    if constexpr (AR::kZq) {
      const T q = C.mod.q;
      int64_t r = y[(size_t)j * k + limb] % (int64_t)q;
      if (r < 0) r += q;
      return (T)r;
    } else return y[(size_t)j * k + limb];
  };
  // phase 1
  for (int col = 0; col < COLS; col++) {
    T v[ROWS];
    for (int i = 0; i < ROWS; i++) v[i] = ld(i * COLS + col);
    pp_line<typename SH::PC, INV, 1, SH::OFF_C>(v, 0, C, A);
    for (int i = 0; i < ROWS; i++) tile[(size_t)i * COLS + col] = v[i];
  }
  // phase 1b: the in-tile axis
  constexpr int DD = SH::DD, SUB = SH::SUB;
  if (DD > 1) {
    for (int ic = 0; ic < ROWS; ic++)
      for (int x = 0; x < SUB; x++) {
        T v[DD];
        for (int i = 0; i < DD; i++) v[i] = tile[(size_t)ic * COLS + i * SUB + x];
        pp_line<typename SH::PD, INV, 1, SH::OFF_D>(v, 0, C, A);
        for (int i = 0; i < DD; i++) tile[(size_t)ic * COLS + i * SUB + x] = v[i];
      }
  }
  // phase 2, one block (ic, id) at a time, all LW lanes of the group in lockstep, H column halves per lane
  constexpr int LW = SH::LW, LL = SH::LL, H = SH::H;
  constexpr bool TOP_IN_LANES = H == 1;
  constexpr int NT = TOP_IN_LANES ? LL - 1 : LL;
  for (int row = 0; row < ROWS * DD; row++) {
    T v[H][LW][D2];
    for (int h = 0; h < H; h++)
      for (int l = 0; l < LW; l++) {
        for (int i = 0; i < D2; i++) v[h][l][i] = tile[(size_t)row * SUB + i * L + h * LW + l];
        for (int ib = 0; ib < SH::DB; ib++) pp_line<typename SH::PA, INV, 1, SH::OFF_A>(v[h][l], ib * SH::DA, C, A);
        for (int ia = 0; ia < SH::DA; ia++) pp_line<typename SH::PB, INV, SH::DA, SH::OFF_B>(v[h][l], ia, C, A);
      }
    auto lt = [&](int h, int r, int l) { return lane_tab[(h * kWLaneRows + r) * 32 + l]; };      // lane = l (group 0)
    auto top_round = [&]() {
      for (int l = 0; l < LW; l++)
        for (int i = 0; i < D2; i++) {
          const T u = v[0][l][i], t = v[H - 1][l][i];
          v[0][l][i] = A.fold(A.add(u, t));
          v[H - 1][l][i] = A.fold(A.sub(u, t));
        }
    };
    auto round = [&](int h, bool trivial, int bit, int row_tw) {
      T nv[LW][D2];
      for (int l = 0; l < LW; l++) {
        const bool hi = (l >> bit) & 1;
        const int partner = l ^ (1 << bit);
        for (int j = 0; j < NP; j++) {
          const T keep = hi ? v[h][l][2 * j + 1] : v[h][l][2 * j];
          const bool phi_ = (partner >> bit) & 1;
          const T recv = phi_ ? v[h][partner][2 * j] : v[h][partner][2 * j + 1];      // what the partner sends
          const T tw = lt(h, row_tw, l);
          if (trivial) {
            const T u = hi ? recv : keep, t = hi ? keep : recv;
            nv[l][2 * j] = A.fold(A.add(u, t));
            nv[l][2 * j + 1] = A.fold(A.sub(u, t));
          } else if (!INV) {
            nv[l][2 * j] = A.fold(A.add(keep, recv));
            nv[l][2 * j + 1] = A.red(A.mul(tw, A.sub(keep, recv)));
          } else {
            const T t = A.red(A.mul(tw, hi ? keep : recv));
            const T u = hi ? recv : keep;
            nv[l][2 * j] = A.fold(A.add(u, t));
            nv[l][2 * j + 1] = A.fold(A.sub(u, t));
          }
        }
      }
      for (int l = 0; l < LW; l++) for (int i = 0; i < D2; i++) v[h][l][i] = nv[l][i];
    };
    if (LL >= 1) {
      if (H == 2 && INV) top_round();
      for (int h = 0; h < H; h++) {
        if (!INV) {
          for (int l = 0; l < LW; l++) for (int i = 0; i < D2; i++) v[h][l][i] = A.red(A.mul(lt(h, 0, l), v[h][l][i]));
          for (int r = 0; r < NT; r++) round(h, false, r, 1 + r);
          if (TOP_IN_LANES) round(h, true, LL - 1, 0);
        } else {
          if (TOP_IN_LANES && LL >= 2) round(h, true, LL - 1, 0);
          for (int r = NT - 1; r >= 1; r--) round(h, false, r, r);
          T nv[LW][D2];
          for (int l = 0; l < LW; l++) {
            const bool hi = l & 1;
            const int partner = l ^ 1;
            for (int j = 0; j < NP; j++) {
              const T keep = hi ? v[h][l][2 * j + 1] : v[h][l][2 * j];
              const T recv = (partner & 1) ? v[h][partner][2 * j] : v[h][partner][2 * j + 1];
              const T t = hi ? keep : recv, u = hi ? recv : keep;
              nv[l][2 * j] = A.red(A.mad(A.mul(lt(h, 6, l), u), lt(h, 7, l), t));
              nv[l][2 * j + 1] = A.red(A.mad(A.mul(lt(h, 8, l), u), lt(h, 9, l), t));
            }
          }
          for (int l = 0; l < LW; l++) for (int i = 0; i < D2; i++) v[h][l][i] = nv[l][i];
        }
      }
      if (H == 2 && !INV) top_round();
    }
    for (int h = 0; h < H; h++)
      for (int l = 0; l < LW; l++)
        for (int j = 0; j < NP; j++)
          for (int s = 0; s < 2; s++) {
            int pos;
            if (LL == 0) pos = 2 * j + s;
            else if (!INV) pos = (2 * j + (l & 1)) * L + ((s << (LL - 1)) | (l >> 1)) + LW * h;
            else pos = (2 * j + (l >> (LL - 1))) * L + 2 * (l & (LW / 2 - 1)) + s + LW * h;
            if constexpr (AR::kZq) y[((size_t)row * SUB + pos) * k + limb] = (int64_t)A.canon(v[h][l][2 * j + s]);
            else y[((size_t)row * SUB + pos) * k + limb] = v[h][l][2 * j + s];
          }
  }
}

template <class SH>
void emulate_dispatch(const FusedW* F, bool inverse, int k, int64_t* y)
{
  for (int t = 0; t < k; t++) {
    const bool m = F->cls[t] == WC_M;
    if (F->cls[t] == WC_S6) {
      if (inverse) emulate_shape<SH, true, WS6>(F, t, k, y); else emulate_shape<SH, false, WS6>(F, t, k, y);
      continue;
    }
    if (inverse) { if (m) emulate_shape<SH, true, WM>(F, t, k, y); else emulate_shape<SH, true, WS>(F, t, k, y); }
    else { if (m) emulate_shape<SH, false, WM>(F, t, k, y); else emulate_shape<SH, false, WS>(F, t, k, y); }
  }
}

#define W_FOR_SHAPE(F, CALL)                                       \
  switch ((F)->shape) {                                            \
    case 0: { typedef SH_64_27 SH; CALL; } break;                  \
    case 1: { typedef SH_64_81 SH; CALL; } break;                  \
    case 2: { typedef SH_32_7_13 SH; CALL; } break;                \
    case 3: { typedef SH_8_7_13 SH; CALL; } break;                 \
    case 4: { typedef SH_8_5_7_13 SH; CALL; } break;               \
    case 5: { typedef SH_32_9_7 SH; CALL; } break;                 \
    case 6: { typedef SH_64_7_13 SH; CALL; } break;                \
    case 7: { typedef SH_128_7_13 SH; CALL; } break;               \
    case 8: { typedef SH_4_3_5_7_13 SH; CALL; } break;             \
    case 9: { typedef SH_9_5_7_13 SH; CALL; } break;               \
    case 10: { typedef SH_64_7 SH; CALL; } break;                  \
    default: break;                                                \
  }

}  // namespace

// ------------------------------------------------------------------ the kernels are instantiated in five translation units
// (a single one took six minutes of nvcc): Z_q shapes 0-1 / 2-5 / 6-9 in parts 0 / 1 / 2, complex shapes 0-4 / 5-9 in parts 3 / 4.
// Part 0 also holds the host side (constants, selection, dispatch, the device-free emulation).  Plans cross the parts as void*.
int fused_w_launch_p0(const lolb_plan* pl, const void* F, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);
int fused_w_launch_p1(const lolb_plan* pl, const void* F, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);
int fused_w_launch_p2(const lolb_plan* pl, const void* F, bool inverse, int64_t* y, int64_t batch, cudaStream_t st);
int fused_w_launch_c_p3(const lolb_plan* pl, const void* F, bool inverse, double2* y, int64_t batch, cudaStream_t st);
int fused_w_launch_c_p4(const lolb_plan* pl, const void* F, bool inverse, double2* y, int64_t batch, cudaStream_t st);

#define W_CASE(N, SHT, CALL) case N: { typedef SHT SH; CALL; } break;
#if LOLB_W_PART == 0
#define W_PART_FN fused_w_launch_p0
#define W_PART_CASES(CALL) W_CASE(0, SH_64_27, CALL) W_CASE(1, SH_64_81, CALL) W_CASE(10, SH_64_7, CALL)
#elif LOLB_W_PART == 1
#define W_PART_FN fused_w_launch_p1
#define W_PART_CASES(CALL) W_CASE(2, SH_32_7_13, CALL) W_CASE(3, SH_8_7_13, CALL) W_CASE(4, SH_8_5_7_13, CALL) W_CASE(5, SH_32_9_7, CALL)
#elif LOLB_W_PART == 2
#define W_PART_FN fused_w_launch_p2
#define W_PART_CASES(CALL) W_CASE(6, SH_64_7_13, CALL) W_CASE(7, SH_128_7_13, CALL) W_CASE(8, SH_4_3_5_7_13, CALL) W_CASE(9, SH_9_5_7_13, CALL)
#elif LOLB_W_PART == 3
#define W_PART_FN fused_w_launch_c_p3
#define W_PART_CASES(CALL) W_CASE(0, SH_64_27, CALL) W_CASE(1, SH_64_81, CALL) W_CASE(2, SH_32_7_13, CALL) W_CASE(3, SH_8_7_13, CALL) W_CASE(4, SH_8_5_7_13, CALL) W_CASE(10, SH_64_7, CALL)
#else
#define W_PART_FN fused_w_launch_c_p4
#define W_PART_CASES(CALL) W_CASE(5, SH_32_9_7, CALL) W_CASE(6, SH_64_7_13, CALL) W_CASE(7, SH_128_7_13, CALL) W_CASE(8, SH_4_3_5_7_13, CALL) W_CASE(9, SH_9_5_7_13, CALL)
#endif

#if LOLB_W_PART <= 2
int W_PART_FN(const lolb_plan* pl, const void* Fv, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  const FusedW* F = (const FusedW*)Fv;
  int rc = LOLB_FUSED_UNAVAILABLE;
  switch (F->shape) { W_PART_CASES(rc = launch_shape<SH>(pl, F, inverse, y, batch, st)) default: break; }
  return rc;
}
#else
int W_PART_FN(const lolb_plan* pl, const void* Fv, bool inverse, double2* y, int64_t batch, cudaStream_t st)
{
  const FusedWC* F = (const FusedWC*)Fv;
  int rc = LOLB_FUSED_UNAVAILABLE;
  switch (F->shape) { W_PART_CASES(rc = launch_shape_c<SH>(pl, F, inverse, y, batch, st)) default: break; }
  return rc;
}
#endif

#if LOLB_W_PART == 0
namespace {
template <class FW, class BUILD>
int select_impl(lolb_plan* pl, void** slot, BUILD build)
{
  FW* F = (FW*)*slot;
  FW tmp;
  int rc = build(pl, &tmp);
  if (rc) return rc;
  if (tmp.shape < 0) {
    if (F) { if (F->d_lane) cudaFree(F->d_lane); delete F; *slot = nullptr; }
    return LOLB_OK;
  }
  if (!F) { F = new FW(); *slot = F; }
  auto* old = F->d_lane;
  *F = tmp;
  F->d_lane = nullptr;
  if (old) cudaFree(old);
  LOLB_CUDA(cudaMalloc((void**)&F->d_lane, F->h_lane.size() * sizeof(F->h_lane[0])));
  LOLB_CUDA(cudaMemcpy(F->d_lane, F->h_lane.data(), F->h_lane.size() * sizeof(F->h_lane[0]), cudaMemcpyHostToDevice));
  return LOLB_OK;
}
template <class FW>
void release_impl(void* slot)
{
  FW* F = (FW*)slot;
  if (!F) return;
  if (F->d_lane) cudaFree(F->d_lane);
  delete F;
}
}  // namespace

int fused_w_select(lolb_plan* pl, void** slot) { return select_impl<FusedW>(pl, slot, build_fused_w); }
int fused_wc_select(lolb_plan* pl, void** slot) { return select_impl<FusedWC>(pl, slot, build_fused_wc); }
void fused_w_release(void* slot) { release_impl<FusedW>(slot); }
void fused_wc_release(void* slot) { release_impl<FusedWC>(slot); }

bool fused_w_available(const void* slot, bool inverse)
{
  const FusedW* F = (const FusedW*)slot;
  return F && F->shape >= 0 && (inverse ? F->ok_inv : F->ok_fwd);
}
bool fused_wc_available(const void* slot, bool inverse)
{
  const FusedWC* F = (const FusedWC*)slot;
  return F && F->shape >= 0 && (inverse ? F->ok_inv : F->ok_fwd);
}

int fused_w_crt(const lolb_plan* pl, const void* slot, bool inverse, int64_t* y, int64_t batch, cudaStream_t st)
{
  const FusedW* F = (const FusedW*)slot;
  if (!fused_w_available(slot, inverse)) return LOLB_FUSED_UNAVAILABLE;
  if (batch <= 0) return LOLB_OK;
  return (F->shape <= 1 || F->shape == 10) ? fused_w_launch_p0(pl, F, inverse, y, batch, st)
       : F->shape <= 5 ? fused_w_launch_p1(pl, F, inverse, y, batch, st) : fused_w_launch_p2(pl, F, inverse, y, batch, st);
}

int fused_wc_crt(const lolb_plan* pl, const void* slot, bool inverse, double2* y, int64_t batch, cudaStream_t st)
{
  const FusedWC* F = (const FusedWC*)slot;
  if (!fused_wc_available(slot, inverse)) return LOLB_FUSED_UNAVAILABLE;
  if (batch <= 0) return LOLB_OK;
  return (F->shape <= 4 || F->shape == 10) ? fused_w_launch_c_p3(pl, F, inverse, y, batch, st) : fused_w_launch_c_p4(pl, F, inverse, y, batch, st);
}

}  // namespace lolb

// Device-free: the fused_w schedule (host-built constants, the same line code compiled for the host, a lane-by-lane
// replica of the exchange network) applied to ONE ring element in host memory, tables derived like lolb_plan_create_rq
// does (ZqBasic.hs:144-171).  Test hook for `-m "not gpu"` (tests/test_fused_w_emulation.py); returns LOLB_ERR_ARG when
// the index has no fused_w kernel and LOLB_ERR_NO_CRT when Z_q has no CRT of that index.
extern "C" int lolb_fused_w_emulate(const PrimeExponent* peArr, hShort_t sizeOfPE, hShort_t tupSize, const hInt_t* qs, int inverse,
                                    hInt_t* y)
{
  using namespace lolb;
  if (!peArr || !qs || !y) { set_error("lolb_fused_w_emulate: NULL argument"); return LOLB_ERR_ARG; }
  lolb_plan pl;
  pl.kind = PLAN_RQ;
  int rc = plan_build_common(&pl, peArr, sizeOfPE, tupSize);
  if (rc) return rc;
  pl.qs.assign(qs, qs + tupSize);
  rc = plan_derive_rq_roots(&pl);
  if (rc) return rc;
  FusedW F;
  rc = build_fused_w(&pl, &F);
  if (rc) return rc;
  if (F.shape < 0) { set_error("lolb_fused_w_emulate: no fused_w kernel for this index / modulus"); return LOLB_ERR_ARG; }
  W_FOR_SHAPE(&F, emulate_dispatch<SH>(&F, inverse != 0, tupSize, y));
  return LOLB_OK;
}

// The same for the complex plans: roots derived like lolb_plan_create_c does (CRTrans.hs:88-95); y = [phi(m)][tupSize] complex.
extern "C" int lolb_fused_w_emulate_c(const PrimeExponent* peArr, hShort_t sizeOfPE, hShort_t tupSize, int inverse, lolb_complex* y)
{
  using namespace lolb;
  if (!peArr || !y) { set_error("lolb_fused_w_emulate_c: NULL argument"); return LOLB_ERR_ARG; }
  lolb_plan pl;
  pl.kind = PLAN_C;
  int rc = plan_build_common(&pl, peArr, sizeOfPE, tupSize);
  if (rc) return rc;
  plan_derive_c_roots(&pl);
  FusedWC F;
  rc = build_fused_wc(&pl, &F);
  if (rc) return rc;
  if (F.shape < 0) { set_error("lolb_fused_w_emulate_c: no fused_w kernel for this index"); return LOLB_ERR_ARG; }
  for (int t = 0; t < tupSize; t++) {
    if (inverse) { W_FOR_SHAPE(&F, (emulate_shape<SH, true, WC>(&F, t, tupSize, (double2*)y))); }
    else { W_FOR_SHAPE(&F, (emulate_shape<SH, false, WC>(&F, t, tupSize, (double2*)y))); }
  }
  return LOLB_OK;
}

#else
}  // namespace lolb
#endif  // LOLB_W_PART == 0
