// plan.cu -- per-(m, moduli) plans: factor bookkeeping, root / diagonal tables on the device and the
// pass lists the generic engine executes.
//
// What the reference does per call in C (crt.cpp:459-560: ppDFT / ppcrt schedule, bitrev per twiddle) and
// per (m, r) in Haskell (CPP.hs:422-442 root tables; ZqBasic.hs:144-171 omega and mhat^-1;
// Tensor.hs:319-337 gCRT) happens here once.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "lolb_internal.cuh"
#include "numtheory.h"

namespace lolb {

uint64_t hash_bytes(const void* p, size_t bytes, uint64_t seed)
{
  const unsigned char* b = (const unsigned char*)p;
  uint64_t h = 0x9E3779B97F4A7C15ull ^ seed;
  size_t i = 0;
  for (; i + 8 <= bytes; i += 8) {
    uint64_t w;
    memcpy(&w, b + i, 8);
    h = (h ^ w) * 0xff51afd7ed558ccdull;
    h ^= h >> 32;
  }
  for (; i < bytes; i++) { h = (h ^ b[i]) * 0x100000001b3ull; }
  return h;
}

void* plan_ws(const lolb_plan* pl, cudaStream_t st, size_t bytes)
{
  std::lock_guard<std::mutex> lock(pl->ws_mu);
  lolb_plan::WsSlot* slot = nullptr;
  for (auto& s : pl->ws) if (s.st == st) { slot = &s; break; }
  if (!slot) { pl->ws.push_back(lolb_plan::WsSlot{st, nullptr, 0}); slot = &pl->ws.back(); }
  if (bytes <= slot->bytes && slot->p) return slot->p;
  if (slot->p) {      // kernels of this stream may still use the old block
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cuda_fail(e, "cudaStreamSynchronize(workspace)"); return nullptr; }
    cudaFree(slot->p);
    slot->p = nullptr; slot->bytes = 0;
  }
  cudaError_t e = cudaMalloc(&slot->p, bytes);
  if (e != cudaSuccess) { slot->p = nullptr; cuda_fail(e, "cudaMalloc(workspace)"); return nullptr; }
  slot->bytes = bytes;
  return slot->p;
}

// the auxiliary stream (highest priority: it runs the finishing half of a split schedule) and eight untimed events of
// the slot of stream `st`, created on first use.  The events live as long as the plan: the returned pointer stays valid
// because slots are only ever appended while ws_mu is held and std::vector growth moves the handles, not the objects
// they name -- callers copy the handles before the next plan_ws* call.
int plan_ws_aux(const lolb_plan* pl, cudaStream_t st, cudaStream_t* aux, cudaEvent_t** events)
{
  std::lock_guard<std::mutex> lock(pl->ws_mu);
  lolb_plan::WsSlot* slot = nullptr;
  for (auto& s : pl->ws) if (s.st == st) { slot = &s; break; }
  if (!slot) { pl->ws.push_back(lolb_plan::WsSlot{st, nullptr, 0}); slot = &pl->ws.back(); }
  if (!slot->aux) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    LOLB_CUDA(cudaStreamCreateWithPriority(&slot->aux, cudaStreamNonBlocking, hi));
    for (auto& e : slot->ev) LOLB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  *aux = slot->aux;
  *events = slot->ev;
  return LOLB_OK;
}

int plan_reserve_stage(const lolb_plan* pl, size_t bytes)
{
  if (bytes <= pl->stage_bytes) return LOLB_OK;
  if (pl->d_stage) { LOLB_CUDA(cudaDeviceSynchronize()); cudaFree(pl->d_stage); pl->d_stage = nullptr; pl->stage_bytes = 0; }
  LOLB_CUDA(cudaMalloc(&pl->d_stage, bytes));
  pl->stage_bytes = bytes;
  return LOLB_OK;
}

static int64_t phi_pp(const PrimeExponent& pe) { return (int64_t)(pe.prime - 1) * ipow64(pe.prime, pe.exponent - 1); }

int plan_build_common(lolb_plan* pl, const PrimeExponent* pe, int npe, int k)
{
  if (npe < 0 || (npe > 0 && !pe) || k < 1 || k > kMaxLimbs) { set_error("plan: bad pe / tupSize"); return LOLB_ERR_ARG; }
  pl->pe.assign(pe, pe + npe);
  pl->k = k;
  int64_t m = 1, n = 1, rad = 1;
  int prev = 1;
  for (int i = 0; i < npe; i++) {
    const int p = pe[i].prime, e = pe[i].exponent;
    if (p < 2 || e < 1 || p <= prev) { set_error("plan: prime powers must be increasing primes with exponent >= 1"); return LOLB_ERR_ARG; }
    for (int d = 2; d * d <= p; d++) if (p % d == 0) { set_error("plan: composite 'prime'"); return LOLB_ERR_ARG; }
    prev = p;
    m *= ipow64(p, e);
    n *= phi_pp(pe[i]);
    if (p != 2) rad *= p;
    if (m > (int64_t)1 << 30 || n > (int64_t)1 << 30) { set_error("plan: index too large"); return LOLB_ERR_ARG; }
  }
  pl->m = m; pl->n = (int32_t)n; pl->odd_rad = rad;
  if (cudaGetDevice(&pl->device) != cudaSuccess) pl->device = 0;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device) == cudaSuccess && sms > 0) pl->num_sms = sms;

  // line pass lists (tensor.h:39-74): prime power i runs at (lts*p^(e-1), rts) with axis length p-1
  for (int kind = PASS_L; kind <= PASS_GAUSS; kind++) {
    PassList& a = pl->line[kind];
    PassList& f = pl->line_folded[kind];
    a.count = f.count = 0;
    a.needs_alt = f.needs_alt = (kind == PASS_GAUSS);
    int64_t rts = 1;
    int64_t cru_off = 0;
    for (int i = 0; i < npe; i++) {
      const int p = pe[i].prime, e = pe[i].exponent;
      if (p != 2) {
        if (a.count >= kMaxPasses) { set_error("plan: too many passes"); return LOLB_ERR_ARG; }
        Pass ps{};
        ps.kind = kind; ps.p = p; ps.d = p - 1; ps.R = (int32_t)rts;
        ps.rustride = (int32_t)ipow64(p, e - 1);
        ps.tab = (int32_t)cru_off;
        a.pass[a.count++] = ps;
        ps.R = (int32_t)(rts * k);
        f.pass[f.count++] = ps;
      }
      cru_off += ipow64(p, e);
      rts *= phi_pp(pe[i]);
    }
  }
  return LOLB_OK;
}

// ---------------------------------------------------------------- CRT pass lists + tables, generic in the table element

// Builds the pass list of tensorFuserCRT(ppcrt | ppcrtinv) and the per-limb table block
//   [ root table of pp 0 | root table of pp 1 | ... | diagonal tables in pass order ]
// `root(i, j, limb)` returns entry j of the root table of prime power i for a limb.
template <class W, class RootFn, class OneFn>
static int build_crt_dir(const lolb_plan* pl, bool inverse, RootFn root, OneFn one, PassList* out_pl,
                         std::vector<W>* out_tab, int32_t* out_stride)
{
  const int npe = (int)pl->pe.size(), k = pl->k;
  PassList& PL = *out_pl;
  PL.count = 0; PL.needs_alt = 0;
  std::vector<int64_t> ru_off(npe);
  int64_t off = 0;
  for (int i = 0; i < npe; i++) { ru_off[i] = off; off += ipow64(pl->pe[i].prime, pl->pe[i].exponent); }

  // diagonal tables are described first (per pass), materialised per limb afterwards
  struct Diag { int pp; int64_t off; int dim; bool crt; int p; int digits; int64_t twstride; };
  std::vector<Diag> diags;
  auto push = [&](const Pass& ps) -> int {
    if (PL.count >= kMaxPasses) { set_error("plan: too many passes"); return LOLB_ERR_ARG; }
    PL.pass[PL.count++] = ps;
    return LOLB_OK;
  };
  auto push_diag = [&](int pp, int dim, int64_t R, bool crt, int p, int digits, int64_t twstride) -> int {
    Pass ps{};
    ps.kind = PASS_DIAG; ps.p = p; ps.d = dim; ps.R = (int32_t)R; ps.tab = (int32_t)off;
    diags.push_back(Diag{pp, off, dim, crt, p, digits, twstride});
    off += dim;
    return push(ps);
  };

  int64_t rts = 1;
  for (int i = 0; i < npe; i++) {
    const int p = pl->pe[i].prime, e = pl->pe[i].exponent;
    const int64_t mprime = ipow64(p, e - 1);
    const int64_t phi = (p - 1) * mprime;
    const int64_t rts_dft = rts * (p - 1);
    int rc = LOLB_OK;
    auto dense = [&](int kind, int d, int64_t R) {
      Pass ps{};
      ps.kind = kind; ps.p = p; ps.d = d; ps.R = (int32_t)R; ps.rustride = (int32_t)mprime; ps.tab = (int32_t)ru_off[i];
      if (!(kind == PASS_DFT && p == 2)) PL.needs_alt = 1;
      return push(ps);
    };
    if (!inverse) {
      // crt.cpp:518-538: crtp ; crtTwiddle ; ppDFT(p^(e-1)) at rts*(p-1), rustride p
      if (p != 2) rc = dense(PASS_CRT, p - 1, rts);
      if (!rc && mprime > 1) rc = push_diag(i, (int)phi, rts, true, p, e - 1, 1);
      for (int round = 0; !rc && round < e - 1; round++) {       // crt.cpp:476-485
        const int64_t Rr = rts_dft * ipow64(p, round);
        rc = dense(PASS_DFT, p, Rr);
        const int64_t dim = ipow64(p, e - 1 - round);
        if (!rc && dim / p > 1) rc = push_diag(i, (int)dim, Rr, false, p, e - 2 - round, (int64_t)p * ipow64(p, round));
      }
    } else {
      // crt.cpp:540-560: ppDFTInv ; crtTwiddle ; crtpinv, all with inverse roots
      for (int round = 0; !rc && round < e - 1; round++) {       // crt.cpp:505-515
        const int64_t Rr = rts_dft * ipow64(p, e - 2 - round);
        const int64_t dim = ipow64(p, round + 1);
        if (dim / p > 1) rc = push_diag(i, (int)dim, Rr, false, p, round, mprime / ipow64(p, round));
        if (!rc) rc = dense(PASS_DFT, p, Rr);
      }
      if (!rc && mprime > 1) rc = push_diag(i, (int)phi, rts, true, p, e - 1, 1);
      if (!rc && p != 2) rc = dense(PASS_CRTINV, p - 1, rts);
    }
    if (rc) return rc;
    rts *= phi;
  }

  const int64_t stride = off;
  if (stride * (int64_t)k > ((int64_t)1 << 31)) { set_error("plan: tables too large"); return LOLB_ERR_ARG; }
  out_tab->assign((size_t)(stride * k), one());
  for (int limb = 0; limb < k; limb++) {
    W* T = out_tab->data() + (size_t)limb * stride;
    for (int i = 0; i < npe; i++) {
      const int64_t pp = ipow64(pl->pe[i].prime, pl->pe[i].exponent);
      for (int64_t j = 0; j < pp; j++) T[ru_off[i] + j] = root(i, j, limb);
    }
    for (const Diag& dg : diags) {
      const int p = dg.p;
      if (dg.crt) {
        // crt.cpp:35-81: entry i0*(p-1)+i1 <- ru[rev(i0)*(i1+1)], i0 >= 1
        const int64_t mprime = dg.dim / (p - 1);
        for (int64_t i0 = 1; i0 < mprime; i0++) {
          const int64_t rev = digit_rev(p, dg.digits, i0);
          for (int i1 = 0; i1 < p - 1; i1++) T[dg.off + i0 * (p - 1) + i1] = root(dg.pp, rev * (i1 + 1), limb);
        }
      } else {
        // crt.cpp:84-126: entry i0*p+i1 <- ru[rev(i0)*i1*twstride], i0 >= 1, i1 >= 1
        const int64_t mprime = dg.dim / p;
        for (int64_t i0 = 1; i0 < mprime; i0++) {
          const int64_t rev = digit_rev(p, dg.digits, i0);
          for (int i1 = 1; i1 < p; i1++) T[dg.off + i0 * p + i1] = root(dg.pp, rev * i1 * dg.twstride, limb);
        }
      }
    }
  }
  *out_stride = (int32_t)stride;
  return LOLB_OK;
}

template <class W>
static int upload(W** dptr, const std::vector<W>& host)
{
  if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
  if (host.empty()) return LOLB_OK;
  LOLB_CUDA(cudaMalloc((void**)dptr, host.size() * sizeof(W)));
  LOLB_CUDA(cudaMemcpy(*dptr, host.data(), host.size() * sizeof(W), cudaMemcpyHostToDevice));
  return LOLB_OK;
}

// ---------------------------------------------------------------- Zq

int plan_derive_rq_roots(lolb_plan* pl)
{
  const int npe = (int)pl->pe.size(), k = pl->k;
  std::vector<uint64_t> w(k);
  for (int t = 0; t < k; t++) {
    w[t] = principal_root((uint64_t)pl->m, (uint64_t)pl->qs[t]);
    if (w[t] == 0) return LOLB_ERR_NO_CRT;
  }
  pl->ru.assign(npe, {});
  pl->ruinv.assign(npe, {});
  for (int i = 0; i < npe; i++) {
    const int64_t pp = ipow64(pl->pe[i].prime, pl->pe[i].exponent);
    pl->ru[i].resize((size_t)pp * k);
    pl->ruinv[i].resize((size_t)pp * k);
    for (int t = 0; t < k; t++) {
      const uint64_t q = (uint64_t)pl->qs[t];
      const uint64_t base = powmod64(w[t], (uint64_t)(pl->m / pp), q);          // CPP.hs:427-431
      const uint64_t ibase = powmod64(base, (uint64_t)(pp - 1), q);              // base^-1 (order divides pp)
      uint64_t a = 1 % q, b = 1 % q;
      for (int64_t j = 0; j < pp; j++) {
        pl->ru[i][(size_t)j * k + t] = (int64_t)a;
        pl->ruinv[i][(size_t)j * k + t] = (int64_t)b;
        a = mulmod64(a, base, q);
        b = mulmod64(b, ibase, q);
      }
    }
  }
  pl->mhatinv.resize(k);
  const int64_t mhat = (pl->m % 2 == 0) ? pl->m / 2 : pl->m;                      // FactoredDefs.hs:374-376
  for (int t = 0; t < k; t++) {
    pl->mhatinv[t] = mod_inverse(pl->qs[t], mhat % pl->qs[t]);
    if (pl->mhatinv[t] == 0 && pl->qs[t] != 1) return LOLB_ERR_NO_CRT;
  }
  return LOLB_OK;
}

static uint32_t canon_u32(int64_t v, int64_t q)
{
  int64_t r = v % q;
  if (r < 0) r += q;
  return (uint32_t)r;
}

int plan_upload_rq_dir(lolb_plan* pl, bool inverse)
{
  const auto& tabs = inverse ? pl->ruinv : pl->ru;
  const int k = pl->k;
  std::vector<uint32_t> host;
  int32_t stride = 0;
  auto root = [&](int i, int64_t j, int limb) -> uint32_t { return canon_u32(tabs[i][(size_t)j * k + limb], pl->qs[limb]); };
  auto one = [&]() -> uint32_t { return 1u; };
  int rc = build_crt_dir<uint32_t>(pl, inverse, root, one, inverse ? &pl->crt_inv : &pl->crt_fwd, &host, &stride);
  if (rc) return rc;
  // Montgomery-form copy for engine_axis (odd moduli below 2^28: a row of <= 13 products then fits one REDC)
  std::vector<uint32_t> hm;
  bool mont_ok = true;
  for (int t = 0; t < k; t++) if (!(pl->qs[t] & 1) || pl->qs[t] >= ((int64_t)1 << 28)) mont_ok = false;
  if (mont_ok) {
    hm.resize(host.size());
    for (int t = 0; t < k; t++) {
      const uint64_t q = (uint64_t)pl->qs[t];
      for (int32_t j = 0; j < stride; j++) hm[(size_t)t * stride + j] = (uint32_t)((((uint64_t)host[(size_t)t * stride + j]) << 32) % q);
    }
  }
  {
    uint32_t** dm = inverse ? &pl->d_tab_inv_m : &pl->d_tab_fwd_m;
    int rcm = upload(dm, hm);          // empty vector frees a stale copy
    if (rcm) return rcm;
  }
  // q = 1 would make "1" non-canonical; moduli are >= 2 (checked at creation)
  if (inverse) {
    rc = upload(&pl->d_tab_inv, host);
    pl->tab_stride_inv = stride;
    for (int t = 0; t < k; t++) pl->zq_mhat.scale[t] = canon_u32(pl->mhatinv[t], pl->qs[t]);
    pl->has_inv = rc == LOLB_OK;
  } else {
    rc = upload(&pl->d_tab_fwd, host);
    pl->tab_stride_fwd = stride;
    pl->has_fwd = rc == LOLB_OK;
  }
  return rc;
}

int plan_upload_rq_gcrt(lolb_plan* pl)
{
  // Tensor.hs:264-337 with w_p^j = ru_i[j * p^(e-1)]; Kronecker product, first prime power fastest (indexK)
  const int npe = (int)pl->pe.size(), k = pl->k, n = pl->n;
  std::vector<int64_t> g((size_t)n * k), gi((size_t)n * k);
  for (int t = 0; t < k; t++) {
    const uint64_t q = (uint64_t)pl->qs[t];
    std::vector<uint64_t> vg(1, 1 % q), vgi(1, 1 % q);
    for (int i = 0; i < npe; i++) {
      const int p = pl->pe[i].prime, e = pl->pe[i].exponent;
      const int64_t mprime = ipow64(p, e - 1), phi = (p - 1) * mprime;
      std::vector<uint64_t> pg(p > 2 ? p - 1 : 1, 1 % q), pgi(p > 2 ? p - 1 : 1, 1 % q);
      if (p != 2) {
        auto wp = [&](int64_t j) -> uint64_t { return (uint64_t)canon_u32(pl->ru[i][(size_t)((j % p) * mprime) * k + t], (int64_t)q); };
        const uint64_t phat_inv = (uint64_t)mod_inverse((int64_t)q, p % (int64_t)q);
        for (int a = 0; a < p - 1; a++) {
          pg[a] = (1 + q - wp(a + 1)) % q;
          uint64_t s = 0;
          for (int j = 1; j <= p - 1; j++) s = (s + mulmod64((uint64_t)j % q, wp((int64_t)(a + 1) * (p - 1 - j)), q)) % q;
          pgi[a] = mulmod64(phat_inv, s, q);
        }
      }
      std::vector<uint64_t> ng(vg.size() * phi), ngi(vg.size() * phi);
      for (int64_t b = 0; b < phi; b++)
        for (size_t a = 0; a < vg.size(); a++) {
          const size_t idx = (size_t)b * vg.size() + a;
          const int64_t sel = p != 2 ? b % (p - 1) : 0;        // ppKron: index mod (p-1)
          ng[idx] = mulmod64(vg[a], pg[sel], q);
          ngi[idx] = mulmod64(vgi[a], pgi[sel], q);
        }
      vg.swap(ng); vgi.swap(ngi);
    }
    for (int j = 0; j < n; j++) { g[(size_t)j * k + t] = (int64_t)vg[j]; gi[(size_t)j * k + t] = (int64_t)vgi[j]; }
  }
  int rc = upload(&pl->d_gcrt, g);
  if (!rc) rc = upload(&pl->d_gcrtinv, gi);
  return rc;
}

// ---------------------------------------------------------------- complex

void plan_derive_c_roots(lolb_plan* pl)
{
  const int npe = (int)pl->pe.size(), k = pl->k;
  pl->cru.assign(npe, {});
  pl->cruinv.assign(npe, {});
  const double m = (double)pl->m;
  for (int i = 0; i < npe; i++) {
    const int64_t pp = ipow64(pl->pe[i].prime, pl->pe[i].exponent), step = pl->m / pp;
    pl->cru[i].resize((size_t)pp * k);
    pl->cruinv[i].resize((size_t)pp * k);
    for (int64_t j = 0; j < pp; j++) {
      // CRTrans.hs:94-95: cis (2*pi*i/m) with i = +-j*m/p^e
      const double ang = 2.0 * M_PI * (double)(j * step) / m;
      const double nang = 2.0 * M_PI * (double)(-(j * step)) / m;
      for (int t = 0; t < k; t++) {
        pl->cru[i][(size_t)j * k + t] = lolb_complex{cos(ang), sin(ang)};
        pl->cruinv[i][(size_t)j * k + t] = lolb_complex{cos(nang), sin(nang)};
      }
    }
  }
  const double mhat = (double)((pl->m % 2 == 0) ? pl->m / 2 : pl->m);
  for (int t = 0; t < kMaxLimbs; t++) pl->c_mhatinv[t] = make_double2(1.0 / mhat, 0.0);   // CRTrans.hs:91
}

int plan_upload_c_dir(lolb_plan* pl, bool inverse)
{
  const auto& tabs = inverse ? pl->cruinv : pl->cru;
  const int k = pl->k;
  std::vector<double2> host;
  int32_t stride = 0;
  auto root = [&](int i, int64_t j, int limb) -> double2 {
    const lolb_complex c = tabs[i][(size_t)j * k + limb];
    return make_double2(c.real, c.imag);
  };
  auto one = [&]() -> double2 { return make_double2(1.0, 0.0); };
  int rc = build_crt_dir<double2>(pl, inverse, root, one, inverse ? &pl->crt_inv : &pl->crt_fwd, &host, &stride);
  if (rc) return rc;
  if (inverse) { rc = upload(&pl->d_ctab_inv, host); pl->ctab_stride_inv = stride; pl->has_inv = rc == LOLB_OK; }
  else { rc = upload(&pl->d_ctab_fwd, host); pl->ctab_stride_fwd = stride; pl->has_fwd = rc == LOLB_OK; }
  return rc;
}

}  // namespace lolb
