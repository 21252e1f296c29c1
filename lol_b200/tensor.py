"""Host-side mirror of the reference's `Tensor` class for the B200 hot path.

`class Tensor t` (lol/Crypto/Lol/Cyclotomic/Tensor.hs:86-193) is the plug-in interface a
back end implements; `instance Tensor CT` (lol-cpp/Crypto/Lol/Cyclotomic/Tensor/CPP.hs:204-264)
is the C++ one.  The methods below keep the class's names and meaning for the operators that
live on this path, lifted from one ring element to a BATCH of them resident on the GPU:

    Tensor method              reference dispatch (CPP.hs)            here
    -------------------------  -------------------------------------  ---------------------------
    l / lInv                   basicDispatch dl / dlinv      :219-220  CudaTensorRq.l / lInv
    mulGPow / mulGDec          basicDispatch dmulgpow/dec    :222-223  .mulGPow / .mulGDec
    divGPow / divGDec          dispatchGInv                  :225-226  .divGPow / .divGDec  (None = Nothing)
    crtFuncs: scalarCRT        repl                          :229      .scalarCRT
              mulGCRT/divGCRT  cZipDispatch dmul <$> gCRT    :230-231  .mulGCRT / .divGCRT
              crt / crtInv     ctCRT / ctCRTInv              :232-233  .crt / .crtInv
    scalarPow                  scalarPow'                    :217      .scalarPow
    zipWithT (*)               (host SV.zipWith)             :257-260  .mul      (mulRq on the device)
    tGaussianDec               cDispatchGaussian             :239      CudaTensorReal.tGaussianDec
    gSqNormDec                 gSqNormDec'                   :243-244  CudaTensorReal.gSqNormDec / CudaTensorInt.gSqNormDec

Like the Haskell functions these are pure: each returns a new tensor unless `inplace=True`.
Tensors are torch CUDA tensors of shape [batch, n, k] (int64 / float64 / complex128) in the
reference's element layout; torch supplies device memory and the current stream, nothing else.
"""
from __future__ import annotations

import math

import torch

from . import capi
from .factored import odd_radical_fact, pps_fact, radical_fact, totient_fact


def _stream() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def _require_cuda(x: torch.Tensor, dtype: torch.dtype, n: int, k: int) -> int:
    if not x.is_cuda:
        raise capi.LolB200Error(capi.LOLB_ERR_ARG, "tensor must live on a CUDA device: this back end has no CPU path")
    if x.dtype != dtype or not x.is_contiguous():
        raise capi.LolB200Error(capi.LOLB_ERR_ARG, f"expected a contiguous {dtype} tensor")
    if x.dim() != 3 or x.shape[1] != n or x.shape[2] != k:
        raise capi.LolB200Error(capi.LOLB_ERR_ARG, f"expected shape [batch, {n}, {k}], got {tuple(x.shape)}")
    return int(x.shape[0])


class CudaTensorRq:
    """`Tensor` operations for index m over Z_q1 x ... x Z_qk (the `ZqBasic q Int64` tuples of Backend.hs:122-149)."""

    def __init__(self, m: int, qs, ru=None, ruinv=None, mhatinv=None):
        self.m = int(m)
        self.qs = [int(q) for q in qs]
        self.k = len(self.qs)
        self.pps = pps_fact(self.m)
        self.n = totient_fact(self.m)
        self.plan = capi.PlanRq(self.pps, self.qs, ru, ruinv, mhatinv)
        assert self.plan.n == self.n

    # -- helpers
    def _unary(self, name, x, inplace):
        b = _require_cuda(x, torch.int64, self.n, self.k)
        y = x if inplace else x.clone()
        capi.check(self.plan.op(name, y.data_ptr(), b, _stream()))
        return y

    # -- Tensor methods
    def l(self, x, inplace=False): return self._unary("L", x, inplace)
    def lInv(self, x, inplace=False): return self._unary("LInv", x, inplace)
    def mulGPow(self, x, inplace=False): return self._unary("GPow", x, inplace)
    def mulGDec(self, x, inplace=False): return self._unary("GDec", x, inplace)
    def crt(self, x, inplace=False): return self._unary("CRT", x, inplace)
    def crtInv(self, x, inplace=False): return self._unary("CRTInv", x, inplace)

    def _div(self, name, x, inplace):
        b = _require_cuda(x, torch.int64, self.n, self.k)
        y = x if inplace else x.clone()
        st = self.plan.op(name, y.data_ptr(), b, _stream())
        if st == capi.LOLB_ERR_NOT_INVERTIBLE:
            return None                       # CPP.hs:321-323: `Nothing`
        capi.check(st)
        return y

    def divGPow(self, x, inplace=False): return self._div("GInvPow", x, inplace)
    def divGDec(self, x, inplace=False): return self._div("GInvDec", x, inplace)

    def mul(self, a, b, inplace=False):
        """zipWithT (*) a b  (mul.cpp:27-30).  `b` may hold one element ([1, n, k]) broadcast over the batch."""
        ba = _require_cuda(a, torch.int64, self.n, self.k)
        bb = _require_cuda(b, torch.int64, self.n, self.k)
        y = a if inplace else a.clone()
        capi.check(self.plan.mul(y.data_ptr(), b.data_ptr(), ba, bb, _stream()))
        return y

    def crtMul(self, a, b, inplace=False):
        """crt a `zipWithT (*)` b in one pass: the ring product of a Pow-basis `a` with a CRT-basis `b`
        (Cyc.hs:276-297; crt.cpp:562-566 then mul.cpp:27-30).  `b` may hold one element broadcast over the batch."""
        ba = _require_cuda(a, torch.int64, self.n, self.k)
        bb = _require_cuda(b, torch.int64, self.n, self.k)
        y = a if inplace else a.clone()
        capi.check(self.plan.crt_mul(y.data_ptr(), b.data_ptr(), ba, bb, _stream()))
        return y

    def mulCrtInv(self, a, b, inplace=False):
        """crtInv (a `zipWithT (*)` b) in one pass (mul.cpp:27-30 then crt.cpp:569-581), e.g. the hint products of key
        switching followed by the change back to the powerful basis (SymmSHE.hs:302-314)."""
        ba = _require_cuda(a, torch.int64, self.n, self.k)
        bb = _require_cuda(b, torch.int64, self.n, self.k)
        y = a if inplace else a.clone()
        capi.check(self.plan.mul_crt_inv(y.data_ptr(), b.data_ptr(), ba, bb, _stream()))
        return y

    def _mul_by_plan_vector(self, x, inverse, inplace):
        b = _require_cuda(x, torch.int64, self.n, self.k)
        ptr = self.plan.gcrt_dev(inverse)
        if not ptr:
            raise capi.LolB200Error(capi.LOLB_ERR_NO_CRT, "no CRT basis over this modulus (crtFuncs = Nothing)")
        y = x if inplace else x.clone()
        capi.check(self.plan.mul(y.data_ptr(), ptr, b, 1, _stream()))
        return y

    def mulGCRT(self, x, inplace=False): return self._mul_by_plan_vector(x, False, inplace)
    def divGCRT(self, x, inplace=False): return self._mul_by_plan_vector(x, True, inplace)

    def scalarPow(self, r, batch=1, device="cuda"):
        """Tensor.hs:114 / CPP.hs:409-413: constant-term coefficient first, zeros elsewhere."""
        y = torch.zeros(batch, self.n, self.k, dtype=torch.int64, device=device)
        y[:, 0, :] = torch.as_tensor([int(v) % q for v, q in zip(self._per_limb(r), self.qs)], dtype=torch.int64, device=device)
        return y

    def scalarCRT(self, r, batch=1, device="cuda"):
        """CPP.hs:229 (`repl`): the scalar in every CRT slot."""
        row = torch.as_tensor([int(v) % q for v, q in zip(self._per_limb(r), self.qs)], dtype=torch.int64, device=device)
        return row.expand(batch, self.n, self.k).contiguous()

    def _per_limb(self, r):
        return list(r) if isinstance(r, (list, tuple)) else [r] * self.k

    # -- coefficient-wise maps around the transforms (fmapT lift / reduce / rescale, UCyc.hs:267-300; roundCoset, Prelude.hs:155-162)
    def lift(self, x):
        """fmapT lift: residues -> representatives in [-q/2, q/2), limb by limb."""
        b = _require_cuda(x, torch.int64, self.n, self.k)
        y = torch.empty_like(x)
        capi.check(self.plan.lift(x.data_ptr(), y.data_ptr(), b, _stream()))
        return y

    def reduce(self, z):
        """fmapT reduce: int64 [batch, n, 1] (or [batch, n, k]) -> residues modulo every q_t."""
        kz = int(z.shape[2]) if z.dim() == 3 else -1
        b = _require_cuda(z, torch.int64, self.n, kz)
        y = torch.empty((b, self.n, self.k), dtype=torch.int64, device=z.device)
        capi.check(self.plan.reduce(z.data_ptr(), kz, y.data_ptr(), b, _stream()))
        return y

    def rescaleDrop(self, x, drop=0):
        """rescalePow over `Rescale (a,b) b` (drop = 0) / `Rescale (a,b) a` (drop = k-1): [batch, n, k] -> [batch, n, k-1]."""
        b = _require_cuda(x, torch.int64, self.n, self.k)
        y = torch.empty((b, self.n, self.k - 1), dtype=torch.int64, device=x.device)
        capi.check(self.plan.rescale_drop(int(drop), x.data_ptr(), y.data_ptr(), b, _stream()))
        return y

    def rescaleMod(self, x, qs_new):
        """fmapT rescaleMod: limb t from modulus q_t to qs_new[t]."""
        b = _require_cuda(x, torch.int64, self.n, self.k)
        y = torch.empty_like(x)
        capi.check(self.plan.rescale_mod([int(q) for q in qs_new], x.data_ptr(), y.data_ptr(), b, _stream()))
        return y

    def roundCoset(self, e, zp=None):
        """errorCoset's rounding (zp given: residues modulo this tensor's moduli) or errorRounded's (zp None): float64 -> int64."""
        b = _require_cuda(e, torch.float64, self.n, self.k)
        if zp is not None:
            _require_cuda(zp, torch.int64, self.n, self.k)
        y = torch.empty(e.shape, dtype=torch.int64, device=e.device)
        capi.check(self.plan.round_coset(e.data_ptr(), zp.data_ptr() if zp is not None else 0, y.data_ptr(), b, _stream()))
        return y

    def apply_host(self, ops: str, y_host: torch.Tensor) -> torch.Tensor:
        """Host-buffer call (lolb_rq_apply_host): `y_host` is a pinned or pageable CPU tensor [batch, n, k],
        transformed in place through the chunked H2D -> kernels -> D2H pipeline."""
        if y_host.is_cuda or y_host.dtype != torch.int64 or not y_host.is_contiguous():
            raise capi.LolB200Error(capi.LOLB_ERR_ARG, "apply_host expects a contiguous int64 CPU tensor")
        capi.check(self.plan.apply_host(ops, y_host.data_ptr(), int(y_host.shape[0])))
        return y_host

    def apply_host_u32(self, ops: str, y_host: torch.Tensor) -> torch.Tensor:
        """The same pipeline over the narrow wire format (lolb_rq_apply_host_u32): `y_host` holds the residues as uint32 in an
        int32 / uint32 CPU tensor [batch, n, k]; half the bytes cross PCIe."""
        if y_host.is_cuda or y_host.element_size() != 4 or y_host.is_floating_point() or not y_host.is_contiguous():
            raise capi.LolB200Error(capi.LOLB_ERR_ARG, "apply_host_u32 expects a contiguous 32-bit integer CPU tensor")
        capi.check(self.plan.apply_host_u32(ops, y_host.data_ptr(), int(y_host.shape[0])))
        return y_host


class _CudaTensorPlain:
    dtype = None
    tag = None

    def __init__(self, m: int, k: int = 1):
        self.m = int(m)
        self.k = int(k)
        self.pps = pps_fact(self.m)
        self.n = totient_fact(self.m)
        self.plan = capi.PlanC(self.pps, self.k)

    def _unary(self, name, x, inplace):
        b = _require_cuda(x, self.dtype, self.n, self.k)
        y = x if inplace else x.clone()
        capi.check(self.plan.op(name + self.tag, y.data_ptr(), b, _stream()))
        return y

    def l(self, x, inplace=False): return self._unary("L", x, inplace)
    def lInv(self, x, inplace=False): return self._unary("LInv", x, inplace)


class CudaTensorInt(_CudaTensorPlain):
    """`Tensor` operations over Int64 (Backend.hs:219-265, the `...R` symbols)."""
    dtype = torch.int64
    tag = "R"

    def mulGPow(self, x, inplace=False): return self._unary("GPow", x, inplace)
    def mulGDec(self, x, inplace=False): return self._unary("GDec", x, inplace)

    def _div(self, name, x, inplace):
        b = _require_cuda(x, self.dtype, self.n, self.k)
        y = x if inplace else x.clone()
        ok = torch.empty(b, dtype=torch.int16, device=x.device)
        capi.check(self.plan.ginv_r(name, y.data_ptr(), ok.data_ptr(), b, _stream()))
        return y, ok          # ok[b] == 0  <=>  Nothing for element b

    def divGPow(self, x, inplace=False): return self._div("GInvPow", x, inplace)
    def divGDec(self, x, inplace=False): return self._div("GInvDec", x, inplace)

    def gSqNormDec(self, x):
        b = _require_cuda(x, self.dtype, self.n, self.k)
        out = torch.empty(b, self.k, dtype=torch.int64, device=x.device)
        capi.check(self.plan.normsq("R", x.data_ptr(), out.data_ptr(), b, _stream()))
        return out


class CudaTensorReal(_CudaTensorPlain):
    """`Tensor` operations over Double (Backend.hs:267-283: L, LInv, gSqNormDec, tGaussianDec)."""
    dtype = torch.float64
    tag = "Double"

    def gSqNormDec(self, x):
        b = _require_cuda(x, self.dtype, self.n, self.k)
        out = torch.empty(b, self.k, dtype=torch.float64, device=x.device)
        capi.check(self.plan.normsq("D", x.data_ptr(), out.data_ptr(), b, _stream()))
        return out

    def gaussianDecTransform(self, y, inplace=False):
        """tensorGaussianDec on caller-supplied i.i.d. Gaussians (random.cpp:61-64)."""
        b = _require_cuda(y, self.dtype, self.n, self.k)
        out = y if inplace else y.clone()
        capi.check(self.plan.op("GaussianDec", out.data_ptr(), b, _stream()))
        return out

    def tGaussianDec(self, v: float, batch: int, seed: int = 0, first: int = 0, device="cuda"):
        """`tGaussianDec v` (Tensor.hs:143; CPP.hs:376-389): per element n reals of scaled variance v*m/rad(m) (true variance
        svar/(2 pi), polar Box-Muller as GaussRandom.hs:34-59) drawn ON THE DEVICE from a counter-based generator keyed by
        (seed, first + element index), then the E_m transform -- lolb_tGaussianDec, one pass where the streaming kernel applies."""
        y = torch.empty(batch, self.n, self.k, dtype=torch.float64, device=device)
        capi.check(self.plan.t_gaussian_dec(float(v), int(seed), int(first), y.data_ptr(), batch, _stream()))
        return y

    def realGaussians(self, svar: float, batch: int, seed: int = 0, first: int = 0, device="cuda"):
        """`realGaussians svar n` (GaussRandom.hs:52-59) on the device: [batch, n, k] i.i.d. N(0, svar / (2 pi))."""
        y = torch.empty(batch, self.n, self.k, dtype=torch.float64, device=device)
        capi.check(capi.real_gaussians(float(svar), int(seed), int(first), y.data_ptr(), self.n * self.k, batch, _stream()))
        return y


class CudaTensorComplex(_CudaTensorPlain):
    """`Tensor` operations over Complex Double (Backend.hs:285-302, the `...C` symbols)."""
    dtype = torch.complex128
    tag = "C"

    def mulGPow(self, x, inplace=False): return self._unary("GPow", x, inplace)
    def mulGDec(self, x, inplace=False): return self._unary("GDec", x, inplace)
    def divGPow(self, x, inplace=False): return self._unary("GInvPow", x, inplace)
    def divGDec(self, x, inplace=False): return self._unary("GInvDec", x, inplace)

    def crt(self, x, inplace=False):
        b = _require_cuda(x, self.dtype, self.n, self.k)
        y = x if inplace else x.clone()
        capi.check(self.plan.op("CRTC", y.data_ptr(), b, _stream()))
        return y

    def crtInv(self, x, inplace=False):
        b = _require_cuda(x, self.dtype, self.n, self.k)
        y = x if inplace else x.clone()
        capi.check(self.plan.op("CRTInvC", y.data_ptr(), b, _stream()))
        return y

    def mul(self, a, b, inplace=False):
        ba = _require_cuda(a, self.dtype, self.n, self.k)
        bb = _require_cuda(b, self.dtype, self.n, self.k)
        y = a if inplace else a.clone()
        capi.check(self.plan.mul(y.data_ptr(), b.data_ptr(), ba, bb, _stream()))
        return y


__all__ = ["CudaTensorRq", "CudaTensorInt", "CudaTensorReal", "CudaTensorComplex", "odd_radical_fact"]
