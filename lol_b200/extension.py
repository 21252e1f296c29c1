"""Host-side mirror of the two-index methods of the reference's `Tensor` class (ring extensions O_m'/O_m, m | m').

    Tensor method (Tensor.hs:160-190)   reference (CPP.hs:246-255 -> CPP/Extension.hs)   here
    ---------------------------------  -----------------------------------------------  --------------------------
    twacePowDec                        twacePowDec'  Extension.hs:99-103                 CudaExtension.twacePowDec
    embedPow / embedDec                embedPow', embedDec'  Extension.hs:60-77          .embedPow / .embedDec
    crtExtFuncs: twaceCRT, embedCRT    twaceCRT', embedCRT'  Extension.hs:81-85, 110-129 .twaceCRT / .embedCRT (None = Nothing)
    coeffs                             coeffs'  Extension.hs:90-93                       .coeffs
    powBasisPow                        powBasisPow'  Extension.hs:133-143                .powBasisPow
    crtSetDec                          crtSetDec'  Extension.hs:145-164                  .crtSetDec (host precomputation, crtset.py)

An extension is built from two single-index tensors of `lol_b200.tensor` over the same ring (`CudaTensorRq` with equal
moduli, or two of `CudaTensorInt` / `CudaTensorReal` / `CudaTensorComplex` with equal tupSize).  Operands are torch CUDA
tensors [batch, phi, k] / [batch, phi', k]; every call is one gather kernel of libctensor_b200 (ext_stream.cu) on the
current stream and returns a new tensor.  No CPU path.
"""
from __future__ import annotations

import torch

from . import capi
from .tensor import CudaTensorComplex, CudaTensorInt, CudaTensorReal, CudaTensorRq, _require_cuda, _stream


class CudaExtension:
    def __init__(self, lo, hi):
        if type(lo) is not type(hi):
            raise capi.LolB200Error(capi.LOLB_ERR_ARG, "both tensors must be over the same ring")
        self.lo, self.hi = lo, hi
        self.k = lo.k
        if isinstance(lo, CudaTensorRq):
            self.ring, self.dtype = capi.RING_RQ, torch.int64
        elif isinstance(lo, CudaTensorInt):
            self.ring, self.dtype = capi.RING_R, torch.int64
        elif isinstance(lo, CudaTensorReal):
            self.ring, self.dtype = capi.RING_DOUBLE, torch.float64
        elif isinstance(lo, CudaTensorComplex):
            self.ring, self.dtype = capi.RING_C, torch.complex128
        else:
            raise capi.LolB200Error(capi.LOLB_ERR_ARG, "unsupported tensor type")
        self.ext = capi.Extension(lo.plan, hi.plan)
        self.phi, self.phi2 = self.ext.phi, self.ext.phi2

    def _run(self, name, x, n_in, out_shape):
        b = _require_cuda(x, self.dtype, n_in, self.k)
        y = torch.empty((b, *out_shape), dtype=self.dtype, device=x.device)
        capi.check(self.ext.op(name, self.ring, x.data_ptr(), y.data_ptr(), b, _stream()))
        return y

    def _run_crt(self, name, x, n_in, out_shape):
        try:
            return self._run(name, x, n_in, out_shape)
        except capi.LolB200Error as e:
            if e.status == capi.LOLB_ERR_NO_CRT:
                return None                      # the reference's `Nothing` (crtExtFuncs :: Maybe ..., Tensor.hs:176-180)
            raise

    def twacePowDec(self, x): return self._run("twacePowDec", x, self.phi2, (self.phi, self.k))
    def embedPow(self, x): return self._run("embedPow", x, self.phi, (self.phi2, self.k))
    def embedDec(self, x): return self._run("embedDec", x, self.phi, (self.phi2, self.k))
    def embedCRT(self, x): return self._run_crt("embedCRT", x, self.phi, (self.phi2, self.k))
    def twaceCRT(self, x): return self._run_crt("twaceCRT", x, self.phi2, (self.phi, self.k))

    def powBasisPow(self):
        """[phi'/phi, phi', k]: the vectors of O_m' (powerful basis) that form an O_m-basis of O_m' (Tensor.hs:177)."""
        y = torch.empty((self.phi2 // self.phi, self.phi2, self.k), dtype=self.dtype, device="cuda")
        capi.check(self.ext.pow_basis_pow(self.ring, y.data_ptr(), _stream()))
        return y

    def crtSetDec(self, p: int):
        """[count, phi', 1] int64 residues mod the prime p: the mod-p CRT set of O_m'/O_m in the decoding basis (Tensor.hs:184-186).
        Built once on the host like the reference does (GF(p^d) arithmetic, crtset.py) and uploaded."""
        from . import crtset
        cs = crtset.crt_set_dec(self.lo.m, self.hi.m, int(p))
        return torch.from_numpy(cs).to("cuda").unsqueeze(-1).contiguous()

    def coeffs(self, x):
        """[batch, phi', k] -> [batch, phi'/phi, phi, k]: the O_m coefficients w.r.t. the powerful / decoding extension basis."""
        y = self._run("coeffsPowDec", x, self.phi2, (self.phi2, self.k))
        return y.view(y.shape[0], self.phi2 // self.phi, self.phi, self.k)
