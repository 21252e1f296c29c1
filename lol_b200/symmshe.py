"""Host-side mirror of the SymmSHE ciphertext multiply and quadratic key switch (BASELINE.json configs[3]).

Reference: lol-apps/Crypto/Lol/Applications/SymmSHE.hs

    (*) on CT                  :443-452   CT d2 (k1+k2+1) (l1*l2) (mulG <$> c1 * c2)
    keySwitchQuadCirc          :359-372   [c0,c1] + switch hint c2
    switch                     :312-314   knapsack <$> hint <*> (fmap reduce <$> decompose c)
    knapsack                   :302-305   sum $ zipWith (*>>) (adviseCRT <$> xs) hint

lifted from one ciphertext to a BATCH of ciphertexts resident on the GPU.  A ciphertext is the list of its polynomial
coefficients c_0, c_1, (c_2), each a [batch, n, k] int64 tensor over R_q' = Z_q1 x ... x Z_qk[X]/Phi_m'.  Every
step runs as CUDA kernels of libctensor_b200 (the CRTs of the `Tensor` path plus the streaming passes of
she_stream.cu); there is no host arithmetic and no CPU path.  The MSD/LSD encoding tags, plaintext-modulus scale
factors (k, l in `CT enc k l c`) and the hint generation (`ksQuadCircHint`, needs the secret key and an error
sampler) stay with the caller: they do not touch the coefficient data.
"""
from __future__ import annotations

import torch

from . import capi
from .tensor import CudaTensorRq, _require_cuda, _stream


class CudaSymmSHE:
    """Ciphertext ring R_q' of index `m` over the RNS moduli `qs`, gadget TrivGad (`gad_base=0`) or BaseBGad b."""

    def __init__(self, m: int, qs, gad_base: int = 0):
        self.t = CudaTensorRq(m, qs)
        self.n, self.k = self.t.n, self.t.k
        self.gad_base = int(gad_base)
        self.ell = self.t.plan.gadget_length(self.gad_base)      # digits per coefficient (Gadget.hs:92-101)

    # ------------------------------------------------------------------ ciphertext product
    def mulCT(self, c1, c2, basis: str = "pow", inplace: bool = False):
        """`c1 * c2` for two linear ciphertexts: returns [d0, d1, d2] in the CRT basis, mulG applied (SymmSHE.hs:443-449).

        basis = "pow": components arrive in the powerful basis and are forced to CRT first (the four tensorCRTRq of
        SURVEY.md section 3.5); "crt": they are already there.  inplace=True transforms / overwrites the inputs."""
        if len(c1) != 2 or len(c2) != 2:
            raise capi.LolB200Error(capi.LOLB_ERR_ARG, "mulCT takes two linear ciphertexts (two coefficients each)")
        b = _require_cuda(c1[0], torch.int64, self.n, self.k)
        ops = list(c1) + list(c2)
        for x in ops:
            if _require_cuda(x, torch.int64, self.n, self.k) != b:
                raise capi.LolB200Error(capi.LOLB_ERR_ARG, "all ciphertext components must share one batch size")
        if basis == "pow":
            ops = [self.t.crt(x, inplace=inplace) for x in ops]
        elif basis != "crt":
            raise capi.LolB200Error(capi.LOLB_ERR_ARG, "basis must be 'pow' or 'crt'")
        a0, a1, b0, b1 = ops
        reuse = inplace or basis == "pow"          # the CRT copies are ours to overwrite
        d0 = a0 if reuse else torch.empty_like(a0)
        d1 = b1 if reuse else torch.empty_like(a0)
        d2 = a1 if reuse else torch.empty_like(a0)
        capi.check(self.t.plan.ct_mul(a0.data_ptr(), a1.data_ptr(), b0.data_ptr(), b1.data_ptr(),
                                      d0.data_ptr(), d1.data_ptr(), d2.data_ptr(), b, True, _stream()))
        return [d0, d1, d2]

    # ------------------------------------------------------------------ key switch
    def decompose(self, x):
        """`fmap reduce <$> decompose x` for a Pow-basis x: tensor [ell, batch, n, k] (SymmSHE.hs:314, Cyc.hs:603)."""
        b = _require_cuda(x, torch.int64, self.n, self.k)
        digits = torch.empty(self.ell, b, self.n, self.k, dtype=torch.int64, device=x.device)
        capi.check(self.t.plan.decompose(x.data_ptr(), digits.data_ptr(), b, self.gad_base, _stream()))
        return digits

    def decomposeCRT(self, x):
        """`adviseCRT <$> (fmap reduce <$> decompose x)` (SymmSHE.hs:305, :314): the gadget digits of a Pow-basis x, each
        in the CRT basis, [ell, batch, n, k]; one kernel where the CRT's load stage can form the digits itself."""
        b = _require_cuda(x, torch.int64, self.n, self.k)
        digits = torch.empty(self.ell, b, self.n, self.k, dtype=torch.int64, device=x.device)
        capi.check(self.t.plan.decompose_crt(x.data_ptr(), digits.data_ptr(), b, self.gad_base, _stream()))
        return digits

    def knapsack(self, hint, digits, c0, c1, inplace: bool = False):
        """[c0, c1] + sum_i digits[i] *>> hint[i]  (SymmSHE.hs:302-305, :372).  `hint` is [ell, 2, n, k] in the CRT
        basis (the linear polynomials of `ksHint`, :288-298), `digits` [ell, batch, n, k] in the CRT basis."""
        b = _require_cuda(c0, torch.int64, self.n, self.k)
        _require_cuda(c1, torch.int64, self.n, self.k)
        if tuple(hint.shape) != (self.ell, 2, self.n, self.k) or hint.dtype != torch.int64 or not hint.is_cuda or not hint.is_contiguous():
            raise capi.LolB200Error(capi.LOLB_ERR_ARG, f"hint must be a contiguous CUDA int64 tensor [{self.ell}, 2, {self.n}, {self.k}]")
        if tuple(digits.shape) != (self.ell, b, self.n, self.k) or not digits.is_contiguous() or not digits.is_cuda:
            raise capi.LolB200Error(capi.LOLB_ERR_ARG, f"digits must be a contiguous CUDA tensor [{self.ell}, {b}, {self.n}, {self.k}]")
        o0 = c0 if inplace else c0.clone()
        o1 = c1 if inplace else c1.clone()
        capi.check(self.t.plan.knapsack(digits.data_ptr(), self.ell, hint.data_ptr(), o0.data_ptr(), o1.data_ptr(), b, _stream()))
        return [o0, o1]

    def keySwitchQuadCirc(self, hint, ct, inplace: bool = False):
        """Degree-2 ciphertext [c0, c1, c2] (CRT basis) -> degree 1 under the same key (SymmSHE.hs:359-372):
        c2 to the powerful basis (tensorCRTInvRq), gadget digits each taken back to CRT (lolb_decomposeCrtRq: one launch
        over all ell * batch digits), knapsack with the hint."""
        if len(ct) != 3:
            raise capi.LolB200Error(capi.LOLB_ERR_ARG, "keySwitchQuadCirc takes a ciphertext with three coefficients")
        c0, c1, c2 = ct
        b = _require_cuda(c2, torch.int64, self.n, self.k)
        p = self.t.crtInv(c2, inplace=inplace)
        digits = self.decomposeCRT(p)
        return self.knapsack(hint, digits, c0, c1, inplace=inplace)

    def mulAndSwitch(self, c1, c2, hint, basis: str = "pow", inplace: bool = False):
        """keySwitchQuadCirc hint (c1 * c2): the op sequence BASELINE.json configs[3] counts as one unit of work."""
        return self.keySwitchQuadCirc(hint, self.mulCT(c1, c2, basis=basis, inplace=inplace), inplace=True)

    # ------------------------------------------------------------------ gadget (for tests and hint construction)
    def gadget(self):
        """The gadget vector over the product ring as ell tuples of residues (Gadget.hs:92-94; ZqBasic.hs:227-228,
        250-255): digit block of limb l holds b^i in limb l and 0 elsewhere."""
        out = []
        for l, q in enumerate(self.t.qs):
            if self.gad_base == 0:
                powers = [1]
            else:
                powers, v, qq = [], 1, q
                while qq != 0:
                    powers.append(v % q); v *= self.gad_base; qq //= self.gad_base
            for pw in powers:
                out.append([pw if t == l else 0 for t in range(self.k)])
        return out
