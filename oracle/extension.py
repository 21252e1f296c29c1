"""ORACLE (test infrastructure): ring-extension operators of Lol's `Tensor` class for O_m'/O_m, m | m'.

numpy restatement of the index correspondences of `lol/Crypto/Lol/Cyclotomic/Tensor.hs:380-510`
and of the operators `lol-cpp/Crypto/Lol/Cyclotomic/Tensor/CPP/Extension.hs:54-129` builds on them
(embedPow', embedDec', embedCRT', twacePowDec', twaceCRT', coeffs').  Those operators are Haskell
code; no GHC exists in this container, so this restatement cannot be run against the reference
binary directly.  It is pinned instead through the reference's own two-index properties
(`lol/Crypto/Lol/Tests/TensorTests.hs:133-215`), evaluated in tests/test_oracle_extension.py with the
COMPILED reference (`oracle/_ref/libctensor_ref.so`) supplying crt / crtInv / l / lInv / divG:

    twacePowDec . embedPow == id,  twacePowDec . embedDec == id         (prop_trem_pow, prop_trem_dec)
    embedCRT  == crt . embedPow . crtInv                                (prop_embed_crt)
    twacePowDec == lInv . twacePowDec . l                               (prop_twace_dec)
    twaceCRT  == crt . twacePowDec . crtInv                             (prop_twace_crt)
    twace (mhat'/g') == mhat * (phi'/phi) / g   in Pow, Dec and CRT     (prop_twace_invar1_*)
    twace preserves scalars                                             (prop_twace_invar2_*)
    x == sum_k embed(coeffs x)_k * powBasis_k                           (prop_coeffsBasis, CycTests.hs:71-76)

and, inside a small SymmSHE, through prop_cttwace / prop_ctembed / prop_ringTunnel of lol-apps' SHETests.hs:211-248
(tests/test_oracle_symmshe_scheme.py).

Arrays are one ring element in the ABI layout [phi][k] (int64 residues, int64, double or complex128).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import numpy as np

from . import tables


# ---------------------------------------------------------------- index correspondences
def merge_pps(pps, pps2):
    """Tensor.hs:502-507 (mergePPs): [(p, e in m, e' in m')]; m must divide m'."""
    lo = dict(pps)
    hi = dict(pps2)
    for p, e in lo.items():
        if hi.get(p, 0) < e:
            raise ValueError("m does not divide m'")
    return [(p, lo.get(p, 0), e2) for p, e2 in sorted(hi.items())]


def _tot_pp(p: int, e: int) -> int:
    """FactoredDefs.hs:429 (totientPP), with e = 0 -> 1."""
    return 1 if e == 0 else (p - 1) * p ** (e - 1)


def totients(mpps):
    """Tensor.hs:509-510."""
    return [(_tot_pp(p, e), _tot_pp(p, e2)) for p, e, e2 in mpps]


def to_index_pair(tots, i2: int):
    """Tensor.hs:391-399 (toIndexPair)."""
    if not tots:
        assert i2 == 0
        return (0, 0)
    (phi, phi2), rest = tots[0], tots[1:]
    q, r = divmod(i2, phi2)
    rq, rr = divmod(r, phi)
    q1, q0 = to_index_pair(rest, q)
    return (rq + q1 * (phi2 // phi), rr + q0 * phi)


def from_index_pair(tots, pair):
    """Tensor.hs:401-406 (fromIndexPair)."""
    i1, i0 = pair
    if not tots:
        assert (i1, i0) == (0, 0)
        return 0
    (phi, phi2), rest = tots[0], tots[1:]
    i0q, i0r = divmod(i0, phi)
    i1q, i1r = divmod(i1, phi2 // phi)
    return (i0r + i1r * phi) + from_index_pair(rest, (i1q, i0q)) * phi2


def base_index_dec(mpps, i2: int):
    """Tensor.hs:483-498 (baseIndexDec): None, or (index into O_m's decoding basis, negate?)."""
    if not mpps:
        assert i2 == 0
        return (0, False)
    (p, e, e2), rest = mpps[0], mpps[1:]
    q, r = divmod(i2, _tot_pp(p, e2))
    phi = _tot_pp(p, e)
    if p > 2 and e == 0 and e2 > 0:
        curr = (0, False) if r == 0 else (0, True) if r == 1 else None
    else:
        curr = (r, False) if r < phi else None
    if curr is None:
        return None
    up = base_index_dec(rest, q)
    if up is None:
        return None
    return (curr[0] + phi * up[0], curr[1] != up[1])


class ExtInfo:
    """indexInfo (Tensor.hs:414-424) plus every table derived from it (Tensor.hs:429-478)."""

    def __init__(self, m: int, m2: int):
        self.m, self.m2 = m, m2
        self.mpps = merge_pps(tables.factor_pps(m), tables.factor_pps(m2))
        self.tots = totients(self.mpps)
        self.phi = tables.totient_pps(tables.factor_pps(m))
        self.phi2 = tables.totient_pps(tables.factor_pps(m2))
        self.rel = self.phi2 // self.phi
        tots = self.tots
        # extIndicesPowDec: [phi]
        self.ext_powdec = np.array([from_index_pair(tots, (0, i)) for i in range(self.phi)], dtype=np.int64)
        # extIndicesCRT: [phi'], block i holds the indices lying above CRT index i
        self.ext_crt = np.array([from_index_pair(tots, divmod(j, self.rel)[::-1]) for j in range(self.phi2)], dtype=np.int64)
        # baseIndicesPow: [phi'] pairs (j0, j1)
        pairs = [to_index_pair(tots, j) for j in range(self.phi2)]
        self.base_pow_j0 = np.array([a for a, _ in pairs], dtype=np.int64)
        self.base_pow_j1 = np.array([b for _, b in pairs], dtype=np.int64)
        self.base_crt = self.base_pow_j1
        # baseIndicesDec: [phi'] -> source index or -1, and sign
        dec = [base_index_dec(self.mpps, j) for j in range(self.phi2)]
        self.base_dec_idx = np.array([-1 if d is None else d[0] for d in dec], dtype=np.int64)
        self.base_dec_neg = np.array([False if d is None else d[1] for d in dec], dtype=bool)
        # extIndicesCoeffs: [phi'/phi][phi]
        self.ext_coeffs = np.array([[from_index_pair(tots, (i1, i0)) for i0 in range(self.phi)] for i1 in range(self.rel)],
                                   dtype=np.int64)


# ---------------------------------------------------------------- operators (one element, [phi][k] -> [phi'][k] or back)
def _neg(x, qs):
    if qs is None:
        return -x
    q = np.asarray(qs, dtype=np.int64)
    return (q - x) % q


def twace_powdec(info: ExtInfo, x):
    """Extension.hs:99-103 (twacePowDec' = backpermute extIndicesPowDec)."""
    return x[info.ext_powdec]


def embed_pow(info: ExtInfo, x):
    """Extension.hs:60-70 (embedPow')."""
    y = x[info.base_pow_j1]
    y[info.base_pow_j0 != 0] = 0
    return y


def embed_dec(info: ExtInfo, x, qs=None):
    """Extension.hs:71-77 (embedDec'); `negate` is modulo each q_t for Zq tuples."""
    src = np.where(info.base_dec_idx < 0, 0, info.base_dec_idx)
    y = x[src]
    neg = _neg(y, qs)
    y = np.where(info.base_dec_neg[:, None], neg, y)
    y[info.base_dec_idx < 0] = 0
    return y


def embed_crt(info: ExtInfo, x):
    """Extension.hs:81-85 (embedCRT' = backpermute baseIndicesCRT)."""
    return x[info.base_crt]


def pow_basis_pow(info: ExtInfo, k: int = 1, dtype=np.int64):
    """Extension.hs:133-143 (powBasisPow'): [phi'/phi][phi'][k], vector r = one where baseIndicesPow = (r, 0), zero elsewhere."""
    out = np.zeros((info.rel, info.phi2, k), dtype=dtype)
    for r in range(info.rel):
        out[r, (info.base_pow_j0 == r) & (info.base_pow_j1 == 0), :] = 1
    return out


def coeffs_powdec(info: ExtInfo, x):
    """Extension.hs:90-93 (coeffs'): [phi'/phi][phi][k]."""
    return x[info.ext_coeffs]


def twace_crt_tweak_zq(info: ExtInfo, qs):
    """Extension.hs:113-126: tweak = m'hat^-1 * mhat * embedCRT(gInvCRT_m) * gCRT_m'  per limb, [phi'][k] int64."""
    qs = [int(q) for q in qs]
    _, ginv_lo = tables.g_crt_vectors(info.m, qs)
    g_hi, _ = tables.g_crt_vectors(info.m2, qs)
    out = np.empty((info.phi2, len(qs)), dtype=np.int64)
    for t, q in enumerate(qs):
        ratio = tables.mhat_inv(info.m2, q) * (tables.value_hat(info.m) % q) % q
        emb = ginv_lo[info.base_crt, t].astype(object)
        out[:, t] = np.array([(int(a) * int(b) % q) * ratio % q for a, b in zip(emb, g_hi[:, t])], dtype=np.int64)
    return out


def twace_crt_zq(info: ExtInfo, x, qs):
    """Extension.hs:127-129: multiply by the tweak, gather by extIndicesCRT, sum each block of phi'/phi."""
    tweak = twace_crt_tweak_zq(info, qs)
    q = np.asarray(qs, dtype=object)
    prod = (x.astype(object) * tweak.astype(object)) % q
    v = prod[info.ext_crt].reshape(info.phi, info.rel, len(qs))
    return (v.sum(axis=1) % q).astype(np.int64)


def twace_crt_c(info: ExtInfo, x, g_lo, g_hi):
    """Same over the complex numbers; g_lo / g_hi are the gCRT vectors of O_m / O_m' ([phi], [phi'] complex128)."""
    ratio = tables.value_hat(info.m) / tables.value_hat(info.m2)
    tweak = ratio * g_hi / g_lo[info.base_crt]
    v = (x * tweak[:, None])[info.ext_crt].reshape(info.phi, info.rel, -1)
    return v.sum(axis=1)
