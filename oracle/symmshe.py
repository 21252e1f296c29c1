"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the host-side arithmetic SymmSHE runs between the FFI calls
of the `Tensor` path, used to check lol_b200/symmshe.py and she_stream.cu.  Only tests/, __graft_entry__.smoke() and
bench.py's CPU legs may import this module.

The reference for these steps is Haskell (no GHC in this image), so nothing here can be compared bit for bit with a run
of the reference itself.  The formulas are pinned SEMANTICALLY through the reference's own SymmSHE test properties
(lol-apps/Crypto/Lol/Applications/Tests/SHETests.hs: prop_encDec, prop_ctmul, prop_ksQuad) in
tests/test_oracle_symmshe_scheme.py, where every ring transform is done by the COMPILED reference: a ciphertext product
followed by `mul_and_switch` must decrypt to pt1 * pt2 for TrivGad and BaseBGad gadgets, and a random hint must not.
Bit-level parity with a Haskell run stays UNPINNED.  They restate, line by line,

    lift = decode'            lol/Crypto/Lol/Types/Unsafe/ZqBasic.hs:92-94, 124-125
    TrivGad                   ZqBasic.hs:227-232          gadget = [1], decompose x = [lift x]
    BaseBGad b                ZqBasic.hs:241-264          gadlen / gadgetZ / decomp radices . lift
    decomp, divModCent        lol/Crypto/Lol/Types/Numeric.hs:202-205, 227-234
    product-ring gadget       lol/Crypto/Lol/Gadget.hs:92-101   (concatenate per component)
    Cyc decompose             lol/Crypto/Lol/Cyclotomic/Cyc.hs:592-604   (coefficient-wise in the powerful basis)
    CT product                lol-apps/Crypto/Lol/Applications/SymmSHE.hs:443-449   mulG <$> c1 * c2
    switch / knapsack         SymmSHE.hs:302-314 ; keySwitchQuadCirc :359-372

and are anchored by the gadget identity  sum_i gadget_i * decompose(x)_i = x  (the defining property of `Decompose`,
Gadget.hs:60-66) in tests/test_oracle_pinning.py.  The CRTs inside the composite functions are the pinned ones of
oracle/cpu.py.
"""
from __future__ import annotations

import numpy as np


def lift(x: np.ndarray, q: int) -> np.ndarray:
    """ZqBasic.hs:92-94: representative in [-q/2, q/2): x if 2x < q else x - q."""
    x = np.asarray(x, dtype=np.int64) % q
    return np.where(2 * x < q, x, x - q)


def gadlen(b: int, q: int) -> int:
    """ZqBasic.hs:241-243."""
    return 0 if q == 0 else 1 + gadlen(b, q // b)


def gadget_length(qs, base: int = 0) -> int:
    return sum(1 if base == 0 else gadlen(base, int(q)) for q in qs)


def gadget(qs, base: int = 0):
    """Gadget.hs:92-94 over ZqBasic.hs:227-228 / :250-255: list of ell tuples of residues."""
    out, k = [], len(qs)
    for l, q in enumerate(qs):
        powers = [1] if base == 0 else [pow(base, i, int(q)) for i in range(gadlen(base, int(q)))]
        for pw in powers:
            out.append([pw if t == l else 0 for t in range(k)])
    return out


def div_mod_cent(a: np.ndarray, b: int):
    """Numeric.hs:227-234: remainder in [-b/2, b/2)."""
    shift = b // 2
    quo, r = np.divmod(a + shift, b)          # numpy divmod floors like Haskell divMod
    return quo, r - shift


def decomp(radices, v: np.ndarray):
    """Numeric.hs:202-205."""
    out = []
    for b in radices:
        v, r = div_mod_cent(v, b)
        out.append(r)
    out.append(v)
    return out


def decompose(x: np.ndarray, qs, base: int = 0):
    """Integer digits of a Pow-basis element x[..., n, k] (Cyc.hs:603 coefficient-wise; Gadget.hs:101 concatenation)."""
    digits = []
    for l, q in enumerate(qs):
        c = lift(x[..., l], int(q))
        if base == 0:
            digits.append(c)                                            # ZqBasic.hs:232
        else:
            digits.extend(decomp([base] * (gadlen(base, int(q)) - 1), c))    # ZqBasic.hs:257-264
    return digits


def reduce_digit(d: np.ndarray, qs) -> np.ndarray:
    """`reduce` of an integer ring element into the product ring: [..., n] -> [..., n, k] canonical residues."""
    return np.stack([np.asarray(d, dtype=np.int64) % int(q) for q in qs], axis=-1)


def decompose_reduced(x: np.ndarray, qs, base: int = 0) -> np.ndarray:
    """SymmSHE.hs:314 `fmap reduce <$> decompose c`: array [ell, ..., n, k]."""
    return np.stack([reduce_digit(d, qs) for d in decompose(x, qs, base)])


def _mulmod(a, b, qs):
    out = np.empty_like(a)
    for l, q in enumerate(qs):
        out[..., l] = (a[..., l].astype(object) * b[..., l].astype(object) % int(q)).astype(np.int64) if int(q) >= 2 ** 31 \
            else (a[..., l] % int(q)) * (b[..., l] % int(q)) % int(q)
    return out


def _addmod(a, b, qs):
    return np.stack([(a[..., l] + b[..., l]) % int(q) for l, q in enumerate(qs)], axis=-1)


def ct_mul_crt(c1, c2, g, qs):
    """SymmSHE.hs:443-449 on CRT-basis components: polynomial product (zipWithT (*) per term, UCyc.hs:232), then
    mulG = product with the gCRT vector (CPP.hs:230) on every coefficient."""
    a0, a1 = c1
    b0, b1 = c2
    d0 = _mulmod(a0, b0, qs)
    d1 = _addmod(_mulmod(a0, b1, qs), _mulmod(a1, b0, qs), qs)
    d2 = _mulmod(a1, b1, qs)
    return [_mulmod(d, np.broadcast_to(g, d.shape), qs) for d in (d0, d1, d2)]


def knapsack(hint: np.ndarray, digits_crt: np.ndarray, c0: np.ndarray, c1: np.ndarray, qs):
    """SymmSHE.hs:302-305 and :372: [c0,c1] + sum_i digit_i *>> hint_i; hint [ell, 2, n, k], digits [ell, ..., n, k]."""
    o0, o1 = c0 % np.asarray(qs), c1 % np.asarray(qs)
    for i in range(digits_crt.shape[0]):
        o0 = _addmod(o0, _mulmod(digits_crt[i], np.broadcast_to(hint[i, 0], digits_crt[i].shape), qs), qs)
        o1 = _addmod(o1, _mulmod(digits_crt[i], np.broadcast_to(hint[i, 1], digits_crt[i].shape), qs), qs)
    return [o0, o1]


def mul_and_switch(lib, c1_pow, c2_pow, hint, tables, qs, base: int = 0):
    """keySwitchQuadCirc hint (c1 * c2) for ONE ciphertext pair with Pow-basis inputs; `lib` is an oracle.cpu.CpuLib,
    tables = (pe, ru, ruinv, mhatinv, gcrt)."""
    pe, ru, ruinv, mh, g = tables
    crt = lambda v: lib.tensorCRTRq(v, pe, ru, qs)
    d0, d1, d2 = ct_mul_crt([crt(c1_pow[0]), crt(c1_pow[1])], [crt(c2_pow[0]), crt(c2_pow[1])], g, qs)
    p = lib.tensorCRTInvRq(d2, pe, ruinv, mh, qs)                       # decompose works in the powerful basis (Cyc.hs:604)
    digits = decompose_reduced(p, qs, base)
    digits_crt = np.stack([crt(np.ascontiguousarray(d)) for d in digits])   # adviseCRT (SymmSHE.hs:305)
    return knapsack(hint, digits_crt, d0, d1, qs)
