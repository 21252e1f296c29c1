"""ORACLE (test infrastructure): the coefficient-wise maps Lol applies around the tensor transforms when it switches
moduli or rounds an error term -- exact Python-integer / numpy restatement of

    lift / reduce      lol/Crypto/Lol/Types/Unsafe/ZqBasic.hs:88-94 (reduce', decode'); fmapT, UCyc.hs:267-296
    rescale (drop)     lol/Crypto/Lol/Prelude.hs:226-232 (`Rescale (a,b) b`), :259-265 (`Rescale (a,b) a`)
    rescaleMod         lol/Crypto/Lol/Prelude.hs:143-153 with divModCent, Types/Numeric.hs:227-234
    roundCoset         lol/Crypto/Lol/Prelude.hs:155-162 with roundMult, Types/Numeric.hs:207-210

Bit-level parity with a reference run is UNPINNED: these are Haskell closures and no GHC exists in this container.  The
restatement is anchored by the identities tests/test_oracle_coeffwise.py checks with independent big-integer arithmetic
(exact division after removing the centred remainder; nearest rounding of q'/q * lift x; the coset property of
roundCoset) and, semantically, by tests/test_oracle_symmshe_scheme.py: `round_coset` produces the encryption error
(errorCoset), `lift` the decryption, `rescale_mod` the plaintext rescale of prop_modSwPT (SHETests.hs:179-189), and
`rescale_drop` the ciphertext modulus switch (modSwitch, SymmSHE.hs:236-248),
with the compiled reference doing every ring transform -- the switched ciphertext must still decrypt to the plaintext.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import numpy as np


def lift(x: np.ndarray, qs) -> np.ndarray:
    """decode' per limb: x [..., k] residues -> representatives in [-q/2, q/2)."""
    q = np.asarray(qs, dtype=np.int64)
    x = np.asarray(x, dtype=np.int64) % q
    return np.where(2 * x < q, x, x - q)


def reduce(z: np.ndarray, qs) -> np.ndarray:
    """reduce' into every limb: z [..., 1] or [..., k] int64 -> [..., k] residues (Haskell `mod`: non-negative)."""
    q = np.asarray(qs, dtype=np.int64)
    return np.asarray(z, dtype=np.int64) % q


def rescale_drop(x: np.ndarray, qs, drop: int) -> np.ndarray:
    """q_d^-1 * (x_t - reduce(lift x_d)) for every t != d."""
    qs = [int(q) for q in qs]
    qd = qs[drop]
    z = lift(x[..., drop:drop + 1], [qd])[..., 0].astype(object)
    out = []
    for t, qt in enumerate(qs):
        if t == drop:
            continue
        inv = pow(qd % qt, -1, qt)
        out.append((((x[..., t].astype(object) - z) % qt) * inv % qt).astype(np.int64))
    return np.stack(out, axis=-1)


def div_mod_cent(a, b):
    """Numeric.hs:227-234 on Python integers / object arrays: remainder in [-b/2, b/2)."""
    shift = b // 2
    q = (a + shift) // b
    r = (a + shift) - q * b
    return q, r - shift


def rescale_mod(x: np.ndarray, qs, qs_new) -> np.ndarray:
    out = []
    for t, (q, qn) in enumerate(zip(qs, qs_new)):
        l = lift(x[..., t:t + 1], [q])[..., 0].astype(object)
        quot, _ = div_mod_cent(int(qn) * l, int(q))
        out.append((quot % int(qn)).astype(np.int64))
    return np.stack(out, axis=-1)


def round_coset(e: np.ndarray, zp, ps) -> np.ndarray:
    """rep + p * round((e - rep) / p) with round-half-even; zp None -> round e (roundMult 1)."""
    e = np.asarray(e, dtype=np.float64)
    if zp is None:
        return np.rint(e).astype(np.int64)
    p = np.asarray(ps, dtype=np.int64)
    rep = lift(zp, ps)
    return rep + p * np.rint((e - rep.astype(np.float64)) / p.astype(np.float64)).astype(np.int64)
