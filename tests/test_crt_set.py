"""crtSetDec' (CPP/Extension.hs:145-164) on the host (lol_b200/crtset.py): the reference's own property for it,
prop_crtSet_pairs (lol/Crypto/Lol/Tests/CycTests.hs:78-86: c_i * c_j = delta_ij * c_i), plus the two facts that pin the set
itself: the elements sum to 1 and there are (phi(m')/d') / (phi(m)/d) of them, d = ord_m(p), d' = ord_m'(p).

The ring arithmetic of the check is independent of the code under test: decoding -> powerful basis with the oracle's L
(tensorLRq, modulus p), then multiplication in O_m' = (x)_i Z_p[x_i]/Phi_{p_i^e_i}(x_i), whose monomial basis IS the
powerful basis (Tensor.hs:359-362 for the index order inside a prime power).
"""
import numpy as np
import pytest

from lol_b200 import crtset
from oracle import tables as T

CASES = [(1, 7, 2), (1, 21, 2), (3, 21, 2), (7, 21, 2), (1, 15, 2), (5, 45, 2), (3, 9, 2), (1, 5, 3), (4, 20, 3), (1, 7, 3),
         (7, 7, 2), (8, 40, 3), (1, 13, 3), (1, 16, 3), (2, 16, 7)]


def _pow_index_to_exponents(pps, n):
    """tensor index -> per-prime-power exponent of zeta_{p^e} in the powerful basis (first prime power fastest)."""
    out = np.zeros((n, len(pps)), dtype=np.int64)
    for j in range(n):
        r = j
        for a, (p, e) in enumerate(pps):
            ph = (p - 1) * p ** (e - 1)
            r, jj = divmod(r, ph)
            out[j, a] = crtset._index_to_pow(p, e, jj)
    return out


def _ring_mul_pow(pps, p, a, b):
    """a * b in O_m' / p, both in the powerful basis (tensor index order)."""
    n = len(a)
    ex = _pow_index_to_exponents(pps, n)
    dims = [q ** e for q, e in pps]
    A = np.zeros(dims, dtype=np.int64)
    B = np.zeros(dims, dtype=np.int64)
    for j in range(n):
        A[tuple(ex[j])] = a[j]
        B[tuple(ex[j])] = b[j]
    # cyclic convolution modulo x_i^(p_i^e_i) - 1 on every axis (Phi divides it), by brute force
    Cc = np.zeros(dims, dtype=np.int64)
    for ia in np.argwhere(A):
        Cc = (Cc + A[tuple(ia)] * np.roll(B, tuple(ia), axis=tuple(range(len(dims))))) % p
    # reduce modulo Phi_{q^e}(x) = sum_{k<q} x^(k q^(e-1)) on every axis: x^(t + (q-1) q^(e-1)) = -sum_{k<q-1} x^(t + k q^(e-1))
    for ax, (q, e) in enumerate(pps):
        Cc = np.moveaxis(Cc, ax, 0).copy()
        s = q ** (e - 1)
        for t in range(s):
            top = Cc[t + (q - 1) * s].copy()
            for k in range(q - 1):
                Cc[t + k * s] = (Cc[t + k * s] - top) % p
            Cc[t + (q - 1) * s] = 0
        Cc = np.moveaxis(Cc, 0, ax)
    return np.array([Cc[tuple(ex[j])] for j in range(n)], dtype=np.int64)


@pytest.mark.parametrize("m,m2,p", CASES, ids=lambda v: str(v))
def test_crt_set_dec_is_a_crt_set(oracle, m, m2, p):
    cs = crtset.crt_set_dec(m, m2, p)
    pps, pe = T.factor_pps(m2), T.pe_array(m2)
    n = T.totient_pps(pps)
    d, d2 = crtset.order(p, m), crtset.order(p, m2)
    phi = T.totient_pps(T.factor_pps(m)) if m > 1 else 1
    assert cs.shape == ((n // d2) // (phi // d), n)
    assert cs.min() >= 0 and cs.max() < p
    pw = [oracle.tensorLRq(c.reshape(n, 1), pe, [p]).reshape(n) % p for c in cs]      # toPow (Dec v) = Pow (l v), UCyc.hs:574
    one = np.zeros(n, dtype=np.int64)
    one[0] = 1
    assert np.array_equal(sum(pw) % p, one)
    for i, a in enumerate(pw):
        for j, b in enumerate(pw):
            prod = _ring_mul_pow(pps, p, a, b)
            assert np.array_equal(prod, a if i == j else np.zeros(n, dtype=np.int64)), (i, j)


def test_partition_cosets_shape():
    """ZmStar.hs:70-90: each part holds one coset of Z_m'^*/<p> above every coset of Z_m^*/<p>."""
    for m, m2, p in CASES:
        parts = crtset.partition_cosets(p, m, m2)
        reps = [r for part in parts for r in part]
        assert len(set(reps)) == len(reps) == T.totient_pps(T.factor_pps(m2)) // crtset.order(p, m2)
        for part in parts:
            below = {min((r * p ** t) % m for t in range(crtset.order(p, m))) if m > 1 else 0 for r in part}
            assert len(below) == len(part)
