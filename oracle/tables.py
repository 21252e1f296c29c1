"""ORACLE (test infrastructure): the Haskell-side inputs of the lol-cpp hot path.

The reference C++ cannot derive its own root tables: `Crypto.Lol.Cyclotomic.Tensor.CPP`
builds them in Haskell and passes pointers (CPP.hs:422-442).  This module restates
those builders with Python integers / numpy so the tests can drive the compiled
reference (`oracle/_ref/libctensor_ref.so`), the C restatement and the CUDA
library with identical arguments, and can check the table builder inside the
product library (lol_b200/csrc/plan.cpp) against an independent implementation.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np


# ---------------------------------------------------------------- factoring
def factor_pps(m: int) -> list[tuple[int, int]]:
    """Prime powers of m in increasing prime order (FactoredDefs.hs:92-94, 360-361)."""
    out, p = [], 2
    while p * p <= m:
        if m % p == 0:
            e = 0
            while m % p == 0:
                m //= p
                e += 1
            out.append((p, e))
        p += 1
    if m > 1:
        out.append((m, 1))
    return out


def totient_pps(pps) -> int:
    """FactoredDefs.hs:429 (totientPP) over all prime powers."""
    return math.prod((p - 1) * p ** (e - 1) for p, e in pps)


def value_pps(pps) -> int:
    return math.prod(p ** e for p, e in pps)


def value_hat(m: int) -> int:
    """m-hat: m for odd m, m/2 otherwise (FactoredDefs.hs:374-376)."""
    return m if m % 2 else m // 2


def radical(m: int) -> int:
    return math.prod(p for p, _ in factor_pps(m))


def odd_radical(m: int) -> int:
    return math.prod(p for p, _ in factor_pps(m) if p != 2)


def is_prime(n: int) -> bool:
    if n < 2:
        return False
    i = 2
    while i * i <= n:
        if n % i == 0:
            return False
        i += 1
    return True


def good_qs(m: int, lower: int):
    """Primes > lower congruent to 1 mod m, ascending (ZqBasic.hs:71-73)."""
    q = lower + ((m - lower) % m) + 1
    while True:
        if is_prime(q):
            yield q
        q += m


# ---------------------------------------------------------------- Zq roots
@lru_cache(maxsize=None)
def smallest_generator(q: int) -> int:
    """Smallest generator of Z_q^* for prime q (ZqBasic.hs:144-163: `head (filter isGen values)`)."""
    if not is_prime(q):
        raise ValueError(f"q={q} is not prime: no CRT over Z_q (ZqBasic.hs:159)")
    order = q - 1
    exps = [order // p for p, _ in factor_pps(order)] if order > 1 else []
    for x in range(q):
        if pow(x, order, q) == 1 % q and all(pow(x, e, q) != 1 for e in exps):
            return x
    raise AssertionError("no generator")


def omega(m: int, q: int) -> int:
    """Principal m-th root of unity mod q: generator^((q-1)/m) (ZqBasic.hs:160-163)."""
    if (q - 1) % m:
        raise ValueError(f"m={m} does not divide q-1={q - 1} (ZqBasic.hs:164)")
    return pow(smallest_generator(q), (q - 1) // m, q)


def mhat_inv(m: int, q: int) -> int:
    """(m-hat)^-1 mod q (ZqBasic.hs:167-171)."""
    return pow(value_hat(m), -1, q)


def ru_tables_zq(m: int, qs, inverse: bool = False) -> list[np.ndarray]:
    """Per prime power: int64 array [p^e, k], entry (j, limb) = w_limb^(+-j * m/p^e)
    (CPP.hs:422-442); the row-major [j][limb] layout is the interleaved tuple
    layout of Backend.hs:80-90."""
    ws = [omega(m, q) for q in qs]
    out = []
    for p, e in factor_pps(m):
        pp = p ** e
        step = m // pp
        t = np.empty((pp, len(qs)), dtype=np.int64)
        for limb, (q, w) in enumerate(zip(qs, ws)):
            base = pow(w, step, q)
            if inverse:
                base = pow(base, -1, q)
            acc = 1 % q
            for j in range(pp):
                t[j, limb] = acc
                acc = acc * base % q
        out.append(t)
    return out


def ru_tables_c(m: int, k: int = 1, inverse: bool = False) -> list[np.ndarray]:
    """Complex root tables cis(+-2 pi j / p^e) (CRTrans.hs:88-95), layout [p^e, k]."""
    out = []
    sgn = -1.0 if inverse else 1.0
    for p, e in factor_pps(m):
        pp = p ** e
        step = m // pp
        j = np.arange(pp, dtype=np.float64) * step
        col = np.exp(sgn * 2j * np.pi * j / m).astype(np.complex128)
        # the Haskell side evaluates cis(2*pi*i/m) with i = j*m/pp reduced mod m; same angle
        out.append(np.repeat(col[:, None], k, axis=1).copy())
    return out


def mhat_inv_c(m: int, k: int = 1) -> np.ndarray:
    return np.full(k, 1.0 / value_hat(m), dtype=np.complex128)


# ---------------------------------------------------------------- g in the CRT basis
def g_crt_vectors(m: int, qs) -> tuple[np.ndarray, np.ndarray]:
    """(gCRT, gInvCRT) as int64 [n, k] (Tensor.hs:264-337, CPP.hs:444-454).

    Prime p: gCRT_p[i] = 1 - w_p^(i+1); gInvCRT_p[i] = phat^-1 * sum_{j=1}^{p-1} j * w_p^((i+1)(p-1-j)).
    Prime power: index i -> i mod (p-1) (ppKron); m: Kronecker product, first prime power fastest (indexK).
    """
    pps = factor_pps(m)
    n = totient_pps(pps)
    g = np.empty((n, len(qs)), dtype=np.int64)
    gi = np.empty((n, len(qs)), dtype=np.int64)
    for limb, q in enumerate(qs):
        w = omega(m, q)
        vec_g, vec_gi = [1], [1]
        for p, e in pps:
            phi = (p - 1) * p ** (e - 1)
            if p == 2:
                fg, fgi = [1 % q] * phi, [1 % q] * phi
            else:
                wp = pow(w, m // p, q)
                phat_inv = pow(p, -1, q)
                pg = [(1 - pow(wp, i + 1, q)) % q for i in range(p - 1)]
                pgi = [phat_inv * sum(j * pow(wp, (i + 1) * (p - 1 - j), q) for j in range(1, p)) % q
                       for i in range(p - 1)]
                fg = [pg[i % (p - 1)] for i in range(phi)]
                fgi = [pgi[i % (p - 1)] for i in range(phi)]
            # earlier prime powers vary fastest (indexK: i = iq * r + ir, ir indexes the LAST snoc'd = first pp)
            vec_g = [a * b % q for b in fg for a in vec_g]
            vec_gi = [a * b % q for b in fgi for a in vec_gi]
        g[:, limb] = vec_g
        gi[:, limb] = vec_gi
    return g, gi


def pe_array(m: int) -> np.ndarray:
    """PrimeExponent[] as int16 [npe, 2] (types.h:27-31; Backend.hs:63-64, 78)."""
    return np.array(factor_pps(m), dtype=np.int16).reshape(-1, 2)


# ---------------------------------------------------------------- synthetic inputs
def gaussian_scaled_variance(m: int, v: float) -> float:
    """svar handed to realGaussians by cDispatchGaussian (CPP.hs:384-387): v * m / rad(m)."""
    return v * (m // radical(m))


def real_gaussians(svar: float, n: int, rng: np.random.Generator) -> np.ndarray:
    """i.i.d. reals with the distribution of GaussRandom.hs:34-59 (true variance svar/(2 pi))."""
    return rng.normal(0.0, math.sqrt(svar / (2 * math.pi)), size=n)
