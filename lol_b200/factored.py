"""Run-time counterpart of the bits of `Crypto.Lol.Factored` the tensor path needs
(lol/Crypto/Lol/FactoredDefs.hs:360-445): the reference reflects these from type-level
integers; here m is a run-time value."""
from __future__ import annotations

import math


def pps_fact(m: int) -> list[tuple[int, int]]:
    """`ppsFact`: prime powers of m in increasing prime order."""
    if m < 1:
        raise ValueError("m must be positive")
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


def totient_fact(m: int) -> int:
    """`totientFact`."""
    return math.prod((p - 1) * p ** (e - 1) for p, e in pps_fact(m))


def value_hat(m: int) -> int:
    """`valueHat`: m for odd m, m/2 otherwise."""
    return m if m % 2 else m // 2


def radical_fact(m: int) -> int:
    """`radicalFact`."""
    return math.prod(p for p, _ in pps_fact(m))


def odd_radical_fact(m: int) -> int:
    """`oddRadicalFact`."""
    return math.prod(p for p, _ in pps_fact(m) if p != 2)
