"""Host-side mirror of `crtSetDec'` (lol-cpp/Crypto/Lol/Cyclotomic/Tensor/CPP/Extension.hs:145-164): the mod-p CRT set of
the extension O_m'/O_m, as vectors over F_p in the decoding basis of O_m'.

Like the reference this is precomputation on the host (a list of phi(m')-vectors built once per (m, m', p)); the vectors are
then ordinary ring elements for the device path (`CudaExtension.crtSetDec` uploads them).  What it follows:

    GF(p^d), d = ord_{m'}(p)            Types/FiniteField.hs:66-121 (polynomials over F_p modulo an irreducible of degree d)
    omega = gen^((p^d - 1)/m')          FiniteField.hs:105-118  (gen = first primitive element)
    twCRTs: Kron of per-prime-power     Tensor.hs:300-313  entry (j, i) = w^(indexToPow j * -indexToZms i) * gCRT(i)
    gCRT of a prime (1 - w_p^(i+1))     Tensor.hs:319-327, ppKron Tensor.hs:264-271
    zmsToIndex                          Tensor.hs:371-379
    partitionCosets                     Types/ZmStar.hs:51-90 (same coset order: min representative; Map order on the
                                        m-cosets; later m'-cosets first within a key)
    trace GF(p^d) -> F_p                FiniteField.hs:179-198 (sum of the Frobenius conjugates)

The reference ships irreducible polynomials for characteristic 2 only (Types/IrreducibleChar2.hs) and takes any user
instance otherwise; here the lexicographically first monic irreducible of degree d is used for every p.  The CRT *set* does
not depend on that choice nor on the choice of generator; the order of its elements can (a different omega permutes the
cosets by an automorphism), so the order is this module's, not pinned to the reference's.
"""
from __future__ import annotations

import itertools
from math import gcd

import numpy as np

from .factored import pps_fact, totient_fact


# ------------------------------------------------------------------ F_p[x] / (f), elements = tuples of d residues (low first)
class GF:
    def __init__(self, p: int, d: int):
        self.p, self.d = p, d
        self.size = p ** d
        self.f = self._first_irreducible()

    # polynomials as tuples, low degree first, no trailing zeros
    def _trim(self, a):
        a = list(a)
        while a and a[-1] % self.p == 0:
            a.pop()
        return tuple(x % self.p for x in a)

    def _polymod(self, a, f):
        a, p = list(a), self.p
        df = len(f) - 1
        inv = pow(f[-1], -1, p)
        for i in range(len(a) - 1, df - 1, -1):
            c = a[i] * inv % p
            if c:
                for j in range(df + 1):
                    a[i - df + j] = (a[i - df + j] - c * f[j]) % p
        return self._trim(a[:df])

    def _polymul(self, a, b):
        if not a or not b:
            return ()
        out = [0] * (len(a) + len(b) - 1)
        for i, x in enumerate(a):
            if x:
                for j, y in enumerate(b):
                    out[i + j] = (out[i + j] + x * y) % self.p
        return self._trim(out)

    def _first_irreducible(self):
        p, d = self.p, self.d
        if d == 1:
            return (0, 1)
        # monic candidates in lexicographic order of (c_{d-1}, ..., c_0); irreducible iff no monic divisor of degree <= d/2
        small = [self._trim(list(c) + [1]) for k in range(1, d // 2 + 1) for c in itertools.product(range(p), repeat=k)]
        for hi in itertools.product(range(p), repeat=d):
            f = tuple(reversed(hi)) + (1,)
            if all(self._polymod(f, g) != () for g in small):
                return f
        raise ArithmeticError("no irreducible polynomial found")

    # field operations on length-d tuples
    def elem(self, a):
        a = self._trim(a)
        return a + (0,) * (self.d - len(a))

    @property
    def zero(self): return (0,) * self.d

    @property
    def one(self): return self.elem((1,))

    def add(self, a, b): return tuple((x + y) % self.p for x, y in zip(a, b))
    def sub(self, a, b): return tuple((x - y) % self.p for x, y in zip(a, b))
    def mul(self, a, b): return self.elem(self._polymod(self._polymul(self._trim(a), self._trim(b)), self.f))

    def pow(self, a, e: int):
        r, e = self.one, int(e)
        while e:
            if e & 1:
                r = self.mul(r, a)
            a = self.mul(a, a)
            e >>= 1
        return r

    def trace(self, a):
        """sum_{i<d} a^(p^i), an element of F_p (FiniteField.hs:193-198)."""
        s, t = self.zero, a
        for _ in range(self.d):
            s = self.add(s, t)
            t = self.pow(t, self.p)
        assert all(c == 0 for c in s[1:])
        return s[0]

    def first_primitive(self):
        n = self.size - 1
        primes, x, q = [], n, 2
        while q * q <= x:
            if x % q == 0:
                primes.append(q)
                while x % q == 0:
                    x //= q
            q += 1
        if x > 1:
            primes.append(x)
        for c in itertools.product(range(self.p), repeat=self.d):
            g = tuple(reversed(c))      # constants first vary slowest: 0, 1, ..., x, x+1, ...
            if any(g) and all(self.pow(g, n // r) != self.one for r in primes):
                return g
        raise ArithmeticError("no primitive element")


# ------------------------------------------------------------------ index maps of Tensor.hs
def _digit_rev(p, e, j):
    r = 0
    for _ in range(e):
        r = r * p + j % p
        j //= p
    return r


def _index_to_pow(p, e, j):      # Tensor.hs:359-362
    jq, jr = divmod(j, p - 1)
    return p ** (e - 1) * jr + _digit_rev(p, e - 1, jq)


def _index_to_zms(p, i):         # Tensor.hs:366-368
    i1, i0 = divmod(i, p - 1)
    return p * i1 + i0 + 1


def _zms_to_index(pps, i):       # Tensor.hs:371-379
    idx, mult = 0, 1
    for p, e in pps:
        i1, i0 = divmod(i % p ** e, p)
        idx += mult * ((p - 1) * i1 + i0 - 1)
        mult *= (p - 1) * p ** (e - 1)
    return idx


def order(p: int, m: int) -> int:
    """Multiplicative order of p modulo m (ZmStar.hs:40-48)."""
    if gcd(p, m) != 1:
        raise ValueError("p and m not coprime")
    if m == 1:
        return 1
    d, x = 1, p % m
    while x != 1:
        x = x * p % m
        d += 1
    return d


def partition_cosets(p: int, m: int, m2: int) -> list[list[int]]:
    """ZmStar.hs:51-90: the cosets of Z_m'^* / <p> by representative, partitioned so that each part holds exactly one coset
    above every coset of Z_m^* / <p>."""
    if gcd(p, m2) != 1:
        raise ValueError("p and m' not coprime")
    remaining = sorted(x for x in range(1, m2 + 1) if gcd(x, m2) == 1) if m2 > 1 else [1]
    remaining = [x % m2 for x in remaining]
    left, cosets = set(remaining), []
    while left:
        x = min(left)
        c, y = {x}, x * p % m2
        while y != x:
            c.add(y)
            y = y * p % m2
        cosets.append(c)
        left -= c
    part: dict[tuple, list] = {}
    for c in cosets:
        key = tuple(sorted({y % m for y in c})) if m > 1 else (0,)
        part[key] = [c] + part.get(key, [])      # insertWith' (++): the newer coset goes first
    cols = [part[k] for k in sorted(part)]
    return [[min(col[r]) for col in cols] for r in range(len(cols[0]))]


def crt_set_dec(m: int, m2: int, p: int) -> np.ndarray:
    """crtSetDec' for O_m'/O_m modulo the prime p (p coprime to m'): int64 [count, phi(m')], residues mod p, decoding basis."""
    if m2 % m:
        raise ValueError("m must divide m'")
    pps = pps_fact(m2)
    phi = totient_fact(m2)
    d = order(p, m2)
    F = GF(p, d)
    gen = F.first_primitive()
    if (F.size - 1) % m2:
        raise ArithmeticError("m' does not divide p^d - 1")
    # per prime power: w_pp = gen^((size-1)/pp) (crtInfo of the prime power), matrix (j, i) -> w^(jToPow j * -iToZms i) * gCRT i
    mats = []
    for pr, e in pps:
        pp, ph = pr ** e, (pr - 1) * pr ** (e - 1)
        w = F.pow(gen, (F.size - 1) // pp)
        wpow = [F.one]
        for _ in range(pp - 1):
            wpow.append(F.mul(wpow[-1], w))
        wp = F.pow(gen, (F.size - 1) // pr)
        wppow = [F.one]
        for _ in range(pr - 1):
            wppow.append(F.mul(wppow[-1], wp))
        gcrt = [F.one if pr == 2 else F.sub(F.one, wppow[(i % (pr - 1) + 1) % pr]) for i in range(ph)]
        mats.append((ph, [[F.mul(wpow[(_index_to_pow(pr, e, j) * -_index_to_zms(pr, i)) % pp], gcrt[i]) for i in range(ph)]
                          for j in range(ph)]))

    def elt(j, i):      # indexK: the first prime power is the fastest index (Tensor.hs:283-288)
        r = F.one
        for ph, mat in mats:
            j, jr = divmod(j, ph)
            i, ir = divmod(i, ph)
            r = F.mul(r, mat[jr][ir])
        return r

    hinv = pow((m2 // 2 if m2 % 2 == 0 else m2) % p, -1, p)
    parts = partition_cosets(p, m, m2)
    out = np.zeros((len(parts), phi), dtype=np.int64)
    for a, reps in enumerate(parts):
        cols = [_zms_to_index(pps, i) for i in reps]
        for j in range(phi):
            s = F.zero
            for c in cols:
                s = F.add(s, elt(j, c))
            out[a, j] = hinv * F.trace(s) % p
    return out
