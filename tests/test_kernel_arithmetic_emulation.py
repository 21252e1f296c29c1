"""CPU suite: the integer-arithmetic arguments the streaming kernels rely on, emulated with Python integers exactly as the
device code computes them (same constants, same correction counts).  These are the checks quoted in DESIGN.md 4.7 / 4.8;
they do not replace the GPU parity tests, they make the bounds reproducible without a GPU.

  * coeff_stream.cu  barrett_u64 / canon64 / floor_div: mu = floor(2^64 / q), at most two corrections, any int64 input
  * coeff_stream.cu  OpRescaleMod: one conditional correction maps divModCent's quotient into [0, q')
  * coeff_stream.cu  OpRescaleDropFast (opt-in variant): 32-bit quotient estimate + Shoup multiplication
  * ext_stream.cu    twaceCRT: exact 64-bit accumulation, reduce once per floor((2^64 - q) / (q-1)^2) products
"""
import random
from math import gcd

M64 = (1 << 64) - 1
MODULI = [2, 3, 17, 257, 65537, 1008001, 2148249601, 4294967291, 4294967295, 1 << 31, (1 << 32) - 5]


def umulhi(a, b):
    return (a * b) >> 64


def barrett_u64(x, q, mu):
    r = (x - umulhi(x, mu) * q) & M64
    for _ in range(2):
        if r >= q:
            r -= q
    assert r < q
    return r


def canon64(x, q, mu):
    if 0 <= x < q:
        return x
    if x >= 0:
        return barrett_u64(x, q, mu)
    r = barrett_u64((-x) & M64, q, mu)
    return 0 if r == 0 else q - r


def floor_div(a, q, mu):
    neg = a < 0
    x = (-a - 1) if neg else a
    qh = umulhi(x, mu)
    r = x - qh * q
    for _ in range(2):
        if r >= q:
            r -= q
            qh += 1
    assert 0 <= r < q
    return -qh - 1 if neg else qh


def lift(c, q):
    return c if 2 * c < q else c - q


def test_barrett_mod_and_floor_division_for_any_int64():
    rnd = random.Random(1)
    for q in MODULI:
        mu = (1 << 64) // q
        edge = [0, 1, -1, q, q - 1, -q, -q - 1, (1 << 63) - 1, -(1 << 63), -(1 << 63) + 1, q * q, -(q * q)]
        for x in edge + [rnd.randint(-(1 << 63), (1 << 63) - 1) for _ in range(4000)]:
            assert canon64(x, q, mu) == x % q
            assert floor_div(x, q, mu) == x // q


def test_rescale_mod_needs_one_correction():
    def emu(c, q, qn):
        quot = floor_div(qn * lift(c, q) + q // 2, q, (1 << 64) // q)
        if quot < 0:
            quot += qn
        if quot >= qn:
            quot -= qn
        return quot

    ref = lambda c, q, qn: ((qn * lift(c, q) + q // 2) // q) % qn
    for q in range(2, 40):
        for qn in range(1, 40):
            assert all(emu(c, q, qn) == ref(c, q, qn) for c in range(q))
    rnd = random.Random(2)
    for _ in range(40000):
        q, qn = rnd.randint(2, 2**32 - 1), rnd.randint(1, 2**32 - 1)
        c = rnd.choice([0, q - 1, q // 2, q // 2 - 1, (q + 1) // 2, rnd.randrange(q)])
        assert emu(c, q, qn) == ref(c, q, qn)


def test_fast_limb_drop_variant():
    def fast(xt, c, qt, qd):
        mu = (1 << 64) // qt
        qh = ((c * (mu >> 32)) >> 32) & 0xFFFFFFFF
        r = c - qh * qt
        assert r >= 0
        for _ in range(3):
            if r >= qt:
                r -= qt
        assert r < qt
        dm = qd % qt
        zr = r
        if 2 * c >= qd:
            zr = zr - dm if zr >= dm else zr + (qt - dm)
        diff = xt - zr if xt >= zr else xt + (qt - zr)
        w = pow(qd % qt, -1, qt)
        sh = (diff * ((w << 32) // qt)) >> 32
        o = diff * w - sh * qt
        assert 0 <= o < 2 * qt
        return o - qt if o >= qt else o

    ref = lambda xt, c, qt, qd: (xt - lift(c, qd)) % qt * pow(qd % qt, -1, qt) % qt
    for qt in range(2, 30):
        for qd in range(2, 30):
            if gcd(qt, qd) == 1:
                assert all(fast(xt, c, qt, qd) == ref(xt, c, qt, qd) for xt in range(qt) for c in range(qd))
    big = [1008001, 1065601, 18869761, 19393921, 2148249601, 2148854401, 4294967291, 4294967279, 65537, 257, 17, 3, 2147483647]
    rnd = random.Random(3)
    for _ in range(40000):
        qt, qd = rnd.choice(big), rnd.choice(big)
        if qt != qd:
            c = rnd.choice([0, qd - 1, qd // 2, qd // 2 - 1, (qd + 1) // 2, rnd.randrange(qd)])
            xt = rnd.choice([0, qt - 1, rnd.randrange(qt)])
            assert fast(xt, c, qt, qd) == ref(xt, c, qt, qd)


def test_twace_crt_chunked_accumulation_never_overflows():
    """acc < q after a reduction, then `chunk` products of residues: the 64-bit accumulator stays below 2^64."""
    rnd = random.Random(4)
    for q in MODULI + [14401, 12289, 537133057]:
        chunk = max(1, ((1 << 64) - q) // ((q - 1) ** 2)) if q > 2 else 1 << 30
        assert (q - 1) + min(chunk, 1 << 31) * (q - 1) ** 2 <= M64 or chunk == 1
        mu = (1 << 64) // q
        for rel in (1, 2, 3, 5, 20, 97):
            xs = [rnd.choice([q - 1, rnd.randrange(q)]) for _ in range(rel)]
            ts = [rnd.choice([q - 1, rnd.randrange(q)]) for _ in range(rel)]
            acc, c = 0, min(chunk, rel)
            for r0 in range(0, rel, c):
                for r in range(r0, min(rel, r0 + c)):
                    acc += xs[r] * ts[r]
                    assert acc <= M64
                acc = barrett_u64(acc, q, mu)
            assert acc == sum(x * t for x, t in zip(xs, ts)) % q


# ---------------------------------------------------------------- fused_stream.cu: the line operators' 32-bit modes
M32 = (1 << 32) - 1


def barrett32(x, q, mu32):
    """fused_stream.cu barrett32: x < 2^32, mu32 = floor(2^32 / q) -> x mod q (one conditional subtraction through min)."""
    assert 0 <= x <= M32
    r = (x - ((x * mu32) >> 32) * q) & M32
    assert r < 2 * q
    return min(r, (r - q) & M32)


def ginv_exact(v, p, kind):
    """line_op over the integers (g.cpp:60-123): GINVPOW / GINVDEC on one line of p - 1 values."""
    d = p - 1
    v = list(v)
    if kind == "pow":
        lo, hi = sum(v), 0
        for a in range(d - 1, -1, -1):
            z = v[a]
            v[a] = (p - 1 - a) * lo - (a + 1) * hi
            lo -= z
            hi += z
    else:
        acc = sum((a + 1) * v[a] for a in range(d))
        for a in range(d - 1, 0, -1):
            keep = acc
            acc -= v[a] * p
            v[a] = keep
        v[0] = acc
    return v


def ginv_red(v, p, kind, q):
    """lineop_ginv_red: the same recurrences with every running value kept in [0, q), all arithmetic in wrapping uint32."""
    d, mu32 = p - 1, (1 << 32) // q
    v = list(v)
    if kind == "pow":
        s = sum(v)
        assert s <= M32
        lo, hi = barrett32(s, q, mu32), 0
        for a in range(d - 1, -1, -1):
            z = v[a]
            x = (p - 1 - a) * lo + p * q - (a + 1) * hi
            assert 0 <= x <= M32
            v[a] = barrett32(x, q, mu32)
            lo = lo - z if lo >= z else lo + q - z
            hi += z
            hi = hi - q if hi >= q else hi
            assert 0 <= lo < q and 0 <= hi < q
    else:
        s = sum((a + 1) * v[a] for a in range(d))
        assert s <= M32
        acc = barrett32(s, q, mu32)
        for a in range(d - 1, 0, -1):
            keep = acc
            x = acc + p * q - p * v[a]
            assert 0 <= x <= M32
            acc = barrett32(x, q, mu32)
            v[a] = keep
        v[0] = acc
    return v


def test_division_by_g_with_reduced_running_sums_stays_in_32_bits():
    """The third arithmetic mode of k_line_tile (line_tile_mode == 2): valid for 78 q < 2^32 at p <= 13; extreme and random lines,
    moduli up to the bound, against the exact integers reduced modulo q."""
    rnd = random.Random(7)
    bound = (1 << 32) // 78
    for q in (3, 8737, 3144961, 12719617, 19393921, 25159681, 55033889, bound):
        for p in (3, 5, 7, 11, 13):
            d = p - 1
            lines = [[q - 1] * d, [0] * d, [q - 1 if a % 2 else 0 for a in range(d)], [0 if a % 2 else q - 1 for a in range(d)]]
            lines += [[rnd.randrange(q) for _ in range(d)] for _ in range(40)]
            for v in lines:
                for kind in ("pow", "dec"):
                    want = [x % q for x in ginv_exact(v, p, kind)]
                    assert ginv_red(v, p, kind, q) == want, (q, p, kind, v)


def test_line_operator_intermediates_fit_the_32_bit_mode():
    """line_mult (fused_stream.cu): L, L^-1, *g keep |intermediates| <= (p + 2) q - the bias added before the reduction and the
    int32 range are sized from it; the divisions by g need p^2 q.  Checked on the extreme lines with exact integers."""
    for p in (3, 5, 7, 11, 13):
        d, q = p - 1, 1000003
        ext = [[q - 1] * d, [q - 1 if a % 2 else 0 for a in range(d)], [0 if a % 2 else q - 1 for a in range(d)]]
        for v0 in ext:
            v = list(v0)                                   # L
            for a in range(1, d):
                v[a] += v[a - 1]
            assert max(abs(x) for x in v) <= (p + 2) * q
            v = list(v0)                                   # L^-1
            for a in range(d - 1, 0, -1):
                v[a] -= v[a - 1]
            assert max(abs(x) for x in v) <= (p + 2) * q
            v = list(v0)                                   # *g, powerful basis
            last = v[d - 1]
            for a in range(d - 1, 0, -1):
                v[a] += last - v[a - 1]
            v[0] += last
            assert max(abs(x) for x in v) <= (p + 2) * q
            v = list(v0)                                   # *g, decoding basis
            acc = v[0]
            for a in range(d - 1, 0, -1):
                acc += v[a]
                v[a] -= v[a - 1]
            v[0] += acc
            assert max(abs(x) for x in v) <= (p + 2) * q and abs(acc) <= (p + 2) * q
            for kind in ("pow", "dec"):
                assert max(abs(x) for x in ginv_exact(v0, p, kind)) <= p * p * q
