"""CPU suite: anchor oracle/coeffwise.py (Haskell closures in the reference, so no binary to compare with) by identities
computed with independent big-integer arithmetic."""
from fractions import Fraction

import numpy as np
import pytest

from conftest import zq_input
from oracle import coeffwise as W

QS_SETS = [[1008001, 1065601], [19393921, 18869761], [19918081, 19393921, 18869761], [2148854401, 2148249601, 2150668801],
           [17, 29], [4294967291, 4294967279]]


def _crt_centered(res, qs):
    """the integer in [-Q/2, Q/2) with the given residues"""
    Q = 1
    for q in qs:
        Q *= q
    x = 0
    for r, q in zip(res, qs):
        Qi = Q // q
        x += int(r) * Qi * pow(Qi, -1, q)
    x %= Q
    return x if 2 * x < Q else x - Q


@pytest.mark.parametrize("qs", QS_SETS, ids=lambda q: f"k{len(q)}_{q[0]}")
def test_lift_reduce_round_trip(qs):
    rng = np.random.default_rng(1)
    x = zq_input(rng, 257, qs)
    x[0], x[1] = 0, [q - 1 for q in qs]
    x[2] = [q // 2 for q in qs]
    l = W.lift(x, qs)
    q = np.asarray(qs, dtype=np.int64)
    assert np.all(2 * l < q) and np.all(2 * l >= -q)
    assert np.array_equal(W.reduce(l, qs), x)
    z = rng.integers(-2**62, 2**62, size=(257, 1))
    r = W.reduce(z, qs)
    assert all(int(r[i, t]) == int(z[i, 0]) % qs[t] for i in range(257) for t in range(len(qs)))


@pytest.mark.parametrize("qs", [q for q in QS_SETS if len(q) >= 2], ids=lambda q: f"k{len(q)}_{q[0]}")
def test_rescale_drop_is_exact_division_of_the_centred_integer(qs):
    """With X the centred integer of (x_d, x_t): rescale = (X - lift x_d) / q_d mod q_t, an exact division."""
    rng = np.random.default_rng(2)
    x = zq_input(rng, 64, qs)
    for drop in (0, len(qs) - 1):
        y = W.rescale_drop(x, qs, drop)
        keep = [t for t in range(len(qs)) if t != drop]
        for i in range(64):
            c = int(W.lift(x[i:i + 1, drop:drop + 1], [qs[drop]])[0, 0])
            for u, t in enumerate(keep):
                X = _crt_centered([x[i, drop], x[i, t]], [qs[drop], qs[t]])
                assert (X - c) % qs[drop] == 0
                assert int(y[i, u]) == ((X - c) // qs[drop]) % qs[t]


@pytest.mark.parametrize("q,qn", [(1008001, 1065601), (18869761, 256), (12289, 2), (2148249601, 18869761), (17, 4294967291)])
def test_rescale_mod_rounds_the_scaled_representative(q, qn):
    rng = np.random.default_rng(3)
    x = zq_input(rng, 500, [q])
    x[0], x[1], x[2] = 0, q - 1, q // 2
    y = W.rescale_mod(x, [q], [qn])
    for i in range(500):
        l = int(W.lift(x[i:i + 1], [q])[0, 0])
        exact = Fraction(qn * l, q)
        cands = [v for v in range(int(exact) - 2, int(exact) + 3) if abs(Fraction(v) - exact) <= Fraction(1, 2)]
        assert any(int(y[i, 0]) == v % qn for v in cands)


@pytest.mark.parametrize("p", [2, 7, 256, 1065601])
def test_round_coset_lands_in_the_coset_and_is_nearest(p):
    rng = np.random.default_rng(4)
    e = rng.standard_normal((300, 1)) * 50.0 * p
    zp = zq_input(rng, 300, [p])
    y = W.round_coset(e, zp, [p])
    assert np.array_equal(y % p, zp)
    assert np.all(np.abs(y - e) <= p / 2 + 1e-6 * p)
    assert np.array_equal(W.round_coset(np.array([[0.5], [1.5], [2.5], [-0.5], [-1.5]]), None, [p]).reshape(-1), [0, 2, 2, 0, -2])
