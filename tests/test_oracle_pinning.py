"""CPU suite: pin the C restatement (oracle/lol_oracle.c).

 (1) against the committed golden fixtures, which are outputs of the UNMODIFIED compiled
     reference (oracle/make_golden.py);
 (2) against the compiled reference itself on fresh random inputs, when oracle/_ref exists;
 (3) against the reference's own algebraic properties (lol/Crypto/Lol/Tests/TensorTests.hs:80-131).
"""
import glob
import hashlib
import os

import numpy as np
import pytest

from conftest import CONFIG_B, NON_CRT_PARAMS, PAPER_PARAMS, REFERENCE_TEST_PARAMS, rel_err, zq_input
from oracle import tables as T

GOLDEN = sorted(g for g in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(g).startswith("ext_"))      # ext_*: ring-extension fixtures, tests/test_*_extension.py
SMALL_GOLDEN = [g for g in GOLDEN if "cfgB" not in g]
FLOAT_TOL = 1e-12   # restatement vs reference, double paths (same formulas, same order)


def test_golden_files_present():
    assert len(SMALL_GOLDEN) >= 9 and any("cfgB" in g for g in GOLDEN)


@pytest.mark.parametrize("path", SMALL_GOLDEN, ids=[os.path.basename(p)[:-4] for p in SMALL_GOLDEN])
def test_restatement_matches_golden(oracle, path):
    g = np.load(path)
    m, qs = int(g["m"]), [int(q) for q in g["qs"]]
    pe = T.pe_array(m)
    ru, rui = T.ru_tables_zq(m, qs), T.ru_tables_zq(m, qs, inverse=True)
    mh = [T.mhat_inv(m, q) for q in qs]
    y, y2 = g["rq_in"], g["rq_in2"]
    assert np.array_equal(oracle.tensorCRTRq(y, pe, ru, qs), g["CRTRq"])
    assert np.array_equal(oracle.tensorCRTInvRq(y, pe, rui, mh, qs), g["CRTInvRq"])
    for nm in ("LRq", "LInvRq", "GPowRq", "GDecRq"):
        assert np.array_equal(getattr(oracle, "tensor" + nm)(y, pe, qs), g[nm]), nm
    for nm in ("GInvPowRq", "GInvDecRq"):
        arr, st = getattr(oracle, "tensor" + nm)(y, pe, qs)
        assert st == int(g[nm + "_status"])
        if st:
            assert np.array_equal(arr, g[nm]), nm
    assert np.array_equal(oracle.mulRq(y, y2, qs), g["mulRq"])
    z = g["r_in"]
    for nm in ("LR", "LInvR", "GPowR", "GDecR"):
        assert np.array_equal(getattr(oracle, "tensor" + nm)(z, pe), g[nm]), nm
    assert oracle.tensorNormSqR(g["norm_in"], pe).reshape(-1)[0] == g["NormSqR"][0]
    d = g["d_in"]
    ruc, ruci = T.ru_tables_c(m), T.ru_tables_c(m, inverse=True)
    for nm in ("LDouble", "LInvDouble"):
        assert rel_err(getattr(oracle, "tensor" + nm)(d, pe), g[nm]) <= FLOAT_TOL
    assert rel_err(oracle.tensorNormSqD(d, pe).reshape(-1)[:1], g["NormSqD"]) <= FLOAT_TOL
    assert rel_err(oracle.tensorGaussianDec(d, pe, ruc), g["GaussianDec"]) <= FLOAT_TOL
    c, c2 = g["c_in"], g["c_in2"]
    assert rel_err(oracle.tensorCRTC(c, pe, ruc), g["CRTC"]) <= FLOAT_TOL
    assert rel_err(oracle.tensorCRTInvC(c, pe, ruci, T.mhat_inv_c(m)), g["CRTInvC"]) <= FLOAT_TOL
    for nm in ("LC", "LInvC", "GPowC", "GDecC"):
        assert rel_err(getattr(oracle, "tensor" + nm)(c, pe), g[nm]) <= FLOAT_TOL
    assert rel_err(oracle.mulC(c, c2), g["mulC"]) <= FLOAT_TOL


def test_restatement_matches_golden_config_b(oracle):
    g = np.load([p for p in GOLDEN if "cfgB" in p][0])
    m, qs = int(g["m"]), [int(q) for q in g["qs"]]
    assert (m, qs) == CONFIG_B
    rng = np.random.default_rng(int(g["seed"]))
    n = T.totient_pps(T.factor_pps(m))
    y, y2 = zq_input(rng, n, qs), zq_input(rng, n, qs)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(y) == str(g["in_digest"]) and sha(y2) == str(g["in2_digest"])
    pe = T.pe_array(m)
    crt = oracle.tensorCRTRq(y, pe, T.ru_tables_zq(m, qs), qs)
    assert np.array_equal(crt[:8], g["CRTRq_head"])
    assert sha(crt) == str(g["CRTRq_digest"])
    mh = [T.mhat_inv(m, q) for q in qs]
    assert sha(oracle.tensorCRTInvRq(y, pe, T.ru_tables_zq(m, qs, inverse=True), mh, qs)) == str(g["CRTInvRq_digest"])
    assert sha(oracle.mulRq(y, y2, qs)) == str(g["mulRq_digest"])


@pytest.mark.parametrize("m,qs", REFERENCE_TEST_PARAMS + PAPER_PARAMS, ids=lambda v: str(v))
def test_restatement_matches_compiled_reference(oracle, reference, m, qs):
    rng = np.random.default_rng(m * 7919 + len(qs))
    pe = T.pe_array(m)
    n = T.totient_pps(T.factor_pps(m))
    ru, rui = T.ru_tables_zq(m, qs), T.ru_tables_zq(m, qs, inverse=True)
    mh = [T.mhat_inv(m, q) for q in qs]
    y, y2 = zq_input(rng, n, qs), zq_input(rng, n, qs)
    assert np.array_equal(oracle.tensorCRTRq(y, pe, ru, qs), reference.tensorCRTRq(y, pe, ru, qs))
    assert np.array_equal(oracle.tensorCRTInvRq(y, pe, rui, mh, qs), reference.tensorCRTInvRq(y, pe, rui, mh, qs))
    for nm in ("tensorLRq", "tensorLInvRq", "tensorGPowRq", "tensorGDecRq"):
        assert np.array_equal(getattr(oracle, nm)(y, pe, qs), getattr(reference, nm)(y, pe, qs)), nm
    for nm in ("tensorGInvPowRq", "tensorGInvDecRq"):
        (a, sa), (b, sb) = getattr(oracle, nm)(y, pe, qs), getattr(reference, nm)(y, pe, qs)
        assert sa == sb and (sa == 0 or np.array_equal(a, b)), nm
    assert np.array_equal(oracle.mulRq(y, y2, qs), reference.mulRq(y, y2, qs))
    z = rng.integers(-(2 ** 62), 2 ** 62, size=(n, 1)).astype(np.int64)     # wrapping arithmetic included
    for nm in ("tensorLR", "tensorLInvR", "tensorGPowR", "tensorGDecR"):
        assert np.array_equal(getattr(oracle, nm)(z, pe), getattr(reference, nm)(z, pe)), nm
    assert oracle.tensorNormSqR(z, pe).flat[0] == reference.tensorNormSqR(z, pe).flat[0]
    d = rng.normal(size=(n, 1))
    c = rng.normal(size=(n, 1)) + 1j * rng.normal(size=(n, 1))
    ruc, ruci = T.ru_tables_c(m), T.ru_tables_c(m, inverse=True)
    assert rel_err(oracle.tensorCRTC(c, pe, ruc), reference.tensorCRTC(c, pe, ruc)) <= FLOAT_TOL
    assert rel_err(oracle.tensorCRTInvC(c, pe, ruci, T.mhat_inv_c(m)), reference.tensorCRTInvC(c, pe, ruci, T.mhat_inv_c(m))) <= FLOAT_TOL
    assert rel_err(oracle.tensorGaussianDec(d, pe, ruc), reference.tensorGaussianDec(d, pe, ruc)) <= FLOAT_TOL
    assert rel_err(oracle.tensorNormSqD(d, pe).flat[0], reference.tensorNormSqD(d, pe).flat[0]) <= FLOAT_TOL
    for nm in ("tensorLDouble", "tensorLInvDouble"):
        assert rel_err(getattr(oracle, nm)(d, pe), getattr(reference, nm)(d, pe)) <= FLOAT_TOL
    for nm in ("tensorLC", "tensorLInvC", "tensorGPowC", "tensorGDecC"):
        assert rel_err(getattr(oracle, nm)(c, pe), getattr(reference, nm)(c, pe)) <= FLOAT_TOL


@pytest.mark.parametrize("m,qs", NON_CRT_PARAMS, ids=lambda v: str(v))
def test_restatement_line_ops_composite_modulus(oracle, reference, m, qs):
    rng = np.random.default_rng(m)
    pe = T.pe_array(m)
    n = T.totient_pps(T.factor_pps(m))
    y = zq_input(rng, n, qs)
    for nm in ("tensorLRq", "tensorLInvRq", "tensorGPowRq", "tensorGDecRq"):
        assert np.array_equal(getattr(oracle, nm)(y, pe, qs), getattr(reference, nm)(y, pe, qs)), nm
    for nm in ("tensorGInvPowRq", "tensorGInvDecRq"):
        (a, sa), (b, sb) = getattr(oracle, nm)(y, pe, qs), getattr(reference, nm)(y, pe, qs)
        assert sa == sb and (sa == 0 or np.array_equal(a, b)), nm


@pytest.mark.parametrize("m,qs", REFERENCE_TEST_PARAMS, ids=lambda v: str(v))
def test_reference_properties_hold_for_restatement(oracle, m, qs):
    """TensorTests.hs: crtInv.crt = id (:115-119), lInv.l = id (:122-123), divG.mulG = id (:87-101),
    mulGDec = lInv.mulGPow.l (:104-105), mulGCRT = crt.mulGPow.crtInv (:107-112), scalarCRT = crt.scalarPow (:126-131)."""
    rng = np.random.default_rng(m + 17)
    pe = T.pe_array(m)
    n = T.totient_pps(T.factor_pps(m))
    ru, rui = T.ru_tables_zq(m, qs), T.ru_tables_zq(m, qs, inverse=True)
    mh = [T.mhat_inv(m, q) for q in qs]
    y = zq_input(rng, n, qs)
    crt = lambda v: oracle.tensorCRTRq(v, pe, ru, qs)
    crtinv = lambda v: oracle.tensorCRTInvRq(v, pe, rui, mh, qs)
    assert np.array_equal(crtinv(crt(y)), y)
    assert np.array_equal(oracle.tensorLInvRq(oracle.tensorLRq(y, pe, qs), pe, qs), y)
    for mul, div in (("tensorGPowRq", "tensorGInvPowRq"), ("tensorGDecRq", "tensorGInvDecRq")):
        out, st = getattr(oracle, div)(getattr(oracle, mul)(y, pe, qs), pe, qs)
        assert st == 1 and np.array_equal(out, y)
    assert np.array_equal(oracle.tensorGDecRq(y, pe, qs),
                          oracle.tensorLInvRq(oracle.tensorGPowRq(oracle.tensorLRq(y, pe, qs), pe, qs), pe, qs))
    g, gi = T.g_crt_vectors(m, qs)
    assert np.array_equal(oracle.mulRq(y, g, qs), crt(oracle.tensorGPowRq(crtinv(y), pe, qs)))
    assert np.array_equal(oracle.mulRq(oracle.mulRq(y, g, qs), gi, qs), y)
    s = np.zeros((n, len(qs)), dtype=np.int64)
    s[0, :] = [5 % q for q in qs]
    assert np.array_equal(crt(s), np.tile(s[0], (n, 1)))


def test_table_anchors():
    # values checked against the running reference at survey time
    assert T.smallest_generator(14401) == 11 and T.omega(14400, 14401) == 11 and T.mhat_inv(14400, 14401) == 14399
    assert T.smallest_generator(786433) == 10 and T.omega(65536, 786433) == 108788
    assert [q for q, _ in zip(T.good_qs(65536, 2 ** 29), range(4))] == CONFIG_B[1]


@pytest.mark.parametrize("e,q", [(5, 97), (8, 257), (11, 12289), (13, 40961)], ids=lambda v: str(v))
def test_pow2_crt_is_negacyclic_evaluation(oracle, e, q):
    """What lol_b200/csrc/fused_pow2_df.cu relies on: for m = 2^e the reference's crtTwiddle + DFT rounds
    (crt.cpp:43-58, 459-486) evaluate f(x) = sum_i y[i] x^rev(i) at psi^(2 pos + 1), psi = ru[0][1], so the
    twist-free Cooley-Tukey rounds with T_r[p] = psi^((2p+1) n / 2^(r+1)) give the same residues."""
    m, n = 1 << e, 1 << (e - 1)
    pe = T.pe_array(m)
    ru, rui = T.ru_tables_zq(m, [q]), T.ru_tables_zq(m, [q], inverse=True)
    y = zq_input(np.random.default_rng(e), n, [q])
    want = oracle.tensorCRTRq(y, pe, ru, [q])
    root = [int(v) for v in ru[0].reshape(-1)]
    v = [int(c) for c in y.reshape(-1)]
    for r in range(e - 1):
        st = 1 << r
        tw = [root[(2 * p + 1) * (n >> (r + 1))] for p in range(st)]
        for pos in range(n):
            if not pos & st:
                u, t = v[pos], v[pos + st] * tw[pos & (st - 1)] % q
                v[pos], v[pos + st] = (u + t) % q, (u - t) % q
    assert v == [int(c) for c in want.reshape(-1)]
    if e <= 8:        # direct evaluation
        rev = lambda i: int(format(i, f"0{e - 1}b")[::-1], 2)
        for pos in (0, 1, n // 2, n - 1):
            x = pow(root[1], 2 * pos + 1, q)
            assert sum(int(y[i, 0]) * pow(x, rev(i), q) for i in range(n)) % q == int(want[pos, 0])
    # inverse: rounds descending with inverse twiddles, then mhat^-1
    rooti = [int(c) for c in rui[0].reshape(-1)]
    w = [int(c) for c in want.reshape(-1)]
    for r in range(e - 2, -1, -1):
        st = 1 << r
        tw = [rooti[(2 * p + 1) * (n >> (r + 1))] for p in range(st)]
        for pos in range(n):
            if not pos & st:
                u, t = w[pos], w[pos + st]
                w[pos], w[pos + st] = (u + t) % q, (u - t) * tw[pos & (st - 1)] % q
    s = T.mhat_inv(m, q)
    assert [c * s % q for c in w] == [int(c) for c in y.reshape(-1)]


# ------------------------------------------------------------------ SymmSHE host-side steps (oracle/symmshe.py)
@pytest.mark.parametrize("qs", [[1008001, 1065601], [14401], [17, 257, 65537], [2148249601]], ids=str)
@pytest.mark.parametrize("base", [0, 2, 3, 16, 1000], ids=lambda b: f"base{b}")
def test_gadget_identity_and_digit_ranges(qs, base):
    """sum_i gadget_i * decompose(x)_i = x (Gadget.hs:60-66), digits centered in [-b/2, b/2) (Numeric.hs:225-234),
    lift in [-q/2, q/2) (ZqBasic.hs:91-94), gadlen = number of base-b digits of q (ZqBasic.hs:241-243)."""
    from oracle import symmshe as S
    rng = np.random.default_rng(len(qs) * 1000 + base)
    x = zq_input(rng, 24, qs, batch=3)
    x[0, 0, :] = 0
    x[0, 1, :] = [q - 1 for q in qs]
    x[0, 2, :] = [q // 2 for q in qs]
    x[0, 3, :] = [(q + 1) // 2 for q in qs]
    g, ints, red = S.gadget(qs, base), S.decompose(x, qs, base), S.decompose_reduced(x, qs, base)
    assert len(g) == len(ints) == red.shape[0] == S.gadget_length(qs, base)
    acc = np.zeros_like(x)
    for gi, di in zip(g, red):
        acc = (acc + di * np.asarray(gi, dtype=np.int64)) % np.asarray(qs)
    assert np.array_equal(acc, x)
    pos = 0
    for l, q in enumerate(qs):
        lf = S.lift(x[..., l], q)
        assert lf.min() >= -(q // 2) - (q % 2 == 0) * 0 - 1 and 2 * lf.max() < q and np.array_equal(lf % q, x[..., l])
        nd = 1 if base == 0 else S.gadlen(base, q)
        if base:
            assert base ** (nd - 1) <= q < base ** nd
            for d in ints[pos:pos + nd - 1]:
                assert d.min() >= -(base // 2) and d.max() < base - base // 2
        pos += nd


def test_div_mod_cent_matches_haskell_semantics():
    from oracle import symmshe as S
    a = np.arange(-50, 50, dtype=np.int64)
    for b in (2, 3, 7, 10):
        quo, r = S.div_mod_cent(a, b)
        assert np.array_equal(quo * b + r, a) and r.min() >= -(b // 2) and r.max() < b - b // 2
