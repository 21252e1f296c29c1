"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

  * every one of the 29 drop-in symbols vs the oracle on the reference's own test parameters
    (bit-exact for Zq / int64, 1e-9 relative for double / complex: BASELINE.json north_star);
  * the batched device-resident entry points vs the oracle, element by element;
  * the committed golden fixtures (outputs of the compiled reference), incl. config B digests;
  * the library's own root tables (ZqBasic.hs restated in C++) vs the oracle's table builder;
  * at BASELINE.json's full sizes: the reference's algebraic properties (TensorTests.hs:80-131)
    -- crtInv.crt = id, linearity, mulGCRT = crt.mulGPow.crtInv -- as size-independent checks.

Nothing here reads /root/reference.
"""
import glob
import hashlib
import os

import numpy as np
import pytest

from conftest import (CONFIG_A, CONFIG_B, CONFIG_C, NON_CRT_PARAMS, PAPER_PARAMS, REFERENCE_TEST_PARAMS, rel_err,
                      zq_input)
from oracle import tables as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle(gpu_oracle):
    """GPU suites compare with the compiled reference when it is present (conftest.gpu_oracle)."""
    return gpu_oracle

FLOAT_TOL = 1e-9      # BASELINE.json: "within a stated relative tolerance (e.g. 1e-9)"
GOLDEN = sorted(g for g in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(g).startswith("ext_"))      # ext_*: ring-extension fixtures, tests/test_*_extension.py
SMALL_GOLDEN = [g for g in GOLDEN if "cfgB" not in g]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU suite needs a CUDA device: libctensor_b200 has no CPU path")
    return torch


@pytest.fixture(scope="module")
def dropin(torch_cuda):
    from lol_b200 import build_library, capi
    build_library()
    assert capi.device_available()
    return capi.DropIn()


def _tables(m, qs):
    return (T.pe_array(m), T.totient_pps(T.factor_pps(m)), T.ru_tables_zq(m, qs), T.ru_tables_zq(m, qs, inverse=True),
            [T.mhat_inv(m, q) for q in qs])


# ------------------------------------------------------------------ drop-in symbols
@pytest.mark.parametrize("m,qs", REFERENCE_TEST_PARAMS + PAPER_PARAMS + [CONFIG_C], ids=lambda v: str(v))
def test_dropin_zq_symbols_bit_exact(dropin, oracle, m, qs):
    from lol_b200 import capi
    before = capi.kernel_launch_count()
    rng = np.random.default_rng(m * 31 + len(qs))
    pe, n, ru, rui, mh = _tables(m, qs)
    y, y2 = zq_input(rng, n, qs), zq_input(rng, n, qs)
    assert np.array_equal(dropin.tensorCRTRq(y, pe, ru, qs), oracle.tensorCRTRq(y, pe, ru, qs))
    assert np.array_equal(dropin.tensorCRTInvRq(y, pe, rui, mh, qs), oracle.tensorCRTInvRq(y, pe, rui, mh, qs))
    for nm in ("tensorLRq", "tensorLInvRq", "tensorGPowRq", "tensorGDecRq"):
        assert np.array_equal(getattr(dropin, nm)(y, pe, qs), getattr(oracle, nm)(y, pe, qs)), nm
    for nm in ("tensorGInvPowRq", "tensorGInvDecRq"):
        (a, sa), (b, sb) = getattr(dropin, nm)(y, pe, qs), getattr(oracle, nm)(y, pe, qs)
        assert sa == sb == 1 and np.array_equal(a, b), nm
    assert np.array_equal(dropin.mulRq(y, y2, qs), oracle.mulRq(y, y2, qs))
    assert capi.kernel_launch_count() >= before + 3      # the CUDA path ran (L, G are the identity without an odd prime: no launch)


@pytest.mark.parametrize("m,qs", NON_CRT_PARAMS, ids=lambda v: str(v))
def test_dropin_line_ops_composite_modulus(dropin, oracle, m, qs):
    rng = np.random.default_rng(m)
    pe = T.pe_array(m)
    n = T.totient_pps(T.factor_pps(m))
    y = zq_input(rng, n, qs)
    for nm in ("tensorLRq", "tensorLInvRq", "tensorGPowRq", "tensorGDecRq"):
        assert np.array_equal(getattr(dropin, nm)(y, pe, qs), getattr(oracle, nm)(y, pe, qs)), nm
    for nm in ("tensorGInvPowRq", "tensorGInvDecRq"):
        (a, sa), (b, sb) = getattr(dropin, nm)(y, pe, qs), getattr(oracle, nm)(y, pe, qs)
        assert sa == sb, nm                      # 0 when rad_odd(m) is not a unit mod q (g.cpp:196-198)
        if sa:
            assert np.array_equal(a, b), nm


@pytest.mark.parametrize("m", [1, 2, 7, 8, 12, 21, 42, 89, 1728, 14400], ids=str)
def test_dropin_int64_symbols_bit_exact(dropin, oracle, m):
    rng = np.random.default_rng(m + 5)
    pe = T.pe_array(m)
    n = T.totient_pps(T.factor_pps(m))
    z = rng.integers(-(2 ** 62), 2 ** 62, size=(n, 1)).astype(np.int64)      # exercises wrap-around
    for nm in ("tensorLR", "tensorLInvR", "tensorGPowR", "tensorGDecR"):
        assert np.array_equal(getattr(dropin, nm)(z, pe), getattr(oracle, nm)(z, pe)), nm
    zs = rng.integers(-8, 9, size=(n, 1)).astype(np.int64)
    assert dropin.tensorNormSqR(zs, pe).flat[0] == oracle.tensorNormSqR(zs, pe).flat[0]
    assert dropin.tensorNormSqR(z, pe).flat[0] == oracle.tensorNormSqR(z, pe).flat[0]
    # divG . mulG = id over Z (TensorTests.hs:87-101), and non-multiples are refused
    for mul, div in (("tensorGPowR", "tensorGInvPowR"), ("tensorGDecR", "tensorGInvDecR")):
        g = getattr(dropin, mul)(zs, pe)
        out, st = getattr(dropin, div)(g, pe)
        ref, rst = getattr(oracle, div)(g, pe)
        assert st == rst == 1 and np.array_equal(out, zs) and np.array_equal(ref, zs)
        if T.odd_radical(m) > 1:
            bad = np.zeros((n, 1), dtype=np.int64)
            bad[0, 0] = 1
            _, st = getattr(dropin, div)(bad, pe)
            _, rst = getattr(oracle, div)(bad, pe)
            assert st == rst == 0


@pytest.mark.parametrize("m", [1, 4, 7, 12, 21, 42, 89, 1728, 14400], ids=str)
def test_dropin_double_and_complex_symbols(dropin, oracle, m):
    rng = np.random.default_rng(m + 11)
    pe = T.pe_array(m)
    n = T.totient_pps(T.factor_pps(m))
    d = rng.normal(size=(n, 1))
    c = rng.normal(size=(n, 1)) + 1j * rng.normal(size=(n, 1))
    c2 = rng.normal(size=(n, 1)) + 1j * rng.normal(size=(n, 1))
    ruc, ruci = T.ru_tables_c(m), T.ru_tables_c(m, inverse=True)
    assert rel_err(dropin.tensorCRTC(c, pe, ruc), oracle.tensorCRTC(c, pe, ruc)) <= FLOAT_TOL
    assert rel_err(dropin.tensorCRTInvC(c, pe, ruci, T.mhat_inv_c(m)), oracle.tensorCRTInvC(c, pe, ruci, T.mhat_inv_c(m))) <= FLOAT_TOL
    assert rel_err(dropin.tensorGaussianDec(d, pe, ruc), oracle.tensorGaussianDec(d, pe, ruc)) <= FLOAT_TOL
    assert rel_err(dropin.tensorNormSqD(d, pe).flat[0], oracle.tensorNormSqD(d, pe).flat[0]) <= FLOAT_TOL
    for nm in ("tensorLDouble", "tensorLInvDouble"):
        assert rel_err(getattr(dropin, nm)(d, pe), getattr(oracle, nm)(d, pe)) <= FLOAT_TOL, nm
    for nm in ("tensorLC", "tensorLInvC", "tensorGPowC", "tensorGDecC"):
        assert rel_err(getattr(dropin, nm)(c, pe), getattr(oracle, nm)(c, pe)) <= FLOAT_TOL, nm
    for nm in ("tensorGInvPowC", "tensorGInvDecC"):
        (a, sa), (b, sb) = getattr(dropin, nm)(c, pe), getattr(oracle, nm)(c, pe)
        assert sa == sb == 1 and rel_err(a, b) <= FLOAT_TOL, nm
    assert rel_err(dropin.mulC(c, c2), oracle.mulC(c, c2)) <= FLOAT_TOL
    # crtInv . crt = id over C (TensorTests.hs:115-119)
    assert rel_err(dropin.tensorCRTInvC(dropin.tensorCRTC(c, pe, ruc), pe, ruci, T.mhat_inv_c(m)), c) <= FLOAT_TOL


def test_dropin_tuple_layout_complex_k2(dropin, oracle):
    m, k = 21, 2
    rng = np.random.default_rng(3)
    pe = T.pe_array(m)
    n = T.totient_pps(T.factor_pps(m))
    c = rng.normal(size=(n, k)) + 1j * rng.normal(size=(n, k))
    d = rng.normal(size=(n, k))
    ruc = T.ru_tables_c(m, k)
    assert rel_err(dropin.tensorCRTC(c, pe, ruc, k), oracle.tensorCRTC(c, pe, ruc, k)) <= FLOAT_TOL
    assert rel_err(dropin.tensorGaussianDec(d, pe, ruc, k), oracle.tensorGaussianDec(d, pe, ruc, k)) <= FLOAT_TOL
    assert rel_err(dropin.tensorNormSqD(d, pe, k).flat[:k], oracle.tensorNormSqD(d, pe, k).flat[:k]) <= FLOAT_TOL
    assert rel_err(dropin.tensorLC(c, pe, k), oracle.tensorLC(c, pe, k)) <= FLOAT_TOL
    z = rng.integers(-100, 100, size=(n, k)).astype(np.int64)
    assert np.array_equal(dropin.tensorGPowR(z, pe, k), oracle.tensorGPowR(z, pe, k))


# ------------------------------------------------------------------ golden fixtures (outputs of the compiled reference)
@pytest.mark.parametrize("path", SMALL_GOLDEN, ids=[os.path.basename(p)[:-4] for p in SMALL_GOLDEN])
def test_dropin_matches_reference_golden(dropin, path):
    g = np.load(path)
    m, qs = int(g["m"]), [int(q) for q in g["qs"]]
    pe, n, ru, rui, mh = _tables(m, qs)
    y, y2 = g["rq_in"], g["rq_in2"]
    assert np.array_equal(dropin.tensorCRTRq(y, pe, ru, qs), g["CRTRq"])
    assert np.array_equal(dropin.tensorCRTInvRq(y, pe, rui, mh, qs), g["CRTInvRq"])
    for nm in ("LRq", "LInvRq", "GPowRq", "GDecRq"):
        assert np.array_equal(getattr(dropin, "tensor" + nm)(y, pe, qs), g[nm]), nm
    for nm in ("GInvPowRq", "GInvDecRq"):
        arr, st = getattr(dropin, "tensor" + nm)(y, pe, qs)
        assert st == int(g[nm + "_status"]) and np.array_equal(arr, g[nm]), nm
    assert np.array_equal(dropin.mulRq(y, y2, qs), g["mulRq"])
    for nm in ("LR", "LInvR", "GPowR", "GDecR"):
        assert np.array_equal(getattr(dropin, "tensor" + nm)(g["r_in"], pe), g[nm]), nm
    assert dropin.tensorNormSqR(g["norm_in"], pe).flat[0] == g["NormSqR"][0]
    ruc, ruci = T.ru_tables_c(m), T.ru_tables_c(m, inverse=True)
    assert rel_err(dropin.tensorGaussianDec(g["d_in"], pe, ruc), g["GaussianDec"]) <= FLOAT_TOL
    assert rel_err(dropin.tensorNormSqD(g["d_in"], pe).flat[0], g["NormSqD"][0]) <= FLOAT_TOL
    assert rel_err(dropin.tensorCRTC(g["c_in"], pe, ruc), g["CRTC"]) <= FLOAT_TOL
    assert rel_err(dropin.tensorCRTInvC(g["c_in"], pe, ruci, T.mhat_inv_c(m)), g["CRTInvC"]) <= FLOAT_TOL
    for nm in ("LDouble", "LInvDouble"):
        assert rel_err(getattr(dropin, "tensor" + nm)(g["d_in"], pe), g[nm]) <= FLOAT_TOL
    for nm in ("LC", "LInvC", "GPowC", "GDecC"):
        assert rel_err(getattr(dropin, "tensor" + nm)(g["c_in"], pe), g[nm]) <= FLOAT_TOL
    assert rel_err(dropin.mulC(g["c_in"], g["c_in2"]), g["mulC"]) <= FLOAT_TOL


def test_config_b_matches_reference_digests(dropin):
    g = np.load([p for p in GOLDEN if "cfgB" in p][0])
    m, qs = int(g["m"]), [int(q) for q in g["qs"]]
    rng = np.random.default_rng(int(g["seed"]))
    pe, n, ru, rui, mh = _tables(m, qs)
    y, y2 = zq_input(rng, n, qs), zq_input(rng, n, qs)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert sha(y) == str(g["in_digest"])
    crt = dropin.tensorCRTRq(y, pe, ru, qs)
    assert np.array_equal(crt[:8], g["CRTRq_head"])
    assert sha(crt) == str(g["CRTRq_digest"])
    assert sha(dropin.tensorCRTInvRq(y, pe, rui, mh, qs)) == str(g["CRTInvRq_digest"])
    assert sha(dropin.mulRq(y, y2, qs)) == str(g["mulRq_digest"])


# ------------------------------------------------------------------ batched device API
# (14400, [429336001]): largest prime = 1 mod 14400 the 64-bit-accumulate fused kernel accepts (10q < 2^32);
# (14400, [43201]): first modulus past the 32-bit-accumulate bound; the triple mixes both arithmetic classes
BATCH_PARAMS = [(7, [29]), (33, [67]), (77, [463]), (143, [859]), (448, [3144961]), (448, [449, 3144961]), (42, [19393921, 18869761]), (42, [2148854401, 2148249601, 2150668801]), (89, [179]),
                (1024, [12289]), (64 * 27, [3457]), CONFIG_A, CONFIG_C, (14400, [429336001]), (14400, [43201]),
                (14400, [14401, 1008001, 429336001]), (14400, [2148249601]),
                # tupSize 3 .. 8: the de-interleaving kernel (k_fused_a_kd), every limb of an element in one CTA iteration
                (14400, [43201, 57601, 100801, 115201]), (14400, [14401, 43201, 57601, 100801, 115201]),
                (14400, [14401, 43201, 57601, 100801, 115201, 172801, 259201]),
                (14400, [14401, 43201, 57601, 100801, 115201, 172801, 259201, 273601]),
                (14400, [43201, 57601, 100801, 115201, 172801, 259201]),      # tupSize 6: odd limb stride in the de-interleaved tile
                # fused_w: the reference's other benchmark rings (Benchmarks/Default.hs:41-48) + m = 2016, both arithmetic classes
                (64 * 81, [10369]), (32 * 7 * 13, [8737]), (8 * 7 * 13, [8737]), (8 * 5 * 7 * 13, [14561]), (2016, [2017]),
                (64 * 27, [3457, 1002241]), (64 * 81, [10031041]), (32 * 7 * 13, [101921]), (8 * 5 * 7 * 13, [1015561, 1026481]),
                (64 * 7 * 13, [3144961]), (64 * 7 * 13, [23297]),
                (64 * 7 * 13, [19393921, 18869761]), (32 * 7 * 13, [25159681, 19918081, 19393921]), (64 * 7 * 13, [25159681, 19918081, 19393921, 18869761]),      # HomomPRF chains: line_tile with tupSize folded
                (128 * 7 * 13, [23297]), (128 * 7 * 13, [3144961]),      # a = 7: two column halves per lane
                (4 * 3 * 5 * 7 * 13, [3144961]), (9 * 5 * 7 * 13, [3144961]), (4 * 3 * 5 * 7 * 13, [21841]), (9 * 5 * 7 * 13, [8191])]      # four odd prime powers      # lol-apps tunnel ring H1 at its modulus, and the Twace-Embed benchmark modulus (Montgomery class)
FUSED_W_INDICES = {64 * 27, 64 * 81, 32 * 7 * 13, 8 * 7 * 13, 8 * 5 * 7 * 13, 2016, 64 * 7 * 13, 128 * 7 * 13, 5460, 4095, 448}


@pytest.mark.parametrize("force_generic", [False, True], ids=["auto", "generic"])
@pytest.mark.parametrize("m,qs", BATCH_PARAMS, ids=lambda v: str(v))
def test_batched_rq_matches_oracle(torch_cuda, oracle, m, qs, force_generic):
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    B = 37                                   # ragged: not a multiple of any tile size
    rng = np.random.default_rng(m + 99)
    pe, n, ru, rui, mh = _tables(m, qs)
    t = CudaTensorRq(m, qs)                  # tables derived inside the library
    if m in FUSED_W_INDICES:
        assert t.plan.kernel_name("CRT") == "fused_w" and t.plan.kernel_name("CRTInv") == "fused_w"
    t.plan.force_generic(force_generic)
    for i in range(len(pe)):
        assert np.array_equal(t.plan.ru_table(i), ru[i]) and np.array_equal(t.plan.ru_table(i, True), rui[i])
    assert list(t.plan.mhatinv()) == mh
    y = zq_input(rng, n, qs, batch=B)
    y2 = zq_input(rng, n, qs, batch=B)
    x, x2 = torch.from_numpy(y).cuda(), torch.from_numpy(y2).cuda()
    per_elem = lambda f: np.stack([f(y[b]) for b in range(B)])
    assert np.array_equal(t.crt(x).cpu().numpy(), per_elem(lambda v: oracle.tensorCRTRq(v, pe, ru, qs)))
    assert np.array_equal(t.crtInv(x).cpu().numpy(), per_elem(lambda v: oracle.tensorCRTInvRq(v, pe, rui, mh, qs)))
    for meth, nm in (("l", "tensorLRq"), ("lInv", "tensorLInvRq"), ("mulGPow", "tensorGPowRq"), ("mulGDec", "tensorGDecRq")):
        assert np.array_equal(getattr(t, meth)(x).cpu().numpy(), per_elem(lambda v: getattr(oracle, nm)(v, pe, qs))), nm
    for meth, nm in (("divGPow", "tensorGInvPowRq"), ("divGDec", "tensorGInvDecRq")):
        assert np.array_equal(getattr(t, meth)(x).cpu().numpy(), per_elem(lambda v: getattr(oracle, nm)(v, pe, qs)[0])), nm
    assert np.array_equal(t.mul(x, x2).cpu().numpy(), np.stack([oracle.mulRq(y[b], y2[b], qs) for b in range(B)]))
    g, gi = T.g_crt_vectors(m, qs)
    assert np.array_equal(t.mulGCRT(x).cpu().numpy(), per_elem(lambda v: oracle.mulRq(v, g, qs)))
    assert np.array_equal(t.divGCRT(x).cpu().numpy(), per_elem(lambda v: oracle.mulRq(v, gi, qs)))
    assert torch.equal(x, torch.from_numpy(y).cuda())      # pure: inputs untouched
    # empty batch is a no-op
    e = torch.empty(0, n, len(qs), dtype=torch.int64, device="cuda")
    assert t.crt(e).shape[0] == 0


@pytest.mark.parametrize("qs", [CONFIG_C[1], [14401, 43201], [14401, 1008001, 429336001]], ids=lambda v: str(v))
def test_deinterleaving_kernel_small_tuples(torch_cuda, oracle, monkeypatch, qs):
    """k_fused_a_kd is the default from tupSize 4 on (BATCH_PARAMS above); LOLB_FUSED_A_KD=1 selects it for tupSize 2 and 3 as
    well, where k_fused_a_k2 / k_fused_a_kn measured faster.  Same parity bar: bit-exact against the oracle, ragged batch."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    monkeypatch.setenv("LOLB_FUSED_A_KD", "1")
    m, B = 14400, 19
    rng = np.random.default_rng(len(qs) + 41)
    pe, n, ru, rui, mh = _tables(m, qs)
    t = CudaTensorRq(m, qs)
    y = zq_input(rng, n, qs, batch=B)
    y[3, :5] = y[3, :5] + np.array(qs) * 3      # non-canonical input is reduced like the reference's `c % q`
    x = torch.from_numpy(y).cuda()
    assert np.array_equal(t.crt(x).cpu().numpy(), np.stack([oracle.tensorCRTRq(y[b], pe, ru, qs) for b in range(B)]))
    assert np.array_equal(t.crtInv(x).cpu().numpy(), np.stack([oracle.tensorCRTInvRq(y[b], pe, rui, mh, qs) for b in range(B)]))


@pytest.mark.parametrize("m,qs", [CONFIG_A, CONFIG_C, (14400, [14401, 1008001, 429336001]), (42, [8191]), (2 ** 13, [537133057])],
                         ids=lambda v: str(v))
def test_fused_crt_mul_pairs(torch_cuda, oracle, m, qs):
    """lolb_crtMulRq / lolb_mulCrtInvRq (one pass at m = 14400, both arithmetic classes and tupSize 1/2/3; two kernels
    back to back elsewhere) are bit-identical to the reference's two calls in sequence: tensorCRTRq then mulRq, mulRq then
    tensorCRTInvRq; full-batch and broadcast second operand; and crtInv(crt(a) * crt(b)) is the negacyclic-style ring
    product both ways round (commutativity as a size-independent check)."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    B = 13
    rng = np.random.default_rng(m + 7)
    pe, n, ru, rui, mh = _tables(m, qs)
    t = CudaTensorRq(m, qs)
    ya, yb = zq_input(rng, n, qs, batch=B), zq_input(rng, n, qs, batch=B)
    a, b = torch.from_numpy(ya).cuda(), torch.from_numpy(yb).cuda()
    b1 = b[:1].contiguous()
    want = np.stack([oracle.mulRq(oracle.tensorCRTRq(ya[i], pe, ru, qs), yb[i], qs) for i in range(B)])
    assert np.array_equal(t.crtMul(a, b).cpu().numpy(), want)
    want1 = np.stack([oracle.mulRq(oracle.tensorCRTRq(ya[i], pe, ru, qs), yb[0], qs) for i in range(B)])
    assert np.array_equal(t.crtMul(a, b1).cpu().numpy(), want1)
    wanti = np.stack([oracle.tensorCRTInvRq(oracle.mulRq(ya[i], yb[i], qs), pe, rui, mh, qs) for i in range(B)])
    assert np.array_equal(t.mulCrtInv(a, b).cpu().numpy(), wanti)
    wanti1 = np.stack([oracle.tensorCRTInvRq(oracle.mulRq(ya[i], yb[0], qs), pe, rui, mh, qs) for i in range(B)])
    assert np.array_equal(t.mulCrtInv(a, b1).cpu().numpy(), wanti1)
    assert torch.equal(a, torch.from_numpy(ya).cuda()) and torch.equal(b, torch.from_numpy(yb).cuda())      # pure
    # ring product a * b = crtInv(crt a . crt b), either way round
    ab = t.mulCrtInv(t.crtMul(a, t.crt(b)), torch.ones_like(b1))
    ba = t.mulCrtInv(t.crtMul(b, t.crt(a)), torch.ones_like(b1))
    assert torch.equal(ab, ba) and torch.equal(ab, t.crtInv(t.mul(t.crt(a), t.crt(b))))
    t.plan.force_generic(True)
    assert np.array_equal(t.crtMul(a, b).cpu().numpy(), want) and np.array_equal(t.mulCrtInv(a, b).cpu().numpy(), wanti)


@pytest.mark.parametrize("e", list(range(5, 17)), ids=lambda e: f"m=2^{e}")
def test_power_of_two_indices(torch_cuda, oracle, e):
    """Every power-of-two index up to config B's 2^16 (fused NTT from 2^7 on, generic below), RNS pair with a
    20-bit and a 30-bit prime, ragged batch; crtInv . crt = id and element-wise parity with the oracle."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    m, qs = 2 ** e, [786433, 537133057]
    B = 3 if e >= 14 else 11
    rng = np.random.default_rng(e)
    pe, n, ru, rui, mh = _tables(m, qs)
    t = CudaTensorRq(m, qs)
    y = zq_input(rng, n, qs, batch=B)
    x = torch.from_numpy(y).cuda()
    f = t.crt(x)
    assert np.array_equal(f.cpu().numpy(), np.stack([oracle.tensorCRTRq(y[b], pe, ru, qs) for b in range(B)]))
    assert np.array_equal(t.crtInv(x).cpu().numpy(), np.stack([oracle.tensorCRTInvRq(y[b], pe, rui, mh, qs) for b in range(B)]))
    assert torch.equal(t.crtInv(f), x)
    t.plan.force_generic(True)
    assert torch.equal(t.crt(x), f)


def test_batched_plain_rings_match_oracle(torch_cuda, oracle):
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorComplex, CudaTensorInt, CudaTensorReal
    m, B = 14400, 5
    rng = np.random.default_rng(77)
    pe = T.pe_array(m)
    n = T.totient_pps(T.factor_pps(m))
    ruc, ruci = T.ru_tables_c(m), T.ru_tables_c(m, inverse=True)
    ti, tr, tc = CudaTensorInt(m), CudaTensorReal(m), CudaTensorComplex(m)
    z = rng.integers(-8, 9, size=(B, n, 1)).astype(np.int64)
    d = rng.normal(size=(B, n, 1))
    c = rng.normal(size=(B, n, 1)) + 1j * rng.normal(size=(B, n, 1))
    zx, dx, cx = torch.from_numpy(z).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda()
    per = lambda f, arr: np.stack([f(arr[b]) for b in range(B)])
    assert np.array_equal(ti.l(zx).cpu().numpy(), per(lambda v: oracle.tensorLR(v, pe), z))
    assert np.array_equal(ti.mulGDec(zx).cpu().numpy(), per(lambda v: oracle.tensorGDecR(v, pe), z))
    assert np.array_equal(ti.gSqNormDec(zx).cpu().numpy().reshape(-1), np.array([oracle.tensorNormSqR(z[b], pe).flat[0] for b in range(B)]))
    gz = ti.mulGPow(zx)
    back, ok = ti.divGPow(gz)
    assert ok.cpu().tolist() == [1] * B and torch.equal(back, zx)
    assert rel_err(tr.gSqNormDec(dx).cpu().numpy().reshape(-1), np.array([oracle.tensorNormSqD(d[b], pe).flat[0] for b in range(B)])) <= FLOAT_TOL
    assert rel_err(tr.gaussianDecTransform(dx).cpu().numpy(), per(lambda v: oracle.tensorGaussianDec(v, pe, ruc), d)) <= FLOAT_TOL
    assert rel_err(tc.crt(cx).cpu().numpy(), per(lambda v: oracle.tensorCRTC(v, pe, ruc), c)) <= FLOAT_TOL
    assert rel_err(tc.crtInv(cx).cpu().numpy(), per(lambda v: oracle.tensorCRTInvC(v, pe, ruci, T.mhat_inv_c(m)), c)) <= FLOAT_TOL
    assert rel_err(tc.mulGPow(cx).cpu().numpy(), per(lambda v: oracle.tensorGPowC(v, pe), c)) <= FLOAT_TOL
    # tGaussianDec: right shape, finite, and gSqNormDec of it is positive (RLWE/Discrete.hs:42-59 usage)
    smp = tr.tGaussianDec(0.1, 8)
    assert smp.shape == (8, n, 1) and torch.isfinite(smp).all()
    assert (tr.gSqNormDec(smp) > 0).all()


@pytest.mark.timeout(120)
@pytest.mark.parametrize("mix", [0, 1, 2], ids=["two-streams", "mixed-launches", "graph"])
@pytest.mark.parametrize("mb", [1, 16], ids=lambda v: f"sub-batch {v} MiB")
@pytest.mark.parametrize("e,k", [(14, 1), (14, 2), (14, 4), (15, 1), (15, 2), (15, 4), (16, 1), (16, 2), (16, 4)], ids=lambda v: str(v))
def test_power_of_two_split_schedule(torch_cuda, oracle, monkeypatch, e, k, mb, mix):
    """fused_pow2_split (m = 2^14 .. 2^16, tupSize 1, 2, 4): chunk kernel and column kernel of one sub-batch on two streams behind
    events, ring slots reused three sub-batches later.  1 MiB sub-batches force many ragged sub-batches and every ring slot to
    be reused; oracle parity on a sample, the generic engine on all elements, crtInv . crt = id, and two calls back to back
    (the second call's first kernel must wait for the first call's last one)."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    monkeypatch.setenv("LOLB_SPLIT_MB", str(mb))
    monkeypatch.setenv("LOLB_SPLIT_MIX", "1" if mix == 1 else "0")      # 1: both kinds of work in one grid per sub-batch, two ring slots, one stream
    monkeypatch.setenv("LOLB_SPLIT_GRAPH", "4" if mix == 2 else "0")    # 2: one CUDA graph, ring of 4 sub-batches
    monkeypatch.setenv("LOLB_DF_SCHEDULE", "split")
    monkeypatch.setenv("LOLB_POW2_MID_OFF", "1")
    m, qs = 2 ** e, CONFIG_B[1][:k]
    B = 29 if mb == 1 else 7
    rng = np.random.default_rng(e * 10 + k + 1)
    pe, n, ru, rui, mh = _tables(m, qs)
    t = CudaTensorRq(m, qs)
    y = zq_input(rng, n, qs, batch=B)
    x = torch.from_numpy(y).cuda()
    f, g = t.crt(x), t.crtInv(x)
    for b in (0, 1, B // 2, B - 1):
        assert np.array_equal(f[b].cpu().numpy(), oracle.tensorCRTRq(y[b], pe, ru, qs)), b
        assert np.array_equal(g[b].cpu().numpy(), oracle.tensorCRTInvRq(y[b], pe, rui, mh, qs)), b
    assert torch.equal(t.crtInv(f), x) and torch.equal(t.crt(g), x)
    z = x.clone()
    st = int(torch.cuda.current_stream().cuda_stream)
    from lol_b200 import capi
    for _ in range(3):      # in place, back to back on one stream
        capi.check(t.plan.op("CRT", z.data_ptr(), B, st))
        capi.check(t.plan.op("CRTInv", z.data_ptr(), B, st))
    assert torch.equal(z, x)
    t.plan.force_generic(True)
    assert torch.equal(t.crt(x), f) and torch.equal(t.crtInv(x), g)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("k", [1, 2, 4], ids=lambda v: f"k={v}")
def test_power_of_two_cluster_kernel(torch_cuda, oracle, monkeypatch, k):
    """fused_pow2_cl (m = 2^16, tupSize 1, 2, 4): the element resident across a thread-block cluster, residues exchanged through
    distributed shared memory.  Oracle parity on a sample, the generic engine on all elements (more elements than clusters fit
    on the device at once), crtInv . crt = id, out-of-range input words, calls back to back in place."""
    torch = torch_cuda
    from lol_b200 import capi
    from lol_b200.tensor import CudaTensorRq
    monkeypatch.setenv("LOLB_DF_SCHEDULE", "cluster")
    m, qs = 2 ** 16, CONFIG_B[1][:k]
    B = 150 // k + 3
    rng = np.random.default_rng(160 + k)
    pe, n, ru, rui, mh = _tables(m, qs)
    t = CudaTensorRq(m, qs)
    y = zq_input(rng, n, qs, batch=B)
    y[1, :64] += np.array(qs) * 3          # outside [0, q): reduced like the reference's constructor
    y[2, 5:9] -= np.array(qs) * 2
    x = torch.from_numpy(y).cuda()
    f, g = t.crt(x), t.crtInv(x)
    for b in (0, 1, 2, B // 2, B - 1):
        assert np.array_equal(f[b].cpu().numpy(), oracle.tensorCRTRq(y[b], pe, ru, qs)), b
        assert np.array_equal(g[b].cpu().numpy(), oracle.tensorCRTInvRq(y[b], pe, rui, mh, qs)), b
    xc = torch.from_numpy(y % np.array(qs)).cuda()
    assert torch.equal(t.crtInv(f), xc) and torch.equal(t.crt(g), xc)
    z = xc.clone()
    st = int(torch.cuda.current_stream().cuda_stream)
    for _ in range(3):
        capi.check(t.plan.op("CRT", z.data_ptr(), B, st))
        capi.check(t.plan.op("CRTInv", z.data_ptr(), B, st))
    assert torch.equal(z, xc)
    t.plan.force_generic(True)
    assert torch.equal(t.crt(x), f) and torch.equal(t.crtInv(x), g)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("schedule", ["paired", "unpaired", "resident"])
@pytest.mark.parametrize("ring,lag", [(3, 1), (5, 4), (48, 12)], ids=lambda v: str(v))
@pytest.mark.parametrize("e,k", [(10, 1), (10, 2), (10, 4), (11, 1), (11, 2), (11, 4), (12, 1), (12, 2), (12, 4), (13, 1), (13, 2), (13, 4), (14, 1),
                                 (14, 4), (15, 2), (16, 1), (16, 2), (16, 4)], ids=lambda v: str(v))
def test_power_of_two_dataflow_kernel(torch_cuda, oracle, monkeypatch, e, k, ring, lag, schedule):
    """fused_pow2_df for tupSize 1, 2, 4: the warp-resident kernel of m = 2^10, 2^11 (ragged groups) and the persistent
    kernel with the whole element in shared memory at m = 2^12, 2^13, and the persistent dataflow kernels with the L2
    exchange ring from m = 2^13 on, both schedules: batches
    larger than the ring so slots are reused, tiny rings so the per-element counters and the slot-reuse ordering are
    exercised (this is the test that caught a lane-0-only fence); oracle parity on a sample of elements, the generic
    engine on all of them, and crtInv . crt = id."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    monkeypatch.setenv("LOLB_DF_RING", str(ring))
    monkeypatch.setenv("LOLB_DF_LAG", str(lag))
    if schedule == "resident":        # element-resident kernels where they exist (m <= 2^12, m = 2^13 with tupSize <= 2)
        if e >= 15 or (e == 14 and k > 1) or ring != 48:
            pytest.skip("no element-resident kernel for this shape / ring size does not apply")
    else:
        if e <= 12 and ring != 48:
            pytest.skip("m <= 2^12 has no exchange ring")
        monkeypatch.setenv("LOLB_DF_SCHEDULE", schedule)
        monkeypatch.setenv("LOLB_POW2_MID_OFF", "1")
    m, qs = 2 ** e, CONFIG_B[1][:k]
    B = 23 if ring < 48 else 61
    rng = np.random.default_rng(e * 10 + k)
    pe, n, ru, rui, mh = _tables(m, qs)
    t = CudaTensorRq(m, qs)
    assert t.plan.kernel_name("CRT") == "fused_pow2_df" and t.plan.kernel_name("CRTInv") == "fused_pow2_df"
    y = zq_input(rng, n, qs, batch=B)
    x = torch.from_numpy(y).cuda()
    f, g = t.crt(x), t.crtInv(x)
    for b in (0, 1, B // 2, B - 1):
        assert np.array_equal(f[b].cpu().numpy(), oracle.tensorCRTRq(y[b], pe, ru, qs)), b
        assert np.array_equal(g[b].cpu().numpy(), oracle.tensorCRTInvRq(y[b], pe, rui, mh, qs)), b
    assert torch.equal(t.crtInv(f), x) and torch.equal(t.crt(g), x)
    t.plan.force_generic(True)
    assert torch.equal(t.crt(x), f) and torch.equal(t.crtInv(x), g)


@pytest.mark.parametrize("k", [1, 2], ids=lambda k: f"k={k}")
def test_complex_crt_fused_kernel_m14400(torch_cuda, oracle, k):
    """tensorCRTC / tensorCRTInvC at m = 14400 run on fused_ac (one HBM round trip, FP64): against the oracle per
    element (1e-9 relative, BASELINE.json), against the generic pass engine on the whole ragged batch, and
    crtInv . crt = id; also through the drop-in symbols with caller-supplied root tables."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorComplex
    m, B = 14400, 19
    rng = np.random.default_rng(41 + k)
    pe = T.pe_array(m)
    t = CudaTensorComplex(m, k)
    assert t.plan.kernel_name("CRTC") == "fused_ac" and t.plan.kernel_name("CRTInvC") == "fused_ac"
    c = rng.normal(size=(B, t.n, k)) + 1j * rng.normal(size=(B, t.n, k))
    x = torch.from_numpy(c).cuda()
    ruc, ruci = T.ru_tables_c(m, k), T.ru_tables_c(m, k, inverse=True)
    f, g = t.crt(x), t.crtInv(x)
    for b in (0, 7, B - 1):
        assert rel_err(f[b].cpu().numpy(), oracle.tensorCRTC(c[b], pe, ruc, k)) <= FLOAT_TOL
        assert rel_err(g[b].cpu().numpy(), oracle.tensorCRTInvC(c[b], pe, ruci, T.mhat_inv_c(m, k), k)) <= FLOAT_TOL
    assert rel_err(t.crtInv(f).cpu().numpy(), c) <= 1e-12
    t.plan.force_generic(True)
    assert t.plan.kernel_name("CRTC") == "generic"
    assert rel_err(t.crt(x).cpu().numpy(), f.cpu().numpy()) <= 1e-12
    assert rel_err(t.crtInv(x).cpu().numpy(), g.cpu().numpy()) <= 1e-12


HOMOM_PRF_QS = [18869761, 19393921, 19918081, 25159681, 3144961, 7338241]      # lol-apps Examples/HomomPRFParams.hs:35-45 (the ones below 2^31)


@pytest.mark.parametrize("m,qs", [(2912, HOMOM_PRF_QS[:2]), (5824, HOMOM_PRF_QS[:3]), (5824, HOMOM_PRF_QS[:4]), (11648, HOMOM_PRF_QS[:4]),
                                  (3640, HOMOM_PRF_QS[:6]), (5460, HOMOM_PRF_QS[:5]), (4095, HOMOM_PRF_QS[:4]), (1728, [3457, 1002241, 10369]),
                                  (5184, [10369, 1073089]), (2912, [8737, 14561, 3144961, 23297])], ids=lambda v: str(v))
def test_fused_w_several_limbs_in_one_launch(torch_cuda, gpu_oracle, monkeypatch, m, qs):
    """tupSize 2 ... 6 on the fused_w rings (the HomomPRF example's modulus chains): k_fused_wm runs up to four limbs of the same
    elements in one launch (limbs of different arithmetic classes are promoted to the widest one).  Oracle parity on a sample,
    bit-identical to one launch per limb and to the generic engine on the whole ragged batch, crtInv . crt = id."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    B = 41
    rng = np.random.default_rng(m + len(qs))
    pe, n, ru, rui, mh = _tables(m, qs)
    t = CudaTensorRq(m, qs)
    assert t.plan.kernel_name("CRT") == "fused_w" and t.plan.kernel_name("CRTInv") == "fused_w"
    y = zq_input(rng, n, qs, batch=B)
    y[3, :5] += np.array(qs) * 2                 # outside [0, q): reduced like the reference's constructor
    x = torch.from_numpy(y).cuda()
    f, g = t.crt(x), t.crtInv(x)
    for b in (0, 3, B - 1):
        assert np.array_equal(f[b].cpu().numpy(), gpu_oracle.tensorCRTRq(y[b], pe, ru, qs)), b
        assert np.array_equal(g[b].cpu().numpy(), gpu_oracle.tensorCRTInvRq(y[b], pe, rui, mh, qs)), b
    xc = torch.from_numpy(y % np.array(qs)).cuda()
    assert torch.equal(t.crtInv(f), xc) and torch.equal(t.crt(g), xc)
    monkeypatch.setenv("LOLB_W_MULTI", "0")
    assert torch.equal(t.crt(x), f) and torch.equal(t.crtInv(x), g)
    monkeypatch.delenv("LOLB_W_MULTI")
    t.plan.force_generic(True)
    assert torch.equal(t.crt(x), f) and torch.equal(t.crtInv(x), g)


@pytest.mark.parametrize("m,k", [(1728, 1), (5184, 1), (2912, 2), (728, 1), (3640, 1), (2016, 3), (5824, 1), (11648, 1), (5460, 2), (4095, 1)],
                         ids=lambda v: str(v))
def test_complex_crt_fused_w_rings(torch_cuda, gpu_oracle, m, k):
    """tensorCRTC / tensorCRTInvC of the fused_w indices (every lol / lol-apps benchmark ring that is not 2^e or 14400) run on
    the fused_w schedule over complex doubles: per element against the oracle (1e-9), the whole ragged batch against the
    generic pass engine, crtInv . crt = id."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorComplex
    B = 37
    rng = np.random.default_rng(m + k)
    pe = T.pe_array(m)
    t = CudaTensorComplex(m, k)
    assert t.plan.kernel_name("CRTC") == "fused_w" and t.plan.kernel_name("CRTInvC") == "fused_w"
    c = rng.normal(size=(B, t.n, k)) + 1j * rng.normal(size=(B, t.n, k))
    x = torch.from_numpy(c).cuda()
    ruc, ruci = T.ru_tables_c(m, k), T.ru_tables_c(m, k, inverse=True)
    f, g = t.crt(x), t.crtInv(x)
    for b in (0, 17, B - 1):
        assert rel_err(f[b].cpu().numpy(), gpu_oracle.tensorCRTC(c[b], pe, ruc, k)) <= FLOAT_TOL
        assert rel_err(g[b].cpu().numpy(), gpu_oracle.tensorCRTInvC(c[b], pe, ruci, T.mhat_inv_c(m, k), k)) <= FLOAT_TOL
    assert rel_err(t.crtInv(f).cpu().numpy(), c) <= 1e-11
    t.plan.force_generic(True)
    assert t.plan.kernel_name("CRTC") == "generic"
    assert rel_err(t.crt(x).cpu().numpy(), f.cpu().numpy()) <= 1e-11
    assert rel_err(t.crtInv(x).cpu().numpy(), g.cpu().numpy()) <= 1e-11


@pytest.mark.parametrize("e", [3, 4, 5, 6, 8, 10, 11, 12, 13, 14], ids=lambda v: f"m=2^{v}")
def test_complex_crt_power_of_two(torch_cuda, gpu_oracle, e):
    """tensorCRTC / tensorCRTInvC for m = 2^e on `k_pow2c` (twist-free Cooley-Tukey rounds over complex doubles, two rounds per
    pass, e - 1 odd and even, one and several elements per CTA, prefetched and single-buffer sizes): per element against the
    oracle (crt.cpp:583-598; 1e-9), the ragged batch against the generic pass engine, crtInv . crt = id; guard zones around the batch."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorComplex
    m = 1 << e
    B = 37 if e <= 11 else 5
    rng = np.random.default_rng(e)
    pe = T.pe_array(m)
    t = CudaTensorComplex(m)
    assert t.plan.kernel_name("CRTC") == "fused_pow2c" and t.plan.kernel_name("CRTInvC") == "fused_pow2c"
    c = rng.normal(size=(B, t.n, 1)) + 1j * rng.normal(size=(B, t.n, 1))
    G = 512
    sent = complex(-7.25e300, 3.5e299)
    buf = torch.full((G + B * t.n + G,), sent, dtype=torch.complex128, device="cuda")
    x = buf[G:G + B * t.n].view(B, t.n, 1)
    x.copy_(torch.from_numpy(c))
    ruc, ruci = T.ru_tables_c(m), T.ru_tables_c(m, inverse=True)
    f, g = t.crt(x), t.crtInv(x)
    for b in (0, B // 2, B - 1):
        assert rel_err(f[b].cpu().numpy(), gpu_oracle.tensorCRTC(c[b], pe, ruc)) <= FLOAT_TOL
        assert rel_err(g[b].cpu().numpy(), gpu_oracle.tensorCRTInvC(c[b], pe, ruci, T.mhat_inv_c(m))) <= FLOAT_TOL
    assert rel_err(t.crtInv(f).cpu().numpy(), c) <= 1e-11
    t.crt(x, inplace=True)
    torch.cuda.synchronize()
    guard = torch.tensor(sent, dtype=torch.complex128, device="cuda")
    assert bool((buf[:G] == guard).all()) and bool((buf[G + B * t.n:] == guard).all())
    assert rel_err(x.cpu().numpy(), f.cpu().numpy()) <= 1e-14
    t.plan.force_generic(True)
    assert t.plan.kernel_name("CRTC") == "generic"
    xs = torch.from_numpy(c).cuda()
    assert rel_err(t.crt(xs).cpu().numpy(), f.cpu().numpy()) <= 1e-11
    assert rel_err(t.crtInv(xs).cpu().numpy(), g.cpu().numpy()) <= 1e-11


@pytest.mark.parametrize("m", [9, 25, 7, 21, 45, 14400, 64 * 27, 89, 91, 77, 33, 728, 2912, 3640, 5460, 4095, 11648], ids=str)
def test_plain_rings_streaming_equals_generic_engine(torch_cuda, oracle, m):
    """The streaming kernels of the modulus-free rings (one or two small odd primes) against the generic pass
    engine on the same batch: bit-identical for int64, <= 1e-12 for double / complex (same operation order)."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorComplex, CudaTensorInt, CudaTensorReal
    B = 9
    rng = np.random.default_rng(m)
    ti, tr, tc = CudaTensorInt(m), CudaTensorReal(m), CudaTensorComplex(m)
    n = ti.n
    z = torch.from_numpy(rng.integers(-(2 ** 62), 2 ** 62, size=(B, n, 1)).astype(np.int64)).cuda()
    d = torch.from_numpy(rng.normal(size=(B, n, 1))).cuda()
    c = torch.from_numpy(rng.normal(size=(B, n, 1)) + 1j * rng.normal(size=(B, n, 1))).cuda()

    def both(t, fn):
        t.plan.force_generic(False)
        a = fn()
        t.plan.force_generic(True)
        b = fn()
        t.plan.force_generic(False)
        return a, b

    for meth in ("l", "lInv", "mulGPow", "mulGDec"):
        a, b = both(ti, lambda: getattr(ti, meth)(z))
        assert torch.equal(a, b), meth
        a, b = both(tc, lambda: getattr(tc, meth)(c))
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-12, meth
    for meth in ("divGPow", "divGDec"):
        a, b = both(tc, lambda: getattr(tc, meth)(c))
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-12, meth
    for meth in ("l", "lInv", "gaussianDecTransform"):
        a, b = both(tr, lambda: getattr(tr, meth)(d))
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-12, meth
    a, b = both(ti, lambda: ti.gSqNormDec(z))
    assert torch.equal(a, b)
    a, b = both(tr, lambda: tr.gSqNormDec(d))
    assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 1e-12
    # and against the oracle for one element
    pe = T.pe_array(m)
    assert oracle.tensorNormSqR(z[2].cpu().numpy(), pe).flat[0] == int(ti.gSqNormDec(z)[2, 0])
    assert rel_err(tr.gaussianDecTransform(d)[1].cpu().numpy(), oracle.tensorGaussianDec(d[1].cpu().numpy(), pe, T.ru_tables_c(m))) <= FLOAT_TOL


@pytest.mark.parametrize("m,qs", [CONFIG_A, CONFIG_C, (42, [8191]), (2 ** 16, [786433]), (2 ** 9, [12289]), (14400, [14401, 43201])],
                         ids=lambda v: str(v))
def test_non_canonical_input_is_reduced_like_the_reference(torch_cuda, oracle, m, qs):
    """Outside the Haskell contract (coefficients in [0,q)), the reference reduces with `c % q` (types.h:62-66) and
    canonicalises at exit (zq.cpp:57-67).  Every kernel family (fused, streaming, generic) must do the same for
    negative, >= q and > 2^32 inputs instead of silently computing on garbage."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    B = 4
    rng = np.random.default_rng(m + 1)
    pe, n, ru, rui, mh = _tables(m, qs)
    y = zq_input(rng, n, qs, batch=B)
    y[0, 0, :] = [q for q in qs]                        # == q
    y[1, n // 2, :] = [-3 for _ in qs]                  # negative
    y[2, n - 1, :] = [q + 5 + (1 << 40) * q for q in qs]   # far above 2^32
    y[3, :, 0] += qs[0]                                 # a whole limb shifted by q
    canon = np.stack([np.mod(y[..., t], q) for t, q in enumerate(qs)], axis=-1)
    t = CudaTensorRq(m, qs)
    x = torch.from_numpy(y).cuda()
    per = lambda f: np.stack([f(canon[b]) for b in range(B)])
    assert np.array_equal(t.crt(x).cpu().numpy(), per(lambda v: oracle.tensorCRTRq(v, pe, ru, qs)))
    assert np.array_equal(t.crtInv(x).cpu().numpy(), per(lambda v: oracle.tensorCRTInvRq(v, pe, rui, mh, qs)))
    if T.odd_radical(m) > 1:
        assert np.array_equal(t.mulGPow(x).cpu().numpy(), per(lambda v: oracle.tensorGPowRq(v, pe, qs)))
        assert np.array_equal(t.lInv(x).cpu().numpy(), per(lambda v: oracle.tensorLInvRq(v, pe, qs)))
    else:
        # power-of-two index: L and G are the identity (l.cpp:35, g.cpp:18) and nothing is launched, so the input
        # comes back untouched (the reference would only add q to negatives, zq.cpp:57-67); documented in DESIGN.md
        assert torch.equal(t.mulGPow(x), x) and torch.equal(t.lInv(x), x)
    assert np.array_equal(t.mul(x, x).cpu().numpy(), per(lambda v: oracle.mulRq(v, v, qs)))
    # the oracle itself agrees: feeding it the raw values gives the same answer as the canonical ones
    assert np.array_equal(oracle.tensorCRTRq(y[1], pe, ru, qs), oracle.tensorCRTRq(canon[1], pe, ru, qs))


def test_host_pipeline_matches_device_path(torch_cuda, oracle):
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    m, qs = CONFIG_A
    t = CudaTensorRq(m, qs)
    B = 300
    rng = np.random.default_rng(5)
    pe, n, ru, rui, mh = _tables(m, qs)
    y = zq_input(rng, n, qs, batch=B)
    h = torch.from_numpy(y.copy()).pin_memory()
    t.apply_host("CRT", h)
    assert np.array_equal(h[3].numpy(), oracle.tensorCRTRq(y[3], pe, ru, qs))
    assert torch.equal(h.cuda(), t.crt(torch.from_numpy(y).cuda()))
    t.apply_host("CRTInv", h)
    assert np.array_equal(h.numpy(), y)
    h2 = torch.from_numpy(y.copy())          # pageable memory works too
    t.apply_host("CRT,CRTInv", h2)
    assert np.array_equal(h2.numpy(), y)


def test_host_pipeline_private_workspaces_pow2(torch_cuda, oracle):
    """The host pipeline runs three stream slots concurrently; kernels that need a workspace (the exchange ring and
    counters of fused_pow2_df) get one per slot.  Config B shape, three chunks of the 96 MiB staging size."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    m, qs = CONFIG_B
    t = CudaTensorRq(m, qs)
    B = 200
    g = torch.Generator()
    g.manual_seed(11)
    y = torch.cat([torch.randint(0, q, (B, t.n, 1), dtype=torch.int64, generator=g) for q in qs], dim=2).contiguous()
    h = y.clone().pin_memory()
    t.apply_host("CRT", h)
    dev = t.crt(y.cuda())
    assert torch.equal(h.cuda(), dev)
    pe, n, ru, rui, mh = _tables(m, qs)
    for b in (0, 97, B - 1):
        assert np.array_equal(h[b].numpy(), oracle.tensorCRTRq(y[b].numpy(), pe, ru, qs))
    t.apply_host("CRTInv", h)
    assert torch.equal(h, y)
    t.apply_host("CRT,CRTInv", h)
    assert torch.equal(h, y)


# ------------------------------------------------------------------ full-size properties (BASELINE.json configs)
def _device_uniform(torch, B, n, qs, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    cols = [torch.randint(0, q, (B, n, 1), dtype=torch.int64, device="cuda", generator=g) for q in qs]
    return torch.cat(cols, dim=2).contiguous()


# the reference's other benchmark rings at the batch sizes bench.py times them at (lol Benchmarks/Default.hs:41-48; lol-apps
# Benchmarks/Default.hs:52-82 tunnel rings at q = 3144961; Examples/HomomPRFParams.hs modulus chain ZQ4 on H1')
FULL_SIZE = [(CONFIG_A, 65536), (CONFIG_C, 8192), (CONFIG_B, 256), ((1728, [3457]), 131072), ((5184, [10369]), 40960), ((2912, [8737]), 61440),
             ((3640, [14561]), 61440), ((11648, [3144961]), 15360), ((5460, [3144961]), 61440), ((4095, [3144961]), 40960),
             ((5824, [25159681, 19918081, 19393921, 18869761]), 8192)]


@pytest.mark.parametrize("cfg,B", FULL_SIZE, ids=["A", "C", "B"] + [f"m={c[0]}/k={len(c[1])}" for c, _ in FULL_SIZE[3:]])
def test_full_size_properties(torch_cuda, oracle, cfg, B):
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    m, qs = cfg
    t = CudaTensorRq(m, qs)
    n, k = t.n, t.k
    x = _device_uniform(torch, B, n, qs, seed=0)
    q = torch.tensor(qs, dtype=torch.int64, device="cuda")
    f = t.crt(x)
    assert int(f.min()) >= 0 and bool((f < q).all())                       # canonical output (zq.cpp:57-67)
    assert torch.equal(t.crtInv(f), x)                                      # crtInv . crt = id
    # linearity: crt(x + x') = crt(x) + crt(x') mod q
    x2 = _device_uniform(torch, B, n, qs, seed=1)
    assert torch.equal(t.crt((x + x2) % q), (f + t.crt(x2)) % q)
    # mulGCRT = crt . mulGPow . crtInv ; divGCRT . mulGCRT = id ; lInv . l = id ; mulGDec = lInv . mulGPow . l
    assert torch.equal(t.mulGCRT(f), t.crt(t.mulGPow(x)))
    assert torch.equal(t.divGCRT(t.mulGCRT(f)), f)
    assert torch.equal(t.lInv(t.l(x)), x)
    assert torch.equal(t.mulGDec(x), t.lInv(t.mulGPow(t.l(x))))
    assert torch.equal(t.divGPow(t.mulGPow(x)), x) and torch.equal(t.divGDec(t.mulGDec(x)), x)
    # scalarCRT = crt . scalarPow
    assert torch.equal(t.crt(t.scalarPow(5, batch=3)), t.scalarCRT(5, batch=3))
    # spot parity against the oracle: first, middle, last element of the batch
    pe, _, ru, rui, mh = _tables(m, qs)
    for b in (0, B // 2, B - 1):
        assert np.array_equal(f[b].cpu().numpy(), oracle.tensorCRTRq(x[b].cpu().numpy(), pe, ru, qs))
    # ring product through the CRT basis: commutes and matches the oracle composition on one element
    prod = t.crtInv(t.mul(f, t.crt(x2)))
    assert torch.equal(prod, t.crtInv(t.mul(t.crt(x2), f)))
    b = B // 3
    ref = oracle.tensorCRTInvRq(oracle.mulRq(oracle.tensorCRTRq(x[b].cpu().numpy(), pe, ru, qs),
                                             oracle.tensorCRTRq(x2[b].cpu().numpy(), pe, ru, qs), qs), pe, rui, mh, qs)
    assert np.array_equal(prod[b].cpu().numpy(), ref)


# ------------------------------------------------------------------ SymmSHE ciphertext multiply + key switch (configs[3])
def _hint(rng, ell, n, qs):
    return np.stack([np.stack([zq_input(rng, n, qs) for _ in range(2)]) for _ in range(ell)])


@pytest.mark.parametrize("m,qs,base", [(14400, [1008001, 1065601], 0), (14400, [1008001, 1065601], 1000), (14400, [1008001, 1065601], 2),
                                       (14400, [14401, 429336001], 1024), (14400, [14401, 14401], 0), (14400, [14401, 429336001], 0), (14400, [1065601, 1008001], 0), (21, [43, 127, 379], 0),
                                       (21, [43, 127, 379], 2), (2, [13, 17, 19], 3), (45, [2148249601], 65536), (16, [97], 0)],
                         ids=lambda v: str(v))
def test_symmshe_steps_match_oracle(torch_cuda, oracle, m, qs, base):
    """lolb_ctMulRq / lolb_decomposeRq / lolb_knapsackRq and their composition (SymmSHE.hs:302-314, 359-372, 443-449)
    against the numpy restatement, step by step and end to end, bit-exact."""
    torch = torch_cuda
    from lol_b200.symmshe import CudaSymmSHE
    from oracle import symmshe as S
    B = 3 if m == 14400 else 11
    rng = np.random.default_rng(m * 7 + base)
    pe, n, ru, rui, mh = _tables(m, qs)
    g, _ = T.g_crt_vectors(m, qs)
    she = CudaSymmSHE(m, qs, gad_base=base)
    assert she.ell == S.gadget_length(qs, base) and she.gadget() == S.gadget(qs, base)
    cts = [zq_input(rng, n, qs, batch=B) for _ in range(4)]
    dev = [torch.from_numpy(c).cuda() for c in cts]
    # ciphertext product on CRT-basis components
    d = she.mulCT(dev[:2], dev[2:], basis="crt")
    ref = S.ct_mul_crt(cts[:2], cts[2:], g, qs)
    for got, want in zip(d, ref):
        assert np.array_equal(got.cpu().numpy(), want)
    assert all(torch.equal(x, torch.from_numpy(c).cuda()) for x, c in zip(dev, cts))      # pure
    # gadget digits of a Pow-basis element
    digits = she.decompose(dev[0])
    assert np.array_equal(digits.cpu().numpy(), S.decompose_reduced(cts[0], qs, base))
    # digits taken to the CRT basis in one call (decomposition inside the CRT kernel's load stage at m = 14400, tupSize 2, TrivGad)
    dc = she.decomposeCRT(dev[0])
    assert torch.equal(dc, she.t.crt(digits.view(she.ell * B, n, len(qs))).view(she.ell, B, n, len(qs)))
    she.t.plan.force_generic(True)
    assert torch.equal(she.decomposeCRT(dev[0]), dc)
    she.t.plan.force_generic(False)
    # knapsack
    hint = _hint(rng, she.ell, n, qs)
    hint_d = torch.from_numpy(hint).cuda()
    o = she.knapsack(hint_d, digits, dev[1], dev[2])
    want = S.knapsack(hint, S.decompose_reduced(cts[0], qs, base), cts[1], cts[2], qs)
    assert np.array_equal(o[0].cpu().numpy(), want[0]) and np.array_equal(o[1].cpu().numpy(), want[1])
    # the whole sequence from Pow-basis ciphertexts, one oracle composition per ciphertext pair
    out = she.mulAndSwitch(dev[:2], dev[2:], hint_d, basis="pow")
    for b in range(B):
        w = S.mul_and_switch(oracle, [cts[0][b], cts[1][b]], [cts[2][b], cts[3][b]], hint, (pe, ru, rui, mh, g), qs, base)
        assert np.array_equal(out[0][b].cpu().numpy(), w[0]) and np.array_equal(out[1][b].cpu().numpy(), w[1])
    assert all(torch.equal(x, torch.from_numpy(c).cuda()) for x, c in zip(dev, cts))
    # in-place variant gives the same result
    work = [x.clone() for x in dev]
    out2 = she.mulAndSwitch(work[:2], work[2:], hint_d, basis="pow", inplace=True)
    assert torch.equal(out2[0], out[0]) and torch.equal(out2[1], out[1])
    # empty batch and argument errors
    e = torch.empty(0, n, len(qs), dtype=torch.int64, device="cuda")
    assert she.decompose(e).shape == (she.ell, 0, n, len(qs))
    from lol_b200 import capi
    with pytest.raises(capi.LolB200Error):
        she.mulCT(dev[:2], dev[2:3])
    with pytest.raises(capi.LolB200Error):
        CudaSymmSHE(m, qs, gad_base=1)


@pytest.mark.parametrize("base", [0, 256], ids=["TrivGad", "BaseBGad256"])
def test_symmshe_full_size_identity_hint(torch_cuda, base):
    """Config C at bench size: with the hint polynomials (gadget_i, 0) the key switch returns (c0 + c2, c1), because
    sum_i gadget_i * decompose(c2)_i = c2 (Gadget.hs:60-66) and CRT . CRT^-1 = id; the ciphertext product commutes."""
    torch = torch_cuda
    from lol_b200.symmshe import CudaSymmSHE
    m, qs = CONFIG_C
    B = 2048
    she = CudaSymmSHE(m, qs, gad_base=base)
    n, k = she.n, she.k
    q = torch.tensor(qs, dtype=torch.int64, device="cuda")
    c = [_device_uniform(torch, B, n, qs, seed=s) for s in range(4)]
    d = she.mulCT(c[:2], c[2:], basis="pow")
    d_sw = she.mulCT(c[2:], c[:2], basis="pow")
    assert all(torch.equal(u, v) for u, v in zip(d, d_sw))
    assert all(int(u.min()) >= 0 and bool((u < q).all()) for u in d)
    hint = torch.zeros(she.ell, 2, n, k, dtype=torch.int64, device="cuda")
    for i, gi in enumerate(she.gadget()):
        hint[i, 0] = she.t.scalarCRT(gi, batch=1)[0]
    o = she.keySwitchQuadCirc(hint, d)
    assert torch.equal(o[0], (d[0] + d[2]) % q) and torch.equal(o[1], d[1])


def test_device_memory_entry_points_without_torch(torch_cuda, oracle):
    """lolb_dev_alloc / upload / copy / download (what the Haskell GT type is built on): a ring element goes to the device, is
    transformed there twice without crossing PCIe, comes back, and equals the oracle's crtInv(mulG(crt x)) composition."""
    from lol_b200 import capi
    m, qs = 14400, [14401]
    rng = np.random.default_rng(8)
    pe, n, ru, rui, mh = _tables(m, qs)
    plan = capi.PlanRq(T.factor_pps(m), qs)
    y = zq_input(rng, n, qs)
    a, b = capi.dev_alloc(y.nbytes), capi.dev_alloc(y.nbytes)
    try:
        capi.dev_upload(a, y)
        capi.dev_copy(b, a, y.nbytes)
        capi.check(plan.op("CRT", b, 1, 0))
        capi.check(plan.op("CRTInv", b, 1, 0))
        capi.check(plan.op("CRT", a, 1, 0))
        back, f = np.empty_like(y), np.empty_like(y)
        capi.dev_download(back, b)
        capi.dev_download(f, a)
    finally:
        capi.dev_free(a)
        capi.dev_free(b)
    assert np.array_equal(back, y)
    assert np.array_equal(f, oracle.tensorCRTRq(y, pe, ru, qs))


# ------------------------------------------------------------------ guard zones (compute-sanitizer is not available on the pool)
GUARDED = [(14400, [14401]), (14400, [1008001, 1065601]), (14400, [14401, 1008001, 1065601, 429336001]), (1728, [3457]), (5184, [10369]),
           (2912, [8737]), (2912, [3144961]), (3640, [14561]), (11648, [3144961]), (5460, [3144961]), (4095, [3144961]), (2016, [2017]), (728, [8737]),
           (5824, HOMOM_PRF_QS[:4]), (2912, HOMOM_PRF_QS[:6]), (1728, [3457, 1002241, 10369]), (2 ** 16, CONFIG_B[1]), (2 ** 15, CONFIG_B[1][:2]),
           (2 ** 11, [12289]), (2 ** 13, CONFIG_B[1]), (21, [43, 127]), (45, [2148249601])]


@pytest.mark.parametrize("m,qs", GUARDED, ids=lambda v: str(v))
@pytest.mark.parametrize("B", [1, 37], ids=lambda v: f"B={v}")
def test_kernels_write_inside_their_batch_only(torch_cuda, m, qs, B):
    """Out-of-bounds WRITES of the in-place kernels, found with sentinel-filled guard zones on both sides of a ragged batch
    (the pool has no compute-sanitizer): CRT, CRT^-1, L, L^-1, the g operators and mulRq must leave the guards untouched."""
    torch = torch_cuda
    from lol_b200 import capi
    from lol_b200.tensor import CudaTensorRq
    t = CudaTensorRq(m, qs)
    n, k = t.n, t.k
    G = 4096                                                 # guard words on each side
    sentinel = -0x5A5A5A5A5A5A5A5B
    buf = torch.full((G + B * n * k + G,), sentinel, dtype=torch.int64, device="cuda")
    body = buf[G:G + B * n * k].view(B, n, k)
    g = torch.Generator(device="cuda")
    g.manual_seed(m)
    for i, q in enumerate(qs):
        body[:, :, i] = torch.randint(0, q, (B, n), dtype=torch.int64, device="cuda", generator=g)
    other = body.clone()
    st = int(torch.cuda.current_stream().cuda_stream)
    ptr = buf.data_ptr() + 8 * G
    for name in ("CRT", "CRTInv", "L", "LInv", "GPow", "GDec", "GInvPow", "GInvDec"):
        capi.check(t.plan.op(name, ptr, B, st))
    capi.check(t.plan.mul(ptr, other.data_ptr(), B, B, st))
    torch.cuda.synchronize()
    assert bool((buf[:G] == sentinel).all()) and bool((buf[G + B * n * k:] == sentinel).all())
    assert int(body.min()) >= 0


@pytest.mark.parametrize("m,k", [(14400, 1), (1728, 1), (5184, 1), (2912, 2), (11648, 1), (5460, 1), (4095, 3), (21, 1)], ids=lambda v: str(v))
def test_complex_kernels_write_inside_their_batch_only(torch_cuda, m, k):
    torch = torch_cuda
    from lol_b200 import capi
    from lol_b200.tensor import CudaTensorComplex
    t = CudaTensorComplex(m, k)
    n, B, G = t.n, 19, 2048
    buf = torch.full((G + B * n * k + G,), complex(-7.25e300, 3.5e299), dtype=torch.complex128, device="cuda")
    body = buf[G:G + B * n * k].view(B, n, k)
    body.copy_(torch.randn(B, n, k, dtype=torch.complex128, device="cuda"))
    st = int(torch.cuda.current_stream().cuda_stream)
    ptr = buf.data_ptr() + 16 * G
    for name in ("CRTC", "CRTInvC"):
        capi.check(t.plan.op(name, ptr, B, st))
    torch.cuda.synchronize()
    guard = torch.tensor(complex(-7.25e300, 3.5e299), dtype=torch.complex128, device="cuda")
    assert bool((buf[:G] == guard).all()) and bool((buf[G + B * n * k:] == guard).all())
    assert bool(torch.isfinite(torch.view_as_real(body)).all()) and float(torch.view_as_real(body).abs().max()) < 1e6


@pytest.mark.parametrize("m,k", [(2912, 1), (5460, 2), (91, 3), (11648, 1), (4095, 1)], ids=lambda v: str(v))
@pytest.mark.parametrize("B", [1, 37], ids=lambda v: f"B={v}")
def test_plain_tile_kernels_write_inside_their_batch_only(torch_cuda, m, k, B):
    """Guard zones around a ragged batch for `k_plain_tile` (cp.async prefetch of the NEXT group, two tile buffers): the line
    operators over int64 / double / complex, the Gaussian transform, the one-pass tGaussianDec and the norm (which must not
    write its input at all)."""
    torch = torch_cuda
    from lol_b200 import capi
    from lol_b200.tensor import CudaTensorComplex, CudaTensorInt, CudaTensorReal
    st = int(torch.cuda.current_stream().cuda_stream)
    G = 2048
    for cls, dtype, sent, names in ((CudaTensorInt, torch.int64, -0x5A5A5A5A5A5A5A5B, ("LR", "LInvR", "GPowR", "GDecR")),
                                    (CudaTensorReal, torch.float64, -7.25e300, ("LDouble", "LInvDouble")),
                                    (CudaTensorComplex, torch.complex128, complex(-7.25e300, 3.5e299), ("LC", "GPowC", "GInvDecC"))):
        t = cls(m, k)
        n = t.n
        buf = torch.full((G + B * n * k + G,), sent, dtype=dtype, device="cuda")
        body = buf[G:G + B * n * k].view(B, n, k)
        body.copy_(torch.randint(-9, 10, (B, n, k), device="cuda").to(dtype))
        ptr = buf.data_ptr() + buf.element_size() * G
        for name in names:
            capi.check(t.plan.op(name, ptr, B, st))
        if cls is CudaTensorReal and k == 1:
            capi.check(t.plan.op("GaussianDec", ptr, B, st))
            before = body.clone()
            out = torch.empty(B, 1, dtype=dtype, device="cuda")
            capi.check(t.plan.normsq("D", ptr, out.data_ptr(), B, st))
            assert torch.equal(body, before) and bool(torch.isfinite(out).all())
            capi.check(t.plan.t_gaussian_dec(0.3, 7, 0, ptr, B, st))
        torch.cuda.synchronize()
        guard = torch.tensor(sent, dtype=dtype, device="cuda")
        assert bool((buf[:G] == guard).all()) and bool((buf[G + B * n * k:] == guard).all()), cls.__name__
        flat = torch.view_as_real(body) if dtype == torch.complex128 else body
        assert bool((flat != (sent.real if isinstance(sent, complex) else sent)).all())


# moduli on both sides of the thresholds of the line kernels' 32-bit mode: (P + 2) q < 2^31 for L, L^-1, *g and P^2 q < 2^31 for /g
LINE_MODE_EDGES = [(2912, 143165569), (2912, 143171393), (2912, 12655553), (2912, 12719617), (14400, 306748801), (14400, 306864001),
                   (14400, 85852801), (14400, 85924801), (1728, 429496129), (1728, 429501313), (1728, 238600513), (1728, 238610881),
                   (2912, 55033889), (2912, 55080481), (5460, 19393921)]      # /g with running sums kept reduced: 78 q < 2^32


@pytest.mark.parametrize("m,q", LINE_MODE_EDGES, ids=lambda v: str(v))
def test_line_operators_at_the_arithmetic_mode_thresholds(torch_cuda, gpu_oracle, m, q):
    """L, L^-1, *g, /g with every coefficient at q - 1 (the largest intermediates), at 0 / q - 1 alternating (the most negative
    ones) and random, for moduli just below and just above the bounds that select int32 or int64 intermediates."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    t = CudaTensorRq(m, [q])
    n, pe = t.n, T.pe_array(m)
    rng = np.random.default_rng(q % 1000)
    y = np.zeros((4, n, 1), dtype=np.int64)
    y[0] = q - 1
    y[1, ::2] = q - 1
    y[2, 1::2] = q - 1
    y[3] = rng.integers(0, q, size=(n, 1))
    x = torch.from_numpy(y).cuda()
    for meth, nm in (("l", "tensorLRq"), ("lInv", "tensorLInvRq"), ("mulGPow", "tensorGPowRq"), ("mulGDec", "tensorGDecRq")):
        got = getattr(t, meth)(x).cpu().numpy()
        for b in range(4):
            assert np.array_equal(got[b], getattr(gpu_oracle, nm)(y[b], pe, [q])), (nm, b)
    for meth, nm in (("divGPow", "tensorGInvPowRq"), ("divGDec", "tensorGInvDecRq")):
        got = getattr(t, meth)(x).cpu().numpy()
        for b in range(4):
            assert np.array_equal(got[b], getattr(gpu_oracle, nm)(y[b], pe, [q])[0]), (nm, b)


@pytest.mark.parametrize("m,k", [(91, 1), (91, 3), (2912, 2), (5460, 1), (385, 2)], ids=lambda v: str(v))
def test_plain_tile_kernel_tuples_and_oracle(torch_cuda, gpu_oracle, m, k):
    """`k_plain_tile` (the modulus-free rings' operators with the element staged in shared memory) on the {7,13}, {5,7,13},
    {3,5,7,13}, {5,7,11} rings with tupSize folded into the strides, more elements than one CTA's group and a ragged last group:
    int64 bit for bit against the oracle (l.cpp:28-98, g.cpp:16-58 through tensor.h:39-74), complex within FLOAT_TOL."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorComplex, CudaTensorInt
    ti, tc = CudaTensorInt(m, k), CudaTensorComplex(m, k)
    assert ti.plan.kernel_name("LR") == "plain_tile"
    n, pe, B = ti.n, T.pe_array(m), 61
    rng = np.random.default_rng(m + k)
    z = rng.integers(-(2 ** 62), 2 ** 62, size=(B, n, k)).astype(np.int64)
    c = rng.normal(size=(B, n, k)) + 1j * rng.normal(size=(B, n, k))
    zx, cx = torch.from_numpy(z).cuda(), torch.from_numpy(c).cuda()
    for meth, nm in (("l", "tensorLR"), ("lInv", "tensorLInvR"), ("mulGPow", "tensorGPowR"), ("mulGDec", "tensorGDecR")):
        got = getattr(ti, meth)(zx).cpu().numpy()
        for b in (0, 1, B // 2, B - 1):
            assert np.array_equal(got[b], getattr(gpu_oracle, nm)(z[b], pe, k)), (nm, b)
    for meth, nm in (("l", "tensorLC"), ("lInv", "tensorLInvC"), ("mulGPow", "tensorGPowC"), ("mulGDec", "tensorGDecC")):
        got = getattr(tc, meth)(cx).cpu().numpy()
        for b in (0, B - 1):
            assert rel_err(got[b], getattr(gpu_oracle, nm)(c[b], pe, k)) <= FLOAT_TOL, (nm, b)
    # round trips on the whole batch
    assert torch.equal(ti.lInv(ti.l(zx)), zx)
    small = torch.from_numpy(rng.integers(-8, 9, size=(B, n, k)).astype(np.int64)).cuda()      # no wrap-around: g | g x exactly
    back, ok = ti.divGPow(ti.mulGPow(small))
    assert ok.cpu().tolist() == [1] * B and torch.equal(back, small)


@pytest.mark.parametrize("m", [91, 2912, 5460, 4095, 11648], ids=str)
def test_tgaussiandec_one_pass_equals_draw_then_transform(torch_cuda, gpu_oracle, m):
    """On the rings served by `k_plain_tile` the one-pass tGaussianDec (inputs drawn into the tile) uses the pair layout of
    realGaussians, so it equals realGaussians followed by tensorGaussianDec bit for bit; the transform itself against the oracle
    (random.cpp:19-64)."""
    torch = torch_cuda
    from lol_b200.factored import radical_fact
    from lol_b200.tensor import CudaTensorReal
    t = CudaTensorReal(m)
    assert t.plan.kernel_name("GaussianDec") == "plain_tile"
    # a batch that starts at an odd word (8-byte aligned only) takes the generic engine instead of the 16-byte accesses
    odd = torch.randn(3 * t.n + 1, dtype=torch.float64, device="cuda")
    view = odd[1:].view(3, t.n, 1)
    assert view.data_ptr() % 16 == 8
    assert rel_err(t.gaussianDecTransform(view.clone()).cpu().numpy(), t.gaussianDecTransform(view, inplace=True).cpu().numpy()) <= 1e-12
    v, B = 0.37, 23
    one = t.tGaussianDec(v, B, seed=9, first=5)
    raw = t.realGaussians(v * (m // radical_fact(m)), B, seed=9, first=5)
    two = t.gaussianDecTransform(raw)
    assert torch.isfinite(one).all() and torch.equal(one, two)
    pe, ruc = T.pe_array(m), T.ru_tables_c(m)
    for b in (0, B - 1):
        assert rel_err(two[b].cpu().numpy(), gpu_oracle.tensorGaussianDec(raw[b].cpu().numpy(), pe, ruc)) <= FLOAT_TOL


@pytest.mark.parametrize("m,qs", [(2912, [143165569, 143171393]), (2912, [12655553, 3144961, 12719617]), (5460, [3144961, 21841]), (91, [547, 911, 1093, 2003, 2549]),
                                  (5824, [25159681, 19918081, 19393921, 18869761]), (2912, [55033889, 8737])],
                         ids=lambda v: str(v))
def test_line_tile_several_limbs(torch_cuda, gpu_oracle, m, qs):
    """`k_line_tile` with tupSize folded into the strides: limbs on both sides of the int32 / int64 thresholds in one element, an odd
    tupSize, more elements than one CTA's group, against the oracle (l.cpp:28-98, g.cpp:16-123 through tensor.h:39-74)."""
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    t = CudaTensorRq(m, qs)
    wide = any(15 * q >= 2 ** 31 for q in qs)      # a limb beyond the 32-bit mode sends the whole plan to the register-tile launches
    assert t.plan.kernel_name("L") == ("line_stream" if wide else "line_tile")
    n, pe, B = t.n, T.pe_array(m), 45
    rng = np.random.default_rng(m)
    y = zq_input(rng, n, qs, batch=B)
    y[0] = np.array(qs) - 1
    y[1, ::2] = np.array(qs) - 1
    x = torch.from_numpy(y).cuda()
    for meth, nm in (("l", "tensorLRq"), ("lInv", "tensorLInvRq"), ("mulGPow", "tensorGPowRq"), ("mulGDec", "tensorGDecRq")):
        got = getattr(t, meth)(x).cpu().numpy()
        for b in (0, 1, 2, 17, B - 1):
            assert np.array_equal(got[b], getattr(gpu_oracle, nm)(y[b], pe, qs)), (nm, b)
    for meth, nm in (("divGPow", "tensorGInvPowRq"), ("divGDec", "tensorGInvDecRq")):
        got = getattr(t, meth)(x).cpu().numpy()
        for b in (0, 1, 2, 17, B - 1):
            assert np.array_equal(got[b], getattr(gpu_oracle, nm)(y[b], pe, qs)[0]), (nm, b)
    # outside the Haskell contract: non-canonical residues are reduced like `c % q` on the way in (every kernel family agrees)
    z = y.copy()
    z[2, :7] = z[2, :7] + np.array(qs) * 5
    z[3, 5:9] = z[3, 5:9] - np.array(qs) * 3
    xz = torch.from_numpy(z).cuda()
    t.plan.force_generic(True)
    ref = t.divGDec(t.l(x))
    t.plan.force_generic(False)
    assert torch.equal(t.divGDec(t.l(x)), ref)
    assert torch.equal(t.mulGPow(xz), t.mulGPow(x))
