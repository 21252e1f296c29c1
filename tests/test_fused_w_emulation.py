"""fused_w (lol_b200/csrc/fused_w_impl.cuh) without a GPU: the library's host-built constants, the kernel's own line code compiled
for the host and a lane-by-lane replica of its exchange network, run on one ring element through
`lolb_fused_w_emulate`, must reproduce the oracle's tensorCRTRq / tensorCRTInvRq bit for bit.

Rings: the reference's benchmark parameters without a fused kernel in round 1 (lol/Crypto/Lol/Benchmarks/Default.hs:41-48:
F64*F27 / 3457, F64*F81 / 10369, the Twace-Embed rings F32*F7*F13 / 8737, F8*F7*F13, F8*F5*F7*F13 / 14561) plus
m = 2016, each also with a modulus of the 64-bit-accumulate (Montgomery) class and an RNS pair.
"""
import numpy as np
import pytest

from conftest import rel_err, zq_input
from lol_b200 import build_library, capi
from oracle import tables as T


def _prime_1_mod(m, above, count=1):
    out, q = [], (above // m) * m + 1
    while len(out) < count:
        q += m
        if all(q % d for d in range(2, int(q ** 0.5) + 1)):
            out.append(q)
    return out


FUSED_W_PARAMS = [
    (1728, [3457]), (5184, [10369]), (2912, [8737]), (728, [8737]), (3640, [14561]), (2016, [2017]),
    (1728, _prime_1_mod(1728, 10 ** 6)), (5184, _prime_1_mod(5184, 10 ** 7)), (2912, _prime_1_mod(2912, 10 ** 5)),
    (3640, _prime_1_mod(3640, 10 ** 6, 2)), (728, _prime_1_mod(728, 3 * 10 ** 7)), (2016, _prime_1_mod(2016, 10 ** 8)),
    (1728, [3457, 1002241]),      # mixed arithmetic classes, tupSize 2
    (5824, [3144961]), (2912, [3144961]), (3640, [3144961]),      # lol-apps tunnel benchmark rings and modulus (Benchmarks/Default.hs:52-82)
    (11648, [23297]), (11648, [3144961]), (11648, [174721]),      # F128*F7*F13: two column halves per lane (a = 7)
    (5460, [3144961]), (4095, [3144961]), (5460, [21841]), (4095, [8191]), (448, [3144961]), (448, [449]),      # tunnel rings H4, H5: four odd prime powers, one in the tile
]


@pytest.fixture(scope="module", autouse=True)
def _lib():
    build_library()


@pytest.mark.parametrize("m,qs", FUSED_W_PARAMS, ids=lambda v: str(v))
def test_fused_w_schedule_matches_oracle(oracle, m, qs):
    rng = np.random.default_rng(m + len(qs))
    pps, pe = T.factor_pps(m), T.pe_array(m)
    n = T.totient_pps(pps)
    ru, rui = T.ru_tables_zq(m, qs), T.ru_tables_zq(m, qs, inverse=True)
    mh = [T.mhat_inv(m, q) for q in qs]
    for trial in range(2):
        y = zq_input(rng, n, qs)
        if trial == 1:      # boundary residues
            y[: n // 2] = np.array(qs) - 1
            y[n // 2:] = 0
            y[1] = 1
        f = capi.fused_w_emulate(pps, qs, y, False)
        assert np.array_equal(f, oracle.tensorCRTRq(y, pe, ru, qs))
        assert np.array_equal(capi.fused_w_emulate(pps, qs, y, True), oracle.tensorCRTInvRq(y, pe, rui, mh, qs))
        assert np.array_equal(capi.fused_w_emulate(pps, qs, f, True), y)      # crtInv . crt = id (TensorTests.hs:115-119)


def test_fused_w_emulation_refuses_other_shapes():
    with pytest.raises(capi.LolB200Error):
        capi.fused_w_emulate(T.factor_pps(14400), [14401], np.zeros((3840, 1), dtype=np.int64))
    with pytest.raises(capi.LolB200Error):      # no CRT of index 1728 over Z_17
        capi.fused_w_emulate(T.factor_pps(1728), [17], np.zeros((576, 1), dtype=np.int64))


FUSED_W_COMPLEX = [1728, 5184, 2912, 728, 3640, 2016, 5824, 11648, 5460, 4095, 448]


@pytest.mark.parametrize("m", FUSED_W_COMPLEX)
@pytest.mark.parametrize("k", [1, 2])
def test_fused_w_schedule_complex_matches_oracle(oracle, m, k):
    """The same schedule over complex doubles (tensorCRTC / tensorCRTInvC, crt.cpp:583-598): 1e-9 relative, the
    floating-point tolerance of BASELINE.json's north_star (operation order and FMA contraction differ from the reference)."""
    rng = np.random.default_rng(m + k)
    pps, pe = T.factor_pps(m), T.pe_array(m)
    n = T.totient_pps(pps)
    ruc, ruci = T.ru_tables_c(m, k), T.ru_tables_c(m, k, inverse=True)
    c = rng.standard_normal((n, k)) + 1j * rng.standard_normal((n, k))
    f = capi.fused_w_emulate_c(pps, c, False)
    assert rel_err(f, oracle.tensorCRTC(c, pe, ruc, k)) <= 1e-9
    assert rel_err(capi.fused_w_emulate_c(pps, c, True), oracle.tensorCRTInvC(c, pe, ruci, T.mhat_inv_c(m, k), k)) <= 1e-9
    assert rel_err(capi.fused_w_emulate_c(pps, f, True), c) <= 1e-9
