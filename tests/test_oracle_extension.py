"""CPU suite: pin oracle/extension.py (the ring-extension index tables and operators, Haskell in the reference)
through the reference's own two-index properties (lol/Crypto/Lol/Tests/TensorTests.hs:133-215), with the COMPILED
reference supplying crt / crtInv / l / lInv / divG, over the reference's two-index parameter list
(lol/Crypto/Lol/Tests/Default.hs:65-77) plus the ring-switching pairs below m = 14400."""
import numpy as np
import pytest

from conftest import zq_input
from oracle import extension as X
from oracle import tables as T

# (m, m', qs): Tests/Default.hs:65-77, then pairs under BASELINE's m = 14400 and a mixed odd pair
TWO_INDEX_PARAMS = [
    (1, 7, [29]), (4, 12, [536871001]), (4, 12, [2148249601]), (2, 8, [17]), (8, 8, [17]), (2, 8, [2148249601]),
    (4, 8, [17]), (3, 21, [8191]), (7, 21, [8191]), (3, 42, [8191]), (3, 21, [18869761]),
    (7, 21, [19393921, 18869761]), (3, 42, [19918081, 19393921, 18869761]),
    (45, 225, [14401]), (64, 14400 // 25, [14401]), (75, 14400 // 64, [1008001, 1065601]), (15, 105, [2311]),
]
IDS = [f"m{m}_m{m2}_k{len(qs)}" for m, m2, qs in TWO_INDEX_PARAMS]


class Ring:
    """The reference's single-index operators over Z_q for one index (compiled lol-cpp through ctypes)."""

    def __init__(self, lib, m, qs):
        self.lib, self.m, self.qs = lib, m, qs
        self.pe = T.pe_array(m)
        self.n = T.totient_pps(T.factor_pps(m))
        self.ru, self.rui = T.ru_tables_zq(m, qs), T.ru_tables_zq(m, qs, inverse=True)
        self.mh = [T.mhat_inv(m, q) for q in qs]

    def crt(self, x): return self.lib.tensorCRTRq(x, self.pe, self.ru, self.qs).reshape(x.shape)
    def crt_inv(self, x): return self.lib.tensorCRTInvRq(x, self.pe, self.rui, self.mh, self.qs).reshape(x.shape)
    def l(self, x): return self.lib.tensorLRq(x, self.pe, self.qs).reshape(x.shape)
    def l_inv(self, x): return self.lib.tensorLInvRq(x, self.pe, self.qs).reshape(x.shape)

    def scalar_pow(self, r):
        x = np.zeros((self.n, len(self.qs)), dtype=np.int64)
        x[0] = [r % q for q in self.qs]
        return x

    def div_g(self, name, x):
        y, ok = getattr(self.lib, name)(x, self.pe, self.qs)
        assert ok == 1
        return y.reshape(x.shape)


@pytest.mark.parametrize("m,m2,qs", TWO_INDEX_PARAMS, ids=IDS)
def test_index_tables_are_consistent(m, m2, qs):
    info = X.ExtInfo(m, m2)
    # toIndexPair / fromIndexPair are inverse bijections [phi'] <-> [phi'/phi] x [phi]
    for j in range(info.phi2):
        assert X.from_index_pair(info.tots, X.to_index_pair(info.tots, j)) == j
    assert sorted(info.ext_crt.tolist()) == list(range(info.phi2))
    assert sorted(info.ext_coeffs.reshape(-1).tolist()) == list(range(info.phi2))
    assert np.array_equal(info.ext_coeffs[0], info.ext_powdec)
    # every base index is hit exactly once with j0 == 0 (embedPow is injective onto the j0 = 0 slots)
    assert sorted(info.base_pow_j1[info.base_pow_j0 == 0].tolist()) == list(range(info.phi))


@pytest.mark.parametrize("m,m2,qs", TWO_INDEX_PARAMS, ids=IDS)
def test_twace_embed_identities(reference, m, m2, qs):
    """prop_trem_pow, prop_trem_dec, prop_twace_dec, prop_twEmID."""
    rng = np.random.default_rng(m * 1000 + m2)
    info = X.ExtInfo(m, m2)
    lo, hi = Ring(reference, m, qs), Ring(reference, m2, qs)
    x = zq_input(rng, info.phi, qs)
    assert np.array_equal(X.twace_powdec(info, X.embed_pow(info, x)), x)
    assert np.array_equal(X.twace_powdec(info, X.embed_dec(info, x, qs)), x)
    y = zq_input(rng, info.phi2, qs)
    assert np.array_equal(X.twace_powdec(info, y), lo.l_inv(X.twace_powdec(info, hi.l(y))))
    # embedDec == lInv . embedPow . l  (both express the same ring element)
    assert np.array_equal(X.embed_dec(info, x, qs), hi.l_inv(X.embed_pow(info, lo.l(x))))
    same = X.ExtInfo(m2, m2)
    assert np.array_equal(X.twace_powdec(same, y), y)
    assert np.array_equal(X.embed_pow(same, y), y)
    assert np.array_equal(X.embed_dec(same, y, qs), y)
    assert np.array_equal(X.embed_crt(same, y), y)
    assert np.array_equal(X.twace_crt_zq(same, y, qs), y)


@pytest.mark.parametrize("m,m2,qs", TWO_INDEX_PARAMS, ids=IDS)
def test_crt_basis_operators_against_compiled_reference_crt(reference, m, m2, qs):
    """prop_embed_crt: embedCRT = crt . embedPow . crtInv;  prop_twace_crt: twaceCRT = crt . twacePowDec . crtInv."""
    rng = np.random.default_rng(m * 977 + m2)
    info = X.ExtInfo(m, m2)
    lo, hi = Ring(reference, m, qs), Ring(reference, m2, qs)
    x = zq_input(rng, info.phi, qs)
    assert np.array_equal(X.embed_crt(info, x), hi.crt(X.embed_pow(info, lo.crt_inv(x))))
    y = zq_input(rng, info.phi2, qs)
    assert np.array_equal(X.twace_crt_zq(info, y, qs), lo.crt(X.twace_powdec(info, hi.crt_inv(y))))


@pytest.mark.parametrize("m,m2,qs", TWO_INDEX_PARAMS, ids=IDS)
def test_twace_invariants(reference, m, m2, qs):
    """prop_twace_invar1_{pow,dec,crt}: twace(mhat'/g') = mhat (phi'/phi) / g;  prop_twace_invar2: scalars are kept."""
    if any(np.gcd(T.odd_radical(m2), q) != 1 for q in qs):
        pytest.skip("g not invertible")
    info = X.ExtInfo(m, m2)
    lo, hi = Ring(reference, m, qs), Ring(reference, m2, qs)
    mhat, mhat2 = T.value_hat(m), T.value_hat(m2)
    out_s, in_s = lo.scalar_pow(mhat * info.rel), hi.scalar_pow(mhat2)
    assert np.array_equal(X.twace_powdec(info, hi.div_g("tensorGInvPowRq", in_s)), lo.div_g("tensorGInvPowRq", out_s))
    assert np.array_equal(X.twace_powdec(info, hi.div_g("tensorGInvDecRq", hi.l_inv(in_s))),
                          lo.div_g("tensorGInvDecRq", lo.l_inv(out_s)))
    _, gi_lo = T.g_crt_vectors(m, qs)
    _, gi_hi = T.g_crt_vectors(m2, qs)
    q = np.asarray(qs, dtype=object)
    crt_in = (hi.crt(in_s).astype(object) * gi_hi.astype(object) % q).astype(np.int64)
    crt_out = (lo.crt(out_s).astype(object) * gi_lo.astype(object) % q).astype(np.int64)
    assert np.array_equal(X.twace_crt_zq(info, crt_in, qs), crt_out)
    assert np.array_equal(X.twace_powdec(info, hi.scalar_pow(1)), lo.scalar_pow(1))
    assert np.array_equal(X.twace_crt_zq(info, hi.crt(hi.scalar_pow(1)), qs), lo.crt(lo.scalar_pow(1)))


@pytest.mark.parametrize("m,m2", [(3, 21), (4, 12), (45, 225), (1, 7)])
def test_coeffs_reassembles_through_the_powerful_basis(m, m2):
    """coeffs' splits an O_m' element into phi'/phi O_m elements; placing them back by extIndicesCoeffs is the identity,
    and coefficient 0 is twacePowDec (Extension.hs:90-103)."""
    rng = np.random.default_rng(5)
    info = X.ExtInfo(m, m2)
    y = rng.integers(-50, 50, size=(info.phi2, 1))
    c = X.coeffs_powdec(info, y)
    assert c.shape == (info.rel, info.phi, 1)
    back = np.empty_like(y)
    back[info.ext_coeffs.reshape(-1)] = c.reshape(-1, 1)
    assert np.array_equal(back, y)
    assert np.array_equal(c[0], X.twace_powdec(info, y))


@pytest.mark.parametrize("m,m2", [(m, m2) for m, m2, _ in TWO_INDEX_PARAMS] + [(225, 14400), (576, 14400), (1, 14400), (1024, 65536)],
                         ids=lambda v: str(v))
def test_library_host_index_tables_match_oracle(m, m2):
    """lolb_ext_index_table (host-only entry of libctensor_b200, ext_stream.cu: the tables the kernels gather through)
    against the restatement of Tensor.hs:391-498 -- no GPU involved."""
    from lol_b200 import build_library, capi
    build_library()
    info = X.ExtInfo(m, m2)
    pps, pps2 = T.factor_pps(m), T.factor_pps(m2)
    dec = np.where(info.base_dec_idx < 0, -1, info.base_dec_idx * 2 + info.base_dec_neg)
    for which, want in ((capi.EXT_INDICES_POWDEC, info.ext_powdec), (capi.EXT_INDICES_CRT, info.ext_crt),
                        (capi.EXT_BASE_POW_J0, info.base_pow_j0), (capi.EXT_BASE_POW_J1, info.base_pow_j1),
                        (capi.EXT_BASE_DEC, dec), (capi.EXT_INDICES_COEFFS, info.ext_coeffs.reshape(-1))):
        assert np.array_equal(capi.ext_index_table(pps, pps2, which), want), which


def test_library_host_index_tables_reject_non_divisors():
    from lol_b200 import build_library, capi
    build_library()
    for m, m2 in ((4, 6), (9, 3), (5, 12)):
        with pytest.raises(capi.LolB200Error):
            capi.ext_index_table(T.factor_pps(m), T.factor_pps(m2), capi.EXT_INDICES_POWDEC)


@pytest.mark.parametrize("m,m2", [(3, 21), (4, 12), (1, 7), (45, 225), (64, 576), (15, 105)], ids=lambda v: str(v))
def test_complex_crt_basis_operators_against_compiled_reference_crt(reference, m, m2):
    """prop_embed_crt / prop_twace_crt over Complex Double with the compiled reference's tensorCRTC / tensorCRTInvC / tensorGPowC:
    pins twace_crt_c (and the g = crt(mulGPow 1) form of the gCRT vector the CUDA path uses) to 1e-9 relative."""
    from conftest import rel_err
    rng = np.random.default_rng(m + m2)
    info = X.ExtInfo(m, m2)

    def ring(mm):
        pe = T.pe_array(mm)
        n = T.totient_pps(T.factor_pps(mm))
        ru, rui, mh = T.ru_tables_c(mm), T.ru_tables_c(mm, inverse=True), T.mhat_inv_c(mm)
        crt = lambda v: reference.tensorCRTC(v, pe, ru).reshape(v.shape)
        crt_inv = lambda v: reference.tensorCRTInvC(v, pe, rui, mh).reshape(v.shape)
        unit = np.zeros((n, 1), dtype=np.complex128)
        unit[0] = 1.0
        g = crt(reference.tensorGPowC(unit, pe).reshape(unit.shape))[:, 0]
        return n, crt, crt_inv, g

    n_lo, crt_lo, crt_inv_lo, g_lo = ring(m)
    n_hi, crt_hi, crt_inv_hi, g_hi = ring(m2)
    x = rng.standard_normal((n_lo, 1)) + 1j * rng.standard_normal((n_lo, 1))
    y = rng.standard_normal((n_hi, 1)) + 1j * rng.standard_normal((n_hi, 1))
    assert rel_err(X.embed_crt(info, x), crt_hi(X.embed_pow(info, crt_inv_lo(x)))) <= 1e-9
    assert rel_err(X.twace_crt_c(info, y, g_lo, g_hi), crt_lo(X.twace_powdec(info, crt_inv_hi(y)))) <= 1e-9


def test_library_host_index_tables_sweep_all_divisor_pairs():
    """Every pair m | m' with m' <= 120 (and a few larger indices with three and four prime factors): the C++ tables of
    ext_stream.cu (iterative mixed-radix forms) equal the recursive restatement of Tensor.hs:391-498."""
    from lol_b200 import build_library, capi
    build_library()
    pairs = [(m, m2) for m2 in list(range(1, 121)) + [420, 1155, 2 * 3 * 5 * 7 * 11, 27 * 25 * 7] for m in range(1, m2 + 1) if m2 % m == 0]
    assert len(pairs) > 600
    for m, m2 in pairs:
        info = X.ExtInfo(m, m2)
        pps, pps2 = T.factor_pps(m), T.factor_pps(m2)
        dec = np.where(info.base_dec_idx < 0, -1, info.base_dec_idx * 2 + info.base_dec_neg)
        for which, want in ((capi.EXT_INDICES_POWDEC, info.ext_powdec), (capi.EXT_INDICES_CRT, info.ext_crt),
                            (capi.EXT_BASE_POW_J0, info.base_pow_j0), (capi.EXT_BASE_DEC, dec),
                            (capi.EXT_INDICES_COEFFS, info.ext_coeffs.reshape(-1))):
            assert np.array_equal(capi.ext_index_table(pps, pps2, which), want), (m, m2, which)


@pytest.mark.parametrize("m,m2,qs", [p for p in TWO_INDEX_PARAMS if p[1] > 1], ids=[i for i, p in zip(IDS, TWO_INDEX_PARAMS) if p[1] > 1])
def test_prop_coeffsBasis(reference, m, m2, qs):
    """lol/Crypto/Lol/Tests/CycTests.hs:71-76: x == sum_k embed(coeffs x)_k * b_k over the powerful extension basis b
    (powBasisPow', Extension.hs:131-141), the products taken in O_m' through the compiled reference's CRT."""
    rng = np.random.default_rng(3 * m + m2)
    info = X.ExtInfo(m, m2)
    hi = Ring(reference, m2, qs)
    x = zq_input(rng, info.phi2, qs)
    q = np.asarray(qs, dtype=object)
    acc = np.zeros((info.phi2, len(qs)), dtype=object)
    cs = X.coeffs_powdec(info, x)
    basis = X.pow_basis_pow(info, len(qs))                                  # powBasisPow': one where (j0, j1) == (k, 0)
    for k in range(info.rel):
        b = basis[k]
        assert b.sum() == len(qs)
        prod = hi.crt(X.embed_pow(info, cs[k])).astype(object) * hi.crt(b).astype(object) % q
        acc = (acc + prod) % q
    assert np.array_equal(hi.crt_inv(acc.astype(np.int64)), x)


import glob as _glob
import os as _os

EXT_GOLDEN = sorted(_glob.glob(_os.path.join(_os.path.dirname(__file__), "golden", "ext_*.npz")))


def test_ext_golden_files_present():
    assert len(EXT_GOLDEN) >= 5


@pytest.mark.parametrize("path", EXT_GOLDEN, ids=[_os.path.basename(p)[:-4] for p in EXT_GOLDEN])
def test_restatement_matches_reference_derived_golden(path):
    """tests/golden/ext_*.npz (oracle/make_golden_ext.py): CRT-basis and Dec-basis vectors computed by the COMPILED reference
    through crt . embedPow . crtInv, crt . twacePowDec . crtInv and lInv . embedPow . l; the direct restatements of
    embedCRT', twaceCRT' and embedDec' (Extension.hs:71-85, 110-129) must reproduce them bit for bit."""
    g = np.load(path)
    m, m2, qs = int(g["m"]), int(g["m2"]), [int(q) for q in g["qs"]]
    info = X.ExtInfo(m, m2)
    for b in range(g["x_in"].shape[0]):
        x, y = g["x_in"][b], g["y_in"][b]
        assert np.array_equal(X.embed_pow(info, x), g["embedPow"][b])
        assert np.array_equal(X.embed_dec(info, x, qs), g["embedDec"][b])
        assert np.array_equal(X.embed_crt(info, x), g["embedCRT"][b])
        assert np.array_equal(X.twace_powdec(info, y), g["twacePowDec"][b])
        assert np.array_equal(X.coeffs_powdec(info, y), g["coeffs"][b])
        assert np.array_equal(X.twace_crt_zq(info, y, qs), g["twaceCRT"][b])


def test_two_index_properties_sweep_all_divisors(reference):
    """prop_trem_pow / prop_trem_dec / prop_embed_crt / prop_twace_crt / prop_twace_dec for EVERY divisor m of
    m' in {12, 21, 36, 42, 45, 63, 75, 105, 225}, over the first prime q = 1 (mod m') above 2^20 (goodQs, ZqBasic.hs:71-73),
    with the compiled reference's crt / crtInv / l / lInv."""
    count = 0
    for m2 in (12, 21, 36, 42, 45, 63, 75, 105, 225):
        q = next(iter(T.good_qs(m2, 1 << 20)))
        hi = Ring(reference, m2, [q])
        for m in (d for d in range(1, m2 + 1) if m2 % d == 0):
            rng = np.random.default_rng(m2 * 1000 + m)
            info = X.ExtInfo(m, m2)
            lo = Ring(reference, m, [q])
            x, y = zq_input(rng, info.phi, [q]), zq_input(rng, info.phi2, [q])
            assert np.array_equal(X.twace_powdec(info, X.embed_pow(info, x)), x), (m, m2)
            assert np.array_equal(X.twace_powdec(info, X.embed_dec(info, x, [q])), x), (m, m2)
            assert np.array_equal(X.embed_crt(info, x), hi.crt(X.embed_pow(info, lo.crt_inv(x)))), (m, m2)
            assert np.array_equal(X.twace_crt_zq(info, y, [q]), lo.crt(X.twace_powdec(info, hi.crt_inv(y)))), (m, m2)
            assert np.array_equal(X.twace_powdec(info, y), lo.l_inv(X.twace_powdec(info, hi.l(y)))), (m, m2)
            assert np.array_equal(X.embed_dec(info, x, [q]), hi.l_inv(X.embed_pow(info, lo.l(x)))), (m, m2)
            count += 1
    assert count >= 60
