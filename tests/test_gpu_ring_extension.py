"""GPU suite (-m gpu): the ring-extension operators (lolb_twacePowDec, lolb_embedPow/Dec/CRT, lolb_coeffsPowDec,
lolb_twaceCRT; ext_stream.cu) through the C ABI against oracle/extension.py, which tests/test_oracle_extension.py pins
to the compiled reference through the reference's two-index properties.

  * bit-exact over Z_q tuples and int64 on the reference's two-index parameters (Tests/Default.hs:65-77) and on pairs
    under BASELINE's m = 14400; 1e-9 relative for double / complex;
  * the device-resident index tables against the oracle's;
  * the reference's properties themselves on the device at full size (prop_trem_*, prop_embed_crt, prop_twace_crt,
    TensorTests.hs:133-170) with the library's own CRT -- size-independent checks at m' = 14400.

Nothing here reads /root/reference.
"""
import numpy as np
import pytest

from conftest import rel_err, zq_input
from oracle import extension as X
from oracle import tables as T
from test_oracle_extension import IDS, TWO_INDEX_PARAMS

pytestmark = pytest.mark.gpu

FLOAT_TOL = 1e-9
BIG_PARAMS = [(225, 14400, [14401]), (14400 // 25, 14400, [1008001, 1065601]), (64, 14400, [14401, 1008001, 1065601]),
              (1, 14400, [14401]),
              # ~2^31 moduli: the exact 64-bit accumulation of twaceCRT reduces every 3 products (rel = 6 and 12)
              (3, 42, [2148854401, 2148249601, 2150668801]), (1, 21, [2148249601])]


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU suite needs a CUDA device: libctensor_b200 has no CPU path")
    from lol_b200 import build_library, capi
    build_library()
    assert capi.device_available()
    return torch


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _batch_input(rng, batch, n, qs):
    return zq_input(rng, n, qs, batch=batch)


@pytest.mark.parametrize("m,m2,qs", TWO_INDEX_PARAMS + BIG_PARAMS,
                         ids=IDS + [f"m{m}_m{m2}_k{len(qs)}" for m, m2, qs in BIG_PARAMS])
def test_extension_zq_bit_exact(torch_cuda, m, m2, qs):
    torch = torch_cuda
    from lol_b200 import capi
    from lol_b200.extension import CudaExtension
    from lol_b200.tensor import CudaTensorRq
    info = X.ExtInfo(m, m2)
    ext = CudaExtension(CudaTensorRq(m, qs), CudaTensorRq(m2, qs))
    assert (ext.phi, ext.phi2) == (info.phi, info.phi2)
    # device-resident tables
    dec = np.where(info.base_dec_idx < 0, -1, info.base_dec_idx * 2 + info.base_dec_neg)
    for which, want in ((capi.EXT_INDICES_POWDEC, info.ext_powdec), (capi.EXT_INDICES_CRT, info.ext_crt),
                        (capi.EXT_BASE_POW_J0, info.base_pow_j0), (capi.EXT_BASE_POW_J1, info.base_pow_j1),
                        (capi.EXT_BASE_DEC, dec), (capi.EXT_INDICES_COEFFS, info.ext_coeffs.reshape(-1))):
        assert np.array_equal(ext.ext.table(which), want), which
    before = capi.kernel_launch_count()
    rng = np.random.default_rng(m * 131 + m2)
    batch = 5 if info.phi2 > 1000 else 9                      # odd: exercises the tail of the 4-way batch loop
    x = _batch_input(rng, batch, info.phi, qs)                # elements of O_m
    y = _batch_input(rng, batch, info.phi2, qs)               # elements of O_m'
    x[0, 0, :] = 0                                            # -0 = 0 in embedDec
    dx, dy = _dev(torch, x), _dev(torch, y)
    got = {"embedPow": ext.embedPow(dx), "embedDec": ext.embedDec(dx), "embedCRT": ext.embedCRT(dx),
           "twacePowDec": ext.twacePowDec(dy), "twaceCRT": ext.twaceCRT(dy), "coeffs": ext.coeffs(dy)}
    torch.cuda.synchronize()
    assert capi.kernel_launch_count() == before + 6
    for b in range(batch):
        assert np.array_equal(got["embedPow"][b].cpu().numpy(), X.embed_pow(info, x[b])), ("embedPow", b)
        assert np.array_equal(got["embedDec"][b].cpu().numpy(), X.embed_dec(info, x[b], qs)), ("embedDec", b)
        assert np.array_equal(got["embedCRT"][b].cpu().numpy(), X.embed_crt(info, x[b])), ("embedCRT", b)
        assert np.array_equal(got["twacePowDec"][b].cpu().numpy(), X.twace_powdec(info, y[b])), ("twacePowDec", b)
        assert np.array_equal(got["coeffs"][b].cpu().numpy(), X.coeffs_powdec(info, y[b])), ("coeffs", b)
        if b < 2 or info.phi2 <= 1000:                        # the object-dtype oracle is slow at phi' = 3840
            assert np.array_equal(got["twaceCRT"][b].cpu().numpy(), X.twace_crt_zq(info, y[b], qs)), ("twaceCRT", b)


@pytest.mark.parametrize("m,m2,k", [(3, 21, 1), (4, 12, 2), (1, 7, 1), (45, 225, 3), (225, 14400, 1), (64, 576, 2)])
def test_extension_plain_rings(torch_cuda, m, m2, k):
    """int64 'R' (bit-exact, wrapping negate), double and complex (copies and sign flips: exact; twaceCRT over C: 1e-9)."""
    torch = torch_cuda
    from lol_b200.extension import CudaExtension
    from lol_b200.tensor import CudaTensorComplex, CudaTensorInt, CudaTensorReal
    info = X.ExtInfo(m, m2)
    rng = np.random.default_rng(m + 7 * m2 + k)
    batch = 6
    for cls, make in ((CudaTensorInt, lambda s: rng.integers(-2**62, 2**62, size=s)),
                      (CudaTensorReal, lambda s: rng.standard_normal(s)),
                      (CudaTensorComplex, lambda s: rng.standard_normal(s) + 1j * rng.standard_normal(s))):
        ext = CudaExtension(cls(m, k), cls(m2, k))
        x, y = make((batch, info.phi, k)), make((batch, info.phi2, k))
        if cls is CudaTensorInt:
            x[0, 0, 0] = -2**63                               # wrapping negate (hInt_t arithmetic)
        dx, dy = _dev(torch, x), _dev(torch, y)
        ep, ed, tw, co = ext.embedPow(dx), ext.embedDec(dx), ext.twacePowDec(dy), ext.coeffs(dy)
        with np.errstate(over="ignore"):
            for b in range(batch):
                assert np.array_equal(ep[b].cpu().numpy(), X.embed_pow(info, x[b]))
                assert np.array_equal(ed[b].cpu().numpy(), X.embed_dec(info, x[b]))
                assert np.array_equal(tw[b].cpu().numpy(), X.twace_powdec(info, y[b]))
                assert np.array_equal(co[b].cpu().numpy(), X.coeffs_powdec(info, y[b]))
        if cls is CudaTensorComplex:
            lo, hi = ext.lo, ext.hi
            ec = ext.embedCRT(dx)
            tc = ext.twaceCRT(dy)
            # gCRT over C: crt(mulGPow(scalarPow 1)), from the library's own (parity-tested) single-index operators
            unit = lambda t: _dev(torch, np.concatenate([np.ones((1, 1, k)), np.zeros((1, t.n - 1, k))], axis=1).astype(np.complex128))
            g_lo = lo.crt(lo.mulGPow(unit(lo)))[0, :, 0].cpu().numpy()
            g_hi = hi.crt(hi.mulGPow(unit(hi)))[0, :, 0].cpu().numpy()
            for b in range(batch):
                assert np.array_equal(ec[b].cpu().numpy(), X.embed_crt(info, x[b]))
                assert rel_err(tc[b].cpu().numpy(), X.twace_crt_c(info, y[b], g_lo, g_hi)) <= FLOAT_TOL
            # prop_twace_crt over C: twaceCRT = crt . twacePowDec . crtInv
            assert rel_err(tc.cpu().numpy(), lo.crt(ext.twacePowDec(hi.crtInv(dy))).cpu().numpy()) <= FLOAT_TOL
        else:
            assert ext.embedCRT(dx) is None and ext.twaceCRT(dy) is None      # no CRT over Z / R: the reference's Nothing


def test_extension_properties_at_full_size(torch_cuda):
    """prop_trem_pow / prop_trem_dec / prop_embed_crt / prop_twace_crt / prop_twace_dec on the device, m = 576 | m' = 14400,
    config C moduli, 4096 elements -- with the library's fused CRT kernels, bit-exact."""
    torch = torch_cuda
    from lol_b200.extension import CudaExtension
    from lol_b200.tensor import CudaTensorRq
    m, m2, qs = 14400 // 25, 14400, [1008001, 1065601]
    lo, hi = CudaTensorRq(m, qs), CudaTensorRq(m2, qs)
    ext = CudaExtension(lo, hi)
    gen = torch.Generator(device="cuda").manual_seed(3)
    batch = 4096
    q = torch.tensor(qs, device="cuda", dtype=torch.int64)
    x = torch.randint(0, 2**40, (batch, lo.n, 2), device="cuda", generator=gen, dtype=torch.int64) % q
    y = torch.randint(0, 2**40, (batch, hi.n, 2), device="cuda", generator=gen, dtype=torch.int64) % q
    assert torch.equal(ext.twacePowDec(ext.embedPow(x)), x)
    assert torch.equal(ext.twacePowDec(ext.embedDec(x)), x)
    assert torch.equal(ext.embedCRT(x), hi.crt(ext.embedPow(lo.crtInv(x))))
    assert torch.equal(ext.twaceCRT(y), lo.crt(ext.twacePowDec(hi.crtInv(y))))
    assert torch.equal(ext.twacePowDec(y), lo.lInv(ext.twacePowDec(hi.l(y))))
    assert torch.equal(ext.embedDec(x), hi.lInv(ext.embedPow(lo.l(x))))
    c = ext.coeffs(y)
    assert torch.equal(c[:, 0], ext.twacePowDec(y))
    # prop_coeffsBasis (CycTests.hs:71-76): y == sum_r embed(coeffs y)_r * powBasis_r, products through the CRT basis of O_m'
    from oracle import extension as X
    basis = ext.powBasisPow()
    assert np.array_equal(basis.cpu().numpy(), X.pow_basis_pow(X.ExtInfo(m, m2), 2))
    acc = torch.zeros_like(y[:256])
    for r in range(basis.shape[0]):
        acc = (acc + hi.crtMul(ext.embedPow(c[:256, r].contiguous()), hi.crt(basis[r:r + 1]))) % q
    assert torch.equal(hi.crtInv(acc), y[:256])


@pytest.mark.parametrize("m,m2,p", [(3, 21, 2), (5, 45, 2), (4, 20, 3)], ids=lambda v: str(v))
def test_crt_set_dec_on_the_device(torch_cuda, gpu_oracle, m, m2, p):
    """crtSetDec (Tensor.hs:184-186; CPP/Extension.hs:145-164): the host-built set uploaded by CudaExtension, taken to the
    powerful basis on the device the way UCyc.crtSet does for e = 1 (UCyc.hs:556-564: toPow . Dec): equals the oracle's L of
    the same vectors and sums to the ring's 1.  The set property itself (c_i c_j = delta_ij c_i) is tests/test_crt_set.py."""
    torch = torch_cuda
    from lol_b200 import crtset
    from lol_b200.extension import CudaExtension
    from lol_b200.tensor import CudaTensorRq
    lo, hi = CudaTensorRq(m, [p]), CudaTensorRq(m2, [p])
    ext = CudaExtension(lo, hi)
    cs = ext.crtSetDec(p)
    host = crtset.crt_set_dec(m, m2, p)
    assert cs.shape == (host.shape[0], hi.n, 1) and np.array_equal(cs.cpu().numpy()[..., 0], host)
    pw = hi.l(cs)
    pe = T.pe_array(m2)
    for r in range(host.shape[0]):
        assert np.array_equal(pw[r].cpu().numpy(), gpu_oracle.tensorLRq(host[r].reshape(-1, 1), pe, [p]))
    total = pw.sum(dim=0) % p
    assert int(total[0, 0]) == 1 and int(total[1:].abs().sum()) == 0


def test_extension_argument_errors(torch_cuda):
    torch = torch_cuda
    from lol_b200 import capi
    from lol_b200.extension import CudaExtension
    from lol_b200.tensor import CudaTensorInt, CudaTensorRq
    with pytest.raises(capi.LolB200Error):                    # 4 does not divide 6
        CudaExtension(CudaTensorRq(4, [13]), CudaTensorRq(6, [13]))
    with pytest.raises(capi.LolB200Error):                    # different moduli
        CudaExtension(CudaTensorRq(3, [13]), CudaTensorRq(6, [7]))
    with pytest.raises(capi.LolB200Error):                    # different rings
        CudaExtension(CudaTensorInt(3), CudaTensorRq(6, [7]))
    # a modulus without a CRT of index m': the gathers work, the CRT-basis pair is Nothing
    ext = CudaExtension(CudaTensorRq(4, [8]), CudaTensorRq(28, [8]))
    x = torch.arange(2 * 2, device="cuda", dtype=torch.int64).reshape(2, 2, 1) % 8
    assert torch.equal(ext.twacePowDec(ext.embedPow(x)), x)
    assert torch.equal(ext.twacePowDec(ext.embedDec(x)), x)
    assert ext.embedCRT(x) is None
    assert ext.twaceCRT(ext.embedPow(x)) is None
    # wrong ring tag / aliasing operands through the raw C ABI
    assert ext.ext.op("embedPow", capi.RING_C, x.data_ptr(), x.data_ptr() + 64, 1) == capi.LOLB_ERR_ARG
    assert ext.ext.op("twacePowDec", capi.RING_RQ, x.data_ptr(), x.data_ptr(), 1) == capi.LOLB_ERR_ARG


import glob as _glob
import os as _os

EXT_GOLDEN = sorted(_glob.glob(_os.path.join(_os.path.dirname(__file__), "golden", "ext_*.npz")))


@pytest.mark.parametrize("path", EXT_GOLDEN, ids=[_os.path.basename(p)[:-4] for p in EXT_GOLDEN])
def test_extension_matches_reference_derived_golden(torch_cuda, path):
    """The CUDA operators against tests/golden/ext_*.npz, whose CRT- and Dec-basis vectors were computed by the compiled
    reference (oracle/make_golden_ext.py)."""
    torch = torch_cuda
    from lol_b200.extension import CudaExtension
    from lol_b200.tensor import CudaTensorRq
    g = np.load(path)
    m, m2, qs = int(g["m"]), int(g["m2"]), [int(q) for q in g["qs"]]
    ext = CudaExtension(CudaTensorRq(m, qs), CudaTensorRq(m2, qs))
    dx, dy = _dev(torch, g["x_in"]), _dev(torch, g["y_in"])
    assert np.array_equal(ext.embedPow(dx).cpu().numpy(), g["embedPow"])
    assert np.array_equal(ext.embedDec(dx).cpu().numpy(), g["embedDec"])
    assert np.array_equal(ext.embedCRT(dx).cpu().numpy(), g["embedCRT"])
    assert np.array_equal(ext.twacePowDec(dy).cpu().numpy(), g["twacePowDec"])
    assert np.array_equal(ext.coeffs(dy).cpu().numpy(), g["coeffs"])
    assert np.array_equal(ext.twaceCRT(dy).cpu().numpy(), g["twaceCRT"])
