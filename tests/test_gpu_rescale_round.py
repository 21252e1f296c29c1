"""GPU suite (-m gpu): the coefficient-wise maps of coeff_stream.cu (lolb_liftRq, lolb_reduceRq, lolb_rescaleDropRq,
lolb_rescaleModRq, lolb_roundCosetRq) through the C ABI against oracle/coeffwise.py -- bit-exact (integer results; the
rounding performs the same IEEE operations as the restatement)."""
import numpy as np
import pytest

from conftest import zq_input
from oracle import coeffwise as W

pytestmark = pytest.mark.gpu

CASES = [(42, [19393921, 18869761]), (42, [2148854401, 2148249601, 2150668801]), (14400, [1008001, 1065601]),
         (14400, [14401, 1008001, 1065601]), (21, [8191, 43]), (8, [17, 4294967291])]


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


@pytest.mark.parametrize("m,qs", CASES, ids=lambda v: str(v))
def test_lift_reduce_rescale_bit_exact(torch_cuda, m, qs):
    torch = torch_cuda
    from lol_b200 import capi
    from lol_b200.tensor import CudaTensorRq
    t = CudaTensorRq(m, qs)
    rng = np.random.default_rng(m + len(qs))
    batch = 7
    x = zq_input(rng, t.n, qs, batch=batch)
    x[0, 0], x[0, 1], x[0, 2] = 0, [q - 1 for q in qs], [q // 2 for q in qs]
    dx = _dev(torch, x)
    before = capi.kernel_launch_count()
    assert np.array_equal(t.lift(dx).cpu().numpy(), W.lift(x, qs))
    z1 = rng.integers(-2**62, 2**62, size=(batch, t.n, 1))
    zk = rng.integers(-2**62, 2**62, size=(batch, t.n, len(qs)))
    assert np.array_equal(t.reduce(_dev(torch, z1)).cpu().numpy(), W.reduce(z1, qs))
    assert np.array_equal(t.reduce(_dev(torch, zk)).cpu().numpy(), W.reduce(zk, qs))
    assert torch.equal(t.reduce(t.lift(dx)), dx)
    for drop in (0, len(qs) - 1) + ((1,) if len(qs) > 2 else ()):
        assert np.array_equal(t.rescaleDrop(dx, drop).cpu().numpy(), W.rescale_drop(x, qs, drop)), drop
    qs_new = [257, 65537, 4294967291][:len(qs)]
    assert np.array_equal(t.rescaleMod(dx, qs_new).cpu().numpy(), W.rescale_mod(x, qs, qs_new))
    assert np.array_equal(t.rescaleMod(dx, qs[::-1]).cpu().numpy(), W.rescale_mod(x, qs, qs[::-1]))
    assert capi.kernel_launch_count() >= before + 8


@pytest.mark.parametrize("m,ps", [(42, [2]), (14400, [256]), (21, [7, 1065601])], ids=lambda v: str(v))
def test_round_coset_bit_exact(torch_cuda, m, ps):
    torch = torch_cuda
    from lol_b200.tensor import CudaTensorRq
    t = CudaTensorRq(m, ps)
    rng = np.random.default_rng(m)
    batch = 5
    e = rng.standard_normal((batch, t.n, len(ps))) * 40.0 * np.asarray(ps, dtype=np.float64)
    e[0, :8, 0] = [0.5, 1.5, 2.5, -0.5, -1.5, 3.0, -3.0, 0.0]
    zp = zq_input(rng, t.n, ps, batch=batch)
    got = t.roundCoset(_dev(torch, e), _dev(torch, zp)).cpu().numpy()
    assert np.array_equal(got, W.round_coset(e, zp, ps))
    assert np.array_equal(got % np.asarray(ps), zp)
    assert np.array_equal(t.roundCoset(_dev(torch, e)).cpu().numpy(), W.round_coset(e, None, ps))


def test_rescale_argument_errors(torch_cuda):
    torch = torch_cuda
    from lol_b200 import capi
    from lol_b200.tensor import CudaTensorRq
    t1 = CudaTensorRq(8, [17])
    x = torch.zeros((1, 4, 1), dtype=torch.int64, device="cuda")
    with pytest.raises(capi.LolB200Error):                    # nothing to drop
        t1.rescaleDrop(x, 0)
    t2 = CudaTensorRq(8, [4, 8])                              # 4 is not a unit modulo 8: `Field b` fails
    y = torch.zeros((1, 4, 2), dtype=torch.int64, device="cuda")
    with pytest.raises(capi.LolB200Error) as ei:
        t2.rescaleDrop(y, 0)
    assert ei.value.status == capi.LOLB_ERR_NOT_INVERTIBLE
    with pytest.raises(capi.LolB200Error):
        t2.rescaleDrop(y, 2)
