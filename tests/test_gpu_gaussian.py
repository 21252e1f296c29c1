"""GPU suite (-m gpu): the on-device Gaussian source (lolb_realGaussians, lolb_tGaussianDec).

The reference draws continuous Gaussians on the host with the polar Box-Muller transform over a `MonadRandom`
(lol/Crypto/Lol/GaussRandom.hs:34-59) and feeds them to tensorGaussianDec (lol-cpp/.../CPP.hs:376-389).  The library runs the
same transform on the device over Philox4x32-10, another uniform source, so parity is distributional:

  * moments and a Kolmogorov-Smirnov test of lolb_realGaussians against N(0, svar / (2 pi)), independence of neighbours;
  * counter-based determinism: same (seed, element) -> same draw whatever the batch split; different seeds differ;
  * lolb_tGaussianDec (one fused pass) against lolb_realGaussians + lolb_tensorGaussianDec (two passes): same distribution
    per coefficient (two-sample KS) and the same mean gSqNorm;
  * the reference's tail bound for tGaussian (lol/Crypto/Lol/RLWE/Continuous.hs:74-84): gSqNormDec of a sample exceeds
    errorBound(v, eps) with probability about eps.
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU suite needs a CUDA device: libctensor_b200 has no CPU path")
    from lol_b200 import build_library
    build_library()
    return torch


def _error_bound(m, n, v, eps):
    """Continuous.hs:74-84 (errorBound): mhat * n * v * stabilize(1 / (2 pi))."""
    mhat = m // 2 if m % 2 == 0 else m
    x = 1.0 / (2.0 * math.pi)
    while True:
        x2 = (0.5 + math.log(2.0 * math.pi * x) / 2.0 - math.log(eps) / n) / math.pi
        if x2 - x < 0.0001:
            return mhat * n * v * x2
        x = x2


def test_real_gaussians_distribution_and_determinism(torch_cuda):
    torch = torch_cuda
    from scipy import stats
    from lol_b200.tensor import CudaTensorReal
    t = CudaTensorReal(14400)
    svar = 3.7
    x = t.realGaussians(svar, 256, seed=11)                      # 256 x 3840 draws
    flat = x.reshape(-1).cpu().numpy()
    sigma = math.sqrt(svar / (2.0 * math.pi))
    n = flat.size
    assert abs(flat.mean()) < 5 * sigma / math.sqrt(n)
    assert abs(flat.var() / sigma ** 2 - 1.0) < 5 * math.sqrt(2.0 / n)
    assert abs(stats.kurtosis(flat)) < 0.05 and abs(stats.skew(flat)) < 0.02
    assert stats.kstest(flat[:200000] / sigma, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(flat[0::2], flat[1::2])[0, 1]) < 5 / math.sqrt(n / 2)      # the two outputs of one Box-Muller pair
    # counter-based: element e of any batch with first = f is element e + f of the stream
    assert torch.equal(t.realGaussians(svar, 256, seed=11), x)
    assert torch.equal(t.realGaussians(svar, 100, seed=11, first=156), x[156:])
    assert not torch.equal(t.realGaussians(svar, 4, seed=12), x[:4])
    # odd n (m = 3 has n = 2; use the raw entry point with n = 7)
    from lol_b200 import capi
    y = torch.empty(5, 7, dtype=torch.float64, device="cuda")
    capi.check(capi.real_gaussians(svar, 1, 0, y.data_ptr(), 7, 5))
    torch.cuda.synchronize()
    assert torch.isfinite(y).all() and (y != 0).all()


@pytest.mark.parametrize("m", [14400, 1728, 21, 42], ids=str)
def test_tgaussiandec_matches_two_pass_path_in_distribution(torch_cuda, m):
    torch = torch_cuda
    from scipy import stats
    from lol_b200.factored import radical_fact
    from lol_b200.tensor import CudaTensorReal
    t = CudaTensorReal(m)
    v, B = 0.1, 4096 if m > 100 else 65536
    fused = t.tGaussianDec(v, B, seed=3)
    assert fused.shape == (B, t.n, 1) and torch.isfinite(fused).all()
    assert torch.equal(t.tGaussianDec(v, B, seed=3), fused)                    # deterministic
    assert torch.equal(t.tGaussianDec(v, B // 2, seed=3, first=B // 2), fused[B // 2:])
    two = t.gaussianDecTransform(t.realGaussians(v * (m // radical_fact(m)), B, seed=4), inplace=True)
    t.plan.force_generic(True)                                                 # the same call on the two-pass generic path
    gen = t.tGaussianDec(v, B, seed=5)
    t.plan.force_generic(False)
    for j in (0, 1, t.n // 2, t.n - 1):                                        # per-coefficient distributions agree
        a, b, c = fused[:, j, 0].cpu().numpy(), two[:, j, 0].cpu().numpy(), gen[:, j, 0].cpu().numpy()
        assert stats.ks_2samp(a, b).pvalue > 1e-4, j
        assert stats.ks_2samp(a, c).pvalue > 1e-4, j
    na, nb = t.gSqNormDec(fused).mean().item(), t.gSqNormDec(two).mean().item()
    assert abs(na / nb - 1.0) < 0.05
    # Continuous.hs:74-84: P[gSqNorm > errorBound(v, eps)] is about eps
    eps = 2.0 ** -5
    bound = _error_bound(m, t.n, v, eps)
    frac = (t.gSqNormDec(fused) > bound).double().mean().item()
    assert frac < 4 * eps
