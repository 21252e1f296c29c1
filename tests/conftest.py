import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# (m, qs): the reference's own tensor-test parameters (lol/Crypto/Lol/Tests/Default.hs:46-77, 128-133),
# the "paper" benchmark sets (lol/Crypto/Lol/Benchmarks/Default.hs:41-46) and BASELINE.json configs A, C
REFERENCE_TEST_PARAMS = [
    (7, [29]), (12, [2148249601]), (1, [17]), (2, [17]), (4, [17]), (8, [17]),
    (21, [8191]), (42, [8191]), (42, [18869761]), (2, [19393921, 18869761]), (3, [19393921, 18869761]),
    (7, [19393921, 18869761]), (6, [19393921, 18869761]), (42, [2148854401, 2148249601, 2150668801]),
    (42, [19393921, 18869761]), (89, [179]),
]
PAPER_PARAMS = [(1024, [12289]), (2048, [12289]), (64 * 27, [3457]), (64 * 81, [10369]), (14400, [14401])]
CONFIG_A = (14400, [14401])
CONFIG_B = (65536, [537133057, 537591809, 537722881, 538116097])
CONFIG_C = (14400, [1008001, 1065601])
# composite / non-CRT moduli the Cyc tests use for L and G (Tests/Default.hs:79-85: Zq PP2/PP4/PP8)
NON_CRT_PARAMS = [(28, [4]), (91, [8]), (7, [2]), (12, [9]), (21, [21])]


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu
    return cpu.restatement()


@pytest.fixture(scope="session")
def reference():
    from oracle import cpu
    if not cpu.have_reference():
        try:
            cpu.build("ref")
        except Exception:
            pass
    if not cpu.have_reference():
        pytest.skip("oracle/_ref/libctensor_ref.so not built (no /root/reference here)")
    return cpu.reference()


class _PinnedOracle:
    """The compiled reference (oracle/_ref) with the four symbols whose reference implementation is defective
    (tensorGInv{Pow,Dec}R: inverted divisibility test, g.cpp:175-182; tensorGInv{Pow,Dec}C: integer 1/oddrad, g.cpp:213-218)
    taken from the restatement, which implements the documented intent."""
    INTENT = {"tensorGInvPowR", "tensorGInvDecR", "tensorGInvPowC", "tensorGInvDecC"}

    def __init__(self, ref, rest):
        self._ref, self._rest = ref, rest
        self.kind = "reference+intent"

    def __getattr__(self, name):
        return getattr(self._rest if name in self.INTENT else self._ref, name)


@pytest.fixture(scope="session")
def gpu_oracle():
    """What the GPU suites compare against: the compiled reference itself when oracle/_ref is on the box (it ships with the
    repo snapshot), one hop instead of two; the restatement (pinned to it by tests/test_oracle_pinning.py) otherwise."""
    from oracle import cpu
    rest = cpu.restatement()
    if cpu.have_reference():
        return _PinnedOracle(cpu.reference(), rest)
    return rest


def zq_input(rng, n, qs, batch=None):
    shape = (n,) if batch is None else (batch, n)
    return np.stack([rng.integers(0, q, size=shape) for q in qs], axis=-1).astype(np.int64)


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(np.abs(b).max(), 1e-300)
    return float(np.abs(a - b).max() / scale)
