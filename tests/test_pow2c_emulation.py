"""CPU suite: the schedule of lol_b200/csrc/fused_pow2c.cu (tensorCRTC / tensorCRTInvC for m = 2^e) restated in numpy -- the same
pass plan (rounds in threes, the remainder in twos), the same single table read per thread and pass with the other twiddles derived
from it (T_{rho-1}[p] = T_rho[p]^2, T_{r+s}[p + j 2^r] = T_{r+s}[p] root^(j n / 2^s)) -- against the oracle's tensorCRTC /
tensorCRTInvC (crt.cpp:583-598).  It makes the kernel's algebra checkable without a GPU; the GPU parity test is
tests/test_gpu_parity.py::test_complex_crt_power_of_two."""
import numpy as np
import pytest

from conftest import rel_err
from oracle import tables as T


def pass_plan(R):
    """fused_pow2c_crt: (first round, rounds) per pass of the forward transform."""
    plan, r = [], 0
    threes, rem = divmod(R, 3)
    if rem == 1 and threes > 0:
        threes -= 1
    for _ in range(threes):
        plan.append((r, 3)); r += 3
    while R - r >= 2:
        plan.append((r, 2)); r += 2
    if R - r == 1:
        plan.append((r, 1)); r += 1
    assert r == R and all(s in (1, 2, 3) for _, s in plan)
    return plan


def twiddle_table(root, n, e):
    """fused_pow2c_select: entry (2^r - 1) + p = root[(2p+1) n / 2^(r+1)]."""
    tw = np.ones(n, dtype=np.complex128)
    for r in range(e - 1):
        for p in range(1 << r):
            tw[(1 << r) - 1 + p] = root[(2 * p + 1) * (n >> (r + 1))]
    return tw


def run_pass(x, n, r, S, tw, rot1, rot2, inverse):
    """pow2c_pass<INV, S> for every thread index i."""
    st, V = 1 << r, 1 << S
    for i in range(n >> S):
        p = i & (st - 1)
        pos = ((i >> r) << (r + S)) | p
        ix = [pos + j * st for j in range(V)]
        v = [x[k] for k in ix]
        W = [[None] * (V // 2) for _ in range(S)]
        W[S - 1][0] = tw[((st << (S - 1)) - 1) + p]
        for s in range(S - 2, -1, -1):
            W[s][0] = W[s + 1][0] * W[s + 1][0]
        if S >= 2:
            W[1][1] = W[1][0] * rot1
        if S >= 3:
            W[2][1] = W[2][0] * rot2
            W[2][2] = W[2][0] * rot1
            W[2][3] = W[2][1] * rot1
        for ss in range(S):
            s = S - 1 - ss if inverse else ss
            half = 1 << s
            for a in range(V):
                if not a & half:
                    u, t, w = v[a], v[a + half], W[s][a & (half - 1)]
                    if not inverse:
                        v[a], v[a + half] = u + t * w, u - t * w
                    else:
                        v[a], v[a + half] = u + t, (u - t) * w
        for k, val in zip(ix, v):
            x[k] = val


@pytest.mark.parametrize("e", [3, 4, 5, 6, 7, 8, 9, 10, 11], ids=lambda v: f"m=2^{v}")
def test_pow2c_schedule_matches_the_oracle(oracle, e):
    m, n = 1 << e, 1 << (e - 1)
    pe = T.pe_array(m)
    ru, rui = T.ru_tables_c(m), T.ru_tables_c(m, inverse=True)
    rng = np.random.default_rng(e)
    y = (rng.normal(size=(n, 1)) + 1j * rng.normal(size=(n, 1))).astype(np.complex128)
    plan = pass_plan(e - 1)
    # forward
    root = ru[0].reshape(-1)
    x = y.reshape(-1).copy()
    tw = twiddle_table(root, n, e)
    for r, S in plan:
        run_pass(x, n, r, S, tw, root[n // 2], root[n // 4], False)
    assert rel_err(x.reshape(n, 1), oracle.tensorCRTC(y, pe, ru)) <= 1e-12
    # inverse: the passes backwards, then mhat^-1
    rooti = rui[0].reshape(-1)
    z = y.reshape(-1).copy()
    twi = twiddle_table(rooti, n, e)
    for r, S in reversed(plan):
        run_pass(z, n, r, S, twi, rooti[n // 2], rooti[n // 4], True)
    z *= T.mhat_inv_c(m)[0]
    assert rel_err(z.reshape(n, 1), oracle.tensorCRTInvC(y, pe, rui, T.mhat_inv_c(m))) <= 1e-12


def test_pow2c_pass_plans():
    assert [pass_plan(R) for R in (2, 3, 4, 5, 7, 10, 13)] == [
        [(0, 2)], [(0, 3)], [(0, 2), (2, 2)], [(0, 3), (3, 2)], [(0, 3), (3, 2), (5, 2)],
        [(0, 3), (3, 3), (6, 2), (8, 2)], [(0, 3), (3, 3), (6, 3), (9, 2), (11, 2)]]
