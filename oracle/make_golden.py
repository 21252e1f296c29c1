"""ORACLE (test infrastructure): generate tests/golden/*.npz from the compiled,
UNMODIFIED reference (oracle/_ref/libctensor_ref.so, built by `make -C oracle ref`).

The reference ships no known-answer vectors, so these are
outputs of the reference itself on seeded inputs.  Run in the build container
(where /root/reference exists):

    python -m oracle.make_golden

Each small case stores inputs and outputs; the m = 2^16, k = 4 case stores
SHA-256 digests of the outputs (the arrays are 1 MiB each) with the input seed.
The tests replay the same calls against the C restatement (CPU suite) and the
CUDA library (GPU suite).
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cpu, tables as T  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (name, m, qs) -- the reference's own test parameters (lol/Crypto/Lol/Tests/Default.hs:46-77,
# 128-133) plus BASELINE.json configs A and C
SMALL_CASES = [
    ("f7_q29", 7, [29]),
    ("f8_q17", 8, [17]),
    ("f12_smooth1", 12, [2148249601]),
    ("f21_q8191", 21, [8191]),
    ("f42_zq2", 42, [19393921, 18869761]),
    ("f42_smooth3", 42, [2148854401, 2148249601, 2150668801]),
    ("f89_q179", 89, [179]),
    ("cfgA_m14400_q14401", 14400, [14401]),
    ("cfgC_m14400_k2", 14400, [1008001, 1065601]),
]
CONFIG_B = ("cfgB_m65536_k4", 65536, [537133057, 537591809, 537722881, 538116097])


def zq_input(rng, n, qs):
    return np.stack([rng.integers(0, q, size=n) for q in qs], axis=1).astype(np.int64)


def make_small(R, name, m, qs, seed):
    rng = np.random.default_rng(seed)
    pe = T.pe_array(m)
    n = T.totient_pps(T.factor_pps(m))
    k = len(qs)
    ru, rui = T.ru_tables_zq(m, qs), T.ru_tables_zq(m, qs, inverse=True)
    mh = np.array([T.mhat_inv(m, q) for q in qs], dtype=np.int64)
    out = {"m": np.int64(m), "qs": np.array(qs, dtype=np.int64), "seed": np.int64(seed)}

    y, y2 = zq_input(rng, n, qs), zq_input(rng, n, qs)
    out["rq_in"], out["rq_in2"] = y, y2
    out["CRTRq"] = R.tensorCRTRq(y, pe, ru, qs)
    out["CRTInvRq"] = R.tensorCRTInvRq(y, pe, rui, mh, qs)
    for nm in ("LRq", "LInvRq", "GPowRq", "GDecRq"):
        out[nm] = getattr(R, "tensor" + nm)(y, pe, qs)
    for nm in ("GInvPowRq", "GInvDecRq"):
        arr, st = getattr(R, "tensor" + nm)(y, pe, qs)
        out[nm], out[nm + "_status"] = arr, np.int64(st)
    out["mulRq"] = R.mulRq(y, y2, qs)

    z = rng.integers(-(2 ** 40), 2 ** 40, size=(n, 1)).astype(np.int64)
    out["r_in"] = z
    for nm in ("LR", "LInvR", "GPowR", "GDecR"):
        out[nm] = getattr(R, "tensor" + nm)(z, pe)
    zs = rng.integers(-8, 9, size=(n, 1)).astype(np.int64)
    out["norm_in"] = zs
    out["NormSqR"] = R.tensorNormSqR(zs, pe).reshape(-1)[:1]

    d = rng.normal(size=(n, 1))
    out["d_in"] = d
    ruc, ruci = T.ru_tables_c(m), T.ru_tables_c(m, inverse=True)
    for nm in ("LDouble", "LInvDouble"):
        out[nm] = getattr(R, "tensor" + nm)(d, pe)
    out["NormSqD"] = R.tensorNormSqD(d, pe).reshape(-1)[:1]
    out["GaussianDec"] = R.tensorGaussianDec(d, pe, ruc)

    c = rng.normal(size=(n, 1)) + 1j * rng.normal(size=(n, 1))
    c2 = rng.normal(size=(n, 1)) + 1j * rng.normal(size=(n, 1))
    out["c_in"], out["c_in2"] = c, c2
    out["CRTC"] = R.tensorCRTC(c, pe, ruc)
    out["CRTInvC"] = R.tensorCRTInvC(c, pe, ruci, T.mhat_inv_c(m))
    for nm in ("LC", "LInvC", "GPowC", "GDecC"):
        out[nm] = getattr(R, "tensor" + nm)(c, pe)
    out["mulC"] = R.mulC(c, c2)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    return sum(v.nbytes for v in out.values())


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_config_b(R, seed):
    name, m, qs = CONFIG_B
    rng = np.random.default_rng(seed)
    pe = T.pe_array(m)
    n = T.totient_pps(T.factor_pps(m))
    ru, rui = T.ru_tables_zq(m, qs), T.ru_tables_zq(m, qs, inverse=True)
    mh = np.array([T.mhat_inv(m, q) for q in qs], dtype=np.int64)
    y, y2 = zq_input(rng, n, qs), zq_input(rng, n, qs)
    out = {"m": np.int64(m), "qs": np.array(qs, dtype=np.int64), "seed": np.int64(seed),
           "in_digest": np.array(digest(y)), "in2_digest": np.array(digest(y2))}
    crt = R.tensorCRTRq(y, pe, ru, qs)
    out["CRTRq_digest"] = np.array(digest(crt))
    out["CRTInvRq_digest"] = np.array(digest(R.tensorCRTInvRq(y, pe, rui, mh, qs)))
    out["mulRq_digest"] = np.array(digest(R.mulRq(y, y2, qs)))
    # a few spot values to make a mismatch debuggable
    out["CRTRq_head"] = crt[:8].copy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    R = cpu.reference()
    total = 0
    for i, (name, m, qs) in enumerate(SMALL_CASES):
        total += make_small(R, name, m, qs, seed=1000 + i)
        print("golden", name)
    assert list(CONFIG_B[2]) == [q for q, _ in zip(T.good_qs(65536, 2 ** 29), range(4))]
    make_config_b(R, seed=2000)
    print("golden", CONFIG_B[0], f"(raw {total / 1e6:.1f} MB before compression)")


if __name__ == "__main__":
    main()
