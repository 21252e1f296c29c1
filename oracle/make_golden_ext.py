"""ORACLE (test infrastructure): generate tests/golden/ext_*.npz for the ring-extension operators.

The CRT-basis vectors are produced BY THE COMPILED REFERENCE (oracle/_ref/libctensor_ref.so) through the reference's own
identities embedCRT = crt . embedPow . crtInv and twaceCRT = crt . twacePowDec . crtInv (TensorTests.hs:145-170), and
embedDec through lInv . embedPow . l -- i.e. every arithmetic step is the reference binary's, only the Pow-basis index
gathers (Tensor.hs:391-498) come from oracle/extension.py.  Run in the build container:

    python -m oracle.make_golden_ext
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cpu, extension as X, tables as T  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CASES = [("ext_f3_f21_q8191", 3, 21, [8191]), ("ext_f4_f12_smooth1", 4, 12, [2148249601]),
         ("ext_f7_f21_zq2", 7, 21, [19393921, 18869761]), ("ext_f45_f225_q14401", 45, 225, [14401]),
         ("ext_f576_f14400_cfgC", 576, 14400, [1008001, 1065601])]
BATCH = 3


class _Ring:
    def __init__(self, R, m, qs):
        self.R, self.qs, self.pe = R, qs, T.pe_array(m)
        self.ru, self.rui = T.ru_tables_zq(m, qs), T.ru_tables_zq(m, qs, inverse=True)
        self.mh = [T.mhat_inv(m, q) for q in qs]

    def crt(self, v): return self.R.tensorCRTRq(v, self.pe, self.ru, self.qs).reshape(v.shape)
    def crt_inv(self, v): return self.R.tensorCRTInvRq(v, self.pe, self.rui, self.mh, self.qs).reshape(v.shape)
    def l(self, v): return self.R.tensorLRq(v, self.pe, self.qs).reshape(v.shape)
    def l_inv(self, v): return self.R.tensorLInvRq(v, self.pe, self.qs).reshape(v.shape)


def make(R, name, m, m2, qs, seed):
    rng = np.random.default_rng(seed)
    info = X.ExtInfo(m, m2)
    lo, hi = _Ring(R, m, qs), _Ring(R, m2, qs)
    zq = lambda n: np.stack([rng.integers(0, q, size=(BATCH, n)) for q in qs], axis=-1).astype(np.int64)
    x, y = zq(info.phi), zq(info.phi2)
    out = {"m": np.int64(m), "m2": np.int64(m2), "qs": np.array(qs, dtype=np.int64), "seed": np.int64(seed), "x_in": x, "y_in": y}
    out["embedPow"] = np.stack([X.embed_pow(info, v) for v in x])
    out["twacePowDec"] = np.stack([X.twace_powdec(info, v) for v in y])
    out["coeffs"] = np.stack([X.coeffs_powdec(info, v) for v in y])
    out["embedDec"] = np.stack([hi.l_inv(X.embed_pow(info, lo.l(v))) for v in x])
    out["embedCRT"] = np.stack([hi.crt(X.embed_pow(info, lo.crt_inv(v))) for v in x])
    out["twaceCRT"] = np.stack([lo.crt(X.twace_powdec(info, hi.crt_inv(v))) for v in y])
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    return out


def main():
    if not cpu.have_reference():
        cpu.build("ref")
    R = cpu.reference()
    for i, (name, m, m2, qs) in enumerate(CASES):
        make(R, name, m, m2, qs, 1000 + i)
        print("wrote", name)


if __name__ == "__main__":
    main()
