"""ORACLE (test infrastructure): ctypes front end for the two CPU implementations.

  * `restatement()` -> oracle/liblol_oracle.so   (plain-C restatement, symbols `lo_*`)
  * `reference()`   -> oracle/_ref/libctensor_ref.so (UNMODIFIED lol-cpp, built by oracle/Makefile)

Both expose the same Python methods (named after the reference's C symbols) over
numpy arrays in the reference's ABI layout: `y[j*k + limb]`, int64 / float64 /
complex128.  All methods work on a copy and return it.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liblol_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libctensor_ref.so")

_i16, _i32, _p = C.c_int16, C.c_int32, C.c_void_p


def build(target: str = "oracle") -> None:
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_p)


def _ptr_array(tables):
    arr = (_p * len(tables))(*[t.ctypes.data for t in tables])
    return arr


class CpuLib:
    """Uniform wrapper; `prefix` is 'lo_' for the restatement and '' for the reference."""

    def __init__(self, path: str, prefix: str, kind: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        self.kind = kind

    def _fn(self, name, restype=None):
        f = getattr(self.lib, self.prefix + name)
        f.restype = restype
        return f

    @staticmethod
    def _args(y, pe, k):
        pe = np.ascontiguousarray(pe, dtype=np.int16).reshape(-1, 2)
        totm = y.size // k
        return pe, totm

    # ---- generic "basic args" call: f(k, y, totm, pe, npe [, qs]) -----------------
    def _basic(self, name, y, pe, k, dtype, qs=None, status=False):
        y = np.array(y, dtype=dtype, order="C", copy=True)
        pe, totm = self._args(y, pe, k)
        f = self._fn(name, _i16 if status else None)
        args = [_i16(k), _ptr(y), _i32(totm), _ptr(pe), _i16(len(pe))]
        if qs is not None:
            qarr = np.ascontiguousarray(qs, dtype=np.int64)
            args.append(_ptr(qarr))
        ret = f(*args)
        return (y, int(ret)) if status else y

    # ---- CRT ---------------------------------------------------------------------
    def tensorCRTRq(self, y, pe, ru, qs):
        k = len(qs)
        y = np.array(y, dtype=np.int64, order="C", copy=True)
        pe, totm = self._args(y, pe, k)
        ru = [np.ascontiguousarray(t, dtype=np.int64) for t in ru]
        qarr = np.ascontiguousarray(qs, dtype=np.int64)
        self._fn("tensorCRTRq")(_i16(k), _ptr(y), _i32(totm), _ptr(pe), _i16(len(pe)), _ptr_array(ru), _ptr(qarr))
        return y

    def tensorCRTInvRq(self, y, pe, ruinv, mhatinv, qs):
        k = len(qs)
        y = np.array(y, dtype=np.int64, order="C", copy=True)
        pe, totm = self._args(y, pe, k)
        ruinv = [np.ascontiguousarray(t, dtype=np.int64) for t in ruinv]
        qarr = np.ascontiguousarray(qs, dtype=np.int64)
        mh = np.ascontiguousarray(mhatinv, dtype=np.int64)
        self._fn("tensorCRTInvRq")(_i16(k), _ptr(y), _i32(totm), _ptr(pe), _i16(len(pe)), _ptr_array(ruinv), _ptr(mh), _ptr(qarr))
        return y

    def tensorCRTC(self, y, pe, ru, k=1):
        y = np.array(y, dtype=np.complex128, order="C", copy=True)
        pe, totm = self._args(y, pe, k)
        ru = [np.ascontiguousarray(t, dtype=np.complex128) for t in ru]
        self._fn("tensorCRTC")(_i16(k), _ptr(y), _i32(totm), _ptr(pe), _i16(len(pe)), _ptr_array(ru))
        return y

    def tensorCRTInvC(self, y, pe, ruinv, mhatinv, k=1):
        y = np.array(y, dtype=np.complex128, order="C", copy=True)
        pe, totm = self._args(y, pe, k)
        ruinv = [np.ascontiguousarray(t, dtype=np.complex128) for t in ruinv]
        mh = np.ascontiguousarray(mhatinv, dtype=np.complex128)
        self._fn("tensorCRTInvC")(_i16(k), _ptr(y), _i32(totm), _ptr(pe), _i16(len(pe)), _ptr_array(ruinv), _ptr(mh))
        return y

    def tensorGaussianDec(self, y, pe, ru, k=1):
        y = np.array(y, dtype=np.float64, order="C", copy=True)
        pe, totm = self._args(y, pe, k)
        ru = [np.ascontiguousarray(t, dtype=np.complex128) for t in ru]
        self._fn("tensorGaussianDec")(_i16(k), _ptr(y), _i32(totm), _ptr(pe), _i16(len(pe)), _ptr_array(ru))
        return y

    # ---- L / G / norm --------------------------------------------------------------
    def tensorLRq(self, y, pe, qs): return self._basic("tensorLRq", y, pe, len(qs), np.int64, qs)
    def tensorLInvRq(self, y, pe, qs): return self._basic("tensorLInvRq", y, pe, len(qs), np.int64, qs)
    def tensorGPowRq(self, y, pe, qs): return self._basic("tensorGPowRq", y, pe, len(qs), np.int64, qs)
    def tensorGDecRq(self, y, pe, qs): return self._basic("tensorGDecRq", y, pe, len(qs), np.int64, qs)
    def tensorGInvPowRq(self, y, pe, qs): return self._basic("tensorGInvPowRq", y, pe, len(qs), np.int64, qs, status=True)
    def tensorGInvDecRq(self, y, pe, qs): return self._basic("tensorGInvDecRq", y, pe, len(qs), np.int64, qs, status=True)

    def tensorLR(self, y, pe, k=1): return self._basic("tensorLR", y, pe, k, np.int64)
    def tensorLInvR(self, y, pe, k=1): return self._basic("tensorLInvR", y, pe, k, np.int64)
    def tensorGPowR(self, y, pe, k=1): return self._basic("tensorGPowR", y, pe, k, np.int64)
    def tensorGDecR(self, y, pe, k=1): return self._basic("tensorGDecR", y, pe, k, np.int64)
    def tensorGInvPowR(self, y, pe, k=1): return self._basic("tensorGInvPowR", y, pe, k, np.int64, status=True)
    def tensorGInvDecR(self, y, pe, k=1): return self._basic("tensorGInvDecR", y, pe, k, np.int64, status=True)
    def tensorNormSqR(self, y, pe, k=1): return self._basic("tensorNormSqR", y, pe, k, np.int64)

    def tensorLDouble(self, y, pe, k=1): return self._basic("tensorLDouble", y, pe, k, np.float64)
    def tensorLInvDouble(self, y, pe, k=1): return self._basic("tensorLInvDouble", y, pe, k, np.float64)
    def tensorNormSqD(self, y, pe, k=1): return self._basic("tensorNormSqD", y, pe, k, np.float64)

    def tensorLC(self, y, pe, k=1): return self._basic("tensorLC", y, pe, k, np.complex128)
    def tensorLInvC(self, y, pe, k=1): return self._basic("tensorLInvC", y, pe, k, np.complex128)
    def tensorGPowC(self, y, pe, k=1): return self._basic("tensorGPowC", y, pe, k, np.complex128)
    def tensorGDecC(self, y, pe, k=1): return self._basic("tensorGDecC", y, pe, k, np.complex128)
    def tensorGInvPowC(self, y, pe, k=1): return self._basic("tensorGInvPowC", y, pe, k, np.complex128, status=True)
    def tensorGInvDecC(self, y, pe, k=1): return self._basic("tensorGInvDecC", y, pe, k, np.complex128, status=True)

    # ---- pointwise -----------------------------------------------------------------
    def mulRq(self, a, b, qs):
        k = len(qs)
        a = np.array(a, dtype=np.int64, order="C", copy=True)
        b = np.ascontiguousarray(b, dtype=np.int64)
        qarr = np.ascontiguousarray(qs, dtype=np.int64)
        self._fn("mulRq")(_i16(k), _ptr(a), _ptr(b), _i32(a.size // k), _ptr(qarr))
        return a

    def mulC(self, a, b, k=1):
        a = np.array(a, dtype=np.complex128, order="C", copy=True)
        b = np.ascontiguousarray(b, dtype=np.complex128)
        self._fn("mulC")(_i16(k), _ptr(a), _ptr(b), _i32(a.size // k))
        return a


def restatement() -> CpuLib:
    if not os.path.exists(ORACLE_SO):
        build("oracle")
    return CpuLib(ORACLE_SO, "lo_", "port")


def have_reference() -> bool:
    return os.path.exists(REF_SO)


def reference() -> CpuLib:
    """The compiled, unmodified lol-cpp.  Built from /root/reference when present;
    on the GPU box only the prebuilt oracle/_ref/libctensor_ref.so exists."""
    if not os.path.exists(REF_SO):
        build("ref")
    if not os.path.exists(REF_SO):
        raise FileNotFoundError(REF_SO)
    return CpuLib(REF_SO, "", "reference")
