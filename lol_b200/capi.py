"""ctypes binding of libctensor_b200.so (include/lol_b200.h).

Plain pointers and sizes only: device buffers are passed as integer addresses
(`tensor.data_ptr()`), streams as `cudaStream_t` handles.  Loading the library
needs no GPU; every operator call does and raises `LolB200Error` otherwise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import library_path

_i16, _i32, _i64, _p = C.c_int16, C.c_int32, C.c_int64, C.c_void_p

LOLB_OK = 0
LOLB_ERR_ARG = 1
LOLB_ERR_NO_CRT = 2
LOLB_ERR_CUDA = 3
LOLB_ERR_NOT_INVERTIBLE = 4


class LolB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libctensor_b200 status {status}: {message}")
        self.status = status


_lib = None


def lib() -> C.CDLL:
    """The shared library.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        path = os.environ.get("LOLB_LIBRARY") or library_path()      # override: kernel-tuning builds (tools/build_variant.py)
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} is missing: build it with `python -m lol_b200.build` "
                "(libctensor_b200 has no CPU or pure-Python fallback)")
        _lib = C.CDLL(path)
        _lib.lolb_last_error.restype = C.c_char_p
        _lib.lolb_kernel_launch_count.restype = _i64
        _lib.lolb_plan_gcrt_dev.restype = _p
        _lib.lolb_plan_kernel_name.restype = C.c_char_p
        _lib.lolb_host_alloc.restype = _p
        _lib.lolb_host_alloc.argtypes = [C.c_uint64]
        _lib.lolb_host_free.argtypes = [_p]
        _lib.lolb_ext_index_table.restype = _i64
        _lib.lolb_ext_totient.restype = _i32
    return _lib


def last_error() -> str:
    return lib().lolb_last_error().decode()


def kernel_launch_count() -> int:
    return int(lib().lolb_kernel_launch_count())


def device_available() -> bool:
    return bool(lib().lolb_device_available())


def check(status: int) -> None:
    if status != LOLB_OK:
        raise LolB200Error(status, last_error())


def pe_array(pps) -> np.ndarray:
    """[(p, e), ...] -> PrimeExponent[] (int16 pairs)."""
    return np.ascontiguousarray(np.array(list(pps), dtype=np.int16).reshape(-1, 2))


def _ptr_array(tables):
    return (_p * len(tables))(*[t.ctypes.data for t in tables])


class PlanRq:
    """lolb_plan over Z_q1 x ... x Z_qk (lolb_plan_create_rq)."""

    def __init__(self, pps, qs, ru=None, ruinv=None, mhatinv=None):
        self.pps = [(int(p), int(e)) for p, e in pps]
        self.qs = [int(q) for q in qs]
        self.k = len(self.qs)
        pe = pe_array(self.pps)
        qarr = np.ascontiguousarray(self.qs, dtype=np.int64)
        keep = []
        ru_p = ruinv_p = mh_p = None
        if ru is not None:
            ru = [np.ascontiguousarray(t, dtype=np.int64) for t in ru]
            keep.append(ru)
            ru_p = _ptr_array(ru)
        if ruinv is not None:
            ruinv = [np.ascontiguousarray(t, dtype=np.int64) for t in ruinv]
            mh = np.ascontiguousarray(mhatinv, dtype=np.int64)
            keep += [ruinv, mh]
            ruinv_p, mh_p = _ptr_array(ruinv), mh.ctypes.data_as(_p)
        self._h = _p()
        check(lib().lolb_plan_create_rq(C.byref(self._h), pe.ctypes.data_as(_p), _i16(len(self.pps)), _i16(self.k),
                                        qarr.ctypes.data_as(_p), ru_p, ruinv_p, mh_p))
        self.n = int(lib().lolb_plan_totient(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().lolb_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def force_generic(self, on: bool = True) -> None:
        lib().lolb_plan_set_force_generic(self._h, int(on))

    def kernel_name(self, op: str) -> str:
        return lib().lolb_plan_kernel_name(self._h, op.encode()).decode()

    def ru_table(self, pp_index: int, inverse: bool = False) -> np.ndarray:
        p, e = self.pps[pp_index]
        out = np.empty((p ** e, self.k), dtype=np.int64)
        check(lib().lolb_plan_get_ru_rq(self._h, int(inverse), pp_index, out.ctypes.data_as(_p)))
        return out

    def mhatinv(self) -> np.ndarray:
        out = np.empty(self.k, dtype=np.int64)
        check(lib().lolb_plan_get_mhatinv_rq(self._h, out.ctypes.data_as(_p)))
        return out

    def gcrt_dev(self, inverse: bool = False) -> int:
        return int(lib().lolb_plan_gcrt_dev(self._h, int(inverse)) or 0)

    # device-pointer operators: `ptr` an integer device address, `stream` a cudaStream_t handle (int)
    def op(self, name: str, ptr: int, batch: int, stream: int = 0) -> int:
        f = getattr(lib(), "lolb_tensor" + name + "Rq")
        return int(f(self._h, _p(ptr), _i64(batch), _p(stream)))

    def mul(self, a_ptr: int, b_ptr: int, batch: int, b_batch: int, stream: int = 0) -> int:
        return int(lib().lolb_mulRq(self._h, _p(a_ptr), _p(b_ptr), _i64(batch), _i64(b_batch), _p(stream)))

    def crt_mul(self, y_ptr: int, b_ptr: int, batch: int, b_batch: int, stream: int = 0) -> int:
        return int(lib().lolb_crtMulRq(self._h, _p(y_ptr), _p(b_ptr), _i64(batch), _i64(b_batch), _p(stream)))

    def mul_crt_inv(self, y_ptr: int, b_ptr: int, batch: int, b_batch: int, stream: int = 0) -> int:
        return int(lib().lolb_mulCrtInvRq(self._h, _p(y_ptr), _p(b_ptr), _i64(batch), _i64(b_batch), _p(stream)))

    # SymmSHE steps between the CRTs (lolb_ctMulRq / lolb_decomposeRq / lolb_knapsackRq)
    def ct_mul(self, a0: int, a1: int, b0: int, b1: int, d0: int, d1: int, d2: int, batch: int, mul_g: bool = True, stream: int = 0) -> int:
        return int(lib().lolb_ctMulRq(self._h, _p(a0), _p(a1), _p(b0), _p(b1), _p(d0), _p(d1), _p(d2), _i64(batch),
                                      C.c_int(int(mul_g)), _p(stream)))

    def gadget_length(self, base: int = 0) -> int:
        ell = int(lib().lolb_gadgetLength(self._h, _i64(base)))
        if ell < 0:
            raise LolB200Error(LOLB_ERR_ARG, last_error())
        return ell

    def decompose(self, x: int, digits: int, batch: int, base: int = 0, stream: int = 0) -> int:
        return int(lib().lolb_decomposeRq(self._h, _p(x), _p(digits), _i64(batch), _i64(base), _p(stream)))

    def decompose_crt(self, x: int, digits: int, batch: int, base: int = 0, stream: int = 0) -> int:
        return int(lib().lolb_decomposeCrtRq(self._h, _p(x), _p(digits), _i64(batch), _i64(base), _p(stream)))

    def knapsack(self, digits: int, ell: int, hints: int, c0: int, c1: int, batch: int, stream: int = 0) -> int:
        return int(lib().lolb_knapsackRq(self._h, _p(digits), C.c_int(ell), _p(hints), _p(c0), _p(c1), _i64(batch), _p(stream)))

    # coefficient-wise maps (coeff_stream.cu)
    def lift(self, x: int, y: int, batch: int, stream: int = 0) -> int:
        return int(lib().lolb_liftRq(self._h, _p(x), _p(y), _i64(batch), _p(stream)))

    def reduce(self, z: int, z_tupsize: int, y: int, batch: int, stream: int = 0) -> int:
        return int(lib().lolb_reduceRq(self._h, _p(z), C.c_int(z_tupsize), _p(y), _i64(batch), _p(stream)))

    def rescale_drop(self, drop: int, x: int, y: int, batch: int, stream: int = 0) -> int:
        return int(lib().lolb_rescaleDropRq(self._h, C.c_int(drop), _p(x), _p(y), _i64(batch), _p(stream)))

    def rescale_mod(self, qs_new, x: int, y: int, batch: int, stream: int = 0) -> int:
        qn = np.ascontiguousarray(qs_new, dtype=np.int64)
        return int(lib().lolb_rescaleModRq(self._h, qn.ctypes.data_as(_p), _p(x), _p(y), _i64(batch), _p(stream)))

    def round_coset(self, e: int, zp: int, y: int, batch: int, stream: int = 0) -> int:
        return int(lib().lolb_roundCosetRq(self._h, _p(e), _p(zp) if zp else None, _p(y), _i64(batch), _p(stream)))

    def apply_host(self, ops: str, host_ptr: int, batch: int) -> int:
        return int(lib().lolb_rq_apply_host(self._h, ops.encode(), _p(host_ptr), _i64(batch)))

    def apply_host_u32(self, ops: str, host_ptr: int, batch: int) -> int:
        """lolb_rq_apply_host_u32: the host pipeline over uint32 residues (half the PCIe bytes); ops = "" copies only."""
        return int(lib().lolb_rq_apply_host_u32(self._h, ops.encode(), _p(host_ptr), _i64(batch)))


class PlanC:
    """lolb_plan for the modulus-free rings: int64 'R', double, complex (lolb_plan_create_c)."""

    def __init__(self, pps, k: int = 1):
        self.pps = [(int(p), int(e)) for p, e in pps]
        self.k = int(k)
        pe = pe_array(self.pps)
        self._h = _p()
        check(lib().lolb_plan_create_c(C.byref(self._h), pe.ctypes.data_as(_p), _i16(len(self.pps)), _i16(self.k)))
        self.n = int(lib().lolb_plan_totient(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().lolb_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def force_generic(self, on: bool = True) -> None:
        lib().lolb_plan_set_force_generic(self._h, int(on))

    def kernel_name(self, op: str) -> str:
        return lib().lolb_plan_kernel_name(self._h, op.encode()).decode()

    def op(self, name: str, ptr: int, batch: int, stream: int = 0) -> int:
        """name: 'LR', 'LInvDouble', 'GPowC', 'CRTC', 'CRTInvC', 'GaussianDec', 'GInvPowC', ..."""
        f = getattr(lib(), "lolb_tensor" + name)
        return int(f(self._h, _p(ptr), _i64(batch), _p(stream)))

    def t_gaussian_dec(self, v: float, seed: int, first: int, ptr: int, batch: int, stream: int = 0) -> int:
        return int(lib().lolb_tGaussianDec(self._h, C.c_double(v), C.c_uint64(seed), C.c_uint64(first), _p(ptr), _i64(batch), _p(stream)))

    def ginv_r(self, name: str, ptr: int, ok_ptr: int, batch: int, stream: int = 0) -> int:
        f = getattr(lib(), "lolb_tensor" + name + "R")
        return int(f(self._h, _p(ptr), _p(ok_ptr), _i64(batch), _p(stream)))

    def normsq(self, tag: str, ptr: int, out_ptr: int, batch: int, stream: int = 0) -> int:
        f = getattr(lib(), "lolb_tensorNormSq" + tag)
        return int(f(self._h, _p(ptr), _p(out_ptr), _i64(batch), _p(stream)))

    def mul(self, a_ptr: int, b_ptr: int, batch: int, b_batch: int, stream: int = 0) -> int:
        return int(lib().lolb_mulC(self._h, _p(a_ptr), _p(b_ptr), _i64(batch), _i64(b_batch), _p(stream)))


# ------------------------------------------------------------------ ring extensions O_m'/O_m (lolb_ext_*)
RING_RQ, RING_R, RING_DOUBLE, RING_C = 0, 1, 2, 3
EXT_INDICES_POWDEC, EXT_INDICES_CRT, EXT_BASE_POW_J0, EXT_BASE_POW_J1, EXT_BASE_DEC, EXT_INDICES_COEFFS = range(6)


def dev_alloc(nbytes: int) -> int:
    lib().lolb_dev_alloc.restype = _p
    lib().lolb_dev_alloc.argtypes = [C.c_uint64]
    p = lib().lolb_dev_alloc(nbytes)
    if not p:
        raise LolB200Error(LOLB_ERR_CUDA, last_error())
    return int(p)


def dev_free(ptr: int) -> None:
    lib().lolb_dev_free.argtypes = [_p]
    lib().lolb_dev_free(_p(ptr))


def dev_upload(dst: int, src: np.ndarray) -> None:
    check(lib().lolb_dev_upload(_p(dst), src.ctypes.data_as(_p), C.c_uint64(src.nbytes), _p(0)))


def dev_download(dst: np.ndarray, src: int) -> None:
    check(lib().lolb_dev_download(dst.ctypes.data_as(_p), _p(src), C.c_uint64(dst.nbytes), _p(0)))


def dev_copy(dst: int, src: int, nbytes: int) -> None:
    check(lib().lolb_dev_copy(_p(dst), _p(src), C.c_uint64(nbytes), _p(0)))


def real_gaussians(svar: float, seed: int, first: int, ptr: int, n: int, batch: int, stream: int = 0) -> int:
    """lolb_realGaussians: [batch][n] doubles on the device, i.i.d. N(0, svar / (2 pi))."""
    return int(lib().lolb_realGaussians(C.c_double(svar), C.c_uint64(seed), C.c_uint64(first), _p(ptr), _i64(n), _i64(batch), _p(stream)))


def fused_w_emulate(pps, qs, y: np.ndarray, inverse: bool = False) -> np.ndarray:
    """lolb_fused_w_emulate: the fused_w schedule (host-built constants, the kernel's line code compiled for the host, a
    lane-by-lane replica of the exchange network) on ONE ring element [n][k] in host memory.  Needs no GPU: CPU test hook."""
    pe = pe_array(pps)
    q = np.ascontiguousarray(qs, dtype=np.int64)
    out = np.ascontiguousarray(y, dtype=np.int64).copy()
    check(lib().lolb_fused_w_emulate(pe.ctypes.data_as(_p), _i16(len(pe)), _i16(len(q)), q.ctypes.data_as(_p), C.c_int(int(inverse)),
                                     out.ctypes.data_as(_p)))
    return out


def fused_w_emulate_c(pps, y: np.ndarray, inverse: bool = False) -> np.ndarray:
    """lolb_fused_w_emulate_c: the same device-free replica over complex doubles; y = [n][k] complex128."""
    pe = pe_array(pps)
    out = np.ascontiguousarray(y, dtype=np.complex128).copy()
    k = 1 if out.ndim == 1 else out.shape[-1]
    check(lib().lolb_fused_w_emulate_c(pe.ctypes.data_as(_p), _i16(len(pe)), _i16(k), C.c_int(int(inverse)), out.ctypes.data_as(_p)))
    return out


def ext_index_table(pps, pps2, which: int) -> np.ndarray:
    """lolb_ext_index_table: one table of Tensor.hs:429-478 computed on the host (needs no GPU)."""
    pe, pe2 = pe_array(pps), pe_array(pps2)
    args = (pe.ctypes.data_as(_p), _i16(len(pe)), pe2.ctypes.data_as(_p), _i16(len(pe2)), C.c_int(which))
    count = int(lib().lolb_ext_index_table(*args, None))
    if count < 0:
        raise LolB200Error(LOLB_ERR_ARG, last_error())
    out = np.empty(count, dtype=np.int32)
    lib().lolb_ext_index_table(*args, out.ctypes.data_as(_p))
    return out


class Extension:
    """lolb_ext over the plans of O_m (`lo`) and O_m' (`hi`); the plans are kept alive by this object."""

    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi
        self._h = _p()
        check(lib().lolb_ext_create(C.byref(self._h), lo.handle, hi.handle))
        self.phi = int(lib().lolb_ext_totient(self._h, 0))
        self.phi2 = int(lib().lolb_ext_totient(self._h, 1))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().lolb_ext_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def table(self, which: int) -> np.ndarray:
        out = np.empty(self.phi if which == EXT_INDICES_POWDEC else self.phi2, dtype=np.int32)
        check(lib().lolb_ext_get_table(self._h, C.c_int(which), out.ctypes.data_as(_p)))
        return out

    def pow_basis_pow(self, ring: int, y_ptr: int, stream: int = 0) -> int:
        return int(lib().lolb_powBasisPow(self._h, C.c_int(ring), _p(y_ptr), _p(stream)))

    def op(self, name: str, ring: int, x_ptr: int, y_ptr: int, batch: int, stream: int = 0) -> int:
        """name: 'twacePowDec', 'embedPow', 'embedDec', 'embedCRT', 'coeffsPowDec', 'twaceCRT'."""
        f = getattr(lib(), "lolb_" + name)
        return int(f(self._h, C.c_int(ring), _p(x_ptr), _p(y_ptr), _i64(batch), _p(stream)))


# ------------------------------------------------------------------ drop-in symbols over numpy (host pointers)
class DropIn:
    """The 29 reference symbols exactly as Backend.hs imports them, driven with numpy host
    arrays (one element per call, in place on a copy)."""

    def _basic(self, name, y, pe, k, dtype, qs=None, status=False):
        y = np.array(y, dtype=dtype, order="C", copy=True)
        pe = pe_array(pe)
        totm = y.size // k
        f = getattr(lib(), name)
        f.restype = _i16 if status else None
        args = [_i16(k), y.ctypes.data_as(_p), _i32(totm), pe.ctypes.data_as(_p), _i16(len(pe))]
        if qs is not None:
            qarr = np.ascontiguousarray(qs, dtype=np.int64)
            args.append(qarr.ctypes.data_as(_p))
        ret = f(*args)
        return (y, int(ret)) if status else y

    def tensorCRTRq(self, y, pe, ru, qs):
        k = len(qs)
        y = np.array(y, dtype=np.int64, order="C", copy=True)
        pe = pe_array(pe)
        ru = [np.ascontiguousarray(t, dtype=np.int64) for t in ru]
        qarr = np.ascontiguousarray(qs, dtype=np.int64)
        f = lib().tensorCRTRq
        f.restype = None
        f(_i16(k), y.ctypes.data_as(_p), _i32(y.size // k), pe.ctypes.data_as(_p), _i16(len(pe)), _ptr_array(ru), qarr.ctypes.data_as(_p))
        return y

    def tensorCRTInvRq(self, y, pe, ruinv, mhatinv, qs):
        k = len(qs)
        y = np.array(y, dtype=np.int64, order="C", copy=True)
        pe = pe_array(pe)
        ruinv = [np.ascontiguousarray(t, dtype=np.int64) for t in ruinv]
        mh = np.ascontiguousarray(mhatinv, dtype=np.int64)
        qarr = np.ascontiguousarray(qs, dtype=np.int64)
        f = lib().tensorCRTInvRq
        f.restype = None
        f(_i16(k), y.ctypes.data_as(_p), _i32(y.size // k), pe.ctypes.data_as(_p), _i16(len(pe)), _ptr_array(ruinv),
          mh.ctypes.data_as(_p), qarr.ctypes.data_as(_p))
        return y

    def tensorCRTC(self, y, pe, ru, k=1):
        y = np.array(y, dtype=np.complex128, order="C", copy=True)
        pe = pe_array(pe)
        ru = [np.ascontiguousarray(t, dtype=np.complex128) for t in ru]
        f = lib().tensorCRTC
        f.restype = None
        f(_i16(k), y.ctypes.data_as(_p), _i32(y.size // k), pe.ctypes.data_as(_p), _i16(len(pe)), _ptr_array(ru))
        return y

    def tensorCRTInvC(self, y, pe, ruinv, mhatinv, k=1):
        y = np.array(y, dtype=np.complex128, order="C", copy=True)
        pe = pe_array(pe)
        ruinv = [np.ascontiguousarray(t, dtype=np.complex128) for t in ruinv]
        mh = np.ascontiguousarray(mhatinv, dtype=np.complex128)
        f = lib().tensorCRTInvC
        f.restype = None
        f(_i16(k), y.ctypes.data_as(_p), _i32(y.size // k), pe.ctypes.data_as(_p), _i16(len(pe)), _ptr_array(ruinv), mh.ctypes.data_as(_p))
        return y

    def tensorGaussianDec(self, y, pe, ru, k=1):
        y = np.array(y, dtype=np.float64, order="C", copy=True)
        pe = pe_array(pe)
        ru = [np.ascontiguousarray(t, dtype=np.complex128) for t in ru]
        f = lib().tensorGaussianDec
        f.restype = None
        f(_i16(k), y.ctypes.data_as(_p), _i32(y.size // k), pe.ctypes.data_as(_p), _i16(len(pe)), _ptr_array(ru))
        return y

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

    def mulRq(self, a, b, qs):
        k = len(qs)
        a = np.array(a, dtype=np.int64, order="C", copy=True)
        b = np.ascontiguousarray(b, dtype=np.int64)
        qarr = np.ascontiguousarray(qs, dtype=np.int64)
        f = lib().mulRq
        f.restype = None
        f(_i16(k), a.ctypes.data_as(_p), b.ctypes.data_as(_p), _i32(a.size // k), qarr.ctypes.data_as(_p))
        return a

    def mulC(self, a, b, k=1):
        a = np.array(a, dtype=np.complex128, order="C", copy=True)
        b = np.ascontiguousarray(b, dtype=np.complex128)
        f = lib().mulC
        f.restype = None
        f(_i16(k), a.ctypes.data_as(_p), b.ctypes.data_as(_p), _i32(a.size // k))
        return a
