"""CPU suite: the C-ABI library loads without a GPU, exports every symbol include/lol_b200.h
declares (and the 29 names Backend.hs:304-337 imports), and fails loudly -- not silently on a CPU
path -- when asked to compute without a device."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lol_b200.h")

# foreign imports of lol-cpp/Crypto/Lol/Cyclotomic/Tensor/CPP/Backend.hs:304-337
BACKEND_HS_IMPORTS = [
    "tensorLR", "tensorLInvR", "tensorLRq", "tensorLInvRq", "tensorLDouble", "tensorLInvDouble", "tensorLC", "tensorLInvC",
    "tensorNormSqR", "tensorNormSqD",
    "tensorGPowR", "tensorGPowRq", "tensorGPowC", "tensorGDecR", "tensorGDecRq", "tensorGDecC",
    "tensorGInvPowR", "tensorGInvPowRq", "tensorGInvPowC", "tensorGInvDecR", "tensorGInvDecRq", "tensorGInvDecC",
    "tensorCRTRq", "tensorCRTC", "tensorCRTInvRq", "tensorCRTInvC",
    "tensorGaussianDec", "mulRq", "mulC",
]


@pytest.fixture(scope="module")
def libpath():
    from lol_b200 import build_library
    return build_library()


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text)
    skip = {"defined", "sizeof"}
    return sorted({n for n in names if n not in skip and (n.startswith("lolb_") or n.startswith("tensor") or n in ("mulRq", "mulC"))})


def test_header_declares_all_backend_imports():
    assert len(BACKEND_HS_IMPORTS) == 29
    decl = set(declared_functions())
    assert set(BACKEND_HS_IMPORTS) <= decl


def test_library_exports_every_declared_symbol(libpath):
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [n for n in declared_functions() if n not in exported]
    assert not missing, missing


def test_library_is_sm100a_only(libpath):
    out = subprocess.run(["cuobjdump", "-lelf", libpath], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_library_does_not_link_oracle(libpath):
    out = subprocess.run(["nm", "-D", libpath], capture_output=True, text=True, check=True).stdout
    assert "lo_tensor" not in out and "lo_mul" not in out
    src = os.path.join(ROOT, "lol_b200")
    for dirpath, _, files in os.walk(src):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                body = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in body, os.path.join(dirpath, f)


def test_loads_and_reports_without_gpu(libpath):
    from lol_b200 import capi
    assert capi.kernel_launch_count() >= 0
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device behaviour is not observable here")
    assert not capi.device_available()
    with pytest.raises(capi.LolB200Error) as ei:
        capi.PlanRq([(2, 6), (3, 2), (5, 2)], [14401])
    assert ei.value.status == capi.LOLB_ERR_CUDA


def test_dropin_aborts_without_gpu(libpath):
    """The drop-in symbols are `void`: with no device they must die loudly (types.h:36-41 ASSERT style),
    never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    code = (
        "import numpy as np, sys; sys.path.insert(0, %r)\n"
        "from lol_b200 import capi\n"
        "y = capi.DropIn().tensorLRq(np.arange(6, dtype=np.int64), [(7, 1)], [29])\n"
        "print('SURVIVED', y)\n" % ROOT)
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert r.returncode != 0 and "SURVIVED" not in r.stdout
    assert "no CPU path" in r.stderr


def test_bad_arguments_rejected(libpath):
    from lol_b200 import capi
    with pytest.raises(capi.LolB200Error) as ei:
        capi.PlanRq([(3, 1), (2, 2)], [13])           # primes must be increasing (ppsFact order)
    assert ei.value.status == capi.LOLB_ERR_ARG
    with pytest.raises(capi.LolB200Error) as ei:
        capi.PlanRq([(2, 2)], [1 << 33])              # modulus beyond what int64 products allow (types.h:79-84)
    assert ei.value.status == capi.LOLB_ERR_ARG


def test_factored_matches_oracle_tables():
    from lol_b200 import factored
    from oracle import tables as T
    for m in (1, 2, 3, 4, 6, 7, 8, 12, 21, 42, 89, 1024, 1728, 5184, 14400, 65536, 7 * 13 * 4):
        assert factored.pps_fact(m) == T.factor_pps(m)
        assert factored.totient_fact(m) == T.totient_pps(T.factor_pps(m))
        assert factored.value_hat(m) == T.value_hat(m)
        assert factored.radical_fact(m) == T.radical(m)
        assert factored.odd_radical_fact(m) == T.odd_radical(m)


def test_header_is_plain_c99(tmp_path):
    """The boundary must be bindable from a C / Haskell FFI: include/lol_b200.h compiles as strict C99 on its own."""
    src = tmp_path / "h.c"
    src.write_text('#include "lol_b200.h"\nint main(void) { return LOLB_EXT_TABLES + LOLB_RING_C + LOLB_ERR_NO_CRT; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_extension_and_coefficient_entry_points_reject_null_handles(libpath):
    """No device needed: the argument checks of the §8f entry points come before any CUDA call."""
    import ctypes as C
    from lol_b200 import capi
    L = capi.lib()
    h = C.c_void_p()
    assert L.lolb_ext_create(C.byref(h), None, None) == capi.LOLB_ERR_ARG and not h.value
    for name in ("lolb_twacePowDec", "lolb_embedPow", "lolb_embedDec", "lolb_embedCRT", "lolb_coeffsPowDec", "lolb_twaceCRT"):
        assert getattr(L, name)(None, C.c_int(0), None, None, C.c_int64(1), None) == capi.LOLB_ERR_ARG, name
    assert L.lolb_liftRq(None, None, None, C.c_int64(1), None) == capi.LOLB_ERR_ARG
    assert L.lolb_reduceRq(None, None, C.c_int(1), None, C.c_int64(1), None) == capi.LOLB_ERR_ARG
    assert L.lolb_rescaleDropRq(None, C.c_int(0), None, None, C.c_int64(1), None) == capi.LOLB_ERR_ARG
    assert L.lolb_rescaleModRq(None, None, None, None, C.c_int64(1), None) == capi.LOLB_ERR_ARG
    assert L.lolb_roundCosetRq(None, None, None, None, C.c_int64(1), None) == capi.LOLB_ERR_ARG
    assert "plan" in capi.last_error() or "NULL" in capi.last_error()
    L.lolb_ext_destroy(None)                                   # destroying nothing is a no-op
    assert capi.lib().lolb_ext_totient(None, 0) == 0


def test_extension_mirror_refuses_cpu_tensors_and_maps_no_crt(monkeypatch):
    """Host logic of lol_b200/extension.py without a device: CPU tensors are refused loudly (no CPU path), the ring tag
    follows the tensor class, and LOLB_ERR_NO_CRT becomes the reference's `Nothing` (crtExtFuncs, Tensor.hs:176-180)."""
    import torch
    from lol_b200 import capi, extension as E
    from lol_b200.tensor import CudaTensorComplex, CudaTensorInt, CudaTensorRq

    class FakeExt:
        phi, phi2 = 2, 6

        def __init__(self, lo, hi):
            self.calls = []

        def op(self, name, ring, x, y, batch, stream=0):
            self.calls.append((name, ring, batch))
            return capi.LOLB_ERR_NO_CRT if name in ("embedCRT", "twaceCRT") else capi.LOLB_OK

    monkeypatch.setattr(capi, "Extension", FakeExt)
    monkeypatch.setattr(capi, "last_error", lambda: "fake")

    def bare(cls, k):
        t = object.__new__(cls)
        t.k, t.plan = k, None
        return t

    for cls, ring, dtype in ((CudaTensorRq, capi.RING_RQ, torch.int64), (CudaTensorInt, capi.RING_R, torch.int64),
                             (CudaTensorComplex, capi.RING_C, torch.complex128)):
        ext = E.CudaExtension(bare(cls, 1), bare(cls, 1))
        assert (ext.ring, ext.dtype, ext.phi, ext.phi2) == (ring, dtype, 2, 6)
        with pytest.raises(capi.LolB200Error) as ei:
            ext.embedPow(torch.zeros((1, 2, 1), dtype=dtype))          # a CPU tensor
        assert ei.value.status == capi.LOLB_ERR_ARG and "no CPU path" in str(ei.value)
    with pytest.raises(capi.LolB200Error):
        E.CudaExtension(bare(CudaTensorRq, 1), bare(CudaTensorInt, 1))     # different rings
    # status mapping, with the device check out of the way
    monkeypatch.setattr(E, "_require_cuda", lambda x, dtype, n, k: int(x.shape[0]))
    monkeypatch.setattr(E, "_stream", lambda: 0)
    ext = E.CudaExtension(bare(CudaTensorRq, 1), bare(CudaTensorRq, 1))
    x, y = torch.zeros((3, 2, 1), dtype=torch.int64), torch.zeros((3, 6, 1), dtype=torch.int64)
    assert ext.embedCRT(x) is None and ext.twaceCRT(y) is None
    assert ext.embedPow(x).shape == (3, 6, 1) and ext.twacePowDec(y).shape == (3, 2, 1) and ext.coeffs(y).shape == (3, 3, 2, 1)
    assert [c[0] for c in ext.ext.calls] == ["embedCRT", "twaceCRT", "embedPow", "twacePowDec", "coeffsPowDec"]


def test_coefficient_mirror_shapes_and_tuple_inference(monkeypatch):
    """Host logic of CudaTensorRq.lift / reduce / rescaleDrop / rescaleMod / roundCoset without a device: output shapes,
    the z tuple size handed to lolb_reduceRq, and the NULL coset of errorRounded."""
    import torch
    from lol_b200 import capi, tensor as Tn

    calls = []

    class FakePlan:
        def lift(self, x, y, b, st=0): calls.append(("lift", b)); return 0
        def reduce(self, z, kz, y, b, st=0): calls.append(("reduce", kz, b)); return 0
        def rescale_drop(self, d, x, y, b, st=0): calls.append(("drop", d, b)); return 0
        def rescale_mod(self, qn, x, y, b, st=0): calls.append(("mod", tuple(qn), b)); return 0
        def round_coset(self, e, zp, y, b, st=0): calls.append(("round", bool(zp), b)); return 0

    t = object.__new__(Tn.CudaTensorRq)
    t.n, t.k, t.qs, t.plan = 4, 2, [17, 29], FakePlan()
    with pytest.raises(capi.LolB200Error):
        t.lift(torch.zeros((1, 4, 2), dtype=torch.int64))                   # CPU tensor: no CPU path
    monkeypatch.setattr(Tn, "_require_cuda", lambda x, dtype, n, k: (_ for _ in ()).throw(AssertionError((x.shape, n, k)))
                        if tuple(x.shape[1:]) != (n, k) or x.dtype != dtype else int(x.shape[0]))
    monkeypatch.setattr(Tn, "_stream", lambda: 0)
    x = torch.zeros((3, 4, 2), dtype=torch.int64)
    assert t.lift(x).shape == (3, 4, 2)
    assert t.reduce(torch.zeros((3, 4, 1), dtype=torch.int64)).shape == (3, 4, 2)
    assert t.reduce(x).shape == (3, 4, 2)
    assert t.rescaleDrop(x, 1).shape == (3, 4, 1)
    assert t.rescaleMod(x, [5, 7]).shape == (3, 4, 2)
    e = torch.zeros((3, 4, 2), dtype=torch.float64)
    assert t.roundCoset(e).dtype == torch.int64 and t.roundCoset(e, x).shape == (3, 4, 2)
    assert calls == [("lift", 3), ("reduce", 1, 3), ("reduce", 2, 3), ("drop", 1, 3), ("mod", (5, 7), 3), ("round", False, 3), ("round", True, 3)]
