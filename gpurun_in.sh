timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_crt_mul" 2>&1 | tail -5
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from lol_b200.tensor import CudaTensorRq
from lol_b200 import capi
for m, qs, B in ((14400, [1008001, 1065601], 32768),):
    t = CudaTensorRq(m, qs); k = len(qs)
    x = torch.cat([torch.randint(0, q, (B, t.n, 1), dtype=torch.int64, device="cuda") for q in qs], dim=2).contiguous()
    b = x.clone()
    st = int(torch.cuda.current_stream().cuda_stream)
    def tm(fn, it=10):
        for _ in range(3): capi.check(fn())
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record()
        for _ in range(it): capi.check(fn())
        e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / it
    for name, fn, seq in (("crtMul", lambda: t.plan.crt_mul(x.data_ptr(), b.data_ptr(), B, B, st), ("CRT",)), ("mulCrtInv", lambda: t.plan.mul_crt_inv(x.data_ptr(), b.data_ptr(), B, B, st), ("CRTInv",))):
        ms = tm(fn)
        ms2 = tm(lambda: t.plan.op(seq[0], x.data_ptr(), B, st)) + tm(lambda: t.plan.mul(x.data_ptr(), b.data_ptr(), B, B, st))
        print(m, qs, name, t.plan.kernel_name("CRTMul"), "fused ms %.3f (frac of 24nk-byte roofline %.3f)  separate ms %.3f" % (ms, 24 * t.n * k * B / ms / 1e6 / 6555.8, ms2))
PY
