"""Time the fused CRT x product pair at config A.  usage: run_mul.py [batch]"""
import sys, torch
sys.path.insert(0, ".")
from lol_b200.tensor import CudaTensorRq
from lol_b200 import capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
t = CudaTensorRq(14400, [14401])
x = torch.randint(0, 14401, (B, t.n, 1), dtype=torch.int64, device="cuda")
b = torch.randint(0, 14401, (B, t.n, 1), dtype=torch.int64, device="cuda")
st = int(torch.cuda.current_stream().cuda_stream)
def timeit(fn, iters=10):
    for _ in range(3): capi.check(fn())
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): capi.check(fn())
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / iters
for rep in range(3):
    for name, fn, bpc in (("CRT", lambda: t.plan.op("CRT", x.data_ptr(), B, st), 16),
                          ("CRTMul", lambda: t.plan.crt_mul(x.data_ptr(), b.data_ptr(), B, B, st), 24),
                          ("CRTMul(b broadcast)", lambda: t.plan.crt_mul(x.data_ptr(), b.data_ptr(), B, 1, st), 16),
                          ("MulCRTInv", lambda: t.plan.mul_crt_inv(x.data_ptr(), b.data_ptr(), B, B, st), 24)):
        ms = timeit(fn)
        print(rep, name, "ms", round(ms, 4), "frac", round(bpc * t.n * B / ms / 1e6 / 6555.8, 4))
