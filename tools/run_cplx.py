"""Time tensorCRTC / tensorCRTInvC on a batch (quick timing).  usage: run_cplx.py m batch [iters]"""
import sys, torch
sys.path.insert(0, ".")
from lol_b200.tensor import CudaTensorComplex
from lol_b200 import capi
m = int(sys.argv[1]); B = int(sys.argv[2]); iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
t = CudaTensorComplex(m)
x = torch.randn(B, t.n, 1, dtype=torch.complex128, device="cuda")
st = int(torch.cuda.current_stream().cuda_stream)
for name in ("CRTC", "CRTInvC"):
    for _ in range(3): capi.check(t.plan.op(name, x.data_ptr(), B, st))
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): capi.check(t.plan.op(name, x.data_ptr(), B, st))
    e.record(); torch.cuda.synchronize(); ms = s.elapsed_time(e) / iters
    print(m, name, t.plan.kernel_name(name), "ms", round(ms, 4), "frac", round(32 * t.n * B / ms / 1e6 / 6555.8, 4))
