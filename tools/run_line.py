"""Time L / LInv / GPow / GDec / GInvPow / GInvDec / mulRq on a batch.  usage: run_line.py m q1,q2,.. batch"""
import sys, torch
sys.path.insert(0, ".")
from lol_b200.tensor import CudaTensorRq
from lol_b200 import capi
m = int(sys.argv[1]); qs = [int(v) for v in sys.argv[2].split(",")]; B = int(sys.argv[3])
t = CudaTensorRq(m, qs); k = len(qs)
x = torch.cat([torch.randint(0, q, (B, t.n, 1), dtype=torch.int64, device="cuda") for q in qs], dim=2).contiguous()
st = int(torch.cuda.current_stream().cuda_stream)
for name in ("L", "LInv", "GPow", "GDec", "GInvPow", "GInvDec"):
    for _ in range(3): capi.check(t.plan.op(name, x.data_ptr(), B, st))
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(10): capi.check(t.plan.op(name, x.data_ptr(), B, st))
    e.record(); torch.cuda.synchronize(); ms = s.elapsed_time(e) / 10
    print(m, qs, name, t.plan.kernel_name(name), "ms", round(ms, 4), "frac", round(16 * t.n * k * B / ms / 1e6 / 6555.8, 4))
