"""Time the modulus-free operators (config 5: complex CRT, Gaussian, norms) at m, batch."""
import sys, torch
sys.path.insert(0, ".")
from lol_b200.tensor import CudaTensorComplex, CudaTensorInt, CudaTensorReal
from lol_b200 import capi
m = int(sys.argv[1]); B = int(sys.argv[2])
tc, tr, ti = CudaTensorComplex(m), CudaTensorReal(m), CudaTensorInt(m)
n = tc.n
c = torch.randn(B, n, 1, dtype=torch.complex128, device="cuda")
d = torch.randn(B, n, 1, dtype=torch.float64, device="cuda")
z = torch.randint(-8, 9, (B, n, 1), dtype=torch.int64, device="cuda")
outd = torch.empty(B, 1, dtype=torch.float64, device="cuda"); outz = torch.empty(B, 1, dtype=torch.int64, device="cuda")
st = int(torch.cuda.current_stream().cuda_stream)
cases = [("CRTC", lambda: tc.plan.op("CRTC", c.data_ptr(), B, st), 32), ("CRTInvC", lambda: tc.plan.op("CRTInvC", c.data_ptr(), B, st), 32),
         ("GaussianDec", lambda: tr.plan.op("GaussianDec", d.data_ptr(), B, st), 16),
         ("NormSqD", lambda: tr.plan.normsq("D", d.data_ptr(), outd.data_ptr(), B, st), 8),
         ("NormSqR", lambda: ti.plan.normsq("R", z.data_ptr(), outz.data_ptr(), B, st), 8),
         ("LDouble", lambda: tr.plan.op("LDouble", d.data_ptr(), B, st), 16), ("GPowC", lambda: tc.plan.op("GPowC", c.data_ptr(), B, st), 32),
         ("LR", lambda: ti.plan.op("LR", z.data_ptr(), B, st), 16),
         ("tGaussianDec", lambda: tr.plan.t_gaussian_dec(0.1, 1, 0, d.data_ptr(), B, st), 8)]
print(m, "kernels:", {o: tr.plan.kernel_name(o) for o in ("GaussianDec", "LDouble")}, {o: tc.plan.kernel_name(o) for o in ("CRTC", "GPowC")})
for name, fn, bpc in cases:
    for _ in range(2): capi.check(fn())
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(20): capi.check(fn())
    e.record(); torch.cuda.synchronize(); ms = s.elapsed_time(e) / 20
    print(f"{name:12s} ms {ms:8.3f}  elems/s {B/ms*1e3:12.0f}  frac {bpc*n*B/ms/1e6/6555.8:.4f}")
