"""Small calls of every new kernel family for compute-sanitizer (memcheck / racecheck): tiny batches, checked against the generic engine."""
import sys, torch
sys.path.insert(0, ".")
from lol_b200.tensor import CudaTensorRq, CudaTensorComplex
Q = [537133057, 537591809, 537722881, 538116097]
cases = [(2 ** 10, Q[:1], 9), (2 ** 11, Q[:2], 5), (2 ** 11, Q[:4], 3), (2 ** 12, Q[:1], 5), (2 ** 12, Q[:4], 3), (2 ** 13, Q[:2], 3),
         (2 ** 14, Q[:1], 3), (2 ** 14, Q[:4], 5), (2 ** 16, Q[:4], 3), (1728, [3457], 7), (42, [8191], 33), (14400, [14401], 5),
         (14400, [1008001, 1065601], 3)]
for m, qs, B in cases:
    t = CudaTensorRq(m, qs)
    x = torch.cat([torch.randint(0, q, (B, t.n, 1), dtype=torch.int64, device="cuda") for q in qs], dim=2).contiguous()
    f, g = t.crt(x), t.crtInv(x)
    ok = torch.equal(t.crtInv(f), x)
    if m == 14400:
        f2 = t.crtMul(x, x.clone()); g2 = t.mulCrtInv(x, x.clone())
    t.plan.force_generic(True)
    ok = ok and torch.equal(t.crt(x), f) and torch.equal(t.crtInv(x), g)
    print(m, len(qs), t.plan.kernel_name("CRT"), "ok" if ok else "MISMATCH", flush=True)
tc = CudaTensorComplex(14400)
c = torch.randn(3, tc.n, 1, dtype=torch.complex128, device="cuda")
print("complex", float((tc.crtInv(tc.crt(c)) - c).abs().max()))
tc = CudaTensorComplex(1728)
c = torch.randn(3, tc.n, 1, dtype=torch.complex128, device="cuda")
print("complex axis", float((tc.crtInv(tc.crt(c)) - c).abs().max()))
torch.cuda.synchronize()
