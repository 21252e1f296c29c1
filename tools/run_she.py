"""Time the SymmSHE streaming steps and the whole multiply + key switch (for ncu / quick timing).
usage: run_she.py [pairs] [gad_base] [iters]"""
import sys, torch
sys.path.insert(0, ".")
from lol_b200 import capi
from lol_b200.symmshe import CudaSymmSHE
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
base = int(sys.argv[2]) if len(sys.argv) > 2 else 0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
m, qs = 14400, [1008001, 1065601]
she = CudaSymmSHE(m, qs, gad_base=base)
mk = lambda *shape: torch.cat([torch.randint(0, q, (*shape, she.n, 1), dtype=torch.int64, device="cuda") for q in qs], dim=-1).contiguous()
cts = [mk(B) for _ in range(4)]
hint = mk(she.ell, 2)
st = int(torch.cuda.current_stream().cuda_stream)
d3 = she.mulCT(cts[:2], cts[2:], basis="crt")
dg = she.decompose(cts[0])
elem = 8 * she.n * she.k
steps = {"ct_mul": (lambda: capi.check(she.t.plan.ct_mul(*[c.data_ptr() for c in cts], *[d.data_ptr() for d in d3], B, True, st)), 7),
         "decompose": (lambda: capi.check(she.t.plan.decompose(cts[0].data_ptr(), dg.data_ptr(), B, base, st)), 1 + she.ell),
         "knapsack": (lambda: capi.check(she.t.plan.knapsack(dg.data_ptr(), she.ell, hint.data_ptr(), d3[0].data_ptr(), d3[1].data_ptr(), B, st)), she.ell + 4),
         "mulAndSwitch": (lambda: she.mulAndSwitch(cts[:2], cts[2:], hint, basis="pow", inplace=True), 22 + 4 * she.ell)}
for name, (fn, passes) in steps.items():
    for _ in range(3): fn()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize(); ms = s.elapsed_time(e) / iters
    print(f"ell={she.ell} {name:13s} {ms:.4f} ms  {passes * elem * B / ms / 1e6:.0f} GB/s  frac {passes * elem * B / ms / 1e6 / 6555.8:.3f}")
