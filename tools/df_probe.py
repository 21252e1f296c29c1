"""Probe one (e, k, B, op) of the pow2 dataflow kernel against the generic engine (hang-safe under `timeout`)."""
import sys, torch, numpy as np
sys.path.insert(0, ".")
from lol_b200.tensor import CudaTensorRq
e, k, B, op = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
qs = [537133057, 537591809, 537722881, 538116097][:k]
t = CudaTensorRq(2 ** e, qs)
x = torch.cat([torch.randint(0, q, (B, t.n, 1), dtype=torch.int64, device="cuda") for q in qs], dim=2).contiguous()
print("start", e, k, B, op, t.plan.kernel_name(op), flush=True)
f = t.crt(x) if op == "CRT" else t.crtInv(x)
torch.cuda.synchronize()
print("ran", flush=True)
t.plan.force_generic(True)
g = t.crt(x) if op == "CRT" else t.crtInv(x)
print("equal", bool(torch.equal(f, g)), flush=True)
if not torch.equal(f, g):
    bad = (f != g)
    per_el = bad.reshape(B, -1).sum(dim=1)
    print("bad elements:", [(i, int(c)) for i, c in enumerate(per_el.tolist()) if c][:24])
    i = int(torch.nonzero(per_el)[0])
    pos = torch.nonzero(bad[i].any(dim=1)).flatten()
    print("element", i, "bad positions", pos.numel(), "first", pos[:8].tolist(), "last", pos[-4:].tolist(), "limbs", bad[i].any(dim=0).tolist())
    ch = torch.unique(pos // 1024); print("bad chunks", ch.tolist()[:40])
