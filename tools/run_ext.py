"""Time the ring-extension operators (ext_stream.cu) on device-resident batches; prints one JSON line per operator.
usage: run_ext.py [m] [m'] [batch of O_m' elements] [iters] [q1,q2,..]      (default moduli: config C's pair)
Algorithmic bytes per element: words read once + words written, 8 k bytes each (tables excluded, L1/L2 resident)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from lol_b200 import capi
from lol_b200.extension import CudaExtension
from lol_b200.tensor import CudaTensorRq

m = int(sys.argv[1]) if len(sys.argv) > 1 else 576
m2 = int(sys.argv[2]) if len(sys.argv) > 2 else 14400
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
qs = [int(v) for v in sys.argv[5].split(",")] if len(sys.argv) > 5 else [1008001, 1065601]
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
lo, hi = CudaTensorRq(m, qs), CudaTensorRq(m2, qs)
ext = CudaExtension(lo, hi)
k, phi, phi2 = ext.k, ext.phi, ext.phi2
q = torch.tensor(qs, device="cuda", dtype=torch.int64)
x = torch.randint(0, 2**40, (B, phi, k), device="cuda", dtype=torch.int64) % q
y = torch.randint(0, 2**40, (B, phi2, k), device="cuda", dtype=torch.int64) % q
ox, oy = torch.empty_like(x), torch.empty_like(y)
st = int(torch.cuda.current_stream().cuda_stream)
dec_src = int((ext.ext.table(capi.EXT_BASE_DEC) >= 0).sum())
ops = {  # name: (src, dst, words moved per element)
    "twacePowDec": (y, ox, 2 * phi), "embedPow": (x, oy, phi + phi2), "embedDec": (x, oy, min(dec_src, phi) + phi2),
    "embedCRT": (x, oy, phi + phi2), "coeffsPowDec": (y, oy, 2 * phi2), "twaceCRT": (y, ox, phi2 + phi),
}
for name, (src, dst, words) in ops.items():
    fn = lambda: capi.check(ext.ext.op(name, capi.RING_RQ, src.data_ptr(), dst.data_ptr(), B, st))
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    gbs = words * 8 * k * B / ms / 1e6
    print(json.dumps({"op": name, "m": m, "m2": m2, "k": k, "batch": B, "ms": round(ms, 4), "elems_per_s": round(B / ms * 1e3),
                      "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3)}))

# coefficient-wise maps (coeff_stream.cu) on the O_m' batch: words = 8-byte words read + written per coefficient word of the output
e = torch.randn((B, phi2, k), device="cuda", dtype=torch.float64) * 1e4
oi = torch.empty_like(y)
od = torch.empty((B, phi2, k - 1), device="cuda", dtype=torch.int64)
P = hi.plan
cops = {
    "liftRq": (lambda: P.lift(y.data_ptr(), oi.data_ptr(), B, st), 2 * phi2 * k),
    "reduceRq": (lambda: P.reduce(oi.data_ptr(), k, oy.data_ptr(), B, st), 2 * phi2 * k),
    "rescaleDropRq": (lambda: P.rescale_drop(0, y.data_ptr(), od.data_ptr(), B, st), phi2 * k + phi2 * (k - 1)),
    "rescaleModRq": (lambda: P.rescale_mod(qs[::-1], y.data_ptr(), oi.data_ptr(), B, st), 2 * phi2 * k),
    "roundCosetRq": (lambda: P.round_coset(e.data_ptr(), y.data_ptr(), oi.data_ptr(), B, st), 3 * phi2 * k),
}
if k == 1:
    del cops["rescaleDropRq"]      # nothing to drop with a single limb
for name, (fn, words) in cops.items():
    for _ in range(3):
        capi.check(fn())
    s, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e2.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e2) / iters
    gbs = words * 8 * B / ms / 1e6
    print(json.dumps({"op": name, "m": m2, "k": k, "batch": B, "ms": round(ms, 4), "elems_per_s": round(B / ms * 1e3),
                      "achieved_gbs": round(gbs, 1), "peak_gbs": peak, "frac": round(gbs / peak, 3)}))
