"""Which kernel family serves each operator on each ring of the reference's benchmark lists (needs a GPU: plans are device objects)."""
import sys
sys.path.insert(0, ".")
from lol_b200.tensor import CudaTensorRq, CudaTensorComplex
rings = [(1024, [12289]), (2048, [12289]), (1728, [3457]), (5184, [10369]), (14400, [14401]), (14400, [1008001, 1065601]), (65536, [537133057, 537591809, 537722881, 538116097]),
         (728, [8737]), (2912, [8737]), (3640, [14561]), (11648, [23297]), (5824, [3144961]), (5460, [3144961]), (4095, [3144961]), (448, [3144961]), (128, [3144961]),
         (2048, [1017857, 1032193]), (5824, [25159681, 19918081, 19393921, 18869761])]
ops = ["CRT", "CRTInv", "L", "GPow", "GInvDec", "mulRq", "CRTMul"]
for m, qs in rings:
    t = CudaTensorRq(m, qs)
    print(m, len(qs), {o: t.plan.kernel_name(o) for o in ops})
for m in (14400, 1728, 2912, 11648, 2048):
    t = CudaTensorComplex(m)
    print(m, "complex", {o: t.plan.kernel_name(o) for o in ("CRTC", "CRTInvC", "LC", "GPowC")})
from lol_b200.tensor import CudaTensorReal
for m in (14400, 1728, 2912, 3640, 5460, 4095, 11648, 2048):
    t = CudaTensorReal(m)
    print(m, "double", {o: t.plan.kernel_name(o) for o in ("GaussianDec", "LDouble")})
