#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on B200: CRT + CRT^-1 ring elements per second at
m = 14400 = 64*9*25 (n = 3840), Z_q with q = 14401, batched over 65536 synthetic ring elements
per GPU (BASELINE.json configs[1]), as an absolute number and as a fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: lolb_tensorCRTRq then lolb_tensorCRTInvRq,
in place, on B ring elements resident in HBM.  value = (N*B) / max-over-ranks step time.  Ring
elements are independent, so N GPUs shard the batch with no collective on the data path (weak
scaling: B per GPU is fixed); NCCL is used for the barrier and the max-reduction of the timing only.

Keys beyond the base contract: `roofline` (dominant kernel vs measured HBM copy bandwidth),
`cpu_baseline` (the reference lol-cpp C++ -- oracle/_ref -- or the C restatement on the host
cores, one process per core because the reference is not thread-safe: `static Zq::q`, types.h:59),
`e2e` (same step through the host-buffer C-ABI call with pinned host memory, H2D and D2H inside
the timed region), `per_op` (other operators of the path, device-resident).

`--impl reference` times the reference's own CPU implementation of the same step on the host
cores and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M, QS = 14400, [14401]
N_COEFF = 3840
BATCH_PER_GPU = 65536
BYTES_PER_ELEM = 16 * N_COEFF * len(QS)          # in-place transform: 8 B read + 8 B written per coefficient
METRIC = "CRT+CRTInv ring elems/sec at m=14400"
UNIT = "ring_elems/s"
FALLBACK_HBM_GBS = 6650.0                         # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index: int, period_s: float = 0.02):
        self.index, self.period = index, period_s
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _names(self, mask):
        nv = self.nv
        table = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        return {k for k, bit in table.items() if mask & bit}

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.reasons |= self._names(int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------ CPU arm
def _cpu_worker(kind: str, pairs: int, seed: int):
    import numpy as np
    from oracle import cpu, tables as T
    lib = cpu.reference() if kind == "reference" else cpu.restatement()
    pe = T.pe_array(M)
    ru, rui = T.ru_tables_zq(M, QS), T.ru_tables_zq(M, QS, inverse=True)
    mh = [T.mhat_inv(M, q) for q in QS]
    rng = np.random.default_rng(seed)
    elems = rng.integers(0, QS[0], size=(256, N_COEFF, 1)).astype(np.int64)     # same distribution as the GPU batch
    # warm-up + correctness of the sample itself
    assert np.array_equal(lib.tensorCRTInvRq(lib.tensorCRTRq(elems[0], pe, ru, QS), pe, rui, mh, QS), elems[0])
    t0 = time.perf_counter()
    for i in range(pairs):
        y = lib.tensorCRTRq(elems[i & 255], pe, ru, QS)
        lib.tensorCRTInvRq(y, pe, rui, mh, QS)
    return time.perf_counter() - t0


def cpu_arm(pairs_per_core: int, cores: int | None = None):
    """One process per core (fork): elems/s = cores*pairs / slowest process."""
    import multiprocessing as mp
    from oracle import cpu
    kind = "reference" if cpu.have_reference() else "port"
    if kind == "port":
        cpu.restatement()
    cores = cores or len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        times = pool.starmap(_cpu_worker, [(kind, pairs_per_core, 1000 + c) for c in range(cores)])
    slowest = max(times)
    return {"value": cores * pairs_per_core / slowest, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{pairs_per_core} CRT+CRTInv pairs per process on {cores} processes (one per core), "
                      f"256 distinct uniform ring elements each, m=14400 q=14401; slowest process {slowest:.2f} s",
            "per_core": pairs_per_core / (sum(times) / len(times))}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # one "step" = a bounded sample of the workload on every host core
    pairs = 384
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_arm(32)
    t0 = time.perf_counter()
    results = [cpu_arm(pairs) for _ in range(max(1, min(args.steps, 3)))]
    wall = time.perf_counter() - t0
    best = max(results, key=lambda r: r["value"])
    line = {
        "impl": "reference", "metric": METRIC, "value": best["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(results), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * wall / len(results),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "configs[1]: m=14400 (n=3840), Zq 14401, CRT then CRTInv per ring element", "host_cores": best["cores"]},
        "cpu_baseline": best,
        "e2e": {"value": best["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------ GPU arm
def time_op(torch, fn, iters):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def ext_section(torch, capi, peak, stream):
    """SURVEY.md section 8f ranks 2-3, device-resident and informational: the ring-extension gathers of O_14400 / O_576 and
    the coefficient-wise maps at the configs[3] moduli (same shapes and byte accounting as tools/run_ext.py,
    DESIGN.md 4.7 / 4.8).  Never lets a failure reach the headline line: the caller records the error text instead."""
    from lol_b200.extension import CudaExtension
    from lol_b200.tensor import CudaTensorRq
    m, m2, qs, Be = 576, 14400, [1008001, 1065601], 16384
    lo, hi = CudaTensorRq(m, qs), CudaTensorRq(m2, qs)
    ext = CudaExtension(lo, hi)
    k, phi, phi2 = ext.k, ext.phi, ext.phi2
    q = torch.tensor(qs, device="cuda", dtype=torch.int64)
    xe = torch.randint(0, 2**40, (Be, phi, k), device="cuda", dtype=torch.int64) % q
    ye = torch.randint(0, 2**40, (Be, phi2, k), device="cuda", dtype=torch.int64) % q
    ox, oy, oi = torch.empty_like(xe), torch.empty_like(ye), torch.empty_like(ye)
    od = torch.empty((Be, phi2, k - 1), device="cuda", dtype=torch.int64)
    ee = torch.randn((Be, phi2, k), device="cuda", dtype=torch.float64) * 1e4
    P = hi.plan
    op = lambda name, src, dst: (lambda: capi.check(ext.ext.op(name, capi.RING_RQ, src.data_ptr(), dst.data_ptr(), Be, stream)))
    ops = {  # name: (launch, 8-byte words read + written per ring element)
        "twacePowDec (100 MB working set: L2-resident)": (op("twacePowDec", ye, ox), 2 * phi * k),
        "embedPow": (op("embedPow", xe, oy), (phi + phi2) * k), "embedDec": (op("embedDec", xe, oy), (phi + phi2) * k),
        "embedCRT": (op("embedCRT", xe, oy), (phi + phi2) * k), "coeffsPowDec": (op("coeffsPowDec", ye, oy), 2 * phi2 * k),
        "twaceCRT": (op("twaceCRT", ye, ox), (phi2 + phi) * k),
        "liftRq": (lambda: capi.check(P.lift(ye.data_ptr(), oi.data_ptr(), Be, stream)), 2 * phi2 * k),
        "reduceRq": (lambda: capi.check(P.reduce(oi.data_ptr(), k, oy.data_ptr(), Be, stream)), 2 * phi2 * k),
        "rescaleDropRq": (lambda: capi.check(P.rescale_drop(0, ye.data_ptr(), od.data_ptr(), Be, stream)), phi2 * (2 * k - 1)),
        "rescaleModRq": (lambda: capi.check(P.rescale_mod(qs[::-1], ye.data_ptr(), oi.data_ptr(), Be, stream)), 2 * phi2 * k),
        "roundCosetRq": (lambda: capi.check(P.round_coset(ee.data_ptr(), ye.data_ptr(), oi.data_ptr(), Be, stream)), 3 * phi2 * k),
    }
    res = {"workload": f"m={m} | m'={m2}, q=(1008001,1065601), {Be} elements of O_m' (1 GiB) resident in HBM"}
    for name, (fn, words) in ops.items():
        fn(); fn(); fn()
        ms = time_op(torch, fn, 10)
        gbs = words * 8 * Be / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "elems_per_s": Be / (ms * 1e-3), "GB/s": gbs, "frac": gbs / peak}
    return res


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from lol_b200 import build_library, capi
    from lol_b200.tensor import CudaTensorRq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libctensor_b200 has no CPU path")
    torch.cuda.set_device(local)
    if rank == 0:
        build_library()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    assert capi.device_available()

    B = args.batch
    t = CudaTensorRq(M, QS)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(rank)
    x = torch.randint(0, QS[0], (B, t.n, 1), dtype=torch.int64, device="cuda", generator=gen)
    x0 = x[:4].clone()
    stream = int(torch.cuda.current_stream().cuda_stream)
    ptr = x.data_ptr()

    def step():
        capi.check(t.plan.op("CRT", ptr, B, stream))
        capi.check(t.plan.op("CRTInv", ptr, B, stream))

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    assert torch.equal(x[:4], x0), "CRTInv . CRT != id"

    K = args.steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = capi.kernel_launch_count()
    with ClockSampler(local) as clk:
        start.record()
        for i in range(K):
            evs[i][0].record()
            capi.check(t.plan.op("CRT", ptr, B, stream))
            evs[i][1].record()
            capi.check(t.plan.op("CRTInv", ptr, B, stream))
            evs[i][2].record()
        end.record()
        torch.cuda.synchronize()
    launches = capi.kernel_launch_count() - launches0
    if world > 1:
        dist.barrier()
    ms_total = start.elapsed_time(end)
    ms_t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_step = float(ms_t.item()) / K
    assert torch.equal(x[:4], x0)
    crt_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / K
    inv_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / K

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, copies inside the timed region);
    # every rank drives its own GPU from its own host buffer, time = max over ranks
    e2e = None
    if not args.no_e2e:
        Be = min(B, args.e2e_batch)
        h = torch.empty(Be, t.n, 1, dtype=torch.int64).pin_memory()
        h.copy_(x[:Be])
        h0 = h[:2].clone()
        t.apply_host("CRT,CRTInv", h)      # warm-up (allocates staging, creates streams)
        ksteps = max(3, min(K, 10))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            capi.check(t.plan.apply_host("CRT,CRTInv", h.data_ptr(), Be))
        dt = (time.perf_counter() - t0) / ksteps
        dt_t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
        dt = float(dt_t.item())
        assert torch.equal(h[:2], h0)
        e2e = {"value": world * Be / dt, "unit": UNIT, "h2d_bytes_per_step": world * Be * N_COEFF * 8,
               "d2h_bytes_per_step": world * Be * N_COEFF * 8, "batch_per_gpu": Be, "ms_per_step": dt * 1e3,
               "api": "lolb_rq_apply_host(plan, \"CRT,CRTInv\", host_ptr, batch) per rank"}
        del h

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    dom_name, dom_ms = ("tensorCRTInvRq", inv_ms) if inv_ms >= crt_ms else ("tensorCRTRq", crt_ms)
    alg_bytes = BYTES_PER_ELEM * B
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    traffic = recorded_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic or {}).get(dom_name), "kernel": dom_name + " [" + t.plan.kernel_name("CRTInv" if dom_name.endswith("InvRq") else "CRT") + "]",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "ms_per_launch": {"tensorCRTRq": crt_ms, "tensorCRTInvRq": inv_ms},
                "frac_per_kernel": {"tensorCRTRq": alg_bytes / (crt_ms * 1e-3) / 1e9 / peak,
                                    "tensorCRTInvRq": alg_bytes / (inv_ms * 1e-3) / 1e9 / peak}}

    # ---- other operators of the path, device-resident (not part of the headline value)
    per_op = {}
    if not args.no_per_op:
        y2 = torch.randint(0, QS[0], (B, t.n, 1), dtype=torch.int64, device="cuda", generator=gen)
        ops = {"L": (lambda: t.plan.op("L", ptr, B, stream), 16), "LInv": (lambda: t.plan.op("LInv", ptr, B, stream), 16),
               "GPow": (lambda: t.plan.op("GPow", ptr, B, stream), 16), "GDec": (lambda: t.plan.op("GDec", ptr, B, stream), 16),
               "GInvPow": (lambda: t.plan.op("GInvPow", ptr, B, stream), 16), "GInvDec": (lambda: t.plan.op("GInvDec", ptr, B, stream), 16),
               "mulRq": (lambda: t.plan.mul(ptr, y2.data_ptr(), B, B, stream), 24),
               "CRTMul": (lambda: t.plan.crt_mul(ptr, y2.data_ptr(), B, B, stream), 24),          # y <- CRT(y) . b, one pass
               "MulCRTInv": (lambda: t.plan.mul_crt_inv(ptr, y2.data_ptr(), B, B, stream), 24)}   # y <- CRTInv(y . b), one pass
        for name, (fn, bpc) in ops.items():
            fn(); fn()
            ms = time_op(torch, fn, 10)
            gbs = bpc * N_COEFF * B / (ms * 1e-3) / 1e9
            per_op[name] = {"ms": ms, "elems_per_s": B / (ms * 1e-3), "GB/s": gbs, "frac": gbs / peak, "kernel": t.plan.kernel_name(name)}
        del y2

    # ---- the other BASELINE.json configurations, device-resident (informational; parity for each is in tests/)
    other = {}
    if not args.no_per_op and world == 1:
        def timed(fn, iters=5):
            fn(); fn()
            return time_op(torch, fn, iters)

        def rq_config(m, qs, Bc, label):
            tc = CudaTensorRq(m, qs)
            kc = len(qs)
            xc = torch.cat([torch.randint(0, q, (Bc, tc.n, 1), dtype=torch.int64, device="cuda", generator=gen) for q in qs], dim=2).contiguous()
            res = {"m": m, "qs": qs, "batch": Bc, "bytes_per_elem_per_transform": 16 * tc.n * kc}
            for name in ("CRT", "CRTInv"):
                ms = timed(lambda: capi.check(tc.plan.op(name, xc.data_ptr(), Bc, stream)))
                gbs = 16 * tc.n * kc * Bc / (ms * 1e-3) / 1e9
                res[name] = {"ms": ms, "elems_per_s": Bc / (ms * 1e-3), "GB/s": gbs, "frac": gbs / peak, "kernel": tc.plan.kernel_name(name)}
            xb = xc.clone()
            ms = timed(lambda: capi.check(tc.plan.mul(xc.data_ptr(), xb.data_ptr(), Bc, Bc, stream)))
            gbs = 24 * tc.n * kc * Bc / (ms * 1e-3) / 1e9
            res["mulRq"] = {"ms": ms, "elems_per_s": Bc / (ms * 1e-3), "GB/s": gbs, "frac": gbs / peak, "kernel": tc.plan.kernel_name("mulRq")}
            other[label] = res

        del x
        torch.cuda.empty_cache()
        rq_config(65536, [537133057, 537591809, 537722881, 538116097], 1024, "configs[2]: m=2^16, four ~30-bit primes")
        rq_config(14400, [1008001, 1065601], 32768, "configs[3] moduli: m=14400, q=(1008001,1065601) (SymmSHE key-switch modulus)")
        rq_config(2048, [12289], 131072, "the reference's own benchmark parameters (lol Benchmarks/Default.hs:41-46): m=2^11, q=12289")
        from lol_b200.tensor import CudaTensorComplex, CudaTensorInt, CudaTensorReal
        Bg = 32768
        tr, ti, tcx = CudaTensorReal(M), CudaTensorInt(M), CudaTensorComplex(M)
        dg = torch.randn(Bg, tr.n, 1, dtype=torch.float64, device="cuda", generator=gen)
        zg = torch.randint(-8, 9, (Bg, tr.n, 1), dtype=torch.int64, device="cuda", generator=gen)
        og = torch.empty(Bg, 1, dtype=torch.int64, device="cuda")
        res = {"m": M, "batch": Bg}
        ms = timed(lambda: capi.check(tr.plan.op("GaussianDec", dg.data_ptr(), Bg, stream)))
        res["tensorGaussianDec"] = {"ms": ms, "elems_per_s": Bg / (ms * 1e-3), "frac": 16 * tr.n * Bg / (ms * 1e-3) / 1e9 / peak}
        ms = timed(lambda: capi.check(ti.plan.normsq("R", zg.data_ptr(), og.data_ptr(), Bg, stream)))
        res["tensorNormSqR"] = {"ms": ms, "elems_per_s": Bg / (ms * 1e-3), "frac": 8 * tr.n * Bg / (ms * 1e-3) / 1e9 / peak}
        cg = torch.randn(Bg, tr.n, 1, dtype=torch.complex128, device="cuda")
        for name in ("CRTC", "CRTInvC"):
            ms = timed(lambda: capi.check(tcx.plan.op(name, cg.data_ptr(), Bg, stream)))
            res["tensor" + name] = {"ms": ms, "elems_per_s": Bg / (ms * 1e-3), "frac": 32 * tr.n * Bg / (ms * 1e-3) / 1e9 / peak,
                                    "kernel": tcx.plan.kernel_name(name)}
        other["configs[4]: m=14400 tensorGaussianDec + tensorNormSqR (double / int64), complex CRT"] = res
        del dg, zg, cg
        x = torch.randint(0, QS[0], (B, t.n, 1), dtype=torch.int64, device="cuda", generator=gen)

    # ---- configs[3]: SymmSHE ciphertext multiply + quadratic key switch (SymmSHE.hs:443-449 then :359-372; op sequence of
    # SURVEY.md section 3.5 with TrivGad over the two limbs, l = 2: 4 CRT, tensor product with mulG, CRTInv, decompose,
    # l CRT, knapsack -- (22 + 4l) x 8nk algorithmic bytes per pair).  Measured on one GPU only: ranks other than 0 have
    # left by now, and ciphertext pairs shard exactly like the ring elements of the headline step (no collective).
    she_res = None
    if not args.no_per_op and world == 1:
        from lol_b200.symmshe import CudaSymmSHE
        del x
        torch.cuda.empty_cache()
        qs_c, Bs = [1008001, 1065601], args.she_pairs
        she = CudaSymmSHE(M, qs_c, gad_base=0)
        cts = [torch.cat([torch.randint(0, q, (Bs, she.n, 1), dtype=torch.int64, device="cuda", generator=gen) for q in qs_c], dim=2).contiguous()
               for _ in range(4)]
        hint = torch.cat([torch.randint(0, q, (she.ell, 2, she.n, 1), dtype=torch.int64, device="cuda", generator=gen) for q in qs_c], dim=3).contiguous()

        def she_step():
            she.mulAndSwitch(cts[:2], cts[2:], hint, basis="pow", inplace=True)

        for _ in range(3):
            she_step()
        l0 = capi.kernel_launch_count()
        she_step()
        she_launches = capi.kernel_launch_count() - l0
        ms = time_op(torch, she_step, 10)
        elem = 8 * she.n * she.k
        alg = (22 + 4 * she.ell) * elem
        she_res = {"workload": f"configs[3]: m=14400, q=(1008001,1065601), TrivGad (l={she.ell}), {Bs} ciphertext pairs per GPU, "
                               "Pow-basis inputs, in place", "ms": ms, "ct_pairs_per_s": Bs / (ms * 1e-3),
                   "algorithmic_bytes_per_pair": alg, "GB/s_per_gpu": alg * Bs / (ms * 1e-3) / 1e9,
                   "frac": alg * Bs / (ms * 1e-3) / 1e9 / peak, "kernel_launches_per_step": int(she_launches)}
        if True:
            d3 = she.mulCT(cts[:2], cts[2:], basis="crt")
            dg = she.decompose(cts[0])
            steps = {"ct_mul": (lambda: capi.check(she.t.plan.ct_mul(*[c.data_ptr() for c in cts], *[d.data_ptr() for d in d3], Bs, True, stream)), 7),
                     "decompose": (lambda: capi.check(she.t.plan.decompose(cts[0].data_ptr(), dg.data_ptr(), Bs, 0, stream)), 1 + she.ell),
                     "knapsack": (lambda: capi.check(she.t.plan.knapsack(dg.data_ptr(), she.ell, hint.data_ptr(), d3[0].data_ptr(), d3[1].data_ptr(), Bs, stream)), she.ell + 4)}
            for name, (fn, passes) in steps.items():
                fn(); fn()
                sms = time_op(torch, fn, 10)
                she_res[name] = {"ms": sms, "GB/s": passes * elem * Bs / (sms * 1e-3) / 1e9, "frac": passes * elem * Bs / (sms * 1e-3) / 1e9 / peak}
            del d3, dg
            if not args.no_cpu:      # the same sequence on one host core: compiled reference CRTs + numpy for the Haskell-side steps
                import time as _time
                import numpy as np
                from oracle import cpu as ocpu, symmshe as osym, tables as T
                lib = ocpu.reference() if ocpu.have_reference() else ocpu.restatement()
                pe = T.pe_array(M)
                tabs = (pe, T.ru_tables_zq(M, qs_c), T.ru_tables_zq(M, qs_c, True), [T.mhat_inv(M, q) for q in qs_c], T.g_crt_vectors(M, qs_c)[0])
                rng = np.random.default_rng(0)
                mk = lambda: np.stack([rng.integers(0, q, size=she.n) for q in qs_c], axis=-1).astype(np.int64)
                hint_h = hint.cpu().numpy()
                sample = 8
                t0 = _time.perf_counter()
                for _ in range(sample):
                    osym.mul_and_switch(lib, [mk(), mk()], [mk(), mk()], hint_h, tabs, qs_c, 0)
                she_res["cpu_baseline"] = {"value": sample / (_time.perf_counter() - t0), "unit": "ct_pairs/s", "cores": 1,
                                           "kind": "reference" if ocpu.have_reference() else "port",
                                           "sample": f"{sample} ciphertext pairs, reference CRTs + numpy host steps"}
        del cts, hint
        torch.cuda.empty_cache()
        x = torch.randint(0, QS[0], (B, t.n, 1), dtype=torch.int64, device="cuda", generator=gen)
    if she_res is not None:
        other["configs[3]: SymmSHE ciphertext multiply + key switch"] = she_res
    if not args.no_per_op and world == 1:
        try:
            other["ring extensions + coefficient-wise maps (SURVEY 8f ranks 2-3)"] = ext_section(torch, capi, peak, stream)
        except Exception as exc:      # informational section: never take the headline line down with it
            other["ring extensions + coefficient-wise maps (SURVEY 8f ranks 2-3)"] = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

    cpu_base = None
    if world == 1 and not args.no_cpu:
        cpu_base = cpu_arm(args.cpu_pairs)

    line = {
        "metric": METRIC, "value": world * B / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 modular (int64 ABI)", "data": "synthetic",
        "config": {"workload": "configs[1]: m=14400=64*9*25 (n=3840), Zq 14401, CRT then CRTInv in place, "
                               f"{B} uniform ring elements per GPU resident in HBM",
                   "batch_per_gpu": B, "l2": "inputs (1.9 GB per GPU) larger than the 126 MB L2; no flush needed",
                   "parallelism": f"batch sharded over {world} GPU(s), no data-path collective"},
        "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clk.summary(), "per_op": per_op, "other_configs": other,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--e2e-batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--cpu-pairs", type=int, default=4096, help="CRT+CRTInv pairs per host process in the cpu_baseline sample")
    ap.add_argument("--she-pairs", type=int, default=4096, help="ciphertext pairs per GPU in the configs[3] section")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-per-op", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
