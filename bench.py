#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on B200: CRT + CRT^-1 ring elements per second at
m = 14400 = 64*9*25 (n = 3840), Z_q with q = 14401, batched over 65536 synthetic ring elements
per GPU (BASELINE.json configs[1]), as an absolute number and as a fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: lolb_tensorCRTRq then lolb_tensorCRTInvRq,
in place, on B ring elements resident in HBM.  value = (N*B) / max-over-ranks step time.  Ring
elements are independent, so N GPUs shard the batch with no collective on the data path (weak
scaling: B per GPU is fixed); NCCL is used for the barrier and the max-reductions of the timings only.

Output: ONE JSON line on stdout (the headline).  Its last key, `summary`, holds the roofline
fractions of the other BASELINE.json configurations in compact form (max over ranks at every N);
the verbose per-operator tables go to stderr as a second JSON line prefixed "DETAIL ".

Keys beyond the base contract: `roofline` (dominant kernel vs measured HBM copy bandwidth),
`cpu_baseline` (the reference lol-cpp C++ -- oracle/_ref -- or the C restatement on the host
cores, timed by the C driver oracle/ref_bench, one process per core because the reference is not
thread-safe: `static Zq::q`, types.h:59), `e2e` (same step through the host-buffer C-ABI call with
pinned host memory, H2D and D2H inside the timed region, plus the copy-only ceiling of the same
pipeline and the uint32 wire format).

`--impl reference` times the reference's own CPU implementation of the same step on the host
cores (same metric, unit and config) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import struct
import subprocess
import sys
import tempfile
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
WORKLOAD = ("configs[1]: m=14400=64*9*25 (n=3840), Zq 14401, CRT then CRTInv in place on uniform synthetic ring elements, "
            "65536 per GPU")


def config_dict(world: int) -> dict:
    """The same dictionary in both arms (the reference arm runs a bounded sample of this workload on the host cores)."""
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH_PER_GPU,
            "l2": "inputs (1.9 GB per GPU) larger than the 126 MB L2; no flush needed",
            "parallelism": f"batch sharded over {world} GPU(s), no data-path collective"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """dram bytes per launch of the dominant kernels from the committed ncu --set full capture (profiles/traffic.json,
    which names the capture it came from), or None.  Not measured in this run: ncu cannot run inside a timed bench."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (pynvml), every 2 ms."""

    def __init__(self, index: int, period_s: float = 0.002):
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


# ------------------------------------------------------------------ CPU arm (oracle/ref_bench, a C driver: no interpreter in the loop)
def _write_tables(path: str):
    """What the Haskell side hands to the C code (CPP.hs:422-442), in ref_bench.c's file layout."""
    import numpy as np
    from oracle import tables as T
    pe = T.pe_array(M)
    ru, rui = T.ru_tables_zq(M, QS), T.ru_tables_zq(M, QS, inverse=True)
    mh = np.array([T.mhat_inv(M, q) for q in QS], dtype=np.int64)
    with open(path, "wb") as f:
        f.write(struct.pack("<iii", len(pe), N_COEFF, len(QS)))
        f.write(np.ascontiguousarray(pe, dtype=np.int16).tobytes())
        f.write(np.array(QS, dtype=np.int64).tobytes())
        for tabs in (ru, rui):
            for t in tabs:
                f.write(np.ascontiguousarray(t, dtype=np.int64).tobytes())
        f.write(mh.tobytes())


def cpu_arm(pairs_per_core: int, steps: int = 1, warmup: int = 0, cores: int | None = None):
    """One process per core (fork inside ref_bench).  Returns the cpu_baseline dictionary plus per-step seconds."""
    from oracle import cpu
    kind = "reference" if cpu.have_reference() else "port"
    if kind == "port":
        cpu.restatement()
    exe = os.path.join(ROOT, "oracle", "ref_bench")
    if not os.path.exists(exe):
        cpu.build("bench")
    lib, prefix = (cpu.REF_SO, "") if kind == "reference" else (cpu.ORACLE_SO, "lo_")
    cores = cores or len(os.sched_getaffinity(0))
    with tempfile.TemporaryDirectory() as td:
        tab = os.path.join(td, "tables.bin")
        _write_tables(tab)
        out = subprocess.run([exe, lib, prefix, tab, str(cores), str(pairs_per_core), str(steps), str(warmup)],
                             check=True, capture_output=True, text=True).stdout
    r = json.loads(out)
    secs = r["step_seconds"]
    total = sum(secs)
    return {"value": cores * pairs_per_core * len(secs) / total, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{len(secs)} steps of {pairs_per_core} CRT+CRTInv pairs per process on {cores} processes (one per core, "
                      f"C driver oracle/ref_bench), 256 distinct uniform ring elements each, m=14400 q=14401; "
                      f"step = slowest process, {1e3 * total / len(secs):.1f} ms on average",
            "per_core": pairs_per_core * len(secs) / sum(r["mean_process_seconds"])}, secs


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # one "step" = a bounded sample of the workload: `ref_pairs` CRT+CRTInv pairs on every host core at once
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    base, secs = cpu_arm(args.ref_pairs, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": config_dict(args.gpus),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


def bind_to_gpu_numa(local: int) -> dict:
    """Pin this process (and therefore its first-touch pinned allocations) to the CPUs of the GPU's NUMA node, so that
    N ranks do not stage through one memory controller.  Best effort: containers often hide the topology."""
    info = {"numa_node": None, "cpus": len(os.sched_getaffinity(0))}
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node >= 0:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpus = set()
                for part in f.read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus |= set(range(int(a), int(b or a) + 1))
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info = {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return info


class Recorder:
    """Timings of the informational sections, keyed in a fixed order so that N ranks can max-reduce them in one tensor."""

    def __init__(self):
        self.rows = []      # (section, name, ms, algorithmic bytes per launch, units per launch, extra)

    def add(self, section, name, ms, alg_bytes, units, **extra):
        self.rows.append([section, name, float(ms), float(alg_bytes), float(units), extra])

    def reduce_max(self, torch, dist, world):
        if world > 1 and self.rows:
            t = torch.tensor([r[2] for r in self.rows], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            for r, v in zip(self.rows, t.tolist()):
                r[2] = v

    def table(self, peak, world):
        out = {}
        for section, name, ms, alg, units, extra in self.rows:
            gbs = alg / (ms * 1e-3) / 1e9
            row = {"ms": ms, "per_s_all_gpus": world * units / (ms * 1e-3), "GB/s_per_gpu": gbs, "frac": gbs / peak}
            row.update(extra)
            out.setdefault(section, {})[name] = row
        return out

    def frac(self, peak, section, name):
        for s, n, ms, alg, _, _ in self.rows:
            if s == section and n == name:
                return round(alg / (ms * 1e-3) / 1e9 / peak, 4)
        return None


def ext_section(torch, capi, rec, stream):
    """SURVEY.md section 8f ranks 2-3, device-resident and informational: the ring-extension gathers of O_14400 / O_576 and
    the coefficient-wise maps at the configs[3] moduli (same shapes and byte accounting as tools/run_ext.py,
    DESIGN.md 4.7 / 4.8)."""
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
    sec = f"ring extensions + coefficient-wise maps (SURVEY 8f ranks 2-3): m={m} | m'={m2}, q=(1008001,1065601), {Be} elements of O_m'"
    for name, (fn, words) in ops.items():
        fn(); fn(); fn()
        rec.add(sec, name, time_op(torch, fn, 10), words * 8 * Be, Be)


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
    numa = bind_to_gpu_numa(local)      # before any pinned allocation
    torch.cuda.set_device(local)
    if rank == 0:
        build_library()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    assert capi.device_available()

    def reduce_max(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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

    W = max(args.warmup, 3)
    for _ in range(W):
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
    ms_step = reduce_max(start.elapsed_time(end)) / K
    assert torch.equal(x[:4], x0)
    crt_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / K
    inv_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / K

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, copies inside the timed region);
    # every rank drives its own GPU from its own host buffer, time = max over ranks.  Beside it: the copy-only ceiling of
    # the same three-slot pipeline (ops = "": one cudaMemcpyAsync per chunk each way, no kernel) and the uint32 wire format.
    e2e = None
    if not args.no_e2e:
        Be = min(B, args.e2e_batch)
        ksteps = max(3, min(K, 10))

        def timed_host(call):
            call()                       # warm-up (allocates staging, creates streams)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                call()
            return reduce_max((time.perf_counter() - t0) / ksteps)

        h = torch.empty(Be, t.n, 1, dtype=torch.int64).pin_memory()
        h.copy_(x[:Be])
        h0 = h[:2].clone()
        dt = timed_host(lambda: capi.check(t.plan.apply_host("CRT,CRTInv", h.data_ptr(), Be)))
        assert torch.equal(h[:2], h0)
        dt_copy = timed_host(lambda: capi.check(t.plan.apply_host("", h.data_ptr(), Be)))
        del h
        h32 = torch.empty(Be, t.n, 1, dtype=torch.int32).pin_memory()
        h32.copy_(x[:Be])
        h32_0 = h32[:2].clone()
        dt32 = timed_host(lambda: capi.check(t.plan.apply_host_u32("CRT,CRTInv", h32.data_ptr(), Be)))
        assert torch.equal(h32[:2], h32_0)
        dt32_copy = timed_host(lambda: capi.check(t.plan.apply_host_u32("", h32.data_ptr(), Be)))
        del h32
        wire = world * Be * N_COEFF * 8
        e2e = {"value": world * Be / dt, "unit": UNIT, "h2d_bytes_per_step": wire, "d2h_bytes_per_step": wire,
               "batch_per_gpu": Be, "ms_per_step": dt * 1e3,
               "api": "lolb_rq_apply_host(plan, \"CRT,CRTInv\", host_ptr, batch) per rank",
               "GB/s_each_way": wire / dt / 1e9,
               "copy_only_ceiling": {"value": world * Be / dt_copy, "ms_per_step": dt_copy * 1e3, "GB/s_each_way": wire / dt_copy / 1e9,
                                     "what": "same 3-slot pipeline, same buffers, no kernels (ops = \"\")"},
               "frac_of_copy_ceiling": dt_copy / dt,
               "u32_wire": {"value": world * Be / dt32, "ms_per_step": dt32 * 1e3, "h2d_bytes_per_step": wire // 2,
                            "d2h_bytes_per_step": wire // 2, "GB/s_each_way": wire / 2 / dt32 / 1e9,
                            "frac_of_copy_ceiling": dt32_copy / dt32, "vs_int64_wire": dt / dt32,
                            "api": "lolb_rq_apply_host_u32 (uint32 residues on the host, widened on the device)"},
               "host_binding": numa}

    peak, peak_src = measured_peak()
    rec = Recorder()
    timed = lambda fn, iters=5: (fn(), fn(), time_op(torch, fn, iters))[2]

    # ---- other operators of the path at config A, device-resident (not part of the headline value)
    if not args.no_per_op:
        y2 = torch.randint(0, QS[0], (B, t.n, 1), dtype=torch.int64, device="cuda", generator=gen)
        ops = {"L": (lambda: t.plan.op("L", ptr, B, stream), 16), "LInv": (lambda: t.plan.op("LInv", ptr, B, stream), 16),
               "GPow": (lambda: t.plan.op("GPow", ptr, B, stream), 16), "GDec": (lambda: t.plan.op("GDec", ptr, B, stream), 16),
               "GInvPow": (lambda: t.plan.op("GInvPow", ptr, B, stream), 16), "GInvDec": (lambda: t.plan.op("GInvDec", ptr, B, stream), 16),
               "mulRq": (lambda: t.plan.mul(ptr, y2.data_ptr(), B, B, stream), 24),
               "CRTMul": (lambda: t.plan.crt_mul(ptr, y2.data_ptr(), B, B, stream), 24),          # y <- CRT(y) . b, one pass
               "MulCRTInv": (lambda: t.plan.mul_crt_inv(ptr, y2.data_ptr(), B, B, stream), 24)}   # y <- CRTInv(y . b), one pass
        for name, (fn, bpc) in ops.items():
            call = (lambda f: (lambda: capi.check(f())))(fn)      # a failing operator must not be timed as a no-op
            rec.add("per_op (config A)", name, timed(call, 10), bpc * N_COEFF * B, B, kernel=t.plan.kernel_name(name))
        del y2
    del x
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configurations and the reference's own benchmark rings, device-resident, on every rank
    # (informational; parity for each is in tests/)
    if not args.no_per_op:
        def rq_config(label, m, qs, Bc):
            tc = CudaTensorRq(m, qs)
            kc = len(qs)
            xc = torch.cat([torch.randint(0, q, (Bc, tc.n, 1), dtype=torch.int64, device="cuda", generator=gen) for q in qs], dim=2).contiguous()
            for name in ("CRT", "CRTInv"):
                ms = timed(lambda: capi.check(tc.plan.op(name, xc.data_ptr(), Bc, stream)))
                rec.add(label, name, ms, 16 * tc.n * kc * Bc, Bc, kernel=tc.plan.kernel_name(name), batch=Bc)
            if tc.plan.kernel_name("L") != "identity":      # the linear operators of the path on this ring (one representative of each cost)
                for name in ("L", "GInvDec"):
                    ms = timed(lambda: capi.check(tc.plan.op(name, xc.data_ptr(), Bc, stream)))
                    rec.add(label, name, ms, 16 * tc.n * kc * Bc, Bc, kernel=tc.plan.kernel_name(name), batch=Bc)
            xb = xc.clone()
            ms = timed(lambda: capi.check(tc.plan.mul(xc.data_ptr(), xb.data_ptr(), Bc, Bc, stream)))
            rec.add(label, "mulRq", ms, 24 * tc.n * kc * Bc, Bc, kernel=tc.plan.kernel_name("mulRq"), batch=Bc)
            del xc, xb, tc
            torch.cuda.empty_cache()

        rq_config("cfgB", 65536, [537133057, 537591809, 537722881, 538116097], 1024)       # configs[2]: m=2^16, four ~30-bit primes
        rq_config("cfgC", 14400, [1008001, 1065601], 32768)                                  # configs[3] moduli (SymmSHE key-switch modulus)
        # the reference's own benchmark parameters (lol Benchmarks/Default.hs:41-48), ~1 GiB batches
        rq_config("F2048/12289", 2048, [12289], 131072)
        rq_config("F64*F27/3457", 1728, [3457], 262144)
        rq_config("F64*F81/10369", 5184, [10369], 81920)
        rq_config("F32*F7*F13/8737", 2912, [8737], 122880)
        rq_config("F8*F5*F7*F13/14561", 3640, [14561], 122880)
        rq_config("F128*F7*F13/23297", 11648, [23297], 30720)
        # the ring-tunnelling chain of lol-apps' benchmarks at its modulus (lol-apps Benchmarks/Default.hs:52-82)
        for mt, bt in ((11648, 30720), (5824, 61440), (2912, 122880), (3640, 122880), (5460, 122880), (4095, 81920)):
            rq_config(f"tunnel m={mt}/3144961", mt, [3144961], bt)
        # the HomomPRF example's modulus chains on its rings (lol-apps Examples/HomomPRFParams.hs:24-45): tupSize 2 and 4
        rq_config("HomomPRF H1'=F64*F7*F13 ZQ2", 5824, [19393921, 18869761], 32768)
        rq_config("HomomPRF H1'=F64*F7*F13 ZQ4", 5824, [25159681, 19918081, 19393921, 18869761], 16384)
        from lol_b200.tensor import CudaTensorComplex, CudaTensorInt, CudaTensorReal
        Bg = 32768
        tr, ti, tcx = CudaTensorReal(M), CudaTensorInt(M), CudaTensorComplex(M)
        dg = torch.randn(Bg, tr.n, 1, dtype=torch.float64, device="cuda", generator=gen)
        zg = torch.randint(-8, 9, (Bg, tr.n, 1), dtype=torch.int64, device="cuda", generator=gen)
        og = torch.empty(Bg, 1, dtype=torch.int64, device="cuda")
        sec = "cfg4"      # configs[4]: m=14400 tensorGaussianDec + tensorNormSqR (double / int64), complex CRT
        rec.add(sec, "tensorGaussianDec", timed(lambda: capi.check(tr.plan.op("GaussianDec", dg.data_ptr(), Bg, stream))), 16 * tr.n * Bg, Bg)
        rec.add(sec, "tensorNormSqR", timed(lambda: capi.check(ti.plan.normsq("R", zg.data_ptr(), og.data_ptr(), Bg, stream))), 8 * tr.n * Bg, Bg)
        cg = torch.randn(Bg, tr.n, 1, dtype=torch.complex128, device="cuda")
        for name in ("CRTC", "CRTInvC"):
            ms = timed(lambda: capi.check(tcx.plan.op(name, cg.data_ptr(), Bg, stream)))
            rec.add(sec, "tensor" + name, ms, 32 * tr.n * Bg, Bg, kernel=tcx.plan.kernel_name(name))
        del dg, zg, cg, og
        torch.cuda.empty_cache()
        # the reference's `error` benchmark (tGaussianDec, lol Benchmarks/TensorBenches.hs) and the line operators over doubles on
        # its other rings: transform alone (16 n bytes per element) and draw + transform in one pass (8 n bytes written)
        for mt, bt in ((2912, 57344), (11648, 14336), (5460, 57344)):
            tg = CudaTensorReal(mt)
            dgt = torch.randn(bt, tg.n, 1, dtype=torch.float64, device="cuda", generator=gen)
            sec_g = f"error m={mt}"
            rec.add(sec_g, "tensorGaussianDec", timed(lambda: capi.check(tg.plan.op("GaussianDec", dgt.data_ptr(), bt, stream))), 16 * tg.n * bt, bt,
                    kernel=tg.plan.kernel_name("GaussianDec"))
            dgt.normal_()
            rec.add(sec_g, "tGaussianDec", timed(lambda: capi.check(tg.plan.t_gaussian_dec(0.1, 1, 0, dgt.data_ptr(), bt, stream))), 8 * tg.n * bt, bt,
                    kernel=tg.plan.kernel_name("GaussianDec"))
            dgt.normal_()
            rec.add(sec_g, "LDouble", timed(lambda: capi.check(tg.plan.op("LDouble", dgt.data_ptr(), bt, stream))), 16 * tg.n * bt, bt,
                    kernel=tg.plan.kernel_name("LDouble"))
            del dgt, tg
            torch.cuda.empty_cache()
        # complex CRT of the other benchmark rings: the fused_w schedule over complex doubles, the pass engine beside it
        for mt, bt in ((1728, 131072), (2912, 65536), (11648, 16384), (2048, 65536)):
            tw = CudaTensorComplex(mt)
            cw = torch.randn(bt, tw.n, 1, dtype=torch.complex128, device="cuda")
            for name in ("CRTC", "CRTInvC"):
                ms = timed(lambda: capi.check(tw.plan.op(name, cw.data_ptr(), bt, stream)))
                rec.add(sec, f"m={mt} tensor{name}", ms, 32 * tw.n * bt, bt, kernel=tw.plan.kernel_name(name))
                tw.plan.force_generic(True)
                ms = timed(lambda: capi.check(tw.plan.op(name, cw.data_ptr(), bt, stream)))
                rec.add(sec, f"m={mt} tensor{name} (pass engine)", ms, 32 * tw.n * bt, bt, kernel="generic")
                tw.plan.force_generic(False)
            del cw
            torch.cuda.empty_cache()

    # ---- configs[3]: SymmSHE ciphertext multiply + quadratic key switch (SymmSHE.hs:443-449 then :359-372; op sequence of
    # SURVEY.md section 3.5 with TrivGad over the two limbs, l = 2: 4 CRT, tensor product with mulG, CRTInv, decompose,
    # l CRT, knapsack -- (22 + 4l) x 8nk algorithmic bytes per pair).  Ciphertext pairs shard over the GPUs exactly like
    # the ring elements of the headline step (hints replicated, no collective): every rank runs its shard.
    she_extra = {}
    if not args.no_per_op:
        from lol_b200.symmshe import CudaSymmSHE
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
        elem = 8 * she.n * she.k
        alg = (22 + 4 * she.ell) * elem
        sec = "she"
        if world > 1:
            dist.barrier()
        rec.add(sec, "mulAndSwitch", time_op(torch, she_step, 10), alg * Bs, Bs, kernel_launches_per_step=int(she_launches),
                workload=f"configs[3]: m=14400, q=(1008001,1065601), TrivGad (l={she.ell}), {Bs} ciphertext pairs per GPU, Pow-basis inputs, in place",
                algorithmic_bytes_per_pair=alg)
        d3 = she.mulCT(cts[:2], cts[2:], basis="crt")
        dg = she.decompose(cts[0])
        steps = {"ct_mul": (lambda: capi.check(she.t.plan.ct_mul(*[c.data_ptr() for c in cts], *[d.data_ptr() for d in d3], Bs, True, stream)), 7),
                 "decompose": (lambda: capi.check(she.t.plan.decompose(cts[0].data_ptr(), dg.data_ptr(), Bs, 0, stream)), 1 + she.ell),
                 "knapsack": (lambda: capi.check(she.t.plan.knapsack(dg.data_ptr(), she.ell, hint.data_ptr(), d3[0].data_ptr(), d3[1].data_ptr(), Bs, stream)), she.ell + 4)}
        for name, (fn, passes) in steps.items():
            rec.add(sec, name, timed(fn, 10), passes * elem * Bs, Bs)
        del d3, dg
        if rank == 0 and world == 1 and not args.no_cpu:      # the same sequence on one host core: compiled reference CRTs + numpy host steps
            import numpy as np
            from oracle import cpu as ocpu, symmshe as osym, tables as T
            lib = ocpu.reference() if ocpu.have_reference() else ocpu.restatement()
            pe = T.pe_array(M)
            tabs = (pe, T.ru_tables_zq(M, qs_c), T.ru_tables_zq(M, qs_c, True), [T.mhat_inv(M, q) for q in qs_c], T.g_crt_vectors(M, qs_c)[0])
            rng = np.random.default_rng(0)
            mk = lambda: np.stack([rng.integers(0, q, size=she.n) for q in qs_c], axis=-1).astype(np.int64)
            hint_h = hint.cpu().numpy()
            sample = 8
            t0 = time.perf_counter()
            for _ in range(sample):
                osym.mul_and_switch(lib, [mk(), mk()], [mk(), mk()], hint_h, tabs, qs_c, 0)
            she_extra["cpu_baseline"] = {"value": sample / (time.perf_counter() - t0), "unit": "ct_pairs/s", "cores": 1,
                                         "kind": "reference" if ocpu.have_reference() else "port",
                                         "sample": f"{sample} ciphertext pairs, reference CRTs + numpy host steps"}
        del cts, hint
        torch.cuda.empty_cache()
        try:
            ext_section(torch, capi, rec, stream)
        except Exception as exc:      # informational section: never take the headline line down with it
            she_extra["ext_section_error"] = f"{type(exc).__name__}: {exc}"
        torch.cuda.empty_cache()

    # ---- N > 1 only: inputs that start on one rank.  NCCL over NVLink scatters a batch from rank 0 and gathers the results
    # (lol_b200/shard.py: grouped ncclSend / ncclRecv); outside the headline step, which generates its data per rank.
    shard_io = None
    if world > 1 and not args.no_per_op:
        from lol_b200.shard import gather_batch, scatter_batch
        Bn = 4096 * world
        full = torch.randint(0, QS[0], (Bn, N_COEFF, 1), dtype=torch.int64, device="cuda", generator=gen) if rank == 0 else None
        loc = scatter_batch(full, Bn, (N_COEFF, 1), torch.int64, torch.device("cuda", local))      # warm-up (NCCL channels)
        gather_batch(loc, Bn)
        ts, tg = [], []
        for _ in range(3):      # ~1 ms transfers timed on the host clock: the best of three (an allocation or a late rank costs as much as the copy)
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter()
            loc = scatter_batch(full, Bn, (N_COEFF, 1), torch.int64, torch.device("cuda", local))
            torch.cuda.synchronize(); dist.barrier()
            t1 = time.perf_counter()
            back = gather_batch(loc, Bn)
            torch.cuda.synchronize(); dist.barrier()
            t2 = time.perf_counter()
            ts.append(reduce_max(t1 - t0)); tg.append(reduce_max(t2 - t1))
        moved = (Bn - 4096) * N_COEFF * 8
        shard_io = {"bytes_over_nvlink": moved, "scatter_GB/s": moved / min(ts) / 1e9, "gather_GB/s": moved / min(tg) / 1e9, "samples": 3,
                    "roundtrip_identical": bool(rank != 0 or torch.equal(back, full))}
        del full, loc, back

    rec.reduce_max(torch, dist, world)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    dom_name, dom_ms = ("tensorCRTInvRq", inv_ms) if inv_ms >= crt_ms else ("tensorCRTRq", crt_ms)
    alg_bytes = BYTES_PER_ELEM * B
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    traffic = recorded_traffic() or {}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get(dom_name), "traffic_source": traffic.get("source"),
                "kernel": dom_name + " [" + t.plan.kernel_name("CRTInv" if dom_name.endswith("InvRq") else "CRT") + "]",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                "ms_per_launch": {"tensorCRTRq": crt_ms, "tensorCRTInvRq": inv_ms},
                "frac_per_kernel": {"tensorCRTRq": alg_bytes / (crt_ms * 1e-3) / 1e9 / peak,
                                    "tensorCRTInvRq": alg_bytes / (inv_ms * 1e-3) / 1e9 / peak}}

    cpu_base = None
    if world == 1 and not args.no_cpu:
        cpu_base, _ = cpu_arm(args.cpu_pairs, steps=3, warmup=1)

    detail = rec.table(peak, world)
    if she_extra:
        detail["extra"] = she_extra
    if shard_io:
        detail["nccl_shard_io"] = shard_io
    sys.stderr.write("DETAIL " + json.dumps({"n_gpus": world, "peak_GB/s": peak, "sections": detail}) + "\n")
    sys.stderr.flush()
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        try:
            with open(os.path.join(out_dir, f"bench_detail_n{world}.json"), "w") as f:
                json.dump({"n_gpus": world, "peak_GB/s": peak, "sections": detail}, f, indent=1)
        except OSError:
            pass

    fr = lambda s, n: rec.frac(peak, s, n)
    # compact: fractions of the measured HBM peak (max-over-ranks times); she_n = ciphertext pairs/s over all GPUs
    summary = {"cfgB_crt": fr("cfgB", "CRT"), "cfgB_inv": fr("cfgB", "CRTInv"), "cfgC_crt": fr("cfgC", "CRT"), "cfgC_inv": fr("cfgC", "CRTInv"),
               "she": fr("she", "mulAndSwitch"), "she_n": None, "m1728": [fr("F64*F27/3457", "CRT"), fr("F64*F27/3457", "CRTInv")],
               "m5184": [fr("F64*F81/10369", "CRT"), fr("F64*F81/10369", "CRTInv")],
               "m2912": [fr("F32*F7*F13/8737", "CRT"), fr("F32*F7*F13/8737", "CRTInv")],
               "m3640": [fr("F8*F5*F7*F13/14561", "CRT"), fr("F8*F5*F7*F13/14561", "CRTInv")],
               "m11648": [fr("F128*F7*F13/23297", "CRT"), fr("F128*F7*F13/23297", "CRTInv")],
               "tunnel": [[fr(f"tunnel m={mt}/3144961", "CRT"), fr(f"tunnel m={mt}/3144961", "CRTInv")] for mt in (11648, 5824, 2912, 3640, 5460, 4095)],
               "prf_k2": [fr("HomomPRF H1'=F64*F7*F13 ZQ2", "CRT"), fr("HomomPRF H1'=F64*F7*F13 ZQ2", "CRTInv")],
               "prf_k4": [fr("HomomPRF H1'=F64*F7*F13 ZQ4", "CRT"), fr("HomomPRF H1'=F64*F7*F13 ZQ4", "CRTInv")],
               "crtC": [fr("cfg4", "tensorCRTC"), fr("cfg4", "tensorCRTInvC")], "gauss": fr("cfg4", "tensorGaussianDec"),
               "gauss2912": fr("error m=2912", "tensorGaussianDec"), "gauss11648": fr("error m=11648", "tensorGaussianDec"),
               "crtC2048": [fr("cfg4", "m=2048 tensorCRTC"), fr("cfg4", "m=2048 tensorCRTInvC")]}
    for s, n, ms, _, units, _ in rec.rows:
        if s == "she" and n == "mulAndSwitch":
            summary["she_n"] = round(world * units / (ms * 1e-3))
    if shard_io:
        summary["nvl_scatter_GBs"] = round(shard_io["scatter_GB/s"], 1)
        summary["nvl_gather_GBs"] = round(shard_io["gather_GB/s"], 1)
    if e2e:
        summary["e2e_copy_frac"] = round(e2e["frac_of_copy_ceiling"], 3)
        summary["e2e_u32"] = round(e2e["u32_wire"]["value"])

    line = {
        "metric": METRIC, "value": world * B / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 modular (int64 ABI)", "data": "synthetic",
        "config": config_dict(world),
        "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": clk.summary(), "summary": summary,
    }
    print(json.dumps(line))
    sys.stdout.flush()
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
    ap.add_argument("--cpu-pairs", type=int, default=2048, help="CRT+CRTInv pairs per host process and step in the cpu_baseline sample")
    ap.add_argument("--ref-pairs", type=int, default=128, help="--impl reference: CRT+CRTInv pairs per host process in one step")
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
