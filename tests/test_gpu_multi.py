"""GPU suite, N > 1 (-m gpu; skipped on a box with fewer than 2 GPUs): the sharded path over NCCL / NVLink.

Ring elements and ciphertext pairs are independent, so the data path has no collective; NCCL carries only
the scatter of inputs that start on one rank and the gather of the results (lol_b200/shard.py: grouped ncclSend / ncclRecv).
Checked here on real devices: scatter -> per-rank CRT (and the SymmSHE multiply + key switch of BASELINE.json configs[3]) ->
gather equals the same call on one GPU, bit for bit, for a ragged batch; the scatter / gather bandwidth over NVLink is printed
and written to gpurun_out/nccl_shard_io.json when that directory exists.
"""
import json
import os
import socket
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from lol_b200.shard import broadcast_replicated, gather_batch, scatter_batch, shard_bounds
        from lol_b200.symmshe import CudaSymmSHE
        from lol_b200.tensor import CudaTensorRq
        dev = torch.device("cuda", rank)
        m, qs = 14400, [1008001, 1065601]
        t = CudaTensorRq(m, qs)
        n, k = t.n, len(qs)
        B = 1001 * world + 3                                   # ragged
        full = want = None
        gen = torch.Generator(device=dev)
        gen.manual_seed(5)
        if rank == 0:
            full = torch.cat([torch.randint(0, q, (B, n, 1), dtype=torch.int64, device=dev, generator=gen) for q in qs], dim=2).contiguous()
            want = t.crtInv(t.mulGCRT(t.crt(full)))           # the whole batch on one GPU
        local = scatter_batch(full, B, (n, k), torch.int64, dev)
        lo, hi = shard_bounds(B, world, rank)
        assert local.shape[0] == hi - lo
        got = gather_batch(t.crtInv(t.mulGCRT(t.crt(local))), B)
        ok = True
        if rank == 0:
            ok = bool(torch.equal(got, want))
        # configs[3]: ciphertext pairs sharded, key-switch hint replicated
        she = CudaSymmSHE(m, qs, gad_base=0)
        Bs = 37 * world + 1
        cts_full = hint = want_she = None
        if rank == 0:
            cts_full = [torch.cat([torch.randint(0, q, (Bs, n, 1), dtype=torch.int64, device=dev, generator=gen) for q in qs], dim=2).contiguous()
                        for _ in range(4)]
            hint = torch.cat([torch.randint(0, q, (she.ell, 2, n, 1), dtype=torch.int64, device=dev, generator=gen) for q in qs], dim=3).contiguous()
            want_she = she.mulAndSwitch([c.clone() for c in cts_full[:2]], [c.clone() for c in cts_full[2:]], hint, basis="pow")
        hint = broadcast_replicated(hint, (she.ell, 2, n, k), torch.int64, dev)
        cts = [scatter_batch(cts_full[i] if rank == 0 else None, Bs, (n, k), torch.int64, dev) for i in range(4)]
        res = she.mulAndSwitch(cts[:2], cts[2:], hint, basis="pow")
        outs = [gather_batch(r.contiguous(), Bs) for r in res]
        if rank == 0:
            ok = ok and all(torch.equal(a, b) for a, b in zip(outs, want_she))
        # bandwidth of the scatter / gather alone (1001 * world elements of 122 880 bytes)
        torch.cuda.synchronize()
        dist.barrier()
        reps, t_sc, t_ga = 5, 0.0, 0.0
        for _ in range(reps):
            t0 = time.perf_counter()
            loc = scatter_batch(full, B, (n, k), torch.int64, dev)
            torch.cuda.synchronize(); dist.barrier()
            t1 = time.perf_counter()
            gather_batch(loc, B)
            torch.cuda.synchronize(); dist.barrier()
            t_sc += t1 - t0
            t_ga += time.perf_counter() - t1
        if rank == 0:
            moved = (B - (hi - lo)) * n * k * 8                 # bytes that leave / enter rank 0
            info = {"ok": ok, "world": world, "bytes_over_nvlink": moved, "scatter_GBps": moved / (t_sc / reps) / 1e9,
                    "gather_GBps": moved / (t_ga / reps) / 1e9}
            with open(out_path, "w") as f:
                json.dump(info, f)
    finally:
        dist.destroy_process_group()


def test_sharded_crt_and_symmshe_over_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available():
        pytest.fail("GPU suite needs a CUDA device: libctensor_b200 has no CPU path")
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs on the box")
    from lol_b200 import build_library
    build_library()
    out = str(tmp_path / "nccl.json")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    info = json.load(open(out))
    print("NCCL shard I/O:", info)
    repo_out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(repo_out):
        with open(os.path.join(repo_out, "nccl_shard_io.json"), "w") as f:
            json.dump(info, f)
    assert info["ok"], "sharded result differs from the single-GPU result"
