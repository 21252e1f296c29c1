"""CPU suite: the N > 1 path (lol_b200/shard.py) with world_size 2 over gloo.

The data-path has no collective; what is tested is the partition (every element owned exactly
once, ragged batches), scatter -> per-rank transform -> gather, and that transforming shards
independently equals transforming the whole batch (the independence the sharding relies on).
The per-rank transform here is the CPU oracle standing in for the GPU operator."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lol_b200.shard import gather_batch, max_shard, scatter_batch, shard_bounds


def test_shard_bounds_partition():
    for batch in (0, 1, 2, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == (max_shard(batch, world) if batch else 0)
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cpu, tables as T
        m, qs = 42, [19393921, 18869761]
        pe, n = T.pe_array(m), T.totient_pps(T.factor_pps(m))
        ru = T.ru_tables_zq(m, qs)
        O = cpu.restatement()
        full = None
        if rank == 0:
            rng = np.random.default_rng(7)
            full = torch.from_numpy(np.stack([rng.integers(0, q, size=(batch, n)) for q in qs], axis=-1).astype(np.int64))
        local = scatter_batch(full, batch, (n, len(qs)), torch.int64, "cpu")
        lo, hi = shard_bounds(batch, world, rank)
        assert local.shape[0] == hi - lo
        out = torch.from_numpy(np.stack([O.tensorCRTRq(local[b].numpy(), pe, ru, qs) for b in range(local.shape[0])])
                               if local.shape[0] else np.zeros((0, n, len(qs)), dtype=np.int64))
        got = gather_batch(out, batch)
        if rank == 0:
            want = np.stack([O.tensorCRTRq(full[b].numpy(), pe, ru, qs) for b in range(batch)])
            ok = got.shape == (batch, n, len(qs)) and np.array_equal(got.numpy(), want)
            with open(result_path, "w") as f:
                f.write("ok" if ok else "mismatch")
        # timing reduction used by bench.py: max over ranks
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == float(world)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [5, 8])
def test_scatter_transform_gather_world2(tmp_path, batch):
    result = tmp_path / "result.txt"
    mp.spawn(_worker, args=(2, _free_port(), batch, str(result)), nprocs=2, join=True)
    assert result.read_text() == "ok"


# ------------------------------------------------------------------ configs[3]: ciphertext pairs sharded, hints replicated
def _she_worker(rank, world, port, batch, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lol_b200.shard import broadcast_replicated
        from oracle import cpu, symmshe as S, tables as T
        m, qs, base = 21, [43, 127], 4
        pe, n = T.pe_array(m), T.totient_pps(T.factor_pps(m))
        tabs = (pe, T.ru_tables_zq(m, qs), T.ru_tables_zq(m, qs, inverse=True), [T.mhat_inv(m, q) for q in qs], T.g_crt_vectors(m, qs)[0])
        ell = S.gadget_length(qs, base)
        O = cpu.restatement()
        comps, hint = [None] * 4, None
        if rank == 0:
            rng = np.random.default_rng(11)
            mk = lambda *shape: torch.from_numpy(np.stack([rng.integers(0, q, size=shape) for q in qs], axis=-1).astype(np.int64))
            comps = [mk(batch, n) for _ in range(4)]
            hint = mk(ell, 2, n)
        local = [scatter_batch(c, batch, (n, len(qs)), torch.int64, "cpu") for c in comps]
        hint_l = broadcast_replicated(hint, (ell, 2, n, len(qs)), torch.int64, "cpu").numpy()
        run = lambda a0, a1, b0, b1: S.mul_and_switch(O, [a0, a1], [b0, b1], hint_l, tabs, qs, base)
        outs = [run(*[c[b].numpy() for c in local]) for b in range(local[0].shape[0])]
        empty = np.zeros((0, n, len(qs)), dtype=np.int64)
        got = [gather_batch(torch.from_numpy(np.stack([o[j] for o in outs]) if outs else empty), batch) for j in range(2)]
        if rank == 0:
            want = [run(*[c[b].numpy() for c in comps]) for b in range(batch)]
            ok = all(np.array_equal(got[j].numpy(), np.stack([w[j] for w in want])) for j in range(2))
            with open(result_path, "w") as f:
                f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [3, 4])
def test_symmshe_pairs_sharded_hints_replicated_world2(tmp_path, batch):
    """keySwitchQuadCirc hint (c1 * c2) per ciphertext pair shards by pair with the hint replicated:
    scatter -> per-rank pipeline -> gather equals the single-rank result."""
    result = tmp_path / "result.txt"
    mp.spawn(_she_worker, args=(2, _free_port(), batch, str(result)), nprocs=2, join=True)
    assert result.read_text() == "ok"


# ------------------------------------------------------------------ ring switching: O_m' elements sharded, output in O_m with one limb fewer
def _ext_worker(rank, world, port, batch, result_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import coeffwise as W, extension as X
        m, m2, qs = 3, 21, [19393921, 18869761]
        info = X.ExtInfo(m, m2)
        full = None
        if rank == 0:
            rng = np.random.default_rng(13)
            full = torch.from_numpy(np.stack([rng.integers(0, q, size=(batch, info.phi2)) for q in qs], axis=-1).astype(np.int64))
        # twaceCRT then the RNS limb drop: [b, phi', k] -> [b, phi, k-1]; the index tables are per-(m, m') data every rank builds itself
        step = lambda y: W.rescale_drop(X.twace_crt_zq(info, y, qs), qs, 0)
        local = scatter_batch(full, batch, (info.phi2, len(qs)), torch.int64, "cpu")
        out = (np.stack([step(local[b].numpy()) for b in range(local.shape[0])]) if local.shape[0]
               else np.zeros((0, info.phi, len(qs) - 1), dtype=np.int64))
        got = gather_batch(torch.from_numpy(out), batch)
        if rank == 0:
            want = np.stack([step(full[b].numpy()) for b in range(batch)])
            ok = got.shape == (batch, info.phi, len(qs) - 1) and np.array_equal(got.numpy(), want)
            with open(result_path, "w") as f:
                f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [1, 5])
def test_ring_switch_elements_sharded_world2(tmp_path, batch):
    """twaceCRT + rescale shard by element like every other operator (batch 1: one rank owns nothing)."""
    result = tmp_path / "result.txt"
    mp.spawn(_ext_worker, args=(2, _free_port(), batch, str(result)), nprocs=2, join=True)
    assert result.read_text() == "ok"
