"""Multi-GPU partitioning of the tensor path: ring elements (and RNS limbs) are independent
(the reference loops over them outside its kernels, tensor.h:61-70, 87-92), so a batch is split
contiguously over ranks and every rank runs the single-GPU operators on its shard.  There is NO
collective inside any transform; `torch.distributed` (NCCL over NVLink on GPUs, gloo in the CPU
tests) is used only to scatter inputs that originate on one rank and to gather results.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Elements [lo, hi) of a batch owned by `rank`: contiguous, sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world) or batch < 0:
        raise ValueError("bad (batch, world, rank)")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_shard(batch: int, world: int) -> int:
    return -(-batch // world)


def scatter_batch(full: torch.Tensor | None, batch: int, tail: tuple[int, ...], dtype: torch.dtype,
                  device, src: int = 0, group=None) -> torch.Tensor:
    """Rank `src` holds `full` ([batch, *tail]); every rank returns its shard ([hi-lo, *tail]).
    Shards travel padded to the largest shard so the collective is one `scatter`."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    pad = max_shard(batch, world)
    recv = torch.empty((pad, *tail), dtype=dtype, device=device)
    chunks = None
    if rank == src:
        assert full is not None and tuple(full.shape) == (batch, *tail)
        chunks = []
        for r in range(world):
            lo, hi = shard_bounds(batch, world, r)
            c = torch.zeros((pad, *tail), dtype=dtype, device=device)
            c[: hi - lo] = full[lo:hi].to(device)
            chunks.append(c)
    dist.scatter(recv, chunks, src=src, group=group)
    lo, hi = shard_bounds(batch, world, rank)
    return recv[: hi - lo].contiguous()


def gather_batch(local: torch.Tensor, batch: int, dst: int = 0, group=None) -> torch.Tensor | None:
    """Inverse of `scatter_batch`: rank `dst` returns the [batch, *tail] tensor, the others None."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    pad = max_shard(batch, world)
    tail = tuple(local.shape[1:])
    send = torch.zeros((pad, *tail), dtype=local.dtype, device=local.device)
    send[: local.shape[0]] = local
    bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        lo, hi = shard_bounds(batch, world, r)
        parts.append(bufs[r][: hi - lo])
    return torch.cat(parts, dim=0)


def broadcast_replicated(t: torch.Tensor | None, shape: tuple[int, ...], dtype: torch.dtype, device, src: int = 0,
                         group=None) -> torch.Tensor:
    """Per-(m, qs) data every rank needs whole -- key-switch hints (SymmSHE.hs:288-298), gCRT vectors: rank `src`
    holds `t`, every rank returns a copy.  KBs to a few MBs, once per key, outside any timed region."""
    rank = dist.get_rank(group)
    buf = t.to(device).contiguous() if rank == src else torch.empty(shape, dtype=dtype, device=device)
    if rank == src:
        assert tuple(buf.shape) == tuple(shape) and buf.dtype == dtype
    dist.broadcast(buf, src=src, group=group)
    return buf
