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
    """Rank `src` holds `full` ([batch, *tail], on `device`); every rank returns its shard ([hi-lo, *tail]).
    One grouped send/receive (ncclSend / ncclRecv over NVLink on GPUs): rank `src` sends the contiguous slices of `full`
    themselves -- no padded or staged copies -- and every other rank receives straight into its shard."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(batch, world, rank)
    if rank == src:
        assert full is not None and tuple(full.shape) == (batch, *tail) and full.dtype == dtype and full.is_contiguous()
        ops = []
        for r in range(world):
            rlo, rhi = shard_bounds(batch, world, r)
            if r != src and rhi > rlo:
                ops.append(dist.P2POp(dist.isend, full[rlo:rhi], r, group))
        for req in (dist.batch_isend_irecv(ops) if ops else []):
            req.wait()
        return full[lo:hi].clone()
    recv = torch.empty((hi - lo, *tail), dtype=dtype, device=device)
    if hi > lo:
        for req in dist.batch_isend_irecv([dist.P2POp(dist.irecv, recv, src, group)]):
            req.wait()
    return recv


def gather_batch(local: torch.Tensor, batch: int, dst: int = 0, group=None) -> torch.Tensor | None:
    """Inverse of `scatter_batch`: rank `dst` returns the [batch, *tail] tensor, the others None.  Rank `dst` receives every
    shard straight into its slice of the result."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(batch, world, rank)
    assert local.shape[0] == hi - lo and local.is_contiguous()
    if rank != dst:
        if hi > lo:
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, local, dst, group)]):
                req.wait()
        return None
    out = torch.empty((batch, *local.shape[1:]), dtype=local.dtype, device=local.device)
    ops = []
    for r in range(world):
        rlo, rhi = shard_bounds(batch, world, r)
        if r != dst and rhi > rlo:
            ops.append(dist.P2POp(dist.irecv, out[rlo:rhi], r, group))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    out[lo:hi] = local
    for req in reqs:
        req.wait()
    return out


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
