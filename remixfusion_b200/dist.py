"""Multi-GPU plumbing for the hot path — one process per GPU, ``torch.distributed`` (NCCL over
NVLink 5 / NVSwitch on the box, gloo in the CPU tests).  The reference has no distributed code at all (two
processes on one GPU), so this layer is new:

  * TSDF integration shards *voxels*: contiguous slabs of the slowest axis (z-slabs for the GBV, x-slabs for the
    local volume).  Each rank integrates the same frame into its slab — no data-path collective except the
    broadcast of the frame itself; G ranks produce the same bits as one (tests/test_tsdf_gpu.py).
  * The ray query shards *rays*: every rank renders its part of the batch against replicated parameters; the loss
    normalisers are batch-global, so the 7 loss sums are all-reduced between forward and backward
    (scene_rep._RayQueryFn) and the table / decoder gradients are summed with one all-reduce per step.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def slab(n: int, rank: int, world: int):
    """Contiguous slab [lo, hi) of an axis of length n owned by `rank` (balanced to within one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(n: int, rank: int, world: int):
    return slab(n, rank, world)


def broadcast_frame(depth: torch.Tensor, color: torch.Tensor, src: int = 0, group=None):
    """Frame broadcast (depth [H,W] + colour, ~13 MB at 1200x680) from the rank that decoded it."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(depth, src=src, group=group)
        dist.broadcast(color, src=src, group=group)
    return depth, color


def allreduce_grads(params, group=None):
    """Sum the gradients of `params` over ranks with ONE all-reduce of a flat buffer (hash table + 4 weight matrices)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def gather_slabs(local: torch.Tensor, sizes, group=None):
    """All-gather variable-size slabs into the full volume on every rank (GBV after a keyframe: 160 MB at R=200)."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return local
    m = max(sizes)                                   # equal-size all-gather of padded slabs, then trim
    pad = local if local.numel() == m else torch.cat([local, local.new_zeros(m - local.numel())])
    outs = [torch.empty(m, dtype=local.dtype, device=local.device) for _ in sizes]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)])
