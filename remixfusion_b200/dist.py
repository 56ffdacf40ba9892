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


class FlatGrads:
    """One persistent flat gradient buffer whose views ARE the parameters' ``.grad``: the per-step gradient sum is a single
    in-place all-reduce of that buffer — no concatenation and no copy back (hash table + the four decoder matrices)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(n, dtype=p0.dtype, device=p0.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        """Instead of ``p.grad = None``: autograd accumulates into the views."""
        self.flat.zero_()

    def allreduce(self, group=None, async_op=False):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None


def allreduce_grads(params, group=None):
    """Sum the gradients of `params` over ranks with ONE all-reduce of a flat buffer (hash table + 4 weight matrices).
    Prefer ``FlatGrads`` in a training loop: it keeps the flat buffer alive and needs no copies."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    base = grads[0]._base if grads[0]._base is not None else None
    if base is not None and all(g._base is base for g in grads) and sum(g.numel() for g in grads) == base.numel():
        dist.all_reduce(base, op=dist.ReduceOp.SUM, group=group)          # already views of one flat buffer
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def gather_slabs(local: torch.Tensor, sizes, group=None, out: torch.Tensor | None = None):
    """All-gather the ranks' slabs into the full volume on every rank (GBV after a keyframe: 160 MB at R = 200).
    Equal slabs (R divisible by the world size) go straight into ``out`` with one ``all_gather_into_tensor`` — no staging,
    no concatenation; unequal slabs are padded to the largest one first."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        if out is not None and out.data_ptr() != local.data_ptr():
            out[:local.numel()].copy_(local)
            return out
        return local
    total = int(sum(sizes))
    if out is None:
        out = torch.empty(total, dtype=local.dtype, device=local.device)
    if len(set(int(s) for s in sizes)) == 1:
        dist.all_gather_into_tensor(out[:total], local[:int(sizes[0])].contiguous(), group=group)
        return out
    m = int(max(sizes))
    pad = local if local.numel() == m else torch.cat([local, local.new_zeros(m - local.numel())])
    stage = torch.empty(m * len(sizes), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(stage, pad, group=group)
    off = 0
    for k, sz in enumerate(sizes):
        out[off:off + int(sz)].copy_(stage[k * m:k * m + int(sz)])
        off += int(sz)
    return out


def frustum_box(K, c2w, H, W, max_depth, box, R):
    """Voxel index range [lo, hi) per axis of the GBV (resolution R over `box` = [[x0,x1],[y0,y1],[z0,z1]]) that the camera
    frustum up to `max_depth` can touch: the axis-aligned hull of the frustum's corners, clamped to the volume."""
    import numpy as np
    K = np.asarray(K, dtype=np.float64).reshape(3, 3); c2w = np.asarray(c2w, dtype=np.float64).reshape(4, 4)
    pts = [np.zeros(3)]
    for px, py in ((-1.0, -1.0), (W + 1.0, -1.0), (-1.0, H + 1.0), (W + 1.0, H + 1.0)):
        pts.append(np.array([(px - K[0, 2]) / K[0, 0] * max_depth, (py - K[1, 2]) / K[1, 1] * max_depth, max_depth]))
    w = np.stack([c2w[:3, :3] @ p + c2w[:3, 3] for p in pts])
    lo, hi = [], []
    for a in range(3):
        b0, b1 = float(box[a][0]), float(box[a][1])
        v0 = (w[:, a].min() - b0) / (b1 - b0) * R
        v1 = (w[:, a].max() - b0) / (b1 - b0) * R
        lo.append(int(min(max(np.floor(v0) - 1, 0), R))); hi.append(int(min(max(np.ceil(v1) + 2, 0), R)))
    return lo, hi


def gather_touched_box(full: torch.Tensor, local: torch.Tensor, R: int, z_slabs, lo, hi, group=None, channels: int = 4):
    """After a keyframe only the voxels inside the frustum's hull changed: all-gather just the [y, x] sub-box of every rank's
    z-slab (equal slabs: one pack, one ``all_gather_into_tensor``, one unpack) instead of the whole volume — at BS3D scale
    (R = 512 / 1024 over 50 x 50 x 10 m) the frustum covers a few per cent of the x-y extent.  `full`: [R^3 * channels] replicated
    copy, `local`: this rank's z-slab [dz * R * R * channels]; lo / hi from ``frustum_box``.  Falls back to ``gather_slabs`` when the slabs differ."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dzs = [int(b - a) for a, b in z_slabs]
    if world == 1 or len(set(dzs)) != 1 or hi[0] <= lo[0] or hi[1] <= lo[1]:
        return gather_slabs(local, [dz * R * R * channels for dz in dzs], group, out=full)
    dz = dzs[0]
    loc = local[:dz * R * R * channels].view(dz, R, R, channels)
    piece = loc[:, lo[1]:hi[1], lo[0]:hi[0], :].contiguous()
    stage = torch.empty((world,) + tuple(piece.shape), dtype=piece.dtype, device=piece.device)
    dist.all_gather_into_tensor(stage.view(-1), piece.view(-1), group=group)
    full[:world * dz * R * R * channels].view(world, dz, R, R, channels)[:, :, lo[1]:hi[1], lo[0]:hi[0], :] = stage
    return full


def _rf_adam_segment(p, g, m, v, hp, step):
    """One fused Adam pass (rf_adam_step) over a contiguous fp32 CUDA segment; clears the gradient in the same pass."""
    import ctypes as C
    from . import abi
    rc = abi.lib().rf_adam_step(abi.dptr(p), abi.dptr(g), abi.dptr(m), abi.dptr(v), C.c_int64(p.numel()), C.c_double(hp["lr"]),
                                C.c_double(hp["betas"][0]), C.c_double(hp["betas"][1]), C.c_double(hp["eps"]),
                                C.c_double(hp["weight_decay"]), C.c_int64(step), None, C.c_int(1), abi.stream_ptr())
    abi.check(rc, "rf_adam_step")


class ShardedAdam:
    """The optimiser step of the mapping loop (mp_slam/mapper.py:417-423 over the groups of mp_slam/slam.py:271-286) for
    replicated parameters and rank-sharded ray batches:  reduce-scatter of the gradients -> fused Adam on the owned 1/W of every
    tensor -> all-gather of the updated parameters.  Same bytes on the wire as the all-reduce it replaces (a ring all-reduce IS
    reduce-scatter + all-gather), but the Adam pass and its two moment buffers shrink by the world size.

    Parameters and gradients are re-homed as views of two flat buffers (each tensor starts on a 16-byte boundary, the total is
    padded to a multiple of 4 W floats): ``p.data`` / ``p.grad`` stay usable by the model and by autograd.  Element-wise, the
    update is ``remixfusion_b200.optim.Adam``'s (torch's dense Adam); a shard that straddles two parameter groups is stepped
    in one segment per group.  ``adam_fn`` replaces the CUDA kernel in the CPU tests of the host logic."""

    def __init__(self, param_groups, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, group=None, adam_fn=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._adam = adam_fn or _rf_adam_segment
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        spans, off, first = [], 0, None                        # (param, start, numel, hyper-parameters)
        for gdict in param_groups:
            hp = {k: gdict.get(k, d) for k, d in defaults.items()}
            for p in gdict["params"]:
                if not p.requires_grad:
                    continue
                first = p if first is None else first
                spans.append((p, off, p.numel(), hp))
                off += (p.numel() + 3) & ~3
        if first is None:
            raise ValueError("ShardedAdam: no trainable parameter")
        quantum = 4 * self.world
        total = (off + quantum - 1) // quantum * quantum
        self.pflat = torch.zeros(total, dtype=first.dtype, device=first.device)
        self.gflat = torch.zeros(total, dtype=first.dtype, device=first.device)
        for p, s, n, _ in spans:
            self.pflat[s:s + n].copy_(p.data.reshape(-1))
            p.data = self.pflat[s:s + n].view_as(p)
            p.grad = self.gflat[s:s + n].view_as(p)
        self.shard = total // self.world
        self.lo = self.rank * self.shard
        self.gshard = self.gflat[self.lo:self.lo + self.shard] if self.world == 1 else torch.zeros_like(self.gflat[:self.shard])
        self.exp_avg = torch.zeros_like(self.gshard)
        self.exp_avg_sq = torch.zeros_like(self.gshard)
        self.segments = []                                     # (start inside the shard, numel, hyper-parameters)
        for _, s, n, hp in spans:
            a, b = max(s, self.lo), min(s + n, self.lo + self.shard)
            if b > a:
                self.segments.append((a - self.lo, b - a, hp))
        self.steps = 0
        self._rs = None
        if self.world > 1:
            try:                                               # gloo (the CPU tests) has no reduce-scatter: all-reduce and keep the shard
                dist.reduce_scatter_tensor(torch.zeros_like(self.gshard), torch.zeros_like(self.gflat), group=group)
                self._rs = True
            except (RuntimeError, NotImplementedError):
                self._rs = False

    def zero_grad(self):
        self.gflat.zero_()

    @torch.no_grad()
    def step(self):
        """Gradients summed over the ranks, parameters stepped and replicated again, gradients cleared."""
        if self.world > 1:
            if self._rs:
                dist.reduce_scatter_tensor(self.gshard, self.gflat, op=dist.ReduceOp.SUM, group=self.group)
            else:
                dist.all_reduce(self.gflat, op=dist.ReduceOp.SUM, group=self.group)
                self.gshard.copy_(self.gflat[self.lo:self.lo + self.shard])
        self.steps += 1
        own = self.pflat[self.lo:self.lo + self.shard]
        for a, n, hp in self.segments:
            self._adam(own[a:a + n], self.gshard[a:a + n], self.exp_avg[a:a + n], self.exp_avg_sq[a:a + n], hp, self.steps)
        if self.world > 1:
            dist.all_gather_into_tensor(self.pflat, own.clone() if not self._rs else own, group=self.group)
            self.gflat.zero_()
