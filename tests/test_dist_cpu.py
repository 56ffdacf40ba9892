"""CPU suite, world_size 2 over gloo: the host-side sharding logic of remixfusion_b200/dist.py (slab partition,
frame broadcast, gradient all-reduce, slab all-gather) and the loss-sum all-reduce arithmetic that makes a sharded
ray batch reproduce the single-process losses."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from remixfusion_b200 import dist as rdist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = rdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # frame broadcast
    depth = torch.full((4, 5), float(rank + 1)); color = torch.full((4, 5, 3), float(10 * (rank + 1)))
    rdist.broadcast_frame(depth, color, src=0)
    ok = bool((depth == 1).all() and (color == 10).all())
    # gradient all-reduce through one flat buffer
    p1 = torch.nn.Parameter(torch.zeros(7)); p2 = torch.nn.Parameter(torch.zeros(3, 2)); p3 = torch.nn.Parameter(torch.zeros(2))
    p1.grad = torch.arange(7.) * (rank + 1); p2.grad = torch.ones(3, 2) * (rank + 1)
    rdist.allreduce_grads([p1, p2, p3])
    ok &= bool(torch.equal(p1.grad, torch.arange(7.) * 3) and torch.equal(p2.grad, torch.ones(3, 2) * 3) and p3.grad is None)
    # slab all-gather of unequal slabs
    lo, hi = rdist.slab(5, rank, world)
    full = rdist.gather_slabs(torch.arange(lo * 4, hi * 4, dtype=torch.float32), [(rdist.slab(5, k, world)[1] - rdist.slab(5, k, world)[0]) * 4 for k in range(world)])
    ok &= bool(torch.equal(full, torch.arange(20.)))
    # equal slabs go straight into the replicated buffer (all_gather_into_tensor, no staging)
    out = torch.full((24 + 8,), -1.0)
    rdist.gather_slabs(torch.arange(rank * 12, rank * 12 + 12, dtype=torch.float32), [12, 12], out=out)
    ok &= bool(torch.equal(out[:24], torch.arange(24.)) and (out[24:] == -1).all())
    # gradients that live in ONE flat buffer (FlatGrads): in-place all-reduce, .grad stays a view
    q1 = torch.nn.Parameter(torch.zeros(5)); q2 = torch.nn.Parameter(torch.zeros(2, 3))
    fg = rdist.FlatGrads([q1, q2])
    (q1 * torch.arange(5.)).sum().backward(); (q2 * (rank + 1.0)).sum().backward()
    fg.allreduce()
    ok &= bool(torch.equal(q1.grad, 2 * torch.arange(5.)) and torch.equal(q2.grad, torch.full((2, 3), 3.0)) and q1.grad._base is fg.flat)
    rdist.allreduce_grads([q1, q2])                              # recognises the shared flat buffer
    ok &= bool(torch.equal(q1.grad, 4 * torch.arange(5.)))
    fg.zero()
    ok &= bool(float(q2.grad.abs().sum()) == 0.0)
    # touched sub-box gather of a z-slab-sharded volume: only the [y, x] box changes on the replicated copy
    R, C = 4, 2
    zs = [rdist.slab(R, k, world) for k in range(world)]
    truth = torch.arange(R * R * R * C, dtype=torch.float32)
    mine = truth.view(R, R, R, C)[zs[rank][0]:zs[rank][1]].reshape(-1).clone()
    full = torch.zeros(R * R * R * C)
    rdist.gather_touched_box(full, mine, R, zs, lo=[1, 0, 0], hi=[3, 2, R], channels=C)
    exp = torch.zeros(R, R, R, C); exp[:, 0:2, 1:3, :] = truth.view(R, R, R, C)[:, 0:2, 1:3, :]
    ok &= bool(torch.equal(full.view(R, R, R, C), exp))
    # sharded loss sums: all-reduced partial sums reproduce the single-process normalised losses
    g = torch.Generator().manual_seed(0)
    vals = torch.rand(10, 7, generator=g, dtype=torch.float64)
    lo, hi = rdist.shard_rays(10, rank, world)
    part = vals[lo:hi].sum(0)
    dist.all_reduce(part)
    ok &= bool(torch.allclose(part, vals.sum(0)))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_slab_partition_covers_axis():
    for n in (1, 7, 200, 300):
        for w in (1, 2, 3, 4, 8):
            cuts = [rdist.slab(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_frustum_box_contains_every_projected_voxel():
    """The hull used to restrict the GBV all-gather must contain every voxel the integrate kernel can touch."""
    from remixfusion_b200 import synth
    rng = np.random.default_rng(0)
    R, box = 64, [[-1.0, 7.0], [-1.3, 3.7], [-1.7, 1.4]]
    K = synth.intrinsics(60.0, 60.0, 59.5, 33.5); H, W = 68, 120
    scene = synth.make_scene(box, 0)
    for c2w in synth.loop_trajectory(scene, 6):
        lo, hi = rdist.frustum_box(K, c2w, H, W, 6.0, box, R)
        g = (np.stack(np.meshgrid(*[np.arange(R)] * 3, indexing="ij"), -1).reshape(-1, 3) / R)
        pts = np.array([b[0] for b in box]) + g * np.array([b[1] - b[0] for b in box])
        cam = (pts - c2w[:3, 3]) @ c2w[:3, :3]
        z = cam[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            px = np.rint(K[0, 0] * cam[:, 0] / z + K[0, 2]); py = np.rint(K[1, 1] * cam[:, 1] / z + K[1, 2])
        vis = (z > 0) & (z <= 6.0) & (px >= 0) & (px < W) & (py >= 0) & (py < H)
        idx = np.rint(g[vis] * R).astype(int)
        assert vis.sum() > 0
        for a in range(3):
            assert idx[:, a].min() >= lo[a] and idx[:, a].max() < hi[a], (a, lo, hi, idx[:, a].min(), idx[:, a].max())


def _torch_adam_segment(p, g, m, v, hp, step):
    """torch's dense Adam on a segment (stand-in for rf_adam_step in the CPU test of the sharding logic); clears g."""
    b1, b2 = hp["betas"]
    gg = g + hp["weight_decay"] * p if hp["weight_decay"] else g.clone()
    m.lerp_(gg, 1 - b1)
    v.mul_(b2).addcmul_(gg, gg, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p.addcdiv_(m, v.sqrt() / (bc2 ** 0.5) + hp["eps"], value=-hp["lr"] / bc1)
    g.zero_()


def _sharded_adam_worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    rdist.init_from_env("gloo")
    torch.manual_seed(1)                                         # replicated parameters
    table = torch.nn.Parameter(torch.randn(1001)); w0 = torch.nn.Parameter(torch.randn(7, 5)); w1 = torch.nn.Parameter(torch.randn(3, 7))
    ref = [torch.nn.Parameter(t.detach().clone()) for t in (table, w0, w1)]
    groups = lambda ps: [{"params": [ps[1], ps[2]], "weight_decay": 1e-6, "lr": 1e-2}, {"params": [ps[0]], "eps": 1e-15, "lr": 1e-2}]
    opt = rdist.ShardedAdam(groups([table, w0, w1]), betas=(0.9, 0.99), adam_fn=_torch_adam_segment)
    ref_opt = torch.optim.Adam(groups(ref), betas=(0.9, 0.99))
    ok = opt.world == world and table.grad._base is opt.gflat and table.data.untyped_storage().data_ptr() == opt.pflat.untyped_storage().data_ptr()
    g = torch.Generator().manual_seed(7)
    for it in range(4):
        x = [torch.randn(world, *t.shape, generator=g) for t in (table, w0, w1)]      # every rank's loss weights (same stream on all ranks)
        sum((p * xi[rank]).sum() for p, xi in zip((table, w0, w1), x)).backward()      # this rank's gradient
        opt.step()
        for r, xi in zip(ref, x):
            r.grad = xi.sum(0)                                                          # the summed gradient, unsharded
        ref_opt.step()
        for a, b in zip((table, w0, w1), ref):
            ok &= bool(torch.allclose(a.detach(), b.detach(), rtol=1e-6, atol=1e-7))
        ok &= bool(float(opt.gflat.abs().sum()) == 0.0)
    q.put((rank, ok, len(opt.segments)))
    dist.destroy_process_group()


def test_sharded_adam_world2_matches_unsharded():
    """reduce-scatter -> Adam on the owned shard -> all-gather reproduces torch.optim.Adam on the summed gradients, with the
    reference's two parameter groups (mp_slam/slam.py:271-286) and a shard boundary inside the hash table."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_sharded_adam_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps: p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sorted(n for _, _, n in res)[-1] >= 2, res              # some shard spans more than one tensor
