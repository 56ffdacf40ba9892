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
