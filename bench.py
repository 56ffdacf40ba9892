#!/usr/bin/env python
"""bench.py — the mapping hot path on BASELINE.json's config 2 (Replica-shaped 1200x680 sequence, 2 cm voxels).

One *step* is one frame of the synthetic sequence pushed through the whole hot path:
  1. TSDF fuse:   integrate the frame into the tracker's local moving volume (400 x 400 x 300 @ 2 cm) and into the
                  mapper's global coarse volume (GBV, R = 200 over the Replica room0 bound);
  2. mixed ray render, forward + backward: all 816 000 pixels of the frame as rays x (48 + 11) samples through
                  JointEncoding.mapping (hash grid 16 x 2^16, OneBlob, GBV trilerp, 2 x 32 decoder, SDF compositing,
                  the four mapping losses) and loss.backward() into the hash table and the decoder.
Units per step = touched voxels (both volumes) + ray samples; metric = units / s (BASELINE.json: "TSDF
voxel-updates/s + ray-samples/s (fwd+bwd)"), with the two parts also reported separately under "parts".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (torchrun): volumes are slab-sharded (x-slabs local, z-slabs GBV) with the frame broadcast from rank 0 over
NCCL; every rank renders its own frame's rays (weak scaling) and the hash/decoder gradients are all-reduced.
`--impl reference`: the reference's CPU path (oracle/: the C restatement of its TSDF kernels and the reference's own
scene_rep semantics on a PyTorch stand-in for tiny-cuda-nn) on the box's host cores, bounded sample, scaled.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from remixfusion_b200 import configs, synth                      # noqa: E402

METRIC = "TSDF voxel-updates/s + ray-samples/s (fwd+bwd)"
UNIT = "voxel-updates+ray-samples/s"
HBM_FALLBACK_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
N_POOL = 4                         # distinct frames of the 200-frame loop kept resident


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
        except Exception:
            pass
    return 2250.0, "nominal dense bf16 (B200_PROFILING.md)"


def make_frames(cfg, n_pool, first=0, stride=1):
    cam = cfg["cam"]
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    scene = synth.make_scene(cfg["mapping"]["bound"], 0)
    poses = synth.loop_trajectory(scene, 200)
    frames = []
    for i in range(n_pool):
        f = (first + i * stride) % 200
        depth, rgb = synth.render_frame(scene, K, cam["H"], cam["W"], poses[f], seed=f)
        frames.append((poses[f], depth, rgb))
    return K, poses, frames


class ClockSampler:
    """SM clock / throttle-reason samples during the timed region.  In-process NVML polling (pynvml, 20 Hz): a spawned
    `nvidia-smi -lms` costs tens of ms of driver stalls per query burst and would perturb the region it observes."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index=0):
        self.rows, self.first, self.stop_flag, self.thread, self.h, self.nv = [], 0, False, None, None, None
        self.gpu = gpu_index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nv = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((float(sm), float(mx), int(rs)))
            except Exception:
                pass
            time.sleep(0.05)

    def mark(self):
        """Rows before this call (warm-up) are not part of the timed region."""
        self.first = len(self.rows)

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"]}
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        rows = self.rows[self.first:] or self.rows[-1:]
        sm = [r[0] for r in rows]; mx = [r[1] for r in rows]
        reasons = sorted({name for r in rows for name, bit in self.REASONS if r[2] & bit})
        busy = sorted(sm)[len(sm) // 2:] if sm else []          # upper half = samples under load
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# =====================================================================================================================
# reference arm / cpu_baseline: the reference's CPU path on the host cores
# =====================================================================================================================
def cpu_reference_step(cfg, K, frame, n_rays, threads, state=None):
    """One bounded CPU step: full TSDF fuse (C oracle, all threads) + mapping fwd+bwd on `n_rays` rays (torch CPU
    restatement of the reference's scene_rep + stand-in encoders).  Returns timings and unit counts."""
    from oracle import tsdf_oracle as O
    from oracle import ray_oracle as RC
    from oracle.ray_oracle import RayOracle
    c2w, depth, rgb = frame
    cam = cfg["cam"]
    if state is None:
        state = {}
        v = cfg["volume"]
        center = np.round(c2w[:3, 3], 0)
        lens = np.array([v["x_config"]["len"], v["y_config"]["len"], v["z_config"]["len"]], dtype=np.float64)
        bnds = np.stack([center - lens, center + lens], 1)
        dims = np.ceil((bnds[:, 1] - bnds[:, 0]) / v["voxel_size"]).astype(int)
        n = int(dims.prod())
        state.update(dims=dims, origin=bnds[:, 0].astype(np.float32), tsdf=np.ones(n, np.float32), w=np.zeros(n, np.float32),
                     col=np.zeros(n, np.float32))
        R = cfg["globalV"]["base_resolution"]
        trgb = np.zeros(4 * R ** 3, np.float32); O.clear_global(trgb)
        state.update(R=R, trgb=trgb, gw=np.zeros(R ** 3, np.float32))
        torch.manual_seed(0)
        h = RC.hash_standin(cfg); g = RC.gbv_standin(cfg)
        hd = cfg["decoder"]["hidden_dim"]
        ws = [torch.nn.Parameter((torch.rand(o, i) * 2 - 1) / np.sqrt(i)) for o, i in ((hd, 81), (16, hd), (hd, 66), (3, hd))]
        bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
        state.update(oracle=RayOracle(cfg, bb, h, g, *ws), hash=h, gbv=g, ws=ws)
    t0 = time.perf_counter()
    packed = O.pack_bgr(np.floor(rgb * 255.0).astype(np.float32))
    nt_l, _ = O.integrate_local(state["tsdf"], state["w"], state["col"], state["dims"], state["origin"], cfg["volume"]["voxel_size"],
                                K, c2w, depth, packed, cfg["volume"]["trunc"], threads=threads)
    box = [v for ax in cfg["mapping"]["bound"] for v in ax]
    nt_g = O.integrate_global(state["trgb"], state["gw"], state["R"], box, K, c2w, depth, rgb, cfg["training"]["c_trunc"], threads=threads)
    t_tsdf = time.perf_counter() - t0
    # rays of this frame (mp_slam/mapper.py:337-344)
    with torch.no_grad():
        state["gbv"].params.copy_(torch.from_numpy(state["trgb"]))
    state["gbv"].params.requires_grad_(False)
    rng = np.random.default_rng(0)
    pix = rng.choice(cam["H"] * cam["W"], n_rays, replace=False)
    dirs = torch.from_numpy(synth.camera_dirs(K, cam["H"], cam["W"]).reshape(-1, 3)[pix])
    c2w_t = torch.from_numpy(c2w.astype(np.float32))
    rays_d = torch.sum(dirs[..., None, :] * c2w_t[:3, :3], -1)
    rays_o = c2w_t[None, :3, -1].repeat(n_rays, 1)
    tgt_d = torch.from_numpy(depth.reshape(-1)[pix])[:, None]
    tgt_c = torch.from_numpy(rgb.reshape(-1, 3)[pix])
    S = cfg["training"]["n_range_d"] + cfg["training"]["n_samples_d"]
    t1 = time.perf_counter()
    orc = state["oracle"]
    for p in [state["hash"].params] + state["ws"]:
        p.grad = None
    ret = orc.mapping(rays_o, rays_d, tgt_c, tgt_d, u=torch.rand(n_rays, S))
    orc.total_loss(ret).backward()
    t_ray = time.perf_counter() - t1
    return state, dict(t_tsdf=t_tsdf, t_ray=t_ray, touched=nt_l + nt_g, samples=n_rays * S)


def run_reference(args, rank, world):
    """The reference's CPU path on the host cores: each step = one full TSDF fuse + mapping fwd+bwd on a BOUNDED ray sample
    (`--cpu-rays`, default 65536 of the frame's 816000).  `ms_per_step` is the time actually measured per step; `value` is the
    full-workload throughput composed from the two measured rates (TSDF at full size, rays scaled linearly by `scale_factor`
    <= 12.5) and flagged `extrapolated`; `measured_units_per_s` is the plain units / time of the sample itself."""
    if rank != 0:
        return
    cfg = configs.replica()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    K, poses, frames = make_frames(cfg, 1)
    cam = cfg["cam"]
    S = cfg["training"]["n_range_d"] + cfg["training"]["n_samples_d"]
    full_samples = cam["H"] * cam["W"] * S
    n_rays = min(args.cpu_rays, cam["H"] * cam["W"])
    state = None
    if args.warmup > 0:
        state, _r = cpu_reference_step(cfg, K, frames[0], min(n_rays, 4096), threads, state)
    ts = []
    for _ in range(max(1, min(args.steps, args.cpu_steps))):
        state, r = cpu_reference_step(cfg, K, frames[0], n_rays, threads, state)
        ts.append(r)
    t_tsdf = statistics.mean(x["t_tsdf"] for x in ts)
    t_ray = statistics.mean(x["t_ray"] for x in ts)
    touched = ts[-1]["touched"]; samples = ts[-1]["samples"]
    scale = full_samples / samples
    t_full = t_tsdf + t_ray * scale                                          # ray part scaled linearly to the full frame
    value = (touched + full_samples) / t_full
    sample = (f"per step: full TSDF fuse of one 1200x680 frame (400x400x300 local + 200^3 GBV, C oracle, {threads} threads: {t_tsdf:.2f} s) + mapping "
              f"fwd+bwd on {n_rays} rays x {S} (torch CPU restatement of the reference's scene_rep + stand-in encoders: {t_ray:.2f} s); value = "
              f"full-workload throughput with the ray time scaled x{scale:.2f} to 816000 rays")
    extra = {"extrapolated": True, "measured_rays": n_rays, "scale_factor": scale, "measured_units_per_s": (touched + samples) / (t_tsdf + t_ray),
             "measured_steps": len(ts), "ms_per_step_full_workload_scaled": t_full * 1e3}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(ts),
            "warmup": min(args.warmup, 1), "ms_per_step": (t_tsdf + t_ray) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic analytic scene (seeded)",
            "config": workload_config(cfg, 1),
            "parts": {"tsdf_voxel_updates_per_s": touched / t_tsdf, "ray_samples_per_s": samples / t_ray},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, **extra},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, **extra}
    print(json.dumps(line), flush=True)


def workload_config(cfg, world):
    cam, t = cfg["cam"], cfg["training"]
    return {"workload": "BASELINE config 2: Replica-shaped 1200x680 frame -> TSDF fuse (local 400x400x300 @ 2 cm + GBV 200^3) "
                        "+ full-frame mixed ray render fwd+bwd (816000 rays x 59 samples)",
            "frame": f"{cam['W']}x{cam['H']}", "rays_per_gpu": cam["H"] * cam["W"], "samples_per_ray": t["n_range_d"] + t["n_samples_d"],
            "hash": f"16 levels x 2^{cfg['grid']['hash_size']}", "hidden": cfg["decoder"]["hidden_dim"],
            "l2": "inputs larger than L2 (576 MB local volume, 160 MB GBV, >2 GB ray buffers per step); no explicit flush",
            "parallelism": f"dp{world}: x-slab/z-slab sharded volumes + frame broadcast; ray batch per rank + grad all-reduce"}


# =====================================================================================================================
# GPU arm
# =====================================================================================================================
def run_gpu(args, rank, world, local_rank):
    import torch.distributed as dist
    from remixfusion_b200 import abi, dist as rdist
    from remixfusion_b200.global_volume import MapVolume
    from remixfusion_b200.scene_rep import JointEncoding
    from remixfusion_b200.volume import moving_volume
    abi.lib()                                                   # fail loudly if the CUDA library is missing
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg = configs.replica(hidden=args.hidden, hash_size=args.hash_size)
    cam = cfg["cam"]
    H, W = cam["H"], cam["W"]
    S = cfg["training"]["n_range_d"] + cfg["training"]["n_samples_d"]
    group = dist.group.WORLD if world > 1 else None

    # frames: the TSDF frame is rank 0's (broadcast); each rank renders rays of its own frame
    # RF_BENCH_SAME_FRAMES=1 (diagnostic): every rank renders the SAME frames, which removes the per-step load imbalance between ranks
    K, poses, frames = make_frames(cfg, N_POOL, first=0 if os.environ.get("RF_BENCH_SAME_FRAMES") else 17 * rank, stride=50)
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
    torch.manual_seed(0)
    model = JointEncoding(cfg, bb, process_group=group, equal_shards=True).to(dev)      # every rank renders one full frame
    with torch.no_grad():
        model.embed_res_fn.params.copy_((torch.rand_like(model.embed_res_fn.params) * 2 - 1) * 1e-2)
    model.train()
    params = [model.embed_res_fn.params] + list(model.decoder_res.fused_weights())
    # The optimiser step of the mapping loop closes every timed step (mp_slam/mapper.py:417-423; groups of mp_slam/slam.py:271-286):
    # gradients and parameters live in two flat buffers; with several ranks the step is reduce-scatter -> fused Adam on the
    # owned shard -> all-gather (remixfusion_b200.dist.ShardedAdam), on one GPU the fused Adam alone.  It also clears the gradients.
    fg = rdist.ShardedAdam([{"params": params[1:], "weight_decay": 1e-6, "lr": 1e-2}, {"params": params[:1], "eps": 1e-15, "lr": 1e-2}],
                           betas=(0.9, 0.99), group=group)

    R = cfg["globalV"]["base_resolution"]
    z_slab = rdist.slab(R, rank, world)
    full_gbv = model.GBV                                        # replicated copy used by the ray query
    if world > 1:
        slab_model = type("M", (), {})()
        slab_model.GBV = type("E", (), {})(); slab_model.GBW = type("E", (), {})()
        slab_model.GBV.params = torch.zeros(4 * (z_slab[1] - z_slab[0]) * R * R, device=dev)
        slab_model.GBW.params = torch.zeros((z_slab[1] - z_slab[0]) * R * R, device=dev)
        mvol = MapVolume(cfg, slab_model, K, z_slab=z_slab)
    else:
        mvol = MapVolume(cfg, model, K)
    mvol.init_mapvolume()
    dx = int(np.ceil(2 * cfg["volume"]["x_config"]["len"] / cfg["volume"]["voxel_size"]))
    x_slab = rdist.slab(dx, rank, world)
    local = moving_volume(cfg, None, poses[0], device=dev, x_slab=x_slab if world > 1 else None)

    # resident inputs
    dirs = torch.from_numpy(synth.camera_dirs(K, H, W).reshape(-1, 3)).to(dev)
    dev_frames = []
    for c2w, depth, rgb in frames:
        d = torch.from_numpy(depth).to(dev); c = torch.from_numpy(rgb).to(dev)
        packed = torch.empty(H * W, device=dev)
        abi.check(abi.lib().rf_pack_bgr(abi.dptr(torch.floor(c * 255.0).contiguous()), abi.dptr(packed), H * W, abi.stream_ptr()), "pack")
        c2w_t = torch.from_numpy(c2w.astype(np.float32)).to(dev)
        rays_d = torch.sum(dirs[..., None, :] * c2w_t[:3, :3], -1).contiguous()          # mp_slam/mapper.py:344
        rays_o = c2w_t[None, :3, -1].repeat(H * W, 1).contiguous()
        dev_frames.append(dict(c2w=c2w, c2w_t=c2w_t, depth=d, rgb=c, packed=packed, rays_o=rays_o, rays_d=rays_d,
                               tgt_d=d.reshape(-1, 1).contiguous(), tgt_c=c.reshape(-1, 3).contiguous()))
    # the frame every rank fuses is rank 0's
    bc = [dict(depth=f["depth"].clone(), rgb=f["rgb"].clone(), packed=f["packed"].clone(), c2w=f["c2w"].copy()) for f in dev_frames]
    if world > 1:
        for f in bc:
            rdist.broadcast_frame(f["depth"], f["rgb"], 0, group); dist.broadcast(f["packed"], 0, group=group)
            pose = torch.from_numpy(f["c2w"]).to(dev); dist.broadcast(pose, 0, group=group); f["c2w"] = pose.cpu().numpy()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    stage_ms = {"tsdf_local": 0.0, "tsdf_global": 0.0, "ray_fwd": 0.0, "ray_bwd": 0.0}
    units = {"touched_local": 0, "touched_global": 0, "samples": 0}

    def step(i, timed):
        f = dev_frames[i % N_POOL]; b = bc[i % N_POOL]
        e = [ev() for _ in range(5)] if timed else None
        if world > 1:                                            # frame broadcast over NVLink (data path of the sharded fuse)
            rdist.broadcast_frame(b["depth"], b["rgb"], 0, group)
        if timed: e[0].record()
        local.integrate_packed(b["depth"], b["packed"], K, b["c2w"], None, 1.0, 0.0)
        if timed: e[1].record()
        mvol.integrate_kf({"rgb": b["rgb"], "depth": b["depth"]}, torch.from_numpy(b["c2w"]).float(), 1.0)
        if world > 1:                                            # replicate the GBV for the ray query: slabs land in place (160 MB)
            sizes = [4 * (rdist.slab(R, k, world)[1] - rdist.slab(R, k, world)[0]) * R * R for k in range(world)]
            rdist.gather_slabs(mvol.model.GBV.params, sizes, group, out=full_gbv.params.data)
        if timed: e[2].record()
        if timed and os.environ.get("RF_BENCH_DEBUG"):
            dbg = torch.cuda.Event(enable_timing=True); dbg.record(); e.append(dbg)
        ret = model.mapping(f["rays_o"], f["rays_d"], f["tgt_c"], f["tgt_d"])
        loss = configs.total_loss(cfg, ret)
        if timed: e[3].record()
        loss.backward()
        fg.step()
        if timed: e[4].record()
        return e, loss

    def count_units(i):
        """(touched local, touched GBV) of this rank's slabs for frame i — same predicate as the integrate kernels."""
        b = bc[i % N_POOL]
        tl, _ = local.count_touched(b["depth"], K, b["c2w"])
        tg = mvol.count_touched(b["depth"], torch.from_numpy(b["c2w"]).float())
        return tl, tg

    # ---- warm-up ------------------------------------------------------------------------------------------------
    # the clock sampler is spawned here: NVML start-up stalls the driver for tens of ms and must not land in the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("RF_BENCH_NO_SAMPLER"):
        sampler.start()
    for i in range(args.warmup):
        # keep the previous step's loss referenced while the next step runs, exactly as the timed loop does: its autograd
        # node owns the 7.7 GB feature workspace, so two workspaces are live at a time and the caching allocator must
        # already hold both blocks (a first cudaMalloc of the second one inside the timed region stalls the host 5-60 ms)
        _, loss = step(i, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    # ---- unit counts (outside the timed region; same frames, same order) -------------------------------------------
    per_frame_units = []
    for i in range(N_POOL):
        per_frame_units.append(count_units(i))

    # ---- timed region ---------------------------------------------------------------------------------------------
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.mark()
    t_start, t_end = ev(), ev()
    t_start.record()
    evs = []
    host_t = []
    for i in range(args.steps):
        h0 = time.perf_counter()
        e, loss = step(args.warmup + i, True)
        host_t.append((time.perf_counter() - h0) * 1e3)
        evs.append(e)
    t_end.record()
    if os.environ.get("RF_BENCH_DEBUG"):
        print(f"[rank {rank}] host ms per step: " + " ".join(f"{x:.2f}" for x in host_t), file=sys.stderr, flush=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = t_start.elapsed_time(t_end)
    for e in evs:
        if len(e) > 5:
            print(f"[rank {rank}] e1->e2 {e[1].elapsed_time(e[2]):.2f} e2->dbg {e[2].elapsed_time(e[5]):.2f} dbg->e3 {e[5].elapsed_time(e[3]):.2f} "
                  f"e3->e4 {e[3].elapsed_time(e[4]):.2f}", file=sys.stderr, flush=True)
        stage_ms["tsdf_local"] += e[0].elapsed_time(e[1]); stage_ms["tsdf_global"] += e[1].elapsed_time(e[2])
        stage_ms["ray_fwd"] += e[2].elapsed_time(e[3]); stage_ms["ray_bwd"] += e[3].elapsed_time(e[4])
    for i in range(args.steps):
        tl, tg = per_frame_units[(args.warmup + i) % N_POOL]
        units["touched_local"] += tl; units["touched_global"] += tg; units["samples"] += H * W * S

    # ---- per-kernel timing of the dominant kernels (CUDA events on the launching stream, resident inputs) ------------
    kern = time_kernels(model, cfg, dev_frames[0], dev, params)

    # ---- e2e: public API with HOST buffers, H2D/D2H inside the timed region ----------------------------------------
    e2e = run_e2e(args, cfg, K, frames, model, mvol, local, params, fg, dev, rank, world, group, per_frame_units, S)

    # ---- the step's collectives timed on their own (events on the launching stream, max over ranks): what the N > 1 step adds ----
    comm_ms = None
    if world > 1:
        def _t(fn, iters=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(); dist.barrier()
            a_, b_ = ev(), ev()
            a_.record()
            for _ in range(iters):
                fn()
            b_.record(); torch.cuda.synchronize()
            t_ = torch.tensor([a_.elapsed_time(b_) / iters], dtype=torch.float64, device=dev)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX, group=group)
            return float(t_)
        sizes_ = [4 * (rdist.slab(R, k, world)[1] - rdist.slab(R, k, world)[0]) * R * R for k in range(world)]
        part_ = torch.zeros(8, dtype=torch.float64, device=dev)
        comm_ms = {"frame_broadcast": _t(lambda: rdist.broadcast_frame(bc[0]["depth"], bc[0]["rgb"], 0, group)),
                   "gbv_slab_all_gather": _t(lambda: rdist.gather_slabs(mvol.model.GBV.params, sizes_, group, out=full_gbv.params.data)),
                   "loss_sums_all_reduce": _t(lambda: dist.all_reduce(part_, group=group)),
                   "grad_reduce_scatter_adam_all_gather": _t(lambda: fg.step()),
                   "bytes": {"frame": int(16 * H * W), "gbv": int(16 * R ** 3), "grads": int(fg.gflat.numel() * 4)}}

    # ---- the other BASELINE configurations (bench_workloads.py), every rank takes part --------------------------------
    extra_parts = {}
    if not args.no_extra:
        import bench_workloads as BW
        del loss
        local = mvol = None
        torch.cuda.empty_cache()
        peak_gbs = measured_peak()[0]
        for name, fn in (("cfg1", (lambda: BW.cfg1_part(dev, peak_gbs)) if world == 1 else None),
                         ("cfg3", lambda: BW.cfg3_part(dev, rank, world, group, frames, K, peak_gbs)),
                         ("cfg4", lambda: BW.cfg4_part(dev, rank, world, group, peak_gbs)),
                         ("cfg5", lambda: BW.cfg5_part(dev, rank, world, group))):
            if fn is None:
                continue
            try:
                extra_parts[name] = fn()
            except Exception as ex:                           # a failing side workload must not take the headline line with it
                if world > 1:
                    raise                                     # (with several ranks a one-sided failure would hang the collectives)
                extra_parts[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    # ---- reduce over ranks ------------------------------------------------------------------------------------------
    vec = torch.tensor([total_ms, e2e["ms"]], dtype=torch.float64, device=dev)
    cnt = torch.tensor([units["touched_local"], units["touched_global"], units["samples"]], dtype=torch.float64, device=dev)
    e2e_units = torch.tensor([float(e2e["units"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(e2e_units, op=dist.ReduceOp.SUM, group=group)
    total_ms, e2e_ms = float(vec[0]), float(vec[1])
    touched = float(cnt[0] + cnt[1]); samples = float(cnt[2])
    if rank != 0:
        return
    secs = total_ms / 1e3
    peak, peak_src = measured_peak()
    P = H * W * S
    tf_peak, tf_src = measured_tensor_peak()
    peaks = (peak, peak_src, tf_peak, tf_src)
    dom = max(kern["kernels_ms"].items(), key=lambda kv: kv[1])
    roof = kernel_roofline(dom[0], dom[1], P, cfg["decoder"]["hidden_dim"], peaks)
    traffic, traffic_src = ncu_traffic()
    if dom[0] in traffic and cfg["decoder"]["hidden_dim"] == 32 and cfg["grid"]["hash_size"] == 16:
        roof["traffic"] = traffic[dom[0]]
        roof["traffic_source"] = f"profiles/{traffic_src} (ncu --set full, same workload)"
    tl0, tg0 = per_frame_units[0]
    hw_bytes = 8.0 * H * W
    line = {
        "metric": METRIC, "value": (touched + samples) / secs, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic analytic scene (seeded), random-init hash table / decoder",
        "config": workload_config(cfg, world),
        "parts": {
            "tsdf_voxel_updates_per_s": touched / ((stage_ms["tsdf_local"] + stage_ms["tsdf_global"]) / 1e3) if touched else None,
            "ray_samples_per_s_fwd_bwd": samples / world / ((stage_ms["ray_fwd"] + stage_ms["ray_bwd"]) / 1e3),
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
            "touched_local_per_frame": tl0, "touched_global_per_frame": tg0,
            # BA mode (gradients w.r.t. rays_o / rays_d as well; SURVEY §8d: 3328 algorithmic bytes per sample), same frame
            "ba_mode": ({"ray_samples_per_s_fwd_bwd": P / ((kern["ba_fwd_ms"] + kern["ba_bwd_ms"]) / 1e3),
                         "launch_ms": [kern["ba_fwd_ms"], kern["ba_bwd_ms"]], "kernels_ms": kern["ba_kernels_ms"],
                         "hbm_form": {"achieved": 3328.0 * P / ((kern["ba_fwd_ms"] + kern["ba_bwd_ms"]) / 1e3) / 1e9, "peak": peak,
                                      "frac": 3328.0 * P / ((kern["ba_fwd_ms"] + kern["ba_bwd_ms"]) / 1e3) / 1e9 / peak}}
                        if "ba_fwd_ms" in kern else None),
        },
        "roofline": roof,
        "roofline_parts": {
            "kernels": {k: kernel_roofline(k, v, P, cfg["decoder"]["hidden_dim"], peaks) for k, v in kern["kernels_ms"].items()},
            "ray_fwd_bwd_hbm_form": {"achieved": 2176.0 * P / ((kern["sample_fwd_ms"] + kern["sample_bwd_ms"]) / 1e3) / 1e9, "peak": peak,
                                     "frac": 2176.0 * P / ((kern["sample_fwd_ms"] + kern["sample_bwd_ms"]) / 1e3) / 1e9 / peak,
                                     "launch_ms": [kern["sample_fwd_ms"], kern["sample_bwd_ms"]]},
            "tsdf_local": {"achieved": (16.0 * kern["touched_local"] + 8.0 * kern["band_local"] + hw_bytes) / (kern["tsdf_local_ms"] / 1e3) / 1e9,
                           "peak": peak, "launch_ms": kern["tsdf_local_ms"],
                           "frac": (16.0 * kern["touched_local"] + 8.0 * kern["band_local"] + hw_bytes) / (kern["tsdf_local_ms"] / 1e3) / 1e9 / peak,
                           "swept_voxels_per_s": kern["swept_local"] / (kern["tsdf_local_ms"] / 1e3)},
            "tsdf_global": {"achieved": (40.0 * kern["touched_global"] + 16.0 * H * W) / (kern["tsdf_global_ms"] / 1e3) / 1e9, "peak": peak,
                            "frac": (40.0 * kern["touched_global"] + 16.0 * H * W) / (kern["tsdf_global_ms"] / 1e3) / 1e9 / peak,
                            "launch_ms": kern["tsdf_global_ms"]},
            "tsdf_recenter": {"achieved": 24.0 * kern["swept_local"] / (kern["tsdf_recenter_ms"] / 1e3) / 1e9, "peak": peak,
                              "frac": 24.0 * kern["swept_local"] / (kern["tsdf_recenter_ms"] / 1e3) / 1e9 / peak,
                              "launch_ms": kern["tsdf_recenter_ms"], "algorithmic_bytes_per_voxel": 24.0},
            "gather_peak_Gops": kern.get("gather_gops"), "atomic_peak_Gops": kern.get("atomic_gops"),
            # SURVEY §8d gather-bound roofline: t_roof = P*136/G_peak + P*256/A_peak (peaks measured above); frac = t_roof / t_measured
            "ray_fwd_bwd_gather_form": gather_form(P, kern),
        },
        "e2e": {"value": float(e2e_units[0]) / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"], "steps": e2e["steps"],
                "ms_per_step": e2e_ms / e2e["steps"], "h2d_gbs_measured": e2e["h2d_gbs"],
                "inputs": "per step: colour + depth frame and the pose from pinned host memory; rays, 0..255 colour and targets derived on the device"},
        # our kernels per step: 2 TSDF integrates, ray_z, encode walk, decoder fwd, composite fwd, loss finalisation, composite bwd,
        # tile liveness, decoder bwd, scatter walk, replica fold, fused Adam x 2 (one launch per parameter-group segment)
        "gpu_launches": 14 * args.steps,
        "clocks": clocks,
    }
    line["parts"].update(extra_parts)
    if comm_ms is not None:
        line["parts"]["collectives_ms"] = comm_ms
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(cfg, K, frames[0], H * W * S, args.cpu_rays)
    print(json.dumps(line), flush=True)


def gather_form(P, kern):
    g, a = kern.get("gather_gops"), kern.get("atomic_gops")
    if not g or not a:
        return None
    t_roof = (P * 136.0 / (g * 1e9) + P * 256.0 / (a * 1e9)) * 1e3
    t_meas = kern["sample_fwd_ms"] + kern["sample_bwd_ms"]
    return {"t_roof_ms": t_roof, "t_measured_ms": t_meas, "frac": t_roof / t_meas}


PROF_NAMES = {4: "encode_walk4_kernel", 5: "mlp_fwd_tc_kernel", 6: "composite_fwd_kernel",
              7: "composite_bwd_kernel", 8: "mlp_bwd_tc2_kernel", 9: "scatter_walk4_kernel", 10: "sample_fwd_kernel",
              11: "sample_bwd_kernel"}


def ncu_traffic():
    """DRAM bytes per launch of each kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the newest committed
    `ncu --set full` summary under profiles/ — taken on this same workload (bench config 2), one capture per kernel."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_full_summary.csv")))
    if not files:
        return {}, None
    rows = list(csv.reader(open(files[-1])))
    names = rows[0][2:]
    val = {r[0]: r for r in rows[1:]}
    out = {}
    try:
        rd, wr = val["dram__bytes_read.sum"], val["dram__bytes_write.sum"]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        for i, n in enumerate(names):
            key = next((k for k in PROF_NAMES.values() if k.replace("_kernel", "") in n or n.startswith(k[:12])), None)
            if key is None and "mlp_fwd" in n: key = "mlp_fwd_tc_kernel"
            if key is None and "mlp_bwd" in n: key = "mlp_bwd_tc2_kernel"
            if key is None and "encode_walk" in n: key = "encode_walk4_kernel"
            if key is None and "scatter_walk" in n: key = "scatter_walk4_kernel"
            if key and key not in out:          # first launch of a kernel = the mapping-mode step (BA-mode launches follow)
                out[key] = float(rd[2 + i]) * scale.get(rd[1], 1.0) + float(wr[2 + i]) * scale.get(wr[1], 1.0)
    except Exception:
        return {}, None
    return out, os.path.basename(files[-1])


def kernel_roofline(name, ms, P, hidden, peaks):
    """Roofline entry of one kernel.  Algorithmic work per sample: gathers 1152 B (16 levels x 8
    corners x 8 B + 8 GBV corners x 16 B), table-gradient reductions 1024 B (16 x 8 x 2 x 4 B), decoder 2*166*h FLOP
    forward and twice that backward."""
    hbm, hbm_src, tf, tf_src = peaks
    if name in ("mlp_fwd_tc_kernel", "mlp_bwd_tc2_kernel"):
        flop = 2.0 * 166 * hidden * (1 if name == "mlp_fwd_tc_kernel" else 2)
        a = flop * P / (ms / 1e3) / 1e12
        return {"bound": "tensor", "kernel": name, "achieved": a, "peak": tf, "unit": "TFLOP/s", "frac": a / tf, "traffic": None,
                "peak_source": tf_src, "algorithmic_flop_per_sample": flop, "launch_ms": ms}
    alg = {"encode_walk4_kernel": 1152.0, "sample_fwd_kernel": 1152.0, "scatter_walk4_kernel": 1024.0, "sample_bwd_kernel": 1024.0}.get(name, 40.0)
    a = alg * P / (ms / 1e3) / 1e9
    out = {"bound": "hbm", "kernel": name, "achieved": a, "peak": hbm, "unit": "GB/s", "frac": a / hbm, "traffic": None,
           "peak_source": hbm_src, "algorithmic_bytes_per_sample": alg, "launch_ms": ms}
    if name in ("encode_walk4_kernel", "scatter_walk4_kernel"):
        # the table (8 MiB at T = 2^16) is L2-resident and a walking thread touches it once per RUN of samples in a cell, not once per
        # sample: SURVEY's per-sample gather / reduction bytes are resolved in L2 (measured peaks: 290 G random 8-byte loads/s,
        # 194 G random RED/s), so this fraction can exceed 1; the HBM bytes the launch really moves are in `traffic`
        out["note"] = "per-sample gather/reduction bytes of SURVEY 8(d); table L2-resident, one access per run of samples in a cell: not an HBM-bound kernel"
    return out


def time_kernels(model, cfg, f, dev, params):
    """Average launch duration of each hot kernel with CUDA events on the launching stream, inputs resident."""
    import ctypes as C
    from remixfusion_b200 import abi
    from remixfusion_b200.global_volume import MapVolume
    from remixfusion_b200.volume import moving_volume
    out = {}
    cam = cfg["cam"]; H, W = cam["H"], cam["W"]
    Kmat = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timeit(fn, iters=5):
        fn(); torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(iters):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    vol = moving_volume(cfg, None, f["c2w"], device=dev)
    # launch durations of the two integrate kernels: events the library records around the launch (the host call around a 30-60 us
    # kernel is Python-bound, so events around the call would time the host)
    from bench_workloads import _kernel_ms
    out["tsdf_local_ms"] = _kernel_ms(lambda i: vol.integrate_packed(f["depth"], f["packed"], Kmat, f["c2w"], None, 1.0, 0.0), 0, iters=5, warm=1)
    tl, tb = vol.count_touched(f["depth"], Kmat, f["c2w"])
    out.update(touched_local=tl, band_local=tb, swept_local=int(np.prod(vol.vol_dim)))
    # N2: re-centring of the moving volume by one metre along x (ping-pong arrays: 12 B read + 12 B written per voxel)
    Lb = abi.lib(); Lb.rf_profile_enable(1)
    b0 = vol.vol_bnds.copy(); acc_r = 0.0
    for i in range(6):
        b1 = b0 + (np.array([[1.0], [0.0], [0.0]]) if i % 2 == 0 else 0.0)
        vol.update_tsdf_swap_rot_trans(b1.copy(), vol.vol_bnds.copy())
        buf = (C.c_float * 64)(); Lb.rf_profile_read(buf)
        if i >= 2: acc_r += buf[13]
    Lb.rf_profile_enable(0)
    out["tsdf_recenter_ms"] = acc_r / 4
    del vol
    m2 = type("M", (), {})(); m2.GBV = type("E", (), {})(); m2.GBW = type("E", (), {})()
    R = cfg["globalV"]["base_resolution"]
    m2.GBV.params = torch.zeros(4 * R ** 3, device=dev); m2.GBW.params = torch.zeros(R ** 3, device=dev)
    gv = MapVolume(cfg, m2, Kmat); gv.init_mapvolume()
    pose = torch.from_numpy(f["c2w"]).float()
    out["touched_global"] = gv.count_touched(f["depth"], pose)
    out["tsdf_global_ms"] = _kernel_ms(lambda i: gv.integrate_kf({"rgb": f["rgb"], "depth": f["depth"]}, pose, 1.0), 1, iters=5, warm=1)
    del gv, m2
    # ray kernels through the C-ABI directly
    n = H * W
    S = cfg["training"]["n_range_d"] + cfg["training"]["n_samples_d"]
    z = model.sample_z(f["tgt_d"], n)
    meta = model._meta(True)
    w = model.decoder_res.fused_weights()
    p = abi.RayParams(abi.dptr(model.embed_res_fn.params.detach()), abi.dptr(model.GBV.params.detach()), abi.dptr(w[0].detach()),
                      abi.dptr(w[1].detach()), abi.dptr(w[2].detach()), abi.dptr(w[3].detach()))
    raw = torch.empty(n, S, 4, device=dev); rgbm = torch.empty(n, 3, device=dev); dm = torch.empty(n, device=dev)
    part = torch.zeros(8, dtype=torch.float64, device=dev)
    td = f["tgt_d"].reshape(-1).contiguous()
    L = abi.lib()
    cfgc = meta["cfg"]; cfgc.n_rays_total = n
    nws = int(L.rf_ray_workspace_floats(C.byref(cfgc), C.byref(meta["hash_desc"]), C.c_int64(n)))
    ws = torch.empty(nws, device=dev) if nws > 0 else None

    def fwd():
        part.zero_()
        abi.check(L.rf_ray_query_forward(C.byref(cfgc), C.byref(meta["hash_desc"]), C.byref(meta["gbv_desc"]), C.byref(p), abi.dptr(f["rays_o"]),
                                         abi.dptr(f["rays_d"]), abi.dptr(td), abi.dptr(f["tgt_c"]), abi.dptr(z), C.c_int64(n), abi.dptr(raw),
                                         abi.dptr(rgbm), abi.dptr(dm), abi.dptr(part), abi.dptr(ws), abi.stream_ptr()), "fwd")
    g_hash = torch.zeros_like(model.embed_res_fn.params); gws = [torch.zeros_like(x) for x in w]
    grads = abi.RayGrads(abi.dptr(g_hash), abi.dptr(gws[0]), abi.dptr(gws[1]), abi.dptr(gws[2]), abi.dptr(gws[3]), None, None)
    scratch = torch.empty(int(L.rf_ray_scratch_floats(C.byref(cfgc), C.byref(meta["hash_desc"]), C.c_int64(n), C.c_int(0))), device=dev)
    lg = torch.tensor([5.0, 0.1, 1000.0, 10.0], device=dev)

    def bwd():
        abi.check(L.rf_ray_query_backward(C.byref(cfgc), C.byref(meta["hash_desc"]), C.byref(meta["gbv_desc"]), C.byref(p), abi.dptr(f["rays_o"]),
                                          abi.dptr(f["rays_d"]), abi.dptr(td), abi.dptr(f["tgt_c"]), C.c_int64(n), abi.dptr(z), abi.dptr(raw),
                                          abi.dptr(rgbm), abi.dptr(dm), None, None, None, abi.dptr(lg), abi.dptr(part), C.byref(grads),
                                          abi.dptr(ws), abi.dptr(scratch), abi.stream_ptr()), "bwd")
    out["sample_fwd_ms"] = timeit(fwd, 3)          # whole forward call: position + encode + decoder + composite
    out["sample_bwd_ms"] = timeit(bwd, 3)          # whole backward call: composite bwd + decoder bwd + scatter
    # per-kernel launch durations: CUDA events recorded by the library around each launch, on the launching stream
    L.rf_profile_enable(1)
    acc = np.zeros(64); reps = 3
    for _ in range(reps):
        fwd(); bwd()
        buf = (C.c_float * 64)()
        L.rf_profile_read(buf)
        acc += np.maximum(np.array(buf[:], dtype=np.float64), 0.0)
    L.rf_profile_enable(0)
    out["kernels_ms"] = {PROF_NAMES[i]: acc[i] / reps for i in PROF_NAMES if acc[i] > 0}
    # BA mode (reported separately): clamp variant + gradients w.r.t. rays_o / rays_d on the same frame
    cfgb = model._ray_cfg(clamp=True)
    if cfgb.mlp_precision == cfgc.mlp_precision:
        cfgb.n_rays_total = n
        g_o = torch.empty(n, 3, device=dev); g_d = torch.empty(n, 3, device=dev)
        grads_ba = abi.RayGrads(abi.dptr(g_hash), abi.dptr(gws[0]), abi.dptr(gws[1]), abi.dptr(gws[2]), abi.dptr(gws[3]), abi.dptr(g_o), abi.dptr(g_d))
        del scratch
        scratch_ba = torch.empty(int(L.rf_ray_scratch_floats(C.byref(cfgb), C.byref(meta["hash_desc"]), C.c_int64(n), C.c_int(1))), device=dev)

        def fwd_ba():
            part.zero_()
            abi.check(L.rf_ray_query_forward(C.byref(cfgb), C.byref(meta["hash_desc"]), C.byref(meta["gbv_desc"]), C.byref(p), abi.dptr(f["rays_o"]),
                                             abi.dptr(f["rays_d"]), abi.dptr(td), abi.dptr(f["tgt_c"]), abi.dptr(z), C.c_int64(n), abi.dptr(raw),
                                             abi.dptr(rgbm), abi.dptr(dm), abi.dptr(part), abi.dptr(ws), abi.stream_ptr()), "fwd_ba")

        def bwd_ba():
            abi.check(L.rf_ray_query_backward(C.byref(cfgb), C.byref(meta["hash_desc"]), C.byref(meta["gbv_desc"]), C.byref(p), abi.dptr(f["rays_o"]),
                                              abi.dptr(f["rays_d"]), abi.dptr(td), abi.dptr(f["tgt_c"]), C.c_int64(n), abi.dptr(z), abi.dptr(raw),
                                              abi.dptr(rgbm), abi.dptr(dm), None, None, None, abi.dptr(lg), abi.dptr(part), C.byref(grads_ba),
                                              abi.dptr(ws), abi.dptr(scratch_ba), abi.stream_ptr()), "bwd_ba")
        out["ba_fwd_ms"] = timeit(fwd_ba, 3)
        out["ba_bwd_ms"] = timeit(bwd_ba, 3)
        L.rf_profile_enable(1)
        accb = np.zeros(64)
        for _ in range(reps):
            bwd_ba()
            buf = (C.c_float * 64)(); L.rf_profile_read(buf)
            accb += np.maximum(np.array(buf[:], dtype=np.float64), 0.0)
        L.rf_profile_enable(0)
        out["ba_kernels_ms"] = {"mlp_bwd_tc2_kernel(BA)": accb[8] / reps, "scatter_walk4_kernel(BA)": accb[9] / reps, "raygrad_walk_kernel": accb[12] / reps,
                                "sample_bwd_kernel": accb[11] / reps}
        out["ba_kernels_ms"] = {k: v for k, v in out["ba_kernels_ms"].items() if v > 0}
    if os.environ.get("RF_DEBUG_PER_LEVEL"):
        print("per-level ms: scatter", [round(acc[16 + l] / reps, 3) for l in range(16)], "encode", [round(acc[40 + l] / reps, 3) for l in range(17)], file=sys.stderr)
    # gather / atomic peaks over a 40 MiB table (SURVEY §8d denominators)
    tab = torch.zeros(40 * 1024 * 1024 // 4, device=dev)
    ms = C.c_float(0)
    nops = 1 << 30
    if L.rf_microbench_gather(abi.dptr(tab), C.c_int64(tab.numel() * 4), C.c_int64(nops), 3, C.byref(ms), abi.stream_ptr()) == 0:
        per = (nops + 148 * 8 * 256 - 1) // (148 * 8 * 256); per = (per + 7) // 8 * 8
        out["gather_gops"] = per * 148 * 8 * 256 / (ms.value / 1e3) / 1e9
    if L.rf_microbench_atomic(abi.dptr(tab), C.c_int64(tab.numel() * 4), C.c_int64(nops), 3, C.byref(ms), abi.stream_ptr()) == 0:
        per = (nops + 148 * 8 * 256 - 1) // (148 * 8 * 256)
        out["atomic_gops"] = per * 148 * 8 * 256 / (ms.value / 1e3) / 1e9
    return out


def run_e2e(args, cfg, K, frames, model, mvol, local, params, fg, dev, rank, world, group, per_frame_units, S):
    """Same step through the public API with HOST buffers: numpy frames into moving_volume.integrate / integrate_kf
    (H2D inside), rays + targets from pinned host memory, loss read back to the host."""
    import torch.distributed as dist
    from remixfusion_b200 import dist as rdist
    cam = cfg["cam"]; H, W = cam["H"], cam["W"]
    # Per-pixel camera directions are a constant of the camera: resident on the device, uploaded once (the reference keeps
    # `batch['direction']` per frame but it never changes).  Per step the host supplies what a frame IS: colour, depth and the pose;
    # the rays (mp_slam/mapper.py: rays_o = c2w[:3, -1], rays_d = sum(dirs[..., None, :] * c2w[:3, :3], -1)), the 0..255 colour of the
    # local volume and the ray targets (views of the same frame) are derived on the device inside the timed region.
    dirs_dev = torch.from_numpy(synth.camera_dirs(K, H, W).reshape(-1, 3).astype(np.float32)).to(dev)
    host = []
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    for c2w, depth, rgb in frames:
        host.append(dict(c2w=c2w, rgb_t=pin(rgb), depth_t=pin(depth), c2w_t=pin(c2w.astype(np.float32))))
    h2d = H * W * 4 + H * W * 12 + 64                      # depth, colour, pose
    d2h = 4 * 4

    # Inputs are staged the way a caller of the public API would: pinned host tensors copied with non_blocking=True on a
    # copy stream, one step ahead, so that the frame of step i+1 crosses PCIe while step i computes.  Every step's copies are
    # issued (and complete) inside the timed region; moving_volume.integrate / integrate_kf / JointEncoding.mapping receive the
    # device tensors (they accept host arrays too: those are staged through the objects' pinned rings on the compute stream).
    copy_stream = torch.cuda.Stream(device=dev)

    def prefetch(i):
        f = host[i % len(host)]
        with torch.cuda.stream(copy_stream):
            t = tuple(f[k].to(dev, non_blocking=True) for k in ("rgb_t", "depth_t", "c2w_t"))
        ev = torch.cuda.Event(); ev.record(copy_stream)
        return t, ev

    def step(i, pre, more):
        f = host[i % len(host)]
        nxt = prefetch(i + 1) if more else None
        (rgb_d, depth_d, c2w_d), ev = pre
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for t in (rgb_d, depth_d, c2w_d):
            t.record_stream(cur)
        rgb255_d = torch.floor(rgb_d * 255.0)
        rd = torch.sum(dirs_dev[:, None, :] * c2w_d[None, :3, :3], -1)
        ro = c2w_d[None, :3, 3].expand(rd.shape).contiguous()
        tc = rgb_d.reshape(-1, 3); td = depth_d.reshape(-1, 1)
        local.integrate(rgb255_d, depth_d, K, f["c2w"], None, 1.0, 0.0)
        mvol.integrate_kf({"rgb": rgb_d, "depth": depth_d}, torch.from_numpy(f["c2w"]).float(), 1.0)
        if world > 1:
            R = cfg["globalV"]["base_resolution"]
            sizes = [4 * (rdist.slab(R, k, world)[1] - rdist.slab(R, k, world)[0]) * R * R for k in range(world)]
            rdist.gather_slabs(mvol.model.GBV.params, sizes, group, out=model.GBV.params.data)
        ret = model.mapping(ro, rd, tc, td)
        loss = configs.total_loss(cfg, ret)
        loss.backward()
        fg.step()                                                # the optimiser step (ShardedAdam of the resident loop)
        # the step's result goes to page-locked host memory; the host reads it one step later, so that it is already enqueueing
        # the next step while this one runs (every step's losses are read inside the timed region, the last one before it ends)
        slot = res_host[i % 2]
        slot.copy_(torch.stack([ret["rgb_res_loss"], ret["depth_res_loss"], ret["sdf_res_loss"], ret["fs_res_loss"]]).detach(), non_blocking=True)
        done = torch.cuda.Event(); done.record()
        return (slot, done), nxt

    def read_result(res):
        slot, done = res
        done.synchronize()
        return float(slot.sum())

    blocking_read = bool(os.environ.get("RF_E2E_BLOCKING_READ"))     # A/B switch: read every step's losses before enqueueing the next step

    n_steps = max(2, min(args.steps, 8))
    res_host = [torch.empty(4, dtype=torch.float32).pin_memory() for _ in range(2)]
    # diagnostic: the host->device rate this box delivers for one frame (pinned, copy stream, nothing else running)
    prefetch(0); torch.cuda.synchronize()
    ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ca.record(copy_stream); _keep = prefetch(0); cb.record(copy_stream); torch.cuda.synchronize()
    h2d_gbs = h2d / (ca.elapsed_time(cb) * 1e-3) / 1e9
    del _keep
    pre = prefetch(0)
    for w in range(3):                                     # warm-up: allocator blocks of two live steps, pinned staging, NCCL channels
        r, pre = step(w, pre, w < 2)
        read_result(r)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    pre = prefetch(0)                                      # the first step's inputs are copied inside the timed region too
    prev = None
    for i in range(n_steps):
        res, pre = step(i, pre, i + 1 < n_steps)
        if blocking_read:
            read_result(res)
            continue
        if prev is not None:
            read_result(prev)
        prev = res
    if prev is not None:
        read_result(prev)
    b.record(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    units = 0
    for i in range(n_steps):
        tl, tg = per_frame_units[i % N_POOL]
        units += tl + tg + H * W * S
    return {"ms": a.elapsed_time(b), "steps": n_steps, "units": units, "h2d": h2d, "d2h": d2h, "h2d_gbs": h2d_gbs}


def cpu_baseline(cfg, K, frame, full_samples, n_rays=65536):
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    state, _ = cpu_reference_step(cfg, K, frame, 2048, threads, None)       # warm-up (allocations, first-touch)
    state, r = cpu_reference_step(cfg, K, frame, n_rays, threads, state)
    scale = full_samples / r["samples"]
    t_full = r["t_tsdf"] + r["t_ray"] * scale
    return {"value": (r["touched"] + full_samples) / t_full, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"full TSDF fuse of one 1200x680 frame (C oracle, {threads} threads: {r['t_tsdf']:.2f} s) + mapping fwd+bwd on "
                      f"{n_rays} rays x {r['samples'] // n_rays} (torch CPU restatement of the reference + stand-in encoders: {r['t_ray']:.2f} s), "
                      f"ray time scaled x{scale:.2f} to the full frame",
            "extrapolated": True, "measured_rays": n_rays, "scale_factor": scale,
            "measured_units_per_s": (r["touched"] + r["samples"]) / (r["t_tsdf"] + r["t_ray"]),
            "tsdf_voxel_updates_per_s": r["touched"] / r["t_tsdf"], "ray_samples_per_s": r["samples"] / r["t_ray"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--hidden", type=int, default=None, help="decoder width (default: Replica yaml, 32)")
    ap.add_argument("--hash-size", type=int, default=None, help="log2 hash-table size (default: Replica yaml, 16)")
    ap.add_argument("--cpu-rays", type=int, default=65536, help="rays per CPU reference step (816000 in the full frame)")
    ap.add_argument("--cpu-steps", type=int, default=6, help="cap on the timed CPU reference steps (each ~10 s)")
    ap.add_argument("--no-extra", action="store_true", help="skip the cfg 1 / 3 / 4 / 5 parts")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    if world > 1:
        from remixfusion_b200 import dist as rdist
        rdist.init_from_env("nccl")
    run_gpu(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
