"""The BASELINE.json configurations other than the headline one (config 2), measured inside the default `bench.py` run and
reported under `parts` of its JSON line (SURVEY.md §8d "Synthetic inputs"):

  cfg1  one synthetic 640x480 frame into a 256^3 volume — the local (z-fastest, packed colour) layout and the GBV layout at
        R = 256 — as achieved GB/s on the algorithmic bytes B1 = 16 N_touched + 8 N_band + 8 HW (local), 40 N_touched + 16 HW (GBV);
  cfg3  2^20 rays x 48 samples, hash 16 x 2^19 at resolution 512, 2 x 64 decoder, mapping mode and BA mode; with N ranks the
        FIXED batch is cut N ways (strong scaling) and the 40 MiB table gradient + decoder gradients are all-reduced;
  cfg4  BS3D-scale scene (50 x 50 x 10 m), GBV R = 512 and 1024, z-slabs over the ranks, 1280x720 frame broadcast, and the
        replicated copy refreshed by all-gathering only the frustum's [y, x] box of every slab;
  cfg5  the online mapping loop (mp_slam/mapper.py:366-520) on a uHumans2-shaped stream: per keyframe {integrate_kf; keyframe
        rays into the device ray store; 10 mapping iterations of 2048 rays (sample -> forward -> backward -> Adam, one CUDA graph);
        5 BA iterations (clamp variant + ray gradients)}.

Every timing is CUDA events on the launching stream between device-wide synchronisations, max over ranks."""
from __future__ import annotations

import numpy as np
import torch

from remixfusion_b200 import configs, synth


def _sync(world):
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def _timed(fn, iters, world, dev, warm=2):
    """Average milliseconds per call of fn(i): events around `iters` calls, max over ranks."""
    for i in range(warm):
        fn(i)
    _sync(world)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(warm + i)
    b.record()
    _sync(world)
    ms = torch.tensor([a.elapsed_time(b) / iters], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def _kernel_ms(fn, slot, iters=10, warm=3):
    """Average launch duration (ms) of the library kernel behind profile slot `slot` over `iters` calls of fn(i): CUDA events the
    library records around the launch on the launching stream (rf_profile_enable / rf_profile_read) — the host call around
    a 20-100 us kernel is Python-bound, so events around the call would time the host."""
    import ctypes as C
    from remixfusion_b200 import abi
    L = abi.lib()
    L.rf_profile_enable(1)
    acc = 0.0
    for i in range(warm + iters):
        fn(i)
        buf = (C.c_float * 64)(); L.rf_profile_read(buf)
        if i >= warm:
            acc += max(float(buf[slot]), 0.0)
    L.rf_profile_enable(0)
    return acc / iters


def _sum_over_ranks(x, world, dev):
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t)
    return float(t)


class _Grids:
    """Stand-in for the two GBV encoders of a model: just the parameter tensors MapVolume writes."""

    def __init__(self, n_vox, dev):
        self.GBV = type("E", (), {})(); self.GBW = type("E", (), {})()
        self.GBV.params = torch.zeros(4 * n_vox, device=dev)
        self.GBW.params = torch.zeros(n_vox, device=dev)


# =====================================================================================================================
def cfg1_part(dev, peak_gbs):
    """BASELINE config 1 on one GPU: 640x480 frame, camera at the origin looking along +z, 256^3 voxels of 2 cm."""
    from remixfusion_b200 import abi
    from remixfusion_b200.global_volume import MapVolume
    from remixfusion_b200.volume import moving_volume
    cam = synth.CFG1_CAM
    H, W = cam["H"], cam["W"]
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    bound = [[-3.0, 2.12], [-3.0, 2.12], [-2.0, 3.12]]                 # origin (-3, -3, -2), 256 voxels of 2 cm
    scene = synth.make_scene([[-2.9, 2.0], [-2.9, 2.0], [-1.9, 3.0]], 1)
    c2w = np.eye(4)
    depth, rgb = synth.render_frame(scene, K, H, W, c2w, seed=1)
    cfg = configs.replica()
    cfg["volume"].update(voxel_size=0.02, trunc=0.06)
    cfg["training"]["c_trunc"] = 0.06
    cfg["mapping"]["bound"] = bound
    cfg["globalV"]["base_resolution"] = 256
    for ax in ("x_config", "y_config", "z_config"):
        cfg["volume"][ax] = {"fix": 0, "len": 2.56, "range": [0, 1]}
    mv = moving_volume(cfg, None, c2w, device=dev)                     # the constructor centres a 5.12 m cube on the pose ...
    mv.vol_bnds = np.array(bound); mv.vol_origin = mv.vol_bnds[:, 0].astype(np.float32)      # ... cfg 1 places it at (-3, -3, -2)
    assert tuple(int(d) for d in mv.vol_dim) == (256, 256, 256), mv.vol_dim
    d = torch.from_numpy(depth).to(dev); c = torch.from_numpy(rgb).to(dev)
    packed = torch.empty(H * W, device=dev)
    abi.check(abi.lib().rf_pack_bgr(abi.dptr(torch.floor(c * 255.0).contiguous()), abi.dptr(packed), H * W, abi.stream_ptr()), "pack")
    call_local = _timed(lambda i: mv.integrate_packed(d, packed, K, c2w, None, 1.0, 0.0), 20, 1, dev, warm=3)
    ms_local = _kernel_ms(lambda i: mv.integrate_packed(d, packed, K, c2w, None, 1.0, 0.0), 0)
    nt, nb = mv.count_touched(d, K, c2w)
    b1 = 16.0 * nt + 8.0 * nb + 8.0 * H * W
    grids = _Grids(256 ** 3, dev)
    gv = MapVolume(cfg, grids, K); gv.init_mapvolume()
    pose = torch.from_numpy(c2w).float()
    call_g = _timed(lambda i: gv.integrate_kf({"rgb": c, "depth": d}, pose, 1.0), 20, 1, dev, warm=3)
    ms_g = _kernel_ms(lambda i: gv.integrate_kf({"rgb": c, "depth": d}, pose, 1.0), 1)
    ntg = gv.count_touched(d, pose)
    b1g = 40.0 * ntg + 16.0 * H * W
    # the streaming upper bound of the same volume (SURVEY §8d "full-touch"): camera 6 m in front of the cube with the whole
    # cube inside its frustum and a wall behind it, so EVERY voxel is free space in front of the surface and is updated
    c2w_f = np.eye(4); c2w_f[:3, 3] = [-0.44, -0.44, -8.0]
    d_f = torch.full((H, W), 11.6, device=dev)
    call_full = _timed(lambda i: mv.integrate_packed(d_f, packed, K, c2w_f, None, 1.0, 0.0), 20, 1, dev, warm=3)
    ms_full = _kernel_ms(lambda i: mv.integrate_packed(d_f, packed, K, c2w_f, None, 1.0, 0.0), 0)
    ntf, nbf = mv.count_touched(d_f, K, c2w_f)
    b1f = 16.0 * ntf + 8.0 * nbf + 8.0 * H * W
    full = {"ms": ms_full, "call_ms": call_full, "touched": ntf, "band": nbf, "voxel_updates_per_s": ntf / (ms_full / 1e3), "algorithmic_bytes": b1f,
            "achieved_gbs": b1f / (ms_full / 1e3) / 1e9, "frac": b1f / (ms_full / 1e3) / 1e9 / peak_gbs,
            "note": "camera at (-0.44, -0.44, -8) looking along +z, constant depth 11.6 m: all 256^3 voxels lie in the frustum in front of the surface"}
    out = {"workload": "640x480 frame -> 256^3 voxels of 2 cm (camera at the origin, +z)", "local_full_touch": full,
           "timing": "ms = the kernel's launch duration (events the library records around the launch); call_ms = the host call (Python-bound at these sizes)",
           "local": {"ms": ms_local, "call_ms": call_local, "touched": nt, "band": nb, "voxel_updates_per_s": nt / (ms_local / 1e3), "swept_voxels_per_s": 256 ** 3 / (ms_local / 1e3),
                     "algorithmic_bytes": b1, "achieved_gbs": b1 / (ms_local / 1e3) / 1e9, "frac": b1 / (ms_local / 1e3) / 1e9 / peak_gbs},
           "gbv_R256": {"ms": ms_g, "call_ms": call_g, "touched": ntg, "voxel_updates_per_s": ntg / (ms_g / 1e3), "swept_voxels_per_s": 256 ** 3 / (ms_g / 1e3),
                        "algorithmic_bytes": b1g, "achieved_gbs": b1g / (ms_g / 1e3) / 1e9, "frac": b1g / (ms_g / 1e3) / 1e9 / peak_gbs}}
    del mv, gv, grids
    torch.cuda.empty_cache()
    return out


# =====================================================================================================================
def _frame_rays(cfg, K, frame, dev):
    cam = cfg["cam"]; H, W = cam["H"], cam["W"]
    c2w, depth, rgb = frame
    dirs = torch.from_numpy(synth.camera_dirs(K, H, W).reshape(-1, 3)).to(dev)
    c2w_t = torch.from_numpy(c2w.astype(np.float32)).to(dev)
    rays_d = torch.sum(dirs[..., None, :] * c2w_t[:3, :3], -1).contiguous()
    rays_o = c2w_t[None, :3, -1].repeat(H * W, 1).contiguous()
    return rays_o, rays_d, torch.from_numpy(rgb).to(dev).reshape(-1, 3).contiguous(), torch.from_numpy(depth).to(dev).reshape(-1, 1).contiguous()


def cfg3_part(dev, rank, world, group, frames, K, peak_gbs):
    """BASELINE config 3: the 2^20-ray batch is FIXED; rank r renders rays [r N/G, (r+1) N/G) and the gradients are all-reduced."""
    from remixfusion_b200 import dist as rdist
    from remixfusion_b200.global_volume import MapVolume
    from remixfusion_b200.scene_rep import JointEncoding
    cfg = configs.replica(hidden=64, hash_size=19, n_range_d=48, n_samples_d=0, voxel_sdf=8.0 / 512)
    S = 48
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
    torch.manual_seed(0)
    model = JointEncoding(cfg, bb, process_group=group, equal_shards=True).to(dev)
    with torch.no_grad():
        model.embed_res_fn.params.copy_((torch.rand_like(model.embed_res_fn.params) * 2 - 1) * 1e-2)
    model.train()
    mv = MapVolume(cfg, model, K); mv.init_mapvolume()
    for c2w, depth, rgb in frames[:2]:
        mv.integrate_kf({"rgb": torch.from_numpy(rgb), "depth": torch.from_numpy(depth)}, torch.from_numpy(c2w).float())
    ro, rd, tc, td = _frame_rays(cfg, K, frames[0], dev)
    n_total = 1 << 20
    pick = torch.randint(0, ro.shape[0], (n_total,), generator=torch.Generator().manual_seed(5)).to(dev)
    lo, hi = rdist.shard_rays(n_total, rank, world)
    assert (hi - lo) * world == n_total
    sel = pick[lo:hi]
    ro, rd, tc, td = (t[sel].contiguous() for t in (ro, rd, tc, td))
    params = [model.embed_res_fn.params] + list(model.decoder_res.fused_weights())
    fg = rdist.FlatGrads(params)

    def step(ba):
        def run(i):
            fg.zero()
            o = ro.clone().requires_grad_(True) if ba else ro
            d = rd.clone().requires_grad_(True) if ba else rd
            ret = model.mapping(o, d, tc, td, clamp=ba)
            configs.total_loss(cfg, ret).backward()
            fg.allreduce(group)
        return run
    P = float(n_total) * S
    out = {"workload": "2^20 rays x 48 samples, hash 16 x 2^19 (res 512, 40 MiB), hidden 64, GBV R = 200", "scaling": "strong",
           "rays_per_rank": hi - lo, "table_grad_allreduce_bytes": int(fg.flat.numel() * 4)}
    for name, ba, bytes_per_sample in (("mapping", False, 2176.0), ("ba", True, 3328.0)):
        ms = _timed(step(ba), 4, world, dev, warm=2)
        out[name] = {"ms": ms, "ray_samples_per_s": P / (ms / 1e3), "hbm_form_gbs": bytes_per_sample * P / (ms / 1e3) / 1e9,
                     "hbm_form_frac": bytes_per_sample * P / (ms / 1e3) / 1e9 / (peak_gbs * world)}
    del model, fg, params
    torch.cuda.empty_cache()
    return out


# =====================================================================================================================
BS3D_CAM = dict(H=720, W=1280, fx=663.72497136, fy=661.02307276, cx=638.02311807, cy=357.98090814)   # configs/BS3D/BS3D.yaml:86-97


def cfg4_part(dev, rank, world, group, peak_gbs, resolutions=(512, 1024)):
    """BASELINE config 4: 50 x 50 x 10 m bound, GBV z-slabs over the ranks; frame broadcast from rank 0; the replicated copy
    every rank's ray query reads is refreshed from the slabs over the frustum's [y, x] box only."""
    from remixfusion_b200 import dist as rdist
    from remixfusion_b200.global_volume import MapVolume
    cam = BS3D_CAM; H, W = cam["H"], cam["W"]
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    bound = [[0.0, 50.0], [0.0, 50.0], [0.0, 10.0]]
    scene = synth.make_scene(bound, 2)
    eye = np.array([14.0, 18.0, 4.0])
    c2w = synth.look_at(eye, eye + np.array([8.0, 5.0, -0.6]))
    max_depth = 12.0
    depth, rgb = synth.render_frame(scene, K, H, W, c2w, seed=4, max_depth=max_depth)
    d = torch.from_numpy(depth).to(dev); c = torch.from_numpy(rgb).to(dev)
    pose = torch.from_numpy(c2w).float()
    cfg = configs.replica()
    cfg["mapping"]["bound"] = bound
    cfg["training"]["c_trunc"] = 0.25
    out = {"workload": "50 x 50 x 10 m bound, 1280x720 frame (depth <= 12 m), GBV z-slabs over the ranks + frame broadcast + "
                       "all-gather of the frustum's [y, x] box of every slab", "frame_bytes_broadcast": int(16 * H * W)}
    for R in resolutions:
        cfg["globalV"]["base_resolution"] = R
        z = rdist.slab(R, rank, world)
        n_own = (z[1] - z[0]) * R * R
        need = 20.0 * n_own + (16.0 * R ** 3 if world > 1 else 0.0)
        free, _ = torch.cuda.mem_get_info(dev)
        if need > 0.8 * free:
            out[f"R{R}"] = {"skipped": f"needs {need / 1e9:.1f} GB of {free / 1e9:.1f} GB free"}
            continue
        grids = _Grids(n_own, dev)
        gv = MapVolume(cfg, grids, K, z_slab=z if world > 1 else None)
        gv.init_mapvolume()
        full = torch.zeros(4 * R ** 3, device=dev) if world > 1 else None
        lo, hi = rdist.frustum_box(K, c2w, H, W, max_depth + cfg["training"]["c_trunc"], bound, R)
        zs = [rdist.slab(R, k, world) for k in range(world)]

        def step(i):
            if world > 1:
                rdist.broadcast_frame(d, c, 0, group)
            gv.integrate_kf({"rgb": c, "depth": d}, pose, 1.0)
            if world > 1:
                rdist.gather_touched_box(full, grids.GBV.params, R, zs, lo, hi, group)
        ms = _timed(step, 6, world, dev, warm=2)
        touched = _sum_over_ranks(gv.count_touched(d, pose), world, dev)
        b1g = 40.0 * touched + 16.0 * H * W * world
        box_frac = (hi[0] - lo[0]) * (hi[1] - lo[1]) / float(R * R)
        out[f"R{R}"] = {"ms": ms, "touched": touched, "voxel_updates_per_s": touched / (ms / 1e3), "swept_voxels_per_s": float(R) ** 3 / (ms / 1e3),
                        "algorithmic_bytes": b1g, "achieved_gbs": b1g / (ms / 1e3) / 1e9, "frac": b1g / (ms / 1e3) / 1e9 / (peak_gbs * world),
                        "gbv_bytes": 16.0 * R ** 3, "gathered_box_fraction": box_frac if world > 1 else None}
        del gv, grids, full
        torch.cuda.empty_cache()
    return out


# =====================================================================================================================
UHUMANS_CAM = dict(H=460, W=700, fx=415.69219, fy=415.69219, cx=350.0, cy=230.0)     # configs/uhumans/uhumans.yaml: 720x480, crop_edge 10
UHUMANS_BOUND = [[-19.0, 7.0], [-9.0, 5.0], [-1.0, 7.0]]                             # configs/uhumans/apartment.yaml:3


def uhumans_config():
    cfg = configs.replica(hash_size=21, n_range_d=21, n_samples_d=96)
    cfg["cam"].update(H=UHUMANS_CAM["H"], W=UHUMANS_CAM["W"], fx=UHUMANS_CAM["fx"], fy=UHUMANS_CAM["fy"], cx=UHUMANS_CAM["cx"], cy=UHUMANS_CAM["cy"],
                      near=0, far=20)
    cfg["training"].update(range_d=0.5, c_trunc=0.25, trunc=0.06, rgb_missing=0.0)
    cfg["mapping"].update(bound=UHUMANS_BOUND, iters=10, BA_iters=5, sample=2048, keyframe_every=5, n_pixels=0.05, filter_depth=True,
                          lr_embed_res=0.01, lr_decoder=0.01)
    return cfg


def cfg5_part(dev, rank, world, group, n_keyframes=24, n_pool=4):
    """BASELINE config 5, the per-keyframe cycle of mp_slam/mapper.py:366-520 on a uHumans2-shaped stream (700x460, 21 + 96
    samples, hash 16 x 2^21, keyframe every 5 frames of a 1000-frame trajectory = 200 cycles; `n_keyframes` of them are timed
    over a pool of `n_pool` distinct synthetic frames).  With N ranks: GBV z-slabs + frame broadcast + box all-gather per
    keyframe; every rank draws its own 2048 rays per iteration and the gradients are all-reduced (weak scaling)."""
    from remixfusion_b200 import dist as rdist
    from remixfusion_b200.global_volume import MapVolume
    from remixfusion_b200.graph import GraphedMappingStep
    from remixfusion_b200.optim import Adam
    from remixfusion_b200.ray_store import KeyFrameDatabase
    from remixfusion_b200.scene_rep import JointEncoding
    cfg = uhumans_config()
    cam = cfg["cam"]; H, W = cam["H"], cam["W"]
    S = cfg["training"]["n_range_d"] + cfg["training"]["n_samples_d"]
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    scene = synth.make_scene(cfg["mapping"]["bound"], 3)
    poses = synth.loop_trajectory(scene, 1000)
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
    torch.manual_seed(0)
    model = JointEncoding(cfg, bb, process_group=group, equal_shards=True).to(dev)
    with torch.no_grad():
        model.embed_res_fn.params.copy_((torch.rand_like(model.embed_res_fn.params) * 2 - 1) * 1e-4)
    model.train()
    R = cfg["globalV"]["base_resolution"]
    zs = [rdist.slab(R, k, world) for k in range(world)]
    if world > 1:
        slab = _Grids((zs[rank][1] - zs[rank][0]) * R * R, dev)
        gv = MapVolume(cfg, slab, K, z_slab=zs[rank])
    else:
        slab, gv = model, MapVolume(cfg, model, K)
    gv.init_mapvolume()
    if world > 1:
        with torch.no_grad():
            model.GBV.params.view(-1, 4)[:, 0] = 1.0
    dirs = torch.from_numpy(synth.camera_dirs(K, H, W)).to(dev)[None]
    pool = []
    for i in range(n_pool):
        f = (5 * i * (1000 // (5 * n_pool))) % 1000
        depth, rgb = synth.render_frame(scene, K, H, W, poses[f], seed=f, max_depth=float(cam["far"]))
        pool.append((poses[f], torch.from_numpy(depth).to(dev), torch.from_numpy(rgb).to(dev)))
    n_rays = int(cfg["mapping"]["sample"])
    keep = int(H * W * cfg["mapping"]["n_pixels"])
    store = KeyFrameDatabase(cfg, H, W, n_keyframes + 8, keep, dev)
    dec = [p for p in model.decoder_res.parameters()]
    groups = [{"params": dec, "weight_decay": 1e-6, "lr": cfg["mapping"]["lr_decoder"]},
              {"params": [model.embed_res_fn.params], "eps": 1e-15, "lr": cfg["mapping"]["lr_embed_res"]}]
    loss_fn = lambda r: configs.total_loss(cfg, r)
    # one GPU: the whole iteration incl. the fused Adam is one CUDA graph; several ranks: reduce-scatter -> Adam on the owned shard
    # -> all-gather (ShardedAdam), which also clears the gradients
    opt = Adam(groups, betas=(0.9, 0.99), capturable=True) if world == 1 else rdist.ShardedAdam(groups, betas=(0.9, 0.99), group=group)
    graphed = GraphedMappingStep(model, opt, n_rays, loss_fn, eager_steps=2) if world == 1 else None
    graphed_ba = GraphedMappingStep(model, opt, n_rays, loss_fn, eager_steps=2, ray_grads=True) if world == 1 else None

    def rays_from_rows(rows, c2w_t):
        d = torch.sum(rows[:, None, :3] * c2w_t[:3, :3], -1)
        return c2w_t[None, :3, -1].expand(rows.shape[0], 3).contiguous(), d.contiguous(), rows[:, 3:6].contiguous(), rows[:, 6:7].contiguous()

    def iteration(rows, c2w_t, ba):
        ro, rd, tc, td = rays_from_rows(rows, c2w_t)
        if graphed is not None:
            (graphed_ba if ba else graphed)(ro, rd, tc, td)
            return
        if ba:
            ro = ro.clone().requires_grad_(True); rd = rd.clone().requires_grad_(True)
        loss_fn(model.mapping(ro, rd, tc, td, clamp=ba)).backward()
        if world == 1:
            opt.step(zero_grad=True)
        else:
            opt.step()

    def keyframe_cycle(i):
        c2w, depth, rgb = pool[i % n_pool]
        pose = torch.from_numpy(c2w).float()
        if world > 1:
            rdist.broadcast_frame(depth, rgb, 0, group)
        gv.integrate_kf({"rgb": rgb, "depth": depth}, pose, 1.0)
        if world > 1:
            lo, hi = rdist.frustum_box(K, c2w, H, W, float(cam["far"]) + cfg["training"]["c_trunc"], cfg["mapping"]["bound"], R)
            rdist.gather_touched_box(model.GBV.params.data, slab.GBV.params, R, zs, lo, hi, group)
        if len(store) < store.rays.shape[0]:
            store.add_keyframe({"direction": dirs, "rgb": rgb[None], "depth": depth[None], "frame_id": 5 * i}, filter_depth=True)
        c2w_t = pose.to(dev)
        for _ in range(cfg["mapping"]["iters"]):
            rows, _ids = store.sample_global_rays(n_rays)
            iteration(rows, c2w_t, False)
        for _ in range(cfg["mapping"]["BA_iters"]):
            rows, _ids = store.sample_global_rays(n_rays)
            iteration(rows, c2w_t, True)
    ms = _timed(keyframe_cycle, n_keyframes, world, dev, warm=3)
    iters = cfg["mapping"]["iters"] + cfg["mapping"]["BA_iters"]
    out = {"workload": f"uHumans2-shaped stream ({W}x{H}, {S} samples per ray, hash 16 x 2^21, GBV R = {R}): per keyframe integrate_kf + ray store + "
                       f"{cfg['mapping']['iters']} mapping iterations + {cfg['mapping']['BA_iters']} BA iterations of {n_rays} rays (one GPU: every iteration is one CUDA graph incl. the fused Adam; several: eager, reduce-scatter / Adam shard / all-gather)",
           "scaling": "weak", "keyframe_cycles_timed": n_keyframes, "ms_per_keyframe_cycle": ms, "keyframe_cycles_per_s": 1e3 / ms,
           "sequence_1000_frames_s": 200 * ms / 1e3, "ray_samples_per_s_fwd_bwd": world * iters * n_rays * S / (ms / 1e3),
           "iterations_per_cycle": iters, "rays_per_iteration_per_rank": n_rays}
    del model, store, graphed, graphed_ba, opt
    torch.cuda.empty_cache()
    return out
