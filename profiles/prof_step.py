"""One step of the bench workload (BASELINE config 2) for Nsight Compute: TSDF fuse (local + GBV), re-centring,
full-frame ray query fwd+bwd in mapping mode, then the same frame in BA mode (ray gradients).  The first pass is a
warm-up; the second runs between cudaProfilerStart/Stop, so that

    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/<name> python profiles/prof_step.py
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/<name>_launches.csv python profiles/prof_step.py

capture exactly one launch of every kernel of the step.  `profiles/summarise_ncu.py` turns the report into the CSV
committed next to this file.  Numbers printed under ncu are never bench values (bench.py is the measurement)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                        # noqa: E402
from remixfusion_b200 import abi, configs, synth                    # noqa: E402
from remixfusion_b200.global_volume import MapVolume                # noqa: E402
from remixfusion_b200.scene_rep import JointEncoding                # noqa: E402
from remixfusion_b200.volume import moving_volume                   # noqa: E402


def main():
    abi.lib()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = configs.replica(hidden=int(os.environ.get("RF_PROF_HIDDEN", 32)), hash_size=int(os.environ.get("RF_PROF_HASH", 16)))
    cam = cfg["cam"]; H, W = cam["H"], cam["W"]
    K, poses, frames = bench.make_frames(cfg, 1, first=0, stride=50)
    c2w, depth, rgb = frames[0]
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
    torch.manual_seed(0)
    model = JointEncoding(cfg, bb).to(dev)
    with torch.no_grad():
        model.embed_res_fn.params.copy_((torch.rand_like(model.embed_res_fn.params) * 2 - 1) * 1e-2)
    model.train()
    params = [model.embed_res_fn.params] + list(model.decoder_res.fused_weights())
    mvol = MapVolume(cfg, model, K); mvol.init_mapvolume()
    local = moving_volume(cfg, None, poses[0], device=dev)
    d = torch.from_numpy(depth).to(dev); c = torch.from_numpy(rgb).to(dev)
    packed = torch.empty(H * W, device=dev)
    abi.check(abi.lib().rf_pack_bgr(abi.dptr(torch.floor(c * 255.0).contiguous()), abi.dptr(packed), H * W, abi.stream_ptr()), "pack")
    dirs = torch.from_numpy(synth.camera_dirs(K, H, W).reshape(-1, 3)).to(dev)
    c2w_t = torch.from_numpy(c2w.astype(np.float32)).to(dev)
    rays_d = torch.sum(dirs[..., None, :] * c2w_t[:3, :3], -1).contiguous()
    rays_o = c2w_t[None, :3, -1].repeat(H * W, 1).contiguous()
    tgt_d = d.reshape(-1, 1).contiguous(); tgt_c = c.reshape(-1, 3).contiguous()
    pose = torch.from_numpy(c2w).float()

    def step():
        local.integrate_packed(d, packed, K, c2w, None, 1.0, 0.0)
        mvol.integrate_kf({"rgb": c, "depth": d}, pose, 1.0)
        b0 = local.vol_bnds.copy()
        local.update_tsdf_swap_rot_trans(b0 + np.array([[1.0], [0.0], [0.0]]), b0.copy())
        for p in params:
            p.grad = None
        configs.total_loss(cfg, model.mapping(rays_o, rays_d, tgt_c, tgt_d)).backward()
        ro = rays_o.clone().requires_grad_(True); rd = rays_d.clone().requires_grad_(True)
        for p in params:
            p.grad = None
        configs.total_loss(cfg, model.mapping(ro, rd, tgt_c, tgt_d, clamp=True)).backward()
        torch.cuda.synchronize()

    step()
    torch.cuda.profiler.start()
    step()
    torch.cuda.profiler.stop()
    print("prof_step ok")


if __name__ == "__main__":
    main()
