"""Stage-1 launches for Nsight Compute: cfg 1 local / GBV R = 256 / local full-touch, then the cfg-2 frame (local 400x400x300,
GBV 200^3), one launch each between cudaProfilerStart/Stop (warm-up first).

    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/<name> python profiles/prof_tsdf.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_workloads as bw                                        # noqa: E402
from remixfusion_b200 import abi, configs, synth                    # noqa: E402
from remixfusion_b200.global_volume import MapVolume                # noqa: E402
from remixfusion_b200.volume import moving_volume                   # noqa: E402


def main():
    abi.lib()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cam = synth.CFG1_CAM
    H, W = cam["H"], cam["W"]
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    bound = [[-3.0, 2.12], [-3.0, 2.12], [-2.0, 3.12]]
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
    mv = moving_volume(cfg, None, c2w, device=dev)
    mv.vol_bnds = np.array(bound); mv.vol_origin = mv.vol_bnds[:, 0].astype(np.float32)
    d = torch.from_numpy(depth).to(dev); c = torch.from_numpy(rgb).to(dev)
    packed = torch.empty(H * W, device=dev)
    abi.check(abi.lib().rf_pack_bgr(abi.dptr(torch.floor(c * 255.0).contiguous()), abi.dptr(packed), H * W, abi.stream_ptr()), "pack")
    grids = bw._Grids(256 ** 3, dev)
    gv = MapVolume(cfg, grids, K); gv.init_mapvolume()
    pose = torch.from_numpy(c2w).float()
    c2w_f = np.eye(4); c2w_f[:3, 3] = [-0.44, -0.44, -8.0]
    d_f = torch.full((H, W), 11.6, device=dev)

    import time_tsdf
    f2l, f2g = time_tsdf.cfg2_setup(dev)

    def step():
        mv.integrate_packed(d, packed, K, c2w, None, 1.0, 0.0)
        gv.integrate_kf({"rgb": c, "depth": d}, pose, 1.0)
        mv.integrate_packed(d_f, packed, K, c2w_f, None, 1.0, 0.0)
        f2l(); f2g()
        torch.cuda.synchronize()

    step(); step()
    torch.cuda.profiler.start()
    step()
    torch.cuda.profiler.stop()
    print("prof_tsdf ok")


if __name__ == "__main__":
    main()
