"""Stage-1 timings for tuning: cfg 1 (local / GBV R = 256 / full-touch) and the cfg-2 frame (local 400x400x300 + GBV 200^3),
for a list of launch shapes (RF_TSDF_SHAPE = "rows,threads", read by the library at every launch).

    python profiles/time_tsdf.py [rows,threads ...]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                        # noqa: E402
import bench_workloads as bw                                        # noqa: E402
from remixfusion_b200 import abi, configs                           # noqa: E402
from remixfusion_b200.global_volume import MapVolume                # noqa: E402
from remixfusion_b200.scene_rep import JointEncoding                # noqa: E402
from remixfusion_b200.volume import moving_volume                   # noqa: E402


def cfg2_setup(dev):
    """(local, global) closures integrating the first cfg-2 frame (1200x680 -> 400x400x300 @ 2 cm, GBV 200^3)."""
    cfg = configs.replica()
    cam = cfg["cam"]; H, W = cam["H"], cam["W"]
    K, poses, frames = bench.make_frames(cfg, 1, first=0, stride=50)
    c2w, depth, rgb = frames[0]
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
    model = JointEncoding(cfg, bb).to(dev)
    mvol = MapVolume(cfg, model, K); mvol.init_mapvolume()
    local = moving_volume(cfg, None, poses[0], device=dev)
    d = torch.from_numpy(depth).to(dev); c = torch.from_numpy(rgb).to(dev)
    packed = torch.empty(H * W, device=dev)
    abi.check(abi.lib().rf_pack_bgr(abi.dptr(torch.floor(c * 255.0).contiguous()), abi.dptr(packed), H * W, abi.stream_ptr()), "pack")
    pose = torch.from_numpy(c2w).float()
    return (lambda i=0: local.integrate_packed(d, packed, K, c2w, None, 1.0, 0.0)), (lambda i=0: mvol.integrate_kf({"rgb": c, "depth": d}, pose, 1.0))


def cfg2_times(dev):
    fl, fg = cfg2_setup(dev)
    return bw._timed(fl, 20, 1, dev, warm=3), bw._timed(fg, 20, 1, dev, warm=3)


def main():
    abi.lib()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    shapes = sys.argv[1:] or ["8,128"]
    for sh in shapes:
        os.environ["RF_TSDF_SHAPE"] = sh
        p = bw.cfg1_part(dev, 6551.4)
        l2, g2 = cfg2_times(dev)
        print(json.dumps({"shape": sh, "cfg1_local_ms": round(p["local"]["ms"], 4), "cfg1_gbv_ms": round(p["gbv_R256"]["ms"], 4),
                          "cfg1_full_ms": round(p["local_full_touch"]["ms"], 4), "cfg1_full_frac": round(p["local_full_touch"]["frac"], 3),
                          "cfg2_local_ms": round(l2, 4), "cfg2_gbv_ms": round(g2, 4)}), flush=True)


if __name__ == "__main__":
    main()
