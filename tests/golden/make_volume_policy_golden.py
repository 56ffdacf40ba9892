"""Generates tests/golden/volume_policy_golden.npz by driving the REFERENCE's own ``moving_volume.check_move_volume_new`` and
``frameid_to_Vrange`` (model/Volume.py:930-1105, imported from /root/reference with skimage / PyCUDA stubbed — only possible in
the build container) along a seeded camera walk, with the tracker's bookkeeping of model/ROtracker.py:920-934 around it.
The GPU side of a move (``copy_volume`` / ``update_tsdf_swap_rot_trans``) is replaced by its host arithmetic
(model/Volume.py:812-821); the kernels behind it are pinned elsewhere (tests/test_tsdf_gpu.py).
Run:  python tests/golden/make_volume_policy_golden.py"""
import copy
import os
import sys
import types

import numpy as np

for name in ("skimage", "skimage.measure", "pycuda", "pycuda.driver", "pycuda.autoprimaryctx", "pycuda.compiler", "pycuda.gpuarray"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["skimage"].measure = sys.modules["skimage.measure"]
sys.modules["pycuda.compiler"].SourceModule = object
sys.path.insert(0, "/root/reference")
from model.Volume import moving_volume          # noqa: E402


class Traj:
    kfx = kfy = kfz = 0.0


def make(cfg_v, pose0):
    mv = moving_volume.__new__(moving_volume)
    mv.voxel_size = float(cfg_v["voxel_size"])
    mv.fix_x, mv.fix_y, mv.fix_z = (cfg_v[a]["fix"] for a in ("x_config", "y_config", "z_config"))
    mv.x_len, mv.y_len, mv.z_len = (cfg_v[a]["len"] for a in ("x_config", "y_config", "z_config"))
    mv.version, mv.t_treshold = "center", cfg_v["t_treshold"]
    traj = Traj()
    mv.vol_bnds = np.asarray(mv.initialize_vol_bnd(pose0, traj, "center"))
    mv.vol_dim = np.ceil((mv.vol_bnds[:, 1] - mv.vol_bnds[:, 0]) / mv.voxel_size).copy(order="C").astype(int)
    mv.vol_bnds[:, 1] = mv.vol_bnds[:, 0] + mv.vol_dim * mv.voxel_size
    mv.start_id, mv.frame_to_Vrange = 0, {}
    mv.moves = []

    def swap(vol_bnds, old_bnds):                # host part of update_tsdf_swap_rot_trans (model/Volume.py:812-821)
        mv.vol_bnds = vol_bnds
        mv.vol_dim = np.ceil((mv.vol_bnds[:, 1] - mv.vol_bnds[:, 0]) / mv.voxel_size).copy(order="C").astype(int)
        mv.vol_bnds[:, 1] = mv.vol_bnds[:, 0] + mv.vol_dim * mv.voxel_size
        mv.moves.append(np.stack([vol_bnds.copy(), np.asarray(old_bnds).copy()]))
    mv.copy_volume = lambda: None
    mv.update_tsdf_swap_rot_trans = swap
    return mv, traj


out = {}
rng = np.random.default_rng(0)
cases = {
    "free": {"voxel_size": 0.02, "t_treshold": 1, "x_config": {"fix": 0, "len": 4}, "y_config": {"fix": 0, "len": 4}, "z_config": {"fix": 0, "len": 3}},
    "zfix": {"voxel_size": 0.03, "t_treshold": 0.5, "x_config": {"fix": 0, "len": 3}, "y_config": {"fix": 0, "len": 2}, "z_config": {"fix": 1, "len": 2}},
}
for cname, cv in cases.items():
    n = 160
    steps = rng.normal(0.0, 0.12, (n, 3)) + np.array([0.06, -0.03, 0.015])
    steps[40:60] *= 4.0                          # a fast stretch: moves by more than a metre between two checks
    pos = np.cumsum(steps, 0) + np.array([0.4, -0.6, 0.2])
    pose = np.eye(4); pose[:3, 3] = pos[0]
    mv, traj = make(cv, pose)
    flags, olds, cur, kf = [], [], [], []
    for i in range(n):
        pose = np.eye(4); pose[:3, 3] = pos[i]
        flag, old = mv.check_move_volume_new(i, pose, traj, version="center")
        if flag:                                 # model/ROtracker.py:925-934
            start = 0 if mv.start_id == 0 else mv.start_id
            mv.start_id = i
            mv.frame_to_Vrange[(start, i - 1)] = old
        flags.append(flag); olds.append(np.asarray(old).copy()); cur.append(mv.vol_bnds.copy()); kf.append([traj.kfx, traj.kfy, traj.kfz])
    out[f"{cname}_pos"] = pos; out[f"{cname}_flags"] = np.asarray(flags); out[f"{cname}_old"] = np.stack(olds)
    out[f"{cname}_bnds"] = np.stack(cur); out[f"{cname}_kf"] = np.asarray(kf)
    out[f"{cname}_lookup"] = np.stack([np.asarray(mv.frameid_to_Vrange(i)) for i in range(n)])
    out[f"{cname}_ranges"] = np.asarray([[s, e] for (s, e) in mv.frame_to_Vrange.keys()], dtype=np.int64).reshape(-1, 2)
    print(cname, "moves:", int(np.sum(flags)), "ranges:", list(mv.frame_to_Vrange.keys())[:6])
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "volume_policy_golden.npz")
np.savez_compressed(dst, **out)
print("wrote", dst)
