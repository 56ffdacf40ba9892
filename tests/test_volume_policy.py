"""Move policy of the local volume — ``moving_volume.check_move_volume_new`` / ``frameid_to_Vrange`` (model/Volume.py:930-1105)
with the tracker's bookkeeping around it (model/ROtracker.py:920-934) — against tests/golden/volume_policy_golden.npz, which
holds what the REFERENCE's own class decided along the same seeded camera walks (tests/golden/make_volume_policy_golden.py).
CPU: the policy arithmetic with the device side of a move replaced by its host arithmetic; GPU: the real object."""
import os

import numpy as np
import pytest

from remixfusion_b200.volume import moving_volume

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "volume_policy_golden.npz"))
CASES = {
    "free": {"voxel_size": 0.02, "t_treshold": 1, "x_config": {"fix": 0, "len": 4}, "y_config": {"fix": 0, "len": 4}, "z_config": {"fix": 0, "len": 3}},
    "zfix": {"voxel_size": 0.03, "t_treshold": 0.5, "x_config": {"fix": 0, "len": 3}, "y_config": {"fix": 0, "len": 2}, "z_config": {"fix": 1, "len": 2}},
}


class Traj:
    kfx = kfy = kfz = 0.0


def _walk(mv, traj, name):
    pos = G[f"{name}_pos"]
    for i in range(pos.shape[0]):
        pose = np.eye(4); pose[:3, 3] = pos[i]
        flag, old = mv.check_move_volume_new(i, pose, traj, version="center")
        if flag:                                               # the caller's bookkeeping, model/ROtracker.py:925-934
            start = 0 if mv.start_id == 0 else mv.start_id
            mv.start_id = i
            mv.frame_to_Vrange[(start, i - 1)] = old
        assert bool(flag) == bool(G[f"{name}_flags"][i]), f"frame {i}: move decision"
        assert np.array_equal(np.asarray(old), G[f"{name}_old"][i]), f"frame {i}: old bounds"
        assert np.array_equal(mv.vol_bnds, G[f"{name}_bnds"][i]), f"frame {i}: bounds after the check"
        assert [traj.kfx, traj.kfy, traj.kfz] == list(G[f"{name}_kf"][i]), f"frame {i}: reference position"
    assert np.array_equal(np.asarray(list(mv.frame_to_Vrange.keys()), dtype=np.int64).reshape(-1, 2), G[f"{name}_ranges"])
    for i in range(pos.shape[0]):
        assert np.array_equal(np.asarray(mv.frameid_to_Vrange(i)), G[f"{name}_lookup"][i]), f"frame {i}: frameid_to_Vrange"


@pytest.mark.parametrize("name", list(CASES))
def test_policy_matches_reference_class_cpu(name):
    cv = CASES[name]
    mv = moving_volume.__new__(moving_volume)                  # no device: only the policy is under test here
    mv.voxel_size = float(cv["voxel_size"]); mv.t_treshold = cv["t_treshold"]; mv.version = "center"
    mv.fix_x, mv.fix_y, mv.fix_z = (cv[a]["fix"] for a in ("x_config", "y_config", "z_config"))
    mv.x_len, mv.y_len, mv.z_len = (cv[a]["len"] for a in ("x_config", "y_config", "z_config"))
    traj = Traj()
    pose = np.eye(4); pose[:3, 3] = G[f"{name}_pos"][0]
    mv.vol_bnds = np.asarray(mv.initialize_vol_bnd(pose, traj, "center"))
    mv.vol_dim = np.ceil((mv.vol_bnds[:, 1] - mv.vol_bnds[:, 0]) / mv.voxel_size).astype(int)
    mv.vol_bnds[:, 1] = mv.vol_bnds[:, 0] + mv.vol_dim * mv.voxel_size
    mv.start_id, mv.frame_to_Vrange = 0, {}

    def swap(vol_bnds, old_bnds):                              # host arithmetic of the move, model/Volume.py:812-821
        mv.vol_bnds = np.asarray(vol_bnds, dtype=np.float64)
        mv.vol_dim = np.ceil((mv.vol_bnds[:, 1] - mv.vol_bnds[:, 0]) / mv.voxel_size).astype(int)
        mv.vol_bnds[:, 1] = mv.vol_bnds[:, 0] + mv.vol_dim * mv.voxel_size
    mv.copy_volume = lambda: None
    mv.update_tsdf_swap_rot_trans = swap
    _walk(mv, traj, name)


@pytest.mark.gpu
def test_policy_on_the_real_volume_gpu(cuda, rf_lib):
    """The real object (device arrays, rf_tsdf_recenter behind every move): same decisions, and content integrated before a
    move is found again at the same world position after it."""
    import torch
    from remixfusion_b200 import configs, synth
    cfg = configs.replica()
    cfg["volume"].update(voxel_size=0.04, **{k: CASES["free"][k] for k in ("x_config", "y_config", "z_config", "t_treshold")})
    traj = Traj()
    pose = np.eye(4); pose[:3, 3] = G["free_pos"][0]
    mv = moving_volume(cfg, traj, pose, device=cuda)
    cam = cfg["cam"]
    K = synth.intrinsics(cam["fx"] / 4, cam["fy"] / 4, cam["cx"] / 4, cam["cy"] / 4)
    depth = np.full((cam["H"] // 4, cam["W"] // 4), 1.5, np.float32); rgb = np.full(depth.shape + (3,), 128.0, np.float32)
    mv.integrate(rgb, depth, K, pose, None)
    w0 = mv.weight_vol_gpu.reshape(*[int(d) for d in mv.vol_dim]).clone(); o0 = mv.vol_origin.copy()
    assert float(w0.sum()) > 0
    # same voxel size as the golden walk's decisions do not depend on (bounds are whole metres): replay the first 40 frames
    pos = G["free_pos"]
    moved = 0
    for i in range(40):
        p = np.eye(4); p[:3, 3] = pos[i]
        flag, old = mv.check_move_volume_new(i, p, traj, version="center")
        assert bool(flag) == bool(G["free_flags"][i]) and np.array_equal(mv.vol_bnds, G["free_bnds"][i])
        if flag:
            moved += 1
            start = 0 if mv.start_id == 0 else mv.start_id
            mv.start_id = i; mv.frame_to_Vrange[(start, i - 1)] = old
    assert moved >= 2 and len(mv.frame_to_Vrange) == moved
    # a voxel that stayed inside every intermediate volume keeps its weight at the same world position
    w1 = mv.weight_vol_gpu.reshape(*[int(d) for d in mv.vol_dim]); shift = np.rint((mv.vol_origin - o0) / mv.voxel_size).astype(int)
    idx = torch.nonzero(w0 > 0)[::97]
    new_idx = idx - torch.tensor(shift, device=cuda)
    ok = ((new_idx >= 0) & (new_idx < torch.tensor([int(d) for d in mv.vol_dim], device=cuda))).all(1)
    assert int(ok.sum()) > 0
    a = w0[idx[ok, 0], idx[ok, 1], idx[ok, 2]]; b = w1[new_idx[ok, 0], new_idx[ok, 1], new_idx[ok, 2]]
    assert float((a == b).float().mean()) > 0.9            # voxels that left a volume in between were cleared
