"""Config dictionaries in the reference's YAML schema (configs/Replica/replica.yaml + room0.yaml), restricted to the
keys the hot path reads (SURVEY.md §5 "Config / flags").  A reference yaml loaded with its own config.py works too."""
from __future__ import annotations

import copy

from . import synth

REPLICA = {
    "dataset": "replica",
    "data": {"sc_factor": 1, "output": "output/Replica/room0", "exp_name": "b200"},
    "globalV": {"use": 1, "base_resolution": 200, "n_levels": 1, "per_level_scale": 1, "n_features_per_level": 4},
    "mapping": {"sample": 2048, "iters": 5, "BA_iters": 5, "keyframe_every": 5, "map_every": 5, "clamp": 1.0,
                "pose_scale": 0.01, "bound": synth.REPLICA_BOUND},
    "grid": {"enc": "HashGrid", "tcnn_encoding": True, "hash_size": 16, "voxel_color": 0.08, "voxel_sdf": 0.02},
    "pos": {"enc": "OneBlob", "n_bins": 16},
    "decoder": {"geo_feat_dim": 15, "hidden_dim": 32, "num_layers": 2, "num_layers_color": 2, "hidden_dim_color": 32,
                "tcnn_network": False},
    "cam": {"H": 680, "W": 1200, "fx": 600.0, "fy": 600.0, "cx": 599.5, "cy": 339.5, "near": 0.1, "far": 5,
            "depth_trunc": 100.},
    "training": {"rgb_weight": 5.0, "depth_weight": 0.1, "sdf_weight": 1000, "fs_weight": 10, "n_samples_d": 11,
                 "range_d": 0.15, "n_range_d": 48, "perturb": 1, "c_trunc": 0.1, "trunc": 0.05, "rgb_missing": 0.05},
    # BASELINE config 2 uses 2 cm local voxels (the shipped yaml has 0.01): 400 x 400 x 300
    "volume": {"voxel_size": 0.02, "version": "center", "trunc": 0.05, "weight_clamp": 1.0, "t_treshold": 1,
               "x_config": {"fix": 0, "len": 4, "range": [0, 1]}, "y_config": {"fix": 0, "len": 4, "range": [0, 1]},
               "z_config": {"fix": 0, "len": 3, "range": [0, 1]}},
}


def replica(hidden=None, hash_size=None, n_range_d=None, n_samples_d=None, voxel_sdf=None):
    cfg = copy.deepcopy(REPLICA)
    if hidden is not None:
        cfg["decoder"]["hidden_dim"] = cfg["decoder"]["hidden_dim_color"] = hidden
    if hash_size is not None:
        cfg["grid"]["hash_size"] = hash_size
    if n_range_d is not None:
        cfg["training"]["n_range_d"] = n_range_d
    if n_samples_d is not None:
        cfg["training"]["n_samples_d"] = n_samples_d
    if voxel_sdf is not None:
        cfg["grid"]["voxel_sdf"] = voxel_sdf
    return cfg


def total_loss(cfg, ret):
    """mp_slam/slam.py:162-169 (SLAM.get_loss_from_ret with rgb, depth, sdf, fs)."""
    t = cfg["training"]
    return (t["rgb_weight"] * ret["rgb_res_loss"] + t["depth_weight"] * ret["depth_res_loss"]
            + t["sdf_weight"] * ret["sdf_res_loss"] + t["fs_weight"] * ret["fs_res_loss"])
