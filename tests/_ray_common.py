"""Helpers shared by the Stage-2 tests: rebuild configs / oracle from the golden fixture."""
import copy
import os

import numpy as np
import torch

from oracle import tcnn_standin
from oracle.ray_oracle import RayOracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ray_golden.npz")
BOUND = [[-1.0, 7.0], [-1.3, 3.7], [-1.7, 1.4]]


def base_config(hash_size=10, R=32, hidden=32):
    return {
        "grid": {"enc": "HashGrid", "tcnn_encoding": True, "hash_size": hash_size, "voxel_color": 0.08, "voxel_sdf": 0.02},
        "pos": {"enc": "OneBlob", "n_bins": 16},
        "globalV": {"use": 1, "base_resolution": R, "n_levels": 1, "per_level_scale": 1, "n_features_per_level": 4},
        "decoder": {"geo_feat_dim": 15, "hidden_dim": hidden, "num_layers": 2, "num_layers_color": 2,
                    "hidden_dim_color": hidden, "tcnn_network": False},
        "cam": {"near": 0.1, "far": 5, "depth_trunc": 100.},
        "training": {"rgb_weight": 5.0, "depth_weight": 0.1, "sdf_weight": 1000, "fs_weight": 10, "n_samples_d": 11,
                     "range_d": 0.15, "n_range_d": 48, "perturb": 1, "c_trunc": 0.1, "trunc": 0.05, "rgb_missing": 0.05},
        "data": {"sc_factor": 1},
        "mapping": {"bound": BOUND, "clamp": 1.0, "pose_scale": 0.01},
    }


def case_config(name):
    cfg = base_config()
    if name == "C":
        cfg["training"].update(n_samples_d=0, perturb=0, rgb_missing=0.0)
    return cfg


from oracle.ray_oracle import gbv_standin, hash_standin, resolution_sdf   # noqa: E402,F401


def oracle_from_golden(G, name, requires_grad=True):
    cfg = case_config(name)
    h = hash_standin(cfg); g = gbv_standin(cfg)
    with torch.no_grad():
        h.params.copy_(torch.from_numpy(G[f"{name}_hash_params"]))
        g.params.copy_(torch.from_numpy(G["in_gbv"]))
    g.params.requires_grad_(False)
    ws = [torch.from_numpy(G[f"{name}_{k}"]).clone().requires_grad_(requires_grad) for k in ("w_sdf0", "w_sdf1", "w_col0", "w_col1")]
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
    return cfg, RayOracle(cfg, bb, h, g, *ws)


def inputs_from_golden(G, ray_grads=False):
    ro = torch.from_numpy(G["in_rays_o"]).clone().requires_grad_(ray_grads)
    rd = torch.from_numpy(G["in_rays_d"]).clone().requires_grad_(ray_grads)
    return ro, rd, torch.from_numpy(G["in_target_rgb"]), torch.from_numpy(G["in_target_d"])
