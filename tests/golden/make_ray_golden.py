"""Generates tests/golden/ray_golden.npz IN THIS CONTAINER by running the reference's own Stage-2 code
(model/scene_rep.py JointEncoding.mapping / render_rays, model/decoder.py, model/utils.py — imported from
/root/reference through oracle/ref_import.py, with oracle/tcnn_standin.py in place of tiny-cuda-nn) on seeded
inputs, forward and backward.

    python tests/golden/make_ray_golden.py

Three cases: A = mapping mode (Replica sampling 48+11, perturb, rgb_missing 0.05);  B = BA mode (clamp=True,
gradients w.r.t. rays);  C = eval mode, n_samples_d=0, perturb=0, rgb_missing=0 (BS3D-style weights).
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import, tsdf_oracle as O       # noqa: E402
from remixfusion_b200 import synth                     # noqa: E402

BOUND = synth.REPLICA_BOUND


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


def make_inputs(n_rays=96, seed=0):
    """A 150x85 frame of the analytic scene -> GBV (R=32, via the C oracle) + a seeded ray batch."""
    cam = synth.REPLICA_CAM
    s = 8
    H, W = cam["H"] // s, cam["W"] // s
    K = synth.intrinsics(cam["fx"] / s, cam["fy"] / s, (cam["cx"] + .5) / s - .5, (cam["cy"] + .5) / s - .5)
    scene = synth.make_scene(BOUND, 0)
    c2w = synth.loop_trajectory(scene, 8)[1]
    depth, rgb = synth.render_frame(scene, K, H, W, c2w, invalid_frac=0.05, seed=seed)
    R = 32
    trgb = np.zeros(4 * R ** 3, np.float32); O.clear_global(trgb); gw = np.zeros(R ** 3, np.float32)
    box = [v for ax in BOUND for v in ax]
    O.integrate_global(trgb, gw, R, box, K, c2w, depth, rgb, 0.1, 1.0)
    rng = np.random.default_rng(seed)
    pix = rng.choice(H * W, n_rays, replace=False)
    dirs = synth.camera_dirs(K, H, W).reshape(-1, 3)[pix]
    c2w32 = c2w.astype(np.float32)
    rays_d = (dirs[:, None, :] * c2w32[None, :3, :3]).sum(-1).astype(np.float32)       # mp_slam/mapper.py:344
    rays_o = np.broadcast_to(c2w32[:3, 3], rays_d.shape).copy()
    return dict(rays_o=rays_o, rays_d=rays_d, target_d=depth.reshape(-1)[pix][:, None].copy(),
                target_rgb=rgb.reshape(-1, 3)[pix].copy(), gbv=trgb)


def set_params(model, gbv, seed=0):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        model.embed_res_fn.params.copy_((torch.rand(model.embed_res_fn.params.shape, generator=g) * 2 - 1) * 0.05)
        model.GBV.params.copy_(torch.from_numpy(gbv))
        for lin in (model.decoder_res.sdf_net.model[0], model.decoder_res.sdf_net.model[2],
                    model.decoder_res.color_net.model[0], model.decoder_res.color_net.model[2]):
            k = 1.0 / np.sqrt(lin.weight.shape[1])
            lin.weight.copy_((torch.rand(lin.weight.shape, generator=g) * 2 - 1) * k)


def total_loss(cfg, ret):          # mp_slam/slam.py:162-169
    t = cfg["training"]
    return (t["rgb_weight"] * ret["rgb_res_loss"] + t["depth_weight"] * ret["depth_res_loss"]
            + t["sdf_weight"] * ret["sdf_res_loss"] + t["fs_weight"] * ret["fs_res_loss"])


def run_case(cfg, inp, clamp, train, ray_grads, seed):
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)          # run.py:90
    model = ref_import.make_reference_model(cfg, bb)
    set_params(model, inp["gbv"], seed)
    model.train(train)
    ro = torch.from_numpy(inp["rays_o"]).requires_grad_(ray_grads)
    rd = torch.from_numpy(inp["rays_d"]).requires_grad_(ray_grads)
    td = torch.from_numpy(inp["target_d"]); tc = torch.from_numpy(inp["target_rgb"])
    S = cfg["training"]["n_range_d"] + cfg["training"]["n_samples_d"]
    torch.manual_seed(1000 + seed)
    u = torch.rand(ro.shape[0], S)                          # what torch.rand(z_vals.shape) will return (scene_rep.py:441)
    torch.manual_seed(1000 + seed)
    ret = model.mapping(ro, rd, tc, td, clamp=clamp)
    out = {"u": u.numpy()}
    out["hash_params"] = model.embed_res_fn.params.detach().numpy().copy()
    out["w_sdf0"] = model.decoder_res.sdf_net.model[0].weight.detach().numpy().copy()
    out["w_sdf1"] = model.decoder_res.sdf_net.model[2].weight.detach().numpy().copy()
    out["w_col0"] = model.decoder_res.color_net.model[0].weight.detach().numpy().copy()
    out["w_col1"] = model.decoder_res.color_net.model[2].weight.detach().numpy().copy()
    if not train:
        for k in ("rgb_res_map", "depth_res_map", "z_vals", "raw"):
            out[k] = ret[k].detach().numpy()
        return out
    for k in ("rgb_res_loss", "depth_res_loss", "sdf_res_loss", "fs_res_loss", "rgb_res", "depth_res"):
        out[k] = ret[k].detach().numpy()
    loss = total_loss(cfg, ret)
    loss.backward()
    out["loss"] = loss.detach().numpy()
    out["g_hash"] = model.embed_res_fn.params.grad.numpy().copy()
    out["g_w_sdf0"] = model.decoder_res.sdf_net.model[0].weight.grad.numpy().copy()
    out["g_w_sdf1"] = model.decoder_res.sdf_net.model[2].weight.grad.numpy().copy()
    out["g_w_col0"] = model.decoder_res.color_net.model[0].weight.grad.numpy().copy()
    out["g_w_col1"] = model.decoder_res.color_net.model[2].weight.grad.numpy().copy()
    if ray_grads:
        out["g_rays_o"] = ro.grad.numpy().copy(); out["g_rays_d"] = rd.grad.numpy().copy()
    # also the un-jittered render in eval mode for raw/z_vals (same params)
    model.eval()
    torch.manual_seed(1000 + seed)
    r2 = model.mapping(ro.detach(), rd.detach(), tc, td, clamp=clamp)
    for k in ("z_vals", "raw"):
        out[k] = r2[k].detach().numpy()
    return out


def main():
    assert ref_import.available()
    inp = make_inputs()
    res = {f"in_{k}": v for k, v in inp.items()}
    cfgA = base_config()
    cfgC = copy.deepcopy(cfgA)
    cfgC["training"].update(n_samples_d=0, perturb=0, rgb_missing=0.0)
    cases = {"A": (cfgA, False, True, False), "B": (cfgA, True, True, True), "C": (cfgC, False, False, False)}
    for name, (cfg, clamp, train, rg) in cases.items():
        out = run_case(cfg, inp, clamp, train, rg, seed=ord(name))
        for k, v in out.items():
            res[f"{name}_{k}"] = v
    path = os.path.join(ROOT, "tests", "golden", "ray_golden.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, os.path.getsize(path), "bytes")
    for name in cases:
        print(name, {k[2:]: float(res[k]) for k in res if k.startswith(name + "_") and res[k].ndim == 0})


if __name__ == "__main__":
    main()
