"""Stage-2 parity AT THE BASELINE SHAPES (GPU) — the full config-2 frame (816 000 rays x 59 samples, hash 16 x 2^16,
GBV R = 200, hidden 32) and config 3 (2^20 rays x 48 samples, hash 16 x 2^19 at resolution 512, hidden 64) are rendered
on the GPU and compared with the CPU oracle (oracle/ray_oracle.py, the restatement pinned against the reference's own
model/scene_rep.py:407-529) on a seeded SUBSET of 8192 rays of the same batch:

  * per-ray outputs (z_vals, raw, rendered colour and depth) of the subset, taken out of the FULL-batch launch, vs the
    oracle run on those rays alone — per-ray results do not depend on the rest of the batch, so this checks the
    full-size launch itself (its replica count, dense / hashed level split, tile segmentation and accumulator flushes
    all differ from the small cases of tests/test_ray_gpu.py);
  * losses and gradients (hash table, four decoder matrices; in BA mode also rays_o / rays_d) of the subset batch at
    the same table / volume / decoder sizes vs the oracle's autograd.

Bars (BASELINE.json north_star): sample depths bit-exact; raw / colour / depth <= 2e-4 relative (bar 1e-3); gradients
<= 1e-3 relative (element-wise with a 2e-4 max|g| floor; relative L2 <= 1e-3, see tests/test_ray_gpu.py::_close_grad).
"""
import numpy as np
import pytest
import torch

from oracle.ray_oracle import RayOracle, gbv_standin, hash_standin
from remixfusion_b200 import configs, synth
from remixfusion_b200.global_volume import MapVolume
from remixfusion_b200.scene_rep import JointEncoding
from tests.test_ray_gpu import _close, _close_grad, _total

pytestmark = pytest.mark.gpu
N_SUB = 8192


def _setup(cuda, shape):
    cfg = configs.replica() if shape == "cfg2" else configs.replica(hidden=64, hash_size=19, n_range_d=48, n_samples_d=0,
                                                                    voxel_sdf=8.0 / 512)
    cam = cfg["cam"]; H, W = cam["H"], cam["W"]
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    scene = synth.make_scene(cfg["mapping"]["bound"], 0)
    poses = synth.loop_trajectory(scene, 200)
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
    torch.manual_seed(11)
    m = JointEncoding(cfg, bb).to(cuda)
    with torch.no_grad():
        m.embed_res_fn.params.copy_((torch.rand_like(m.embed_res_fn.params) * 2 - 1) * 1e-2)
    # a real GBV: three keyframes of the loop fused by the product kernel (bit-exact vs the reference kernel elsewhere)
    mv = MapVolume(cfg, m, K)
    mv.init_mapvolume()
    for f in (0, 5, 10):
        d, c = synth.render_frame(scene, K, H, W, poses[f], seed=f)
        mv.integrate_kf({"rgb": torch.from_numpy(c), "depth": torch.from_numpy(d)}, torch.from_numpy(poses[f]).float())
    c2w = poses[3].astype(np.float32)
    depth, rgb = synth.render_frame(scene, K, H, W, c2w, seed=3)
    dirs = torch.from_numpy(synth.camera_dirs(K, H, W).reshape(-1, 3)).to(cuda)
    c2w_t = torch.from_numpy(c2w).to(cuda)
    rays_d = torch.sum(dirs[..., None, :] * c2w_t[:3, :3], -1).contiguous()
    rays_o = c2w_t[None, :3, -1].repeat(H * W, 1).contiguous()
    td = torch.from_numpy(depth).to(cuda).reshape(-1, 1).contiguous()
    tc = torch.from_numpy(rgb).to(cuda).reshape(-1, 3).contiguous()
    if shape == "cfg3":                                    # 2^20 rays drawn from the frame's pixels (with repetition)
        pick = torch.randint(0, H * W, (1 << 20,), generator=torch.Generator().manual_seed(5)).to(cuda)
        rays_o, rays_d, td, tc = (t[pick].contiguous() for t in (rays_o, rays_d, td, tc))
    # the oracle twin: same table, volume and decoder
    h = hash_standin(cfg); g = gbv_standin(cfg)
    assert h.params.numel() == m.embed_res_fn.params.numel() and g.params.numel() == m.GBV.params.numel()
    with torch.no_grad():
        h.params.copy_(m.embed_res_fn.params.cpu()); g.params.copy_(m.GBV.params.cpu())
    g.params.requires_grad_(False)
    ws = [w.detach().cpu().clone().requires_grad_(True) for w in m.decoder_res.fused_weights()]
    return cfg, m, RayOracle(cfg, bb, h, g, *ws), h, ws, (rays_o, rays_d, tc, td)


@pytest.mark.parametrize("shape", ["cfg2", "cfg3"])
def test_baseline_shape_subset_matches_oracle(cuda, rf_lib, shape):
    cfg, m, orc, h, ws, (rays_o, rays_d, tc, td) = _setup(cuda, shape)
    n = rays_o.shape[0]
    S = cfg["training"]["n_range_d"] + cfg["training"]["n_samples_d"]
    assert (n, S) == ((816000, 59) if shape == "cfg2" else (1 << 20, 48))
    gen = torch.Generator().manual_seed(23)
    sub = torch.sort(torch.randperm(n, generator=gen)[:N_SUB]).values
    u_sub = torch.rand(N_SUB, S, generator=gen)
    u_full = torch.rand(n, S, device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
    sub_c = sub.to(cuda)
    u_full[sub_c] = u_sub.to(cuda)

    # ---- (1) the FULL-size launch, subset rows vs the oracle -------------------------------------------------------
    m.eval()
    with torch.no_grad():
        ret = m.mapping(rays_o, rays_d, tc, td, u=u_full)
    sl = lambda t: t[sub_c].cpu()
    ro_s, rd_s, tc_s, td_s = (t[sub_c].cpu() for t in (rays_o, rays_d, tc, td))
    orc.training = False
    with torch.no_grad():
        ref = orc.mapping(ro_s, rd_s, tc_s, td_s, u=u_sub)
    assert np.array_equal(sl(ret["z_vals"]).numpy(), ref["z_vals"].numpy()), "z_vals of the full launch: not bit-exact"
    _close(sl(ret["raw"]), ref["raw"].numpy(), 2e-4, "raw (full launch)")
    _close(sl(ret["rgb_res_map"]), ref["rgb_res_map"].numpy(), 2e-4, "rgb_res_map (full launch)")
    _close(sl(ret["depth_res_map"]), ref["depth_res_map"].numpy(), 2e-4, "depth_res_map (full launch)")
    del ret, u_full

    # ---- (2) losses + gradients of the subset batch at the full table / volume / decoder sizes ---------------------
    for mode in ("mapping", "ba"):
        ba = mode == "ba"
        orc.training = True
        for p in [h.params] + ws:
            p.grad = None
        ro_r = ro_s.clone().requires_grad_(ba); rd_r = rd_s.clone().requires_grad_(ba)
        r_ref = orc.mapping(ro_r, rd_r, tc_s, td_s, clamp=ba, u=u_sub)
        orc.total_loss(r_ref).backward()
        m.train()
        params = [m.embed_res_fn.params] + list(m.decoder_res.fused_weights())
        for p in params:
            p.grad = None
        ro_c = ro_s.to(cuda).requires_grad_(ba); rd_c = rd_s.to(cuda).requires_grad_(ba)
        r = m.mapping(ro_c, rd_c, tc_s.to(cuda), td_s.to(cuda), clamp=ba, u=u_sub)
        _total(cfg, r).backward()
        for k in ("rgb_res_loss", "depth_res_loss", "sdf_res_loss", "fs_res_loss", "rgb_res", "depth_res"):
            _close(r[k], r_ref[k].detach().numpy(), 2e-4, f"{mode}: {k}")
        _close_grad(params[0].grad, h.params.grad.numpy(), f"{shape}/{mode}: g_hash", 1, kink_aware=True)
        for w_cuda, w_ref, nm in zip(params[1:], ws, ("sdf0", "sdf1", "col0", "col1")):
            _close_grad(w_cuda.grad, w_ref.grad.numpy(), f"{shape}/{mode}: g_w_{nm}", 1, kink_aware=True)
        if ba:
            _close_grad(ro_c.grad, ro_r.grad.numpy(), f"{shape}/{mode}: g_rays_o", 1, kink_aware=True)
            _close_grad(rd_c.grad, rd_r.grad.numpy(), f"{shape}/{mode}: g_rays_d", 1, kink_aware=True)
