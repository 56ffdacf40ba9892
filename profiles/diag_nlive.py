"""Diagnostic: distribution of n_live (last sample with a non-zero upstream gradient, csrc/ray_query.cu composite_bwd_kernel) on
the bench frame, and the share of 128-ray backward tiles that are entirely dead, in pixel order and valid-depth-first order."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                   # noqa: E402
from remixfusion_b200 import abi, configs, synth               # noqa: E402
from remixfusion_b200.global_volume import MapVolume           # noqa: E402
from remixfusion_b200.scene_rep import JointEncoding           # noqa: E402

dev = torch.device("cuda:0")
cfg = configs.replica()
cam = cfg["cam"]; H, W = cam["H"], cam["W"]
S = cfg["training"]["n_range_d"] + cfg["training"]["n_samples_d"]
K, poses, frames = bench.make_frames(cfg, 2)
bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
torch.manual_seed(0)
model = JointEncoding(cfg, bb).to(dev)
with torch.no_grad():
    model.embed_res_fn.params.copy_((torch.rand_like(model.embed_res_fn.params) * 2 - 1) * 1e-2)
mv = MapVolume(cfg, model, K); mv.init_mapvolume()
for c2w, depth, rgb in frames:
    mv.integrate_kf({"rgb": torch.from_numpy(rgb), "depth": torch.from_numpy(depth)}, torch.from_numpy(c2w).float())
c2w, depth, rgb = frames[0]
dirs = torch.from_numpy(synth.camera_dirs(K, H, W).reshape(-1, 3)).to(dev)
c2w_t = torch.from_numpy(c2w.astype(np.float32)).to(dev)
rays_d = torch.sum(dirs[..., None, :] * c2w_t[:3, :3], -1).contiguous()
rays_o = c2w_t[None, :3, -1].repeat(H * W, 1).contiguous()
td = torch.from_numpy(depth).to(dev).reshape(-1).contiguous(); tc = torch.from_numpy(rgb).to(dev).reshape(-1, 3).contiguous()
n = H * W
L = abi.lib()


def run(order):
    ro, rd, tdd, tcc = rays_o[order].contiguous(), rays_d[order].contiguous(), td[order].contiguous(), tc[order].contiguous()
    z = model.sample_z(tdd.reshape(-1, 1), n)
    meta = model._meta(True); cfgc = meta["cfg"]; cfgc.n_rays_total = n
    w = model.decoder_res.fused_weights()
    p = abi.RayParams(abi.dptr(model.embed_res_fn.params.detach()), abi.dptr(model.GBV.params.detach()), *[abi.dptr(x.detach()) for x in w])
    raw = torch.empty(n, S, 4, device=dev); rgbm = torch.empty(n, 3, device=dev); dm = torch.empty(n, device=dev)
    part = torch.zeros(8, dtype=torch.float64, device=dev)
    ws = torch.empty(int(L.rf_ray_workspace_floats(C.byref(cfgc), C.byref(meta["hash_desc"]), C.c_int64(n))), device=dev)
    abi.check(L.rf_ray_query_forward(C.byref(cfgc), C.byref(meta["hash_desc"]), C.byref(meta["gbv_desc"]), C.byref(p), abi.dptr(ro), abi.dptr(rd),
                                     abi.dptr(tdd), abi.dptr(tcc), abi.dptr(z), C.c_int64(n), abi.dptr(raw), abi.dptr(rgbm), abi.dptr(dm),
                                     abi.dptr(part), abi.dptr(ws), abi.stream_ptr()), "fwd")
    g_hash = torch.zeros_like(model.embed_res_fn.params); gws = [torch.zeros_like(x) for x in w]
    grads = abi.RayGrads(abi.dptr(g_hash), *[abi.dptr(x) for x in gws], None, None)
    scratch = torch.empty(int(L.rf_ray_scratch_floats(C.byref(cfgc), C.byref(meta["hash_desc"]), C.c_int64(n), C.c_int(0))), device=dev)
    lg = torch.tensor([5.0, 0.1, 1000.0, 10.0], device=dev)
    abi.check(L.rf_ray_query_backward(C.byref(cfgc), C.byref(meta["hash_desc"]), C.byref(meta["gbv_desc"]), C.byref(p), abi.dptr(ro), abi.dptr(rd),
                                      abi.dptr(tdd), abi.dptr(tcc), C.c_int64(n), abi.dptr(z), abi.dptr(raw), abi.dptr(rgbm), abi.dptr(dm),
                                      None, None, None, abi.dptr(lg), abi.dptr(part), C.byref(grads), abi.dptr(ws), abi.dptr(scratch),
                                      abi.stream_ptr()), "bwd")
    torch.cuda.synchronize()
    nl = scratch[4 * n * S: 4 * n * S + n].view(torch.int32).clone()
    return nl, tdd


for name, order in (("pixel order", torch.arange(n, device=dev)),
                    ("valid depth first", torch.sort((td <= 0).to(torch.uint8), stable=True).indices)):
    nl, tdd = run(order)
    nlc = nl.cpu().numpy(); v = (tdd > 0).cpu().numpy()
    pad = (-n) % 128
    tiles = np.concatenate([nlc, np.zeros(pad, np.int32)]).reshape(-1, 128).max(1)          # per 128-ray group: max n_live
    dead = sum(int((tiles <= s).sum()) for s in range(S)) / (len(tiles) * S)
    print(f"{name}: mean n_live {nlc.mean():.1f} (valid rays {nlc[v].mean():.1f}, p95 {np.percentile(nlc[v], 95):.0f}; invalid rays "
          f"{nlc[~v].mean() if (~v).any() else 0:.1f}, n = {(~v).sum()}); live rows {nlc.sum() / (n * S):.3f}; dead tiles {dead:.3f}")
    print("  histogram of n_live:", np.bincount(nlc, minlength=S + 1).tolist())
