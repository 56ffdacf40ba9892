"""N4 (first half): device-resident lattice sweep vs the reference's evaluation order (65 536-point chunks built and
normalised on the CPU, utils.py:123-157)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_lattice_sweep_matches_chunked_reference_order(cuda, rf_lib):
    from remixfusion_b200 import configs
    from remixfusion_b200.lattice import getVoxels, query_lattice, query_vertex_colors
    from remixfusion_b200.scene_rep import JointEncoding
    cfg = configs.replica()
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
    torch.manual_seed(3)
    m = JointEncoding(cfg, bb).to(cuda).eval()
    with torch.no_grad():
        m.embed_res_fn.params.copy_((torch.rand_like(m.embed_res_fn.params) * 2 - 1) * 1e-2)
        m.GBV.params.copy_((torch.rand_like(m.GBV.params) * 2 - 1) * 0.5)
        w = torch.zeros_like(m.GBW.params); w[: w.numel() // 2] = 1.0       # observed weight in the lower half (index = x + y R + z R^2)
        m.GBW.params.copy_(w)
    bbd = bb.to(cuda)
    voxel = 0.11                                              # 73 x 46 x 29 lattice = 97 k points: two reference chunks
    tsdf, mask, (tx, ty, tz) = query_lattice(m, cfg, bbd, voxel_size=voxel, slab_points=30000)      # several slabs
    # the reference's order of operations (utils.py:131-157)
    rx, ry, rz = getVoxels(bb[0, 1], bb[0, 0], bb[1, 1], bb[1, 0], bb[2, 1], bb[2, 0], voxel, None)
    assert torch.equal(rx, tx) and torch.equal(ry, ty) and torch.equal(rz, tz)
    pts = torch.stack(torch.meshgrid(rx, ry, rz, indexing="ij"), -1).to(torch.float32)
    flat = pts.reshape(-1, 3)
    flat = (flat - bb[:, 0]) / (bb[:, 1] - bb[:, 0])
    chunk = 1024 * 64
    with torch.no_grad():
        raw = [m.query_sdf_res(flat[i:i + chunk, None, :].to(cuda)).cpu() for i in range(0, flat.shape[0], chunk)]
        w = [m.query_w_res(flat[i:i + chunk, None, :].to(cuda)).cpu() for i in range(0, flat.shape[0], chunk)]
    ref_tsdf = torch.cat(raw, 0).reshape(pts.shape[:-1]); ref_w = torch.cat(w, 0).reshape(pts.shape[:-1])
    assert tsdf.shape == ref_tsdf.shape and tsdf.shape[0] > 60
    assert torch.equal(tsdf.cpu(), ref_tsdf)                  # per-point kernels: batch composition does not change the bits
    assert torch.equal(mask.cpu(), ref_w > 0) and 0.2 < float(mask.float().mean()) < 0.8
    # vertex colours (utils.py:188-203)
    verts = (bb[:, 0] + torch.rand(5000, 3, dtype=torch.float64, generator=torch.Generator().manual_seed(1)) * (bb[:, 1] - bb[:, 0])).numpy()
    col = query_vertex_colors(m, cfg, bbd, verts, slab_points=2048)
    vf = (torch.from_numpy(verts).to(bbd) - bbd[:, 0]) / (bbd[:, 1] - bbd[:, 0])
    with torch.no_grad():
        ref = torch.cat([m.query_color_residual(vf[i:i + chunk, None, :]) for i in range(0, vf.shape[0], chunk)], 0).reshape(-1, 3)
    assert torch.equal(col, torch.clip(ref, 0, 1) * 255)
