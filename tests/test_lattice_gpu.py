"""N4 (first half): device-resident lattice sweep (remixfusion_b200/lattice.py) vs (a) the CPU oracle evaluating the same lattice
the way the reference does (lattice built and normalised on the CPU, utils.py:123-157; query_sdf_res / query_w_res of
model/scene_rep.py:212-282 on the stand-in encoders) and (b) the reference's chunked evaluation order; then the whole
extract_mesh chain against the reference's compiled marching cubes on the oracle's lattice."""
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


def test_lattice_and_mesh_match_cpu_oracle(cuda, rf_lib):
    """tsdf / mask volumes of query_lattice vs the CPU oracle (oracle/ray_oracle.py + stand-in encoders) on the same lattice, and
    extract_mesh vs the reference's C++ marching cubes (oracle/_ref/libmc_ref.so) applied to the oracle's volume."""
    from oracle import mc_oracle, tcnn_standin
    from remixfusion_b200.lattice import extract_mesh, getVoxels, query_lattice
    from tests import _ray_common as R
    from tests._mc_common import same_surface
    from tests.test_ray_gpu import G, _model_from_golden
    cfg, m = _model_from_golden("A", cuda)
    m.eval()
    cfg["data"]["translation"] = 0
    _, orc = R.oracle_from_golden(G, "A", requires_grad=False)
    gbw = tcnn_standin.GridStandIn(1, 1, False, 0, cfg["globalV"]["base_resolution"], 1)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        wv = torch.zeros(gbw.params.shape); wv[: int(0.6 * wv.numel())] = 1.0          # observed weight in the lower 60 % of z (index = x + y R + z R^2)
        gbw.params.copy_(wv)
        m.GBW.params.copy_(gbw.params)
    bb = orc.bounding_box
    voxel = 0.16
    tsdf, mask, (tx, ty, tz) = query_lattice(m, cfg, bb.to(cuda), voxel_size=voxel, slab_points=20000)
    rx, ry, rz = getVoxels(bb[0, 1], bb[0, 0], bb[1, 1], bb[1, 0], bb[2, 1], bb[2, 0], voxel, None)
    pts = torch.stack(torch.meshgrid(rx, ry, rz, indexing="ij"), -1).to(torch.float32)
    flat = ((pts.reshape(-1, 3) - bb[:, 0]) / (bb[:, 1] - bb[:, 0])).float()               # float64 (utils.py:138-139), narrowed at the encoder boundary
    with torch.no_grad():
        ex = orc.GBV(flat)
        t = torch.clamp(ex[..., 0] * cfg["training"]["c_trunc"] / cfg["training"]["trunc"], -1, 1)                # scene_rep.py:230-233
        sdf = (torch.relu(torch.cat([orc.embed_res_fn(flat), orc.embedpos_fn(flat), t[:, None]], -1) @ orc.w_sdf0.t()) @ orc.w_sdf1.t())[:, 0] + t
        w = gbw(flat)[..., 0]
    ref_tsdf = sdf.reshape(pts.shape[:-1]).numpy(); ref_mask = (w.reshape(pts.shape[:-1]) > 0).numpy()
    np.testing.assert_allclose(tsdf.cpu().numpy(), ref_tsdf, rtol=2e-4, atol=2e-4)
    assert (mask.cpu().numpy() != ref_mask).mean() < 1e-3 and 0.2 < ref_mask.mean() < 0.9      # weights ~0 may round either way
    # mesh: product chain vs the compiled reference on the PRODUCT's volume (so that only marching cubes + rescaling are compared)
    assert mc_oracle.available()
    iso = float(np.median(tsdf.cpu().numpy()[mask.cpu().numpy()]))        # a level the random field crosses everywhere
    verts, faces, colours = extract_mesh(m, cfg, bb.to(cuda), voxel_size=voxel, isolevel=iso, truncation=3.0)
    vol = np.where(mask.cpu().numpy(), tsdf.cpu().numpy(), np.nan).astype(np.float32)
    Vr, Fr = mc_oracle.marching_cubes(vol, iso, 3.0)
    Vr = Vr / np.array([[rx.numel() - 1, ry.numel() - 1, rz.numel() - 1]])
    scale = np.array([float(rx[-1] - rx[0]), float(ry[-1] - ry[0]), float(rz[-1] - rz[0])]); off = np.array([float(rx[0]), float(ry[0]), float(rz[0])])
    Vr = scale[None] * Vr + off
    ok, msg = same_surface(verts, faces, Vr, Fr, atol=1e-4)
    assert ok and faces.shape[0] > 100, msg
    assert colours is not None and colours.shape == (verts.shape[0], 3) and float(colours.min()) >= 0 and float(colours.max()) <= 255
