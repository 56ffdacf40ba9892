"""Host logic of remixfusion_b200/lattice.py with a pure-torch stand-in model (no GPU): lattice construction, float64
normalisation, slab loop and output shapes equal the reference's evaluation order (utils.py:78-157, :188-203)."""
import numpy as np
import torch

from remixfusion_b200.lattice import getVoxels, query_lattice, query_vertex_colors


class _Model:
    def query_sdf_res(self, q):
        x = q.to(torch.float32)
        return torch.sin(3 * x[..., 0]) + x[..., 1] * x[..., 2]

    def query_w_res(self, q):
        return (q.to(torch.float32)[..., 2] < 0.5).float()

    def query_color_residual(self, q):
        x = q.to(torch.float32).reshape(-1, 3)
        return torch.stack([x[:, 0] * 2 - 0.5, x[:, 1], 1.5 * x[:, 2]], -1)


def test_lattice_matches_reference_order_on_cpu():
    cfg = {"grid": {"tcnn_encoding": True}}
    bb = torch.tensor([[-1.0, 7.0], [-1.3, 3.7], [-1.7, 1.4]], dtype=torch.float64)
    mcb = torch.tensor([[-0.5, 6.0], [-1.0, 3.0], [-1.5, 1.0]], dtype=torch.float64)
    m = _Model()
    tsdf, mask, (tx, ty, tz) = query_lattice(m, cfg, bb, marching_cube_bound=mcb, voxel_size=0.25, slab_points=500)
    # utils.py:131-157
    x_min, y_min, z_min = mcb[:, 0]; x_max, y_max, z_max = mcb[:, 1]
    rx, ry, rz = getVoxels(x_max, x_min, y_max, y_min, z_max, z_min, 0.25, None)
    assert rx.numel() == round(6.5 / 0.25 + 0.0005) + 1 and torch.equal(rx, tx) and torch.equal(rz, tz)
    pts = torch.stack(torch.meshgrid(rx, ry, rz, indexing="ij"), -1).to(torch.float32)
    flat = pts.reshape(-1, 3)
    flat = (flat - bb[:, 0]) / (bb[:, 1] - bb[:, 0])
    chunk = 1024
    raw = torch.cat([m.query_sdf_res(flat[i:i + chunk, None, :]) for i in range(0, flat.shape[0], chunk)], 0)
    w = torch.cat([m.query_w_res(flat[i:i + chunk, None, :]) for i in range(0, flat.shape[0], chunk)], 0)
    assert tsdf.shape == pts.shape[:-1]
    assert torch.equal(tsdf, raw.reshape(pts.shape[:-1])) and torch.equal(mask, w.reshape(pts.shape[:-1]) > 0)
    # resolution form
    t2, m2, (ax, _, _) = query_lattice(m, cfg, bb, resolution=9)
    assert t2.shape == (9, 9, 9) and ax.numel() == 9 and float(ax[0]) == -1.0 and float(ax[-1]) == 7.0
    # vertex colours: clip(0, 1) * 255 (utils.py:203)
    verts = np.random.default_rng(0).random((777, 3)) * 4 - 1
    col = query_vertex_colors(m, cfg, bb, verts, slab_points=100)
    vf = (torch.from_numpy(verts).to(bb) - bb[:, 0]) / (bb[:, 1] - bb[:, 0])
    ref = torch.clip(m.query_color_residual(vf[:, None, :]), 0, 1) * 255
    assert col.shape == (777, 3) and torch.equal(col, ref)
