"""SURVEY §8f N4 (first half): the dense-lattice sweep that feeds mesh extraction, kept on the device.

The reference's ``extract_mesh_github`` (utils.py:123-212) builds the marching-cubes lattice on the CPU, normalises it
there, and evaluates ``model.query_sdf_res`` / ``model.query_w_res`` in 65 536-point chunks with a host copy and a
``.cpu()`` per chunk (utils.py:111-146).  ``query_lattice`` does the same evaluation — same lattice (``getVoxels``, utils.py:78-
103), same float64 normalisation, same per-point kernels (``rf_point_query_forward`` / grid encode) — in slabs of millions
of points that never leave the GPU, and returns the ``tsdf`` and ``mask`` volumes ``measure.marching_cubes`` is called with
(utils.py:157-173).  Marching cubes itself and the mesh export are not built."""
from __future__ import annotations

import torch


def getVoxels(x_max, x_min, y_max, y_min, z_max, z_min, voxel_size=None, resolution=None):
    """utils.py:78-103: lattice coordinates per axis (CPU float32 linspace, as the reference)."""
    x_max, x_min, y_max, y_min, z_max, z_min = (float(v) for v in (x_max, x_min, y_max, y_min, z_max, z_min))
    if voxel_size is not None:
        Nx = round((x_max - x_min) / voxel_size + 0.0005)
        Ny = round((y_max - y_min) / voxel_size + 0.0005)
        Nz = round((z_max - z_min) / voxel_size + 0.0005)
        return torch.linspace(x_min, x_max, Nx + 1), torch.linspace(y_min, y_max, Ny + 1), torch.linspace(z_min, z_max, Nz + 1)
    return torch.linspace(x_min, x_max, resolution), torch.linspace(y_min, y_max, resolution), torch.linspace(z_min, z_max, resolution)


@torch.no_grad()
def query_lattice(model, config, bounding_box, marching_cube_bound=None, voxel_size=None, resolution=None, slab_points=1 << 23):
    """Returns (tsdf [X,Y,Z] float32, mask [X,Y,Z] bool, (tx, ty, tz)) on ``bounding_box.device``.

    tsdf = ``query_sdf_res`` over the lattice (utils.py:143-157), mask = ``query_w_res > 0`` (:161).  ``slab_points`` bounds
    the points per call (x-slabs of the lattice), i.e. the size of the query workspace (160 B per point)."""
    if marching_cube_bound is None:
        marching_cube_bound = bounding_box
    dev = bounding_box.device
    x_min, y_min, z_min = marching_cube_bound[:, 0]
    x_max, y_max, z_max = marching_cube_bound[:, 1]
    tx, ty, tz = getVoxels(x_max, x_min, y_max, y_min, z_max, z_min, voxel_size, resolution)
    X, Y, Z = tx.numel(), ty.numel(), tz.numel()
    txd, tyd, tzd = tx.to(dev), ty.to(dev), tz.to(dev)
    tsdf = torch.empty(X, Y, Z, dtype=torch.float32, device=dev)
    weight = torch.empty(X, Y, Z, dtype=torch.float32, device=dev)
    b0 = bounding_box[:, 0]; bl = bounding_box[:, 1] - bounding_box[:, 0]          # float64, as the reference's bounding_box
    per = max(1, int(slab_points) // max(1, Y * Z))
    for x0 in range(0, X, per):
        x1 = min(X, x0 + per)
        pts = torch.stack(torch.meshgrid(txd[x0:x1], tyd, tzd, indexing="ij"), -1).to(torch.float32).reshape(-1, 3)
        if config["grid"]["tcnn_encoding"]:
            pts = (pts - b0) / bl                                                      # :138-139 (float32 - float64 -> float64)
        q = pts[:, None, :]
        tsdf[x0:x1] = model.query_sdf_res(q).reshape(x1 - x0, Y, Z)
        weight[x0:x1] = model.query_w_res(q).reshape(x1 - x0, Y, Z)
    return tsdf, weight > 0, (tx, ty, tz)


@torch.no_grad()
def query_vertex_colors(model, config, bounding_box, vertices, slab_points=1 << 23):
    """utils.py:188-203: colours (0..255) of mesh vertices given in the lattice's world frame; ``vertices`` [V,3] tensor or
    array.  One device-resident sweep instead of 65 536-vertex chunks with a ``.cpu()`` each."""
    v = torch.as_tensor(vertices).to(bounding_box)
    if config["grid"]["tcnn_encoding"]:
        v = (v - bounding_box[:, 0]) / (bounding_box[:, 1] - bounding_box[:, 0])
    out = torch.empty(v.shape[0], 3, dtype=torch.float32, device=bounding_box.device)
    for i0 in range(0, v.shape[0], int(slab_points)):
        out[i0:i0 + slab_points] = model.query_color_residual(v[i0:i0 + slab_points, None, :]).reshape(-1, 3)
    return torch.clip(out, 0, 1) * 255
