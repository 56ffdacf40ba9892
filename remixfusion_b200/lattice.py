"""SURVEY §8f N4 (first half): the dense-lattice sweep that feeds mesh extraction, kept on the device.

The reference's ``extract_mesh_github`` (utils.py:123-212) builds the marching-cubes lattice on the CPU, normalises it
there, and evaluates ``model.query_sdf_res`` / ``model.query_w_res`` in 65 536-point chunks with a host copy and a
``.cpu()`` per chunk (utils.py:111-146).  ``query_lattice`` does the same evaluation — same lattice (``getVoxels``, utils.py:78-
103), same float64 normalisation, same per-point kernels (``rf_point_query_forward`` / grid encode) — in slabs of millions
of points that never leave the GPU, and returns the ``tsdf`` and ``mask`` volumes ``measure.marching_cubes`` is called with
(utils.py:157-173).

N4 (second half): ``marching_cubes`` extracts the iso-surface of such a volume on the device (csrc/marching_cubes.cu: the
dual-grid algorithm of the reference's thirdparty/NumpyMarchingCubes), ``extract_mesh`` chains lattice sweep -> marching
cubes -> the reference's rescaling to metric units (utils.py:176-186) -> vertex colours.  Writing the mesh file (trimesh) is
left to the caller."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import abi


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


@torch.no_grad()
def marching_cubes(volume, isovalue=0.0, truncation=3.0, mask=None):
    """``mcubes.marching_cubes(volume, isovalue, truncation)`` (thirdparty/NumpyMarchingCubes; utils.py:169) on the device.

    volume [X,Y,Z] float32 CUDA tensor; voxels with |d| >= truncation (or NaN / -inf) are invalid; ``mask`` (bool [X,Y,Z],
    e.g. ``weight > 0`` as in utils.py:161-170) marks further voxels invalid.  Returns (vertices [V,3] float32 in voxel
    units, faces [F,3] int64), both on the device: the triangle soup is the reference's (same cells, same case table,
    same interpolation, same order); vertices are welded by exact identity (the dual-grid edge, or the lattice corner a
    vertex was snapped to) with the reference's first-occurrence numbering, then degenerate and duplicate faces are dropped
    as in its merge_close_vertices / remove_duplicate_faces."""
    if not volume.is_cuda:
        raise abi.RfError("marching_cubes: a CUDA tensor is required (no CPU fallback)")
    vol = volume.detach().to(torch.float32)
    if mask is not None:
        vol = torch.where(mask.to(vol.device), vol, torch.full_like(vol, float("nan")))
    vol = vol.contiguous()
    X, Y, Z = (int(d) for d in vol.shape)
    dev = vol.device
    L = abi.lib()
    L.rf_mc_corner_floats.restype = C.c_int64
    corner = torch.empty(int(L.rf_mc_corner_floats(X, Y, Z)), dtype=torch.float32, device=dev)
    counts = torch.empty(X * Y * Z, dtype=torch.int32, device=dev)
    abi.check(L.rf_mc_count(abi.dptr(vol), C.c_int(X), C.c_int(Y), C.c_int(Z), C.c_float(isovalue), C.c_float(truncation), abi.dptr(corner),
                            abi.dptr(counts), abi.stream_ptr()), "rf_mc_count")
    ends = torch.cumsum(counts, 0, dtype=torch.int64)
    n_tri = int(ends[-1]) if ends.numel() else 0
    if n_tri == 0:
        return torch.zeros(0, 3, dtype=torch.float32, device=dev), torch.zeros(0, 3, dtype=torch.int64, device=dev)
    offsets = (ends - counts).contiguous()
    tris = torch.empty(n_tri, 3, 3, dtype=torch.float32, device=dev)
    keys = torch.empty(n_tri, 3, dtype=torch.int64, device=dev)
    abi.check(L.rf_mc_emit(abi.dptr(corner), C.c_int(X), C.c_int(Y), C.c_int(Z), C.c_float(isovalue), abi.dptr(offsets), abi.dptr(tris),
                           abi.dptr(keys), abi.stream_ptr()), "rf_mc_emit")
    # weld: one vertex per key, numbered by first occurrence in the soup (the reference's order), value = that occurrence
    flat_keys = keys.reshape(-1)
    uniq, inverse = torch.unique(flat_keys, return_inverse=True)
    slot = torch.arange(flat_keys.numel(), device=dev)
    first = torch.full((uniq.numel(),), flat_keys.numel(), dtype=torch.int64, device=dev).scatter_reduce_(0, inverse, slot, "amin")
    order = torch.argsort(first)                               # unique ids in order of first occurrence
    rank = torch.empty_like(order); rank[order] = torch.arange(order.numel(), device=dev)
    vertices = tris.reshape(-1, 3)[first[order]]
    faces = rank[inverse].reshape(-1, 3)
    # degenerate faces (a repeated vertex), then duplicate faces (same vertex set; the first one stays)
    ok = (faces[:, 0] != faces[:, 1]) & (faces[:, 1] != faces[:, 2]) & (faces[:, 0] != faces[:, 2])
    faces = faces[ok]
    if faces.numel():
        srt, _ = torch.sort(faces, dim=1)
        V = int(vertices.shape[0]) + 1
        code = (srt[:, 0] * V + srt[:, 1]) * V + srt[:, 2] if V < (1 << 20) else None
        if code is not None:
            _, inv = torch.unique(code, return_inverse=True)
        else:
            _, inv = torch.unique(srt, dim=0, return_inverse=True)
        idx = torch.arange(faces.shape[0], device=dev)
        keep = torch.full((int(inv.max()) + 1,), faces.shape[0], dtype=torch.int64, device=dev).scatter_reduce_(0, inv, idx, "amin")
        faces = faces[torch.sort(keep).values]
    return vertices, faces


@torch.no_grad()
def extract_mesh(model, config, bounding_box, marching_cube_bound=None, voxel_size=None, resolution=None, isolevel=0.0, truncation=3.0,
                 color=True, slab_points=1 << 23):
    """utils.py:123-212 without the file export: lattice sweep (``query_lattice``), marching cubes over (tsdf, weight > 0),
    vertices mapped to the lattice's world frame and to metric units (:173-186), optional vertex colours (:188-203).
    Returns (vertices [V,3] float64 numpy, faces [F,3] int64 numpy, colours [V,3] uint8-range float32 numpy or None)."""
    tsdf, mask, (tx, ty, tz) = query_lattice(model, config, bounding_box, marching_cube_bound, voxel_size, resolution, slab_points)
    v, f = marching_cubes(tsdf, isolevel, truncation, mask=mask)
    vertices = v.double().cpu().numpy()
    vertices /= np.array([[tx.shape[0] - 1, ty.shape[0] - 1, tz.shape[0] - 1]])
    txn, tyn, tzn = tx.numpy(), ty.numpy(), tz.numpy()
    scale = np.array([txn[-1] - txn[0], tyn[-1] - tyn[0], tzn[-1] - tzn[0]])
    offset = np.array([txn[0], tyn[0], tzn[0]])
    vertices = scale[np.newaxis, :] * vertices + offset
    vertices = vertices / config["data"]["sc_factor"] - config["data"].get("translation", 0)
    colours = query_vertex_colors(model, config, bounding_box, np.ascontiguousarray(vertices), slab_points).cpu().numpy() if color and len(vertices) else None
    return vertices, f.cpu().numpy(), colours
