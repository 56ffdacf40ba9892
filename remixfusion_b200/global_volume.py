"""Global coarse TSDF volume ("GBV") integration — host-side mirror of the three hot-path methods of the
reference's ``Mapper``: ``create_global_volume`` (mp_slam/mapper.py:213-254), ``init_mapvolume`` (:267-282)
and ``integrate_kf`` (:823-872).  Method names, arguments and the in-place contract on
``model.GBV.params`` / ``model.GBW.params`` are the reference's; the kernels are librf_b200's.

The reference ``Mapper`` can adopt it by delegation:

    self._gv = MapVolume(config, model, K)       # in Mapper.__init__ instead of create_global_volume(...)
    init_mapvolume = lambda self: self._gv.init_mapvolume()
    integrate_kf   = lambda self, batch, pose, obs_weight=1.0: self._gv.integrate_kf(batch, pose, obs_weight)
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import abi


class MapVolume:
    def __init__(self, config, model, K, z_slab=None):
        """config: the reference's config dict; model: object with ``GBV.params`` [4*R^3] and ``GBW.params`` [R^3]
        fp32 CUDA parameters (tcnn Dense-grid layout); K: 3x3 intrinsics.
        z_slab: (z0, z1) owned by this rank when the volume is sharded over GPUs (params then hold the slab only)."""
        self.config = config
        self.model = model
        self.K = np.asarray(K, dtype=np.float64).reshape(3, 3)
        self.create_global_volume(config["globalV"]["base_resolution"])
        self.z_slab = (0, int(self.vol_dim[2])) if z_slab is None else (int(z_slab[0]), int(z_slab[1]))
        self.slab_local = 0 if z_slab is None else 1
        self._check_params()

    def create_global_volume(self, base_resolution):
        """mp_slam/mapper.py:213-254 (launch geometry is the kernels' business here)."""
        self.vol_dim = np.array([base_resolution, base_resolution, base_resolution])
        self.map_box = self.config["mapping"]["bound"]
        self.voxel_size = 1.0 / base_resolution
        self.vol_origin = np.array([self.map_box[0][0], self.map_box[1][0], self.map_box[2][0]])
        self.box_length = np.array([self.map_box[i][1] - self.map_box[i][0] for i in range(3)])
        self.trunc_margin = self.config["training"]["c_trunc"]

    def _n_own(self) -> int:
        R = int(self.vol_dim[0])
        return (self.z_slab[1] - self.z_slab[0]) * R * R

    def _check_params(self):
        """The parameter tensors must match the addressing mode: a z-slab object owns slab-sized tensors (the kernels then
        write at offset 0), the unsharded object the full R^3 grids (tcnn pads to a multiple of 8 entries).  A full-size model
        passed together with a z_slab would silently be written at the wrong offset."""
        R = int(self.vol_dim[0])
        if not (0 <= self.z_slab[0] < self.z_slab[1] <= R):
            raise abi.RfError(f"MapVolume: z_slab {self.z_slab} outside [0, {R}]")
        n = self._n_own()
        nv, nw = self.model.GBV.params.numel(), self.model.GBW.params.numel()
        pad = 1024                                   # allocation slack a caller may add (tcnn pads to 8 entries)
        if self.slab_local:
            ok = 4 * n <= nv <= 4 * (n + pad) and n <= nw <= n + pad
        else:
            ok = 4 * R ** 3 <= nv <= 4 * (R ** 3 + pad) and R ** 3 <= nw <= R ** 3 + pad
        if not ok:
            raise abi.RfError(f"MapVolume: GBV/GBW parameter sizes ({nv}, {nw}) do not match "
                              f"{'the z-slab ' + str(self.z_slab) if self.slab_local else 'the full volume'} at R = {R} "
                              f"(expected {4 * n} and {n} floats, up to {pad} entries of padding)")

    def init_mapvolume(self):
        """GBV[v] = (1, 0, 0, 0) for every voxel (mp_slam/mapper.py:267-282, kernel :161-183)."""
        self._check_params()
        p = self.model.GBV.params
        rc = abi.lib().rf_tsdf_clear_global(abi.dptr(p.data), C.c_int64(self._n_own()), abi.stream_ptr())
        abi.check(rc, "rf_tsdf_clear_global")

    def integrate_kf(self, batch, pose, obs_weight=1.0):
        """Integrate an RGB-D keyframe into the global volume (mp_slam/mapper.py:823-872).

        batch['rgb']: (H,W,3) or (1,H,W,3) float in [0,1]; batch['depth']: (H,W) or (1,H,W) metres (host or device);
        pose: (4,4) camera-to-world tensor (device tensors are read in place, no host sync)."""
        self._check_params()
        dev = self.model.GBV.params.device
        color_im = batch["rgb"].squeeze().to(dev, non_blocking=True).float().contiguous()
        depth_im = batch["depth"].squeeze().to(dev, non_blocking=True).float().contiguous()
        im_h, im_w = depth_im.shape
        R = int(self.vol_dim[0])
        _b, b_p = abi.farr([self.map_box[0][0], self.map_box[0][1], self.map_box[1][0], self.map_box[1][1],
                            self.map_box[2][0], self.map_box[2][1]], 6)
        _k, k_p = abi.farr(self.K, 9)
        if isinstance(pose, torch.Tensor) and pose.is_cuda:
            pose_d = pose.float().reshape(-1).contiguous()
            c_p, on_dev = C.cast(C.c_void_p(pose_d.data_ptr()), C.POINTER(C.c_float)), 1
        else:
            _c, c_p = abi.farr(pose.detach().cpu().numpy() if isinstance(pose, torch.Tensor) else pose, 16)
            on_dev = 0
        rc = abi.lib().rf_tsdf_integrate_global(
            abi.dptr(self.model.GBV.params.data), abi.dptr(self.model.GBW.params.data), C.c_int(R), b_p, k_p,
            c_p, C.c_int(on_dev), abi.dptr(depth_im), abi.dptr(color_im), C.c_int(im_h), C.c_int(im_w),
            C.c_float(self.trunc_margin), C.c_float(obs_weight),
            C.c_int(self.z_slab[0]), C.c_int(self.z_slab[1]), C.c_int(self.slab_local),
            abi.dptr(abi.pixel_lambda(_k, im_h, im_w, dev)),
            abi.dptr(abi.depth_max(depth_im) if self._n_own() >= abi.FAR_PLANE_MIN_VOXELS else None), abi.stream_ptr())
        abi.check(rc, "rf_tsdf_integrate_global")

    def count_touched(self, depth_im, pose, obs_weight=1.0):
        """Voxels the frame would update (metric numerator)."""
        dev = self.model.GBV.params.device
        depth_im = depth_im.squeeze().to(dev).float().contiguous()
        im_h, im_w = depth_im.shape
        R = int(self.vol_dim[0])
        _b, b_p = abi.farr([v for ax in self.map_box for v in ax], 6)
        _k, k_p = abi.farr(self.K, 9)
        _c, c_p = abi.farr(pose.detach().cpu().numpy() if isinstance(pose, torch.Tensor) else pose, 16)
        counts = torch.zeros(2, dtype=torch.int64, device=dev)
        rc = abi.lib().rf_tsdf_count_global(
            C.c_int(R), b_p, k_p, c_p, C.c_int(0), abi.dptr(depth_im), C.c_int(im_h), C.c_int(im_w),
            C.c_float(self.trunc_margin), abi.dptr(self.model.GBV.params.data), abi.dptr(self.model.GBW.params.data),
            C.c_float(obs_weight), C.c_int(self.z_slab[0]), C.c_int(self.z_slab[1]), C.c_int(self.slab_local),
            C.c_void_p(counts.data_ptr()), abi.stream_ptr())
        abi.check(rc, "rf_tsdf_count_global")
        return int(counts.cpu()[0])
