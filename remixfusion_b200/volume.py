"""Local moving TSDF volume — host-side mirror of ``moving_volume`` (model/Volume.py:19-124, :713-757).

Same constructor arguments, attribute names and ``integrate`` signature as the reference so that
model/ROtracker.py:132,939 can call it unchanged; the work runs in librf_b200.so (``rf_tsdf_integrate_local``)
instead of a PyCUDA JIT kernel.  Device memory is held in torch tensors (``tsdf_vol_gpu`` etc. keep the
reference's names); the raw pointers stay valid for the object's lifetime.

Scope: construction (``center`` bounds, model/Volume.py:1133-1149), ``integrate``, ``clean_volume`` and the volume
re-centring ``copy_volume`` / ``update_tsdf_swap_rot_trans`` (N2).  Point-cloud and mesh dumps are out of scope.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import abi


class moving_volume:
    """Moving volume of RGB-D images (reference: model/Volume.py:19)."""
    _STAGE_SLOTS = 3

    def __init__(self, cfg, traj, init_pose, gpu_mode=True, start=0, device=None, x_slab=None):
        self.config = cfg
        v = cfg["volume"]
        self.voxel_size = float(v["voxel_size"])
        self.surface_trunc = cfg["training"]["trunc"]
        self.trunc_margin = v["trunc"]
        self.fix_x, self.fix_y, self.fix_z = v["x_config"]["fix"], v["y_config"]["fix"], v["z_config"]["fix"]
        self.x_len, self.y_len, self.z_len = v["x_config"]["len"], v["y_config"]["len"], v["z_config"]["len"]
        self.version = v["version"]
        self.t_treshold = v.get("t_treshold", 1)
        self.weight_clamp = v["weight_clamp"]
        self.color_const = 256 * 256
        if not gpu_mode or not torch.cuda.is_available():
            raise abi.RfError("moving_volume: a CUDA device is required (the reference has no CPU path either, "
                              "model/Volume.py:613-619)")
        self.device = torch.device(device if device is not None else "cuda")

        init_pose = np.asarray(init_pose, dtype=np.float64)
        self.vol_bnds = np.asarray(self.initialize_vol_bnd(init_pose, traj, self.version))
        assert self.vol_bnds.shape == (3, 2), "[!] `vol_bnds` should be of shape (3, 2)."
        # model/Volume.py:67-71
        self.vol_dim = np.ceil((self.vol_bnds[:, 1] - self.vol_bnds[:, 0]) / self.voxel_size).copy(order="C").astype(int)
        self.vol_bnds[:, 1] = self.vol_bnds[:, 0] + self.vol_dim * self.voxel_size
        self.vol_origin = self.vol_bnds[:, 0].copy(order="C").astype(np.float32)
        self.start_id = 0
        self.frame_to_Vrange = {}

        dx, dy, dz = (int(d) for d in self.vol_dim)
        if dx * dy * dz >= 2 ** 31:
            raise abi.RfError("moving_volume: more than 2^31 voxels")
        # x-slab owned by this process (multi-GPU sharding): arrays hold only the slab
        self.x_slab = (0, dx) if x_slab is None else (int(x_slab[0]), int(x_slab[1]))
        n_own = (self.x_slab[1] - self.x_slab[0]) * dy * dz
        self.tsdf_vol_gpu = torch.ones(n_own, dtype=torch.float32, device=self.device)       # model/Volume.py:85
        self.weight_vol_gpu = torch.zeros(n_own, dtype=torch.float32, device=self.device)    # :86
        self.color_vol_gpu = torch.zeros(n_own, dtype=torch.float32, device=self.device)     # :87
        self._frame = None
        self._in_flight = []

    # ---- bounds (model/Volume.py:910-925, :1133-1149) ----------------------------------------------------
    def initialize_vol_bnd(self, cam_pose_iter, traj, version):
        if version == "center":
            return self.center_volbnd(np.zeros((3, 2)), cam_pose_iter, traj)
        raise NotImplementedError("volume.version != 'center' (angle-based bounds, model/Volume.py:1151-1201) "
                                  "is volume bookkeeping outside the hot path")

    def center_volbnd(self, vol_bnds, cam_pose_iter, tsdf_cam):
        vol_bnds = np.zeros((3, 2))
        if tsdf_cam is not None:
            tsdf_cam.kfx, tsdf_cam.kfy, tsdf_cam.kfz = cam_pose_iter[0, 3], cam_pose_iter[1, 3], cam_pose_iter[2, 3]
        center_cam = np.round(cam_pose_iter[:3, 3], 0)
        for ax, ln in enumerate((self.x_len, self.y_len, self.z_len)):
            vol_bnds[ax, 0] = center_cam[ax] - ln
            vol_bnds[ax, 1] = center_cam[ax] + ln
        return vol_bnds

    # ---- move policy (model/Volume.py:930-1108): called by the tracker every frame (model/ROtracker.py:920-934) -----
    def check_move_volume_new(self, cur_id, cam_pose_iter, traj, version="center", larger_flag=False, get_pc=False, gap=100):
        """Decide whether the camera has left the volume's comfort zone and, if so, re-centre the volume on it.

        ``traj`` carries the camera position at the last move (``kfx``, ``kfy``, ``kfz``; set by ``center_volbnd``).  An axis
        that is not fixed and whose translation since then exceeds ``volume.t_treshold`` shifts the bounds by that translation;
        the shifted bounds are rounded to whole metres and, if they differ from the current ones, the volume is moved
        (``copy_volume`` + ``update_tsdf_swap_rot_trans``).  Returns ``(moved, old_bnds)`` like the reference; the caller
        records ``frame_to_Vrange[(start, end)] = old_bnds`` (model/ROtracker.py:925-934).  ``larger_flag`` / ``get_pc`` /
        ``gap`` only matter to the angle-based "more" policy."""
        if version != "center":
            raise NotImplementedError("volume.version != 'center' (angle-based swap, model/Volume.py:1008-1081) is volume "
                                      "bookkeeping no shipped config selects")
        pose = np.asarray(cam_pose_iter.detach().cpu().numpy() if isinstance(cam_pose_iter, torch.Tensor) else cam_pose_iter)
        old_bnds = self.vol_bnds.copy()
        new_bnds = self.vol_bnds.copy()
        moved_axis = False
        for ax, (name, fixed) in enumerate((("kfx", self.fix_x), ("kfy", self.fix_y), ("kfz", self.fix_z))):
            shift = pose[ax, 3] - getattr(traj, name)
            if np.abs(shift) > self.t_treshold and not fixed:
                new_bnds[ax, :] += shift
                setattr(traj, name, pose[ax, 3])
                moved_axis = True
        if not moved_axis:
            return False, old_bnds
        new_bnds = np.round(new_bnds, 0)                     # whole metres (round half to even, as Python's round on float64)
        if (new_bnds == old_bnds).all():
            return False, old_bnds
        self.copy_volume()
        self.update_tsdf_swap_rot_trans(new_bnds, old_bnds)
        return True, old_bnds

    def frameid_to_Vrange(self, value):
        """Bounds of the volume that was live at frame ``value`` (model/Volume.py:1084-1105): the recorded range that contains
        it, else the current bounds."""
        for (start, end), bnds in self.frame_to_Vrange.items():
            if start <= value <= end:
                return bnds
        return self.vol_bnds

    # ---- re-centring (N2): model/Volume.py:883-908 copy_volume, :796-858 update_tsdf_swap_rot_trans ----------------
    def copy_volume(self):
        """The reference copies the live arrays into backup arrays here and gathers from the backup in
        ``update_tsdf_swap_rot_trans``.  This implementation ping-pongs between two sets of arrays instead, so there is
        nothing to copy: the live arrays ARE the source of the next swap (half the memory traffic)."""
        return None

    def update_tsdf_swap_rot_trans(self, vol_bnds, old_bnds):
        """Move the volume to ``vol_bnds``: every voxel takes the value of the nearest voxel of the old volume
        (bounds ``old_bnds``) at the same world position, or the cleared value (model/Volume.py:796-858, kernel :128-194)."""
        if self.x_slab != (0, int(self.vol_dim[0])):
            raise abi.RfError("re-centring a slab-sharded moving volume is not built (the tracker's volume is single-GPU)")
        old_dims = tuple(int(d) for d in self.vol_dim)
        self.vol_bnds = np.asarray(vol_bnds, dtype=np.float64)
        self.vol_dim = np.ceil((self.vol_bnds[:, 1] - self.vol_bnds[:, 0]) / self.voxel_size).copy(order="C").astype(int)
        self.vol_bnds[:, 1] = self.vol_bnds[:, 0] + self.vol_dim * self.voxel_size
        self.vol_origin = self.vol_bnds[:, 0].copy(order="C").astype(np.float32)
        old_bnds = np.asarray(old_bnds, dtype=np.float64)
        old_origin = old_bnds[:, 0].copy(order="C").astype(np.float32)
        old_vol_dim = np.ceil((old_bnds[:, 1] - old_bnds[:, 0]) / self.voxel_size).astype(int)
        if tuple(int(d) for d in old_vol_dim) != old_dims:
            raise abi.RfError(f"old_bnds describe {tuple(old_vol_dim)} voxels but the volume holds {old_dims}")
        dx, dy, dz = (int(d) for d in self.vol_dim)
        n = dx * dy * dz
        if n >= 2 ** 31:
            raise abi.RfError("moving_volume: more than 2^31 voxels")
        back = getattr(self, "_back", None)
        if back is None or back[0].numel() != n:
            back = [torch.empty(n, dtype=torch.float32, device=self.device) for _ in range(3)]
        _o, o_p = abi.farr(self.vol_origin, 3)
        _oo, oo_p = abi.farr(old_origin, 3)
        rc = abi.lib().rf_tsdf_recenter(abi.dptr(back[0]), abi.dptr(back[1]), abi.dptr(back[2]),
                                        abi.dptr(self.tsdf_vol_gpu), abi.dptr(self.weight_vol_gpu), abi.dptr(self.color_vol_gpu),
                                        C.c_int(dx), C.c_int(dy), C.c_int(dz), o_p,
                                        C.c_int(old_dims[0]), C.c_int(old_dims[1]), C.c_int(old_dims[2]), oo_p,
                                        C.c_float(self.voxel_size), abi.stream_ptr())
        abi.check(rc, "rf_tsdf_recenter")
        live = [self.tsdf_vol_gpu, self.weight_vol_gpu, self.color_vol_gpu]
        self.tsdf_vol_gpu, self.weight_vol_gpu, self.color_vol_gpu = back
        self._back = live if live[0].numel() == n else None
        self.x_slab = (0, dx)

    # ---- hot path ---------------------------------------------------------------------------------------
    def _stage(self, arr, key):
        """Host numpy / CPU tensor -> device fp32 through a ring of pinned staging buffers (async copy); CUDA tensors
        pass through.  A slot is a (pinned host, device, event) triple: the event is recorded after the slot's copy AND
        the kernels that read its device buffer have been enqueued (``_release``), and the host buffer is rewritten only
        after that event has completed — back-to-back integrate() calls queued behind a long-running kernel therefore never
        overwrite a frame that is still in flight (the reference's ``cuda.In`` copies were synchronous, model/Volume.py:733-749)."""
        if isinstance(arr, torch.Tensor) and arr.is_cuda:
            return arr.to(torch.float32).contiguous()
        if isinstance(arr, torch.Tensor) and arr.is_pinned() and arr.dtype == torch.float32 and arr.is_contiguous():
            # already page-locked: one asynchronous copy, no staging pass over the frame on the host (the caller keeps the
            # tensor unchanged until the stream has consumed it, as with any non_blocking copy)
            return arr.to(self.device, non_blocking=True)
        a = arr.numpy() if isinstance(arr, torch.Tensor) else np.asarray(arr)
        a = np.ascontiguousarray(a, dtype=np.float32)
        if self._frame is None:
            self._frame = {}
        ring = self._frame.get(key)
        if ring is None or ring["shape"] != a.shape:
            ring = {"shape": a.shape, "next": 0, "slots": [
                [torch.empty(a.shape, dtype=torch.float32).pin_memory(), torch.empty(a.shape, dtype=torch.float32, device=self.device), None]
                for _ in range(self._STAGE_SLOTS)]}
            self._frame[key] = ring
        slot = ring["slots"][ring["next"]]
        ring["next"] = (ring["next"] + 1) % len(ring["slots"])
        if slot[2] is not None:
            slot[2].synchronize()                    # the previous user of this slot (copy + kernels) has finished
        slot[0].numpy()[...] = a
        slot[1].copy_(slot[0], non_blocking=True)
        self._in_flight.append(slot)
        return slot[1]

    def _release(self):
        """Mark the staged slots of this call as busy until everything enqueued so far on the stream has run."""
        if self._in_flight:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            for slot in self._in_flight:
                slot[2] = ev
            self._in_flight = []

    def integrate(self, color_im, depth_im, cam_intr, cam_pose, old_bnd, obs_weight=1., reintegrate_flag=0.0):
        """Integrate an RGB-D frame into the TSDF volume (model/Volume.py:713-757).

        color_im: (H, W, 3) RGB with channel values 0..255 (model/ROtracker.py:82,896); depth_im: (H, W) metres;
        cam_intr: (3, 3); cam_pose: (4, 4) camera-to-world; old_bnd: (3, 2) or None.
        """
        depth = self._stage(depth_im, "depth")
        im_h, im_w = depth.shape
        rgb = self._stage(color_im, "rgb")
        if rgb.shape != (im_h, im_w, 3):
            raise abi.RfError(f"color_im shape {tuple(rgb.shape)} does not match depth {(im_h, im_w)}")
        if self._frame is None:
            self._frame = {}
        packed = self._frame.get("packed")
        if packed is None or packed.numel() != im_h * im_w:
            packed = torch.empty(im_h * im_w, dtype=torch.float32, device=self.device)
            self._frame["packed"] = packed
        L = abi.lib()
        st = abi.stream_ptr()
        abi.check(L.rf_pack_bgr(abi.dptr(rgb), abi.dptr(packed), C.c_int(im_h * im_w), st), "rf_pack_bgr")
        self.integrate_packed(depth, packed, cam_intr, cam_pose, old_bnd, obs_weight, reintegrate_flag)
        self._release()

    def integrate_packed(self, depth, packed, cam_intr, cam_pose, old_bnd=None, obs_weight=1., reintegrate_flag=0.0):
        """Same as integrate() with the frame already resident: depth [H,W], packed BGR [H*W] CUDA fp32."""
        im_h, im_w = depth.shape
        dx, dy, dz = (int(d) for d in self.vol_dim)
        _o, o_p = abi.farr(self.vol_origin, 3)
        _k, k_p = abi.farr(np.asarray(cam_intr, dtype=np.float64), 9)
        _c, c_p = abi.farr(np.asarray(cam_pose.detach().cpu().numpy() if isinstance(cam_pose, torch.Tensor) else cam_pose,
                                      dtype=np.float64), 16)
        reint = 1 if float(reintegrate_flag) == 1.0 else 0
        if old_bnd is not None:
            _b, b_p = abi.farr(np.asarray(old_bnd, dtype=np.float64), 6)
        else:
            _b, b_p = None, C.POINTER(C.c_float)()
        L = abi.lib()
        rc = L.rf_tsdf_integrate_local(
            abi.dptr(self.tsdf_vol_gpu), abi.dptr(self.weight_vol_gpu), abi.dptr(self.color_vol_gpu),
            C.c_int(dx), C.c_int(dy), C.c_int(dz), o_p, C.c_float(self.voxel_size), k_p, c_p,
            abi.dptr(depth), abi.dptr(packed), C.c_int(im_h), C.c_int(im_w),
            C.c_float(self.trunc_margin), C.c_float(obs_weight),
            C.c_int(1 if float(self.weight_clamp) == 1.0 else 0), C.c_int(reint), b_p,
            C.c_int(self.x_slab[0]), C.c_int(self.x_slab[1]), C.c_int(1),
            abi.dptr(abi.pixel_lambda(_k, im_h, im_w, depth.device)),
            abi.dptr(abi.depth_max(depth) if self.tsdf_vol_gpu.numel() >= abi.FAR_PLANE_MIN_VOXELS else None), abi.stream_ptr())
        abi.check(rc, "rf_tsdf_integrate_local")

    def count_touched(self, depth, cam_intr, cam_pose, old_bnd=None, reintegrate_flag=0.0):
        """(n_touched, n_band) for this frame — the numerator of voxel-updates/s."""
        im_h, im_w = depth.shape
        dx, dy, dz = (int(d) for d in self.vol_dim)
        _o, o_p = abi.farr(self.vol_origin, 3)
        _k, k_p = abi.farr(np.asarray(cam_intr, dtype=np.float64), 9)
        _c, c_p = abi.farr(np.asarray(cam_pose, dtype=np.float64), 16)
        if old_bnd is not None:
            _b, b_p = abi.farr(np.asarray(old_bnd, dtype=np.float64), 6)
        else:
            _b, b_p = None, C.POINTER(C.c_float)()
        counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        rc = abi.lib().rf_tsdf_count_local(
            C.c_int(dx), C.c_int(dy), C.c_int(dz), o_p, C.c_float(self.voxel_size), k_p, c_p,
            abi.dptr(depth), C.c_int(im_h), C.c_int(im_w), C.c_float(self.trunc_margin),
            C.c_int(1 if float(reintegrate_flag) == 1.0 else 0), b_p,
            C.c_int(self.x_slab[0]), C.c_int(self.x_slab[1]), C.c_void_p(counts.data_ptr()), abi.stream_ptr())
        abi.check(rc, "rf_tsdf_count_local")
        c = counts.cpu()
        return int(c[0]), int(c[1])

    def clean_volume(self):
        """Reset tsdf=1, weight=0, colour=0 (model/Volume.py:655-673)."""
        rc = abi.lib().rf_tsdf_clear_local(abi.dptr(self.tsdf_vol_gpu), abi.dptr(self.weight_vol_gpu),
                                           abi.dptr(self.color_vol_gpu), C.c_int64(self.tsdf_vol_gpu.numel()),
                                           abi.stream_ptr())
        abi.check(rc, "rf_tsdf_clear_local")
