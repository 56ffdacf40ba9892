"""Volume-reading half of the random-optimisation tracker: host mirror of the GPU-backed methods of
``RO_tracker`` — ``init_depth_vertex`` (model/ROtracker.py:436-456), ``init_normal`` (:458-470), ``evaluate_tsdf``
(:536-604), ``cal_transform`` (:606-714) — over ``rf_track_vertex_normal`` / ``rf_track_fitness`` / ``rf_track_cal_transform``,
and of the search loop itself, ``random_optimization`` (:716-836, with ``get_PST`` :467-493 and ``update_PST`` :495-531), as ONE
call: ``rf_track_random_optimization`` keeps the search state on the device and enqueues the 20 iterations back to back; the host
reads the pose once per frame instead of two fitness arrays per iteration.  The step-by-step methods stay for callers that drive
the loop themselves: they set ``current_global_R`` / ``current_global_T`` / ``transform_candidate`` / ``search_size`` on this
object exactly as on the reference's tracker."""
from __future__ import annotations

import ctypes as C
import random

import numpy as np
import torch

from . import abi


class ROSearch:
    def __init__(self, MV, im_h, im_w, cut_dist, truncation, sample_range, device=None):
        """MV: the ``moving_volume`` whose ``tsdf_vol_gpu`` the search reads (model/ROtracker.py:132)."""
        self.MV = MV
        self.im_h, self.im_w = int(im_h), int(im_w)
        self.cut_dist, self.truncation, self.sample_range = float(cut_dist), float(truncation), float(sample_range)
        self.device = torch.device(device) if device is not None else MV.tsdf_vol_gpu.device
        n = self.im_h * self.im_w
        self.depth_map_gpu = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.depth_vertex_gpu = torch.zeros(n * 4, dtype=torch.float32, device=self.device)
        self.normal_vertex_gpu = torch.zeros(n * 3, dtype=torch.float32, device=self.device)
        self._row_sample = torch.zeros(self.im_h, dtype=torch.float32, device=self.device)
        self._cam_intr = None
        self.current_global_R = np.eye(3, dtype=np.float32)
        self.current_global_T = np.zeros(3, dtype=np.float32)
        self.transform_candidate = np.zeros((0, 6), dtype=np.float32)
        self.search_size = np.zeros(6, dtype=np.float32)
        self._cand_key, self._cand_dev = None, None
        self._last = None                                  # (value, count, n) of the last evaluate_tsdf, on the device
        self.count_search = 0                              # cfg["RO"]["count_search"] (model/ROtracker.py:58)

    # ---- model/ROtracker.py:436-456 --------------------------------------------------------------------------
    def init_depth_vertex(self, depth_im, cam_intr, seed_num=None):
        if isinstance(depth_im, torch.Tensor):
            self.depth_map_gpu.copy_(depth_im.reshape(-1).to(self.device, torch.float32))
        else:
            self.depth_map_gpu.copy_(torch.from_numpy(np.ascontiguousarray(depth_im.reshape(-1).astype(np.float32))))
        if seed_num is None:
            seed_num = random.randint(1, 1000000)                           # :440
        self._cam_intr = np.ascontiguousarray(np.asarray(cam_intr, dtype=np.float32).reshape(-1))
        rc = abi.lib().rf_track_vertex_normal(abi.dptr(self.depth_map_gpu), self.im_h, self.im_w, abi.fptr(self._cam_intr),
                                              C.c_float(self.cut_dist), C.c_float(self.truncation), C.c_int(int(seed_num)),
                                              C.c_float(self.sample_range), abi.dptr(self._row_sample), abi.dptr(self.depth_vertex_gpu),
                                              abi.dptr(self.normal_vertex_gpu), abi.stream_ptr())
        abi.check(rc, "rf_track_vertex_normal")

    # ---- model/ROtracker.py:458-470: the normal map is produced together with the vertex map ----------------------
    def init_normal(self):
        return None

    # ---- model/ROtracker.py:536-604 --------------------------------------------------------------------------
    def evaluate_tsdf(self, cur_id, level, node_size, cam_intr, level_index, as_numpy=True):
        cand = np.ascontiguousarray(np.asarray(self.transform_candidate, dtype=np.float32).reshape(-1, 6))
        n = int(node_size) // 1024 * 1024                 # the reference launches int(node_size / 1024) blocks of 1024 candidates
        key = (cand.shape[0], hash(cand.tobytes()))          # 245 KB at the largest PST: cheaper than the upload it avoids
        if self._cand_key != key:
            self._cand_dev = torch.from_numpy(cand).to(self.device); self._cand_key = key
        total = cand.shape[0]
        value = torch.zeros(total, dtype=torch.float32, device=self.device)
        count = torch.zeros(total, dtype=torch.float32, device=self.device)
        if n > 0:
            L = abi.lib()
            ns = int(L.rf_track_fitness_scratch_floats(n, self.im_h, self.im_w, int(level)))
            scratch = torch.empty(max(ns, 2), dtype=torch.float32, device=self.device)
            K = np.ascontiguousarray(np.asarray(cam_intr, dtype=np.float32).reshape(-1))
            R = np.ascontiguousarray(np.asarray(self.current_global_R, dtype=np.float32).reshape(-1))
            T = np.ascontiguousarray(np.asarray(self.current_global_T, dtype=np.float32).reshape(-1))
            ss = np.ascontiguousarray(np.asarray(self.search_size, dtype=np.float32).reshape(-1))
            dims = (C.c_int * 3)(int(self.MV.vol_dim[0]), int(self.MV.vol_dim[1]), int(self.MV.vol_dim[2]))
            org = np.ascontiguousarray(np.asarray(self.MV.vol_origin, dtype=np.float32).reshape(-1))
            rc = L.rf_track_fitness(abi.dptr(self.MV.tsdf_vol_gpu), dims, abi.fptr(org), C.c_float(float(self.MV.voxel_size)),
                                    abi.dptr(self.depth_vertex_gpu), abi.dptr(self.normal_vertex_gpu), self.im_h, self.im_w, abi.fptr(K),
                                    abi.fptr(R), abi.fptr(T), abi.dptr(self._cand_dev), n, abi.fptr(ss), int(level), int(level_index),
                                    abi.dptr(value), abi.dptr(count), abi.dptr(scratch), abi.stream_ptr())
            abi.check(rc, "rf_track_fitness")
        self._last = (value, count, n)
        if not as_numpy:
            return value / (count + 1e-6), value, count
        v = value.cpu().numpy(); c = count.cpu().numpy()
        return v / (c + 1e-6), v, c                       # :601-604

    # ---- model/ROtracker.py:606-714 --------------------------------------------------------------------------
    def cal_transform(self, search_value=None):
        """(success, min_tsdf, mean_transform[7]) from the fitness arrays the last ``evaluate_tsdf`` left on the device:
        the reference's Python loop over all candidates as one kernel and a 9-float read-back.  ``search_value`` is
        accepted for signature parity and ignored (it is the first array ``evaluate_tsdf`` returned)."""
        if self._last is None:
            raise abi.RfError("cal_transform: call evaluate_tsdf first")
        value, count, n = self._last
        ss = np.ascontiguousarray(np.asarray(self.search_size, dtype=np.float32).reshape(-1))
        out = torch.empty(9, dtype=torch.float32, device=self.device)
        rc = abi.lib().rf_track_cal_transform(abi.dptr(value), abi.dptr(count), abi.dptr(self._cand_dev), int(max(n, 1)), abi.fptr(ss),
                                              int(self.count_search), abi.dptr(out), abi.stream_ptr())
        abi.check(rc, "rf_track_cal_transform")
        o = out.cpu().numpy()
        return bool(o[0] > 0.5), float(o[1]), o[2:9].copy()

    # ---- model/ROtracker.py:716-836 (+ :467-493, :495-531): the whole search loop on the device -----------------------------
    def configure_search(self, ro_cfg, all_pst, tiff_index=None, depth_level=None):
        """``ro_cfg``: the reference's ``cfg["RO"]`` entries (init_size, scaling_coefficient, particle_iter_lens, PST_size,
        fix_level_index, count_search, iterative_scale); ``all_pst``: ``ALL_PST`` as ``readpst`` builds it — one array
        [n_tables, PST_size[c], 6] per class c; ``tiff_index`` / ``depth_level``: the tracker's two 20-entry lists (:118-123)."""
        self.init_size = float(ro_cfg["init_size"]); self.scaling_coefficient = float(ro_cfg["scaling_coefficient"])
        self.particle_iter_lens = int(ro_cfg["particle_iter_lens"]); self.PST_size = [int(v) for v in ro_cfg["PST_size"]]
        self.fix_level_index = bool(ro_cfg["fix_level_index"]); self.count_search = int(ro_cfg["count_search"])
        self.iterative_scale = bool(ro_cfg["iterative_scale"])
        self.tiff_index = list(tiff_index) if tiff_index is not None else [
            0, 1 + 20, 2 + 40, 3, 4 + 20, 5 + 40, 6 + 0, 7 + 20, 8 + 40, 9 + 0, 10 + 20, 11 + 40, 12 + 0, 13 + 20, 14 + 40,
            15 + 0, 16 + 20, 17 + 40, 18 + 0, 19 + 20]
        self.depth_level = list(depth_level) if depth_level is not None else [32, 16, 8] * 6 + [32, 16]
        self.ALL_PST = [np.ascontiguousarray(np.asarray(a, dtype=np.float32)) for a in all_pst]
        offs, base = [], 0
        for a in self.ALL_PST:
            offs.append(base); base += a.size
        self._pst_dev = torch.from_numpy(np.concatenate([a.reshape(-1) for a in self.ALL_PST])).to(self.device)
        self._pst_off, self._pst_n = [], []
        for k in range(20):
            ti = self.tiff_index[k]
            cls = ti // 20; idx = (ti - cls * 20) // 3                                   # get_PST :488-493
            n = self.PST_size[k % 3] // 1024 * 1024                                      # evaluate_tsdf launches int(node_size / 1024) blocks
            if n != self.ALL_PST[cls].shape[1] or n <= 0:
                raise abi.RfError("configure_search: PST_size must be a positive multiple of 1024 and equal the table length "
                                  "(cal_transform walks the whole table, evaluate_tsdf only full blocks)")
            self._pst_off.append(offs[cls] + idx * self.ALL_PST[cls].shape[1] * 6); self._pst_n.append(n)
        self.initialize_search_size = np.zeros(6)
        self.previous_frame_success = False
        self.previous_search_size = np.zeros(6, dtype=np.float32)
        self.success_mask = 0

    def init_searchsize(self):                                                            # :408-419
        self.search_size = np.zeros(6, dtype=np.float32)
        self.previous_search_size = np.zeros(6, dtype=np.float32)
        self.search_size[...] = self.init_size

    def random_optimization(self, cur_id, cam_pose, rgb_im, depth_im, cam_intr, beta=0.9, inherit=False, seed_num=None):
        """Same arguments and return value (the optimised 4x4 float32 pose) as the reference method; also leaves
        ``current_global_R/T``, ``search_size``, ``previous_search_size``, ``initialize_search_size``, ``previous_frame_success``
        as the reference loop would."""
        if not hasattr(self, "_pst_dev"):
            raise abi.RfError("random_optimization: call configure_search first")
        cam_pose = np.asarray(cam_pose)
        self.current_global_R = cam_pose[:3, :3].copy(); self.current_global_T = cam_pose[:3, 3].copy()
        if inherit is True and self.previous_frame_success:                               # :741-744
            self.search_size = self.initialize_search_size
        else:
            self.init_searchsize()
        self.init_depth_vertex(depth_im, cam_intr, seed_num=seed_num)
        self.init_normal()
        L = abi.lib()
        i32 = lambda v: (C.c_int * 20)(*[int(x) for x in v])
        off, n, lev = i32(self._pst_off), i32(self._pst_n), i32(self.depth_level)
        f32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))
        st_host = torch.zeros(64, dtype=torch.float32).pin_memory() if not hasattr(self, "_st_host") else self._st_host
        self._st_host = st_host
        abi.check(L.rf_track_state_init(C.c_void_p(st_host.data_ptr()), abi.fptr(f32(self.current_global_R)), abi.fptr(f32(self.current_global_T)),
                                        abi.fptr(f32(self.search_size)), abi.fptr(f32(self.previous_search_size)), off, n, lev,
                                        self.im_h, self.im_w), "rf_track_state_init")
        if not hasattr(self, "_st_dev"):
            n_max = max(self._pst_n)
            self._st_dev = torch.zeros(64, dtype=torch.float32, device=self.device)
            self._sv = torch.zeros(n_max, dtype=torch.float32, device=self.device); self._sc = torch.zeros_like(self._sv)
            ns = int(L.rf_track_random_optimization_scratch_floats(n, lev, self.im_h, self.im_w))
            self._ro_scratch = torch.empty(max(ns, 2), dtype=torch.float32, device=self.device)
        self._st_dev.copy_(st_host, non_blocking=True)
        dims = (C.c_int * 3)(int(self.MV.vol_dim[0]), int(self.MV.vol_dim[1]), int(self.MV.vol_dim[2]))
        rc = L.rf_track_random_optimization(abi.dptr(self.MV.tsdf_vol_gpu), dims, abi.fptr(f32(self.MV.vol_origin)), C.c_float(float(self.MV.voxel_size)),
                                            abi.dptr(self.depth_vertex_gpu), abi.dptr(self.normal_vertex_gpu), self.im_h, self.im_w, abi.fptr(f32(cam_intr)),
                                            abi.dptr(self._pst_dev), off, n, lev, int(self.particle_iter_lens), int(self.count_search),
                                            C.c_float(self.scaling_coefficient), int(self.fix_level_index), int(self.iterative_scale), C.c_float(float(beta)),
                                            abi.dptr(self._st_dev), abi.dptr(self._sv), abi.dptr(self._sc), abi.dptr(self._ro_scratch), abi.stream_ptr())
        abi.check(rc, "rf_track_random_optimization")
        st = self._st_dev.cpu().numpy()                                                    # the frame's one read-back
        self.current_global_R = st[0:9].reshape(3, 3).copy(); self.current_global_T = st[9:12].copy()
        self.search_size[...] = st[12:18]; self.previous_search_size = st[18:24].copy()
        words = st.view(np.int32)
        self.previous_frame_success = bool(words[44])
        if self.previous_frame_success:                                                    # :824-827: the SAME array object from now on
            self.initialize_search_size = self.search_size
        self.success_mask = int(words[49])
        cam_pose_iter = np.eye(4, dtype=np.float32)                                       # :832-836
        cam_pose_iter[:3, :3] = self.current_global_R; cam_pose_iter[:3, 3] = self.current_global_T
        return cam_pose_iter
