"""Volume-reading half of the random-optimisation tracker: host mirror of the three GPU-backed methods of
``RO_tracker`` — ``init_depth_vertex`` (model/ROtracker.py:436-456), ``init_normal`` (:458-470) and ``evaluate_tsdf``
(:536-604) — over ``rf_track_vertex_normal`` / ``rf_track_fitness``.  The search policy around them (PST tables,
``cal_transform``, ``update_PST``, ``random_optimization``; :606-831) is host logic of the caller and stays there: it sets
``current_global_R`` / ``current_global_T`` / ``transform_candidate`` / ``search_size`` on this object exactly as it does
on the reference's tracker, and reads back the three arrays ``evaluate_tsdf`` returns."""
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
