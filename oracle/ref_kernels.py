"""TEST INFRASTRUCTURE ONLY.  Launch the literal reference TSDF kernels (cubins built by oracle/build_ref.py
from the strings in /root/reference) on torch CUDA tensors through the CUDA driver API (cuda-python),
reproducing the host wrappers of the reference launch for launch:

  * moving_volume.integrate       model/Volume.py:713-757   (grid/block: model/Volume.py:110-123)
  * Mapper.integrate_kf           mp_slam/mapper.py:823-872 (grid/block: mp_slam/mapper.py:239-251)
  * Mapper.init_mapvolume         mp_slam/mapper.py:267-282
  * RO_tracker.init_depth_vertex / init_normal / evaluate_tsdf   model/ROtracker.py:436-470, :536-604

Every scalar travels as float32 inside small device arrays exactly as PyCUDA's ``cuda.In`` would ship it.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return (torch.cuda.is_available()
            and os.path.exists(os.path.join(REF_DIR, "ref_local_volume.cubin"))
            and os.path.exists(os.path.join(REF_DIR, "ref_global_volume.cubin"))
            and os.path.exists(os.path.join(REF_DIR, "ref_tracker.cubin")))


def require() -> None:
    """On a CUDA box the literal reference kernels are a mandatory part of the parity suite: fail, never skip."""
    missing = [c for c in ("ref_local_volume.cubin", "ref_global_volume.cubin", "ref_tracker.cubin")
               if not os.path.exists(os.path.join(REF_DIR, c))]
    assert not missing, (f"oracle/_ref is missing {missing}: run `python -c 'import __graft_entry__ as g; g.build()'` where "
                         "/root/reference exists (oracle/build_ref.py); the built cubins travel to the GPU box")
    assert torch.cuda.is_available(), "the reference kernels need a CUDA device"


def _check(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)


class _Module:
    def __init__(self, cubin: str):
        from cuda.bindings import driver
        self.drv = driver
        torch.cuda.init()
        torch.zeros(1, device="cuda")          # make sure the primary context is current
        data = open(os.path.join(REF_DIR, cubin), "rb").read()
        self.mod = _check(driver.cuModuleLoadData(data))
        self.fn = {}

    def get(self, name: str):
        if name not in self.fn:
            self.fn[name] = _check(self.drv.cuModuleGetFunction(self.mod, name.encode()))
        return self.fn[name]

    def launch(self, name, grid, block, ptrs):
        """ptrs: list of device pointers (ints)."""
        args = [ctypes.c_void_p(int(p)) for p in ptrs]
        arr = (ctypes.c_void_p * len(args))(*[ctypes.addressof(a) for a in args])
        stream = torch.cuda.current_stream().cuda_stream
        _check(self.drv.cuLaunchKernel(self.get(name), grid[0], grid[1], grid[2], block[0], block[1], block[2],
                                       0, stream, ctypes.addressof(arr), 0))


_mods = {}


def _mod(name):
    if name not in _mods:
        _mods[name] = _Module(name)
    return _mods[name]


def _launch_geometry(n_vox: int, max_threads=1024, max_grid=(2147483647, 65535, 65535)):
    """model/Volume.py:110-123 == mp_slam/mapper.py:239-251."""
    n_blocks = int(np.ceil(float(n_vox) / float(max_threads)))
    gx = min(max_grid[0], int(np.floor(np.cbrt(n_blocks))))
    gy = min(max_grid[1], int(np.floor(np.sqrt(n_blocks / gx))))
    gz = min(max_grid[2], int(np.ceil(float(n_blocks) / float(gx * gy))))
    n_loops = int(np.ceil(float(n_vox) / float(gx * gy * gz * max_threads)))
    return (gx, gy, gz), n_loops


def _dev(a) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(-1))).cuda()


def ref_integrate_local(tsdf, weight, color, vol_dim, vol_origin, voxel_size, K, c2w, depth, packed_bgr,
                        trunc_margin, obs_weight=1.0, reintegrate_flag=0.0, weight_clamp=1.0, old_bnd=None):
    """In place on the three CUDA fp32 tensors.  depth/packed_bgr: CUDA fp32 [H,W]."""
    m = _mod("ref_local_volume.cubin")
    H, W = depth.shape
    grid, n_loops = _launch_geometry(int(np.prod(vol_dim)))
    keep = [_dev(vol_dim), _dev(vol_origin), _dev(K), _dev(c2w),
            _dev(old_bnd if old_bnd is not None else np.zeros(6))]
    for loop in range(n_loops):
        other = _dev([loop, voxel_size, H, W, trunc_margin, obs_weight, reintegrate_flag, weight_clamp])
        keep.append(other)
        m.launch("integrate", grid, (1024, 1, 1),
                 [tsdf.data_ptr(), weight.data_ptr(), color.data_ptr(), keep[0].data_ptr(), keep[1].data_ptr(),
                  keep[2].data_ptr(), keep[3].data_ptr(), other.data_ptr(), keep[4].data_ptr(),
                  packed_bgr.data_ptr(), depth.data_ptr()])
    torch.cuda.synchronize()


def ref_recenter(new, old, vol_dim, vol_origin, old_vol_dim, old_origin, voxel_size):
    """model/Volume.py:796-858: launch the reference `swap_rot_trans` (new <- old).  new / old: lists of three CUDA fp32
    tensors (tsdf, weight, color); the reference's `old` is its backup copy of the volume (copy_volume, :883-908)."""
    m = _mod("ref_local_volume.cubin")
    grid, n_loops = _launch_geometry(int(np.prod(vol_dim)))
    keep = [_dev(vol_dim), _dev(vol_origin), _dev(old_origin), _dev(old_vol_dim)]
    for loop in range(n_loops):
        other = _dev([loop, voxel_size])
        keep.append(other)
        m.launch("swap_rot_trans", grid, (1024, 1, 1),
                 [new[0].data_ptr(), old[0].data_ptr(), new[1].data_ptr(), old[1].data_ptr(), new[2].data_ptr(), old[2].data_ptr(),
                  keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(), keep[3].data_ptr(), other.data_ptr()])
    torch.cuda.synchronize()


def ref_integrate_global(trgb, wgt, R, box, K, c2w, depth, rgb, trunc_margin, obs_weight=1.0):
    """In place on GBV params [R^3*4] and GBW params [R^3] (CUDA fp32).  rgb: CUDA fp32 [H,W,3] in [0,1]."""
    m = _mod("ref_global_volume.cubin")
    H, W = depth.shape
    grid, n_loops = _launch_geometry(R ** 3)
    voxel_size = 1.0 / R
    keep = [_dev([R, R, R]), _dev(K), _dev(c2w)]
    for loop in range(n_loops):
        other = _dev([loop, voxel_size, H, W, trunc_margin, obs_weight,
                      box[0], box[1], box[2], box[3], box[4], box[5]])
        keep.append(other)
        m.launch("integrate", grid, (1024, 1, 1),
                 [trgb.data_ptr(), wgt.data_ptr(), keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(),
                  other.data_ptr(), rgb.data_ptr(), depth.data_ptr()])
    torch.cuda.synchronize()


def ref_clear_global(trgb, R):
    m = _mod("ref_global_volume.cubin")
    grid, n_loops = _launch_geometry(R ** 3)
    keep = [_dev([R, R, R])]
    for loop in range(n_loops):
        other = _dev([loop])
        keep.append(other)
        m.launch("clean_tsdf", grid, (1024, 1, 1), [trgb.data_ptr(), keep[0].data_ptr(), other.data_ptr()])
    torch.cuda.synchronize()


def ref_track_vertex_normal(depth, K, cut_dist, trunc, seed_num, sample_range):
    """model/ROtracker.py:436-470: launch the reference `compute_vertex` and `compute_normal`.  depth: CUDA fp32 [H,W].
    Returns (depth_vertex [H*W*4], normal [H*W*3]) CUDA fp32 (normal zero-initialised, as the reference's allocation)."""
    m = _mod("ref_tracker.cubin")
    H, W = depth.shape
    vertex = torch.zeros(H * W * 4, device="cuda"); normal = torch.zeros(H * W * 3, device="cuda")
    block = (int((H + 32 - 1) / 32), int((W + 32 - 1) / 32), 1)
    k = _dev(K)
    o1 = _dev([H, W, cut_dist, trunc, seed_num, sample_range])
    m.launch("compute_vertex", (32, 32, 1), block, [depth.data_ptr(), vertex.data_ptr(), k.data_ptr(), o1.data_ptr()])
    o2 = _dev([H, W])
    m.launch("compute_normal", (32, 32, 1), block, [vertex.data_ptr(), normal.data_ptr(), o2.data_ptr()])
    torch.cuda.synchronize()
    return vertex, normal


def ref_track_fitness(tsdf, vol_dim, vol_origin, voxel_size, vertex, normal, H, W, K, R, T, cand, search_size, level, level_index):
    """model/ROtracker.py:536-604: launch the reference `compute_tsdf_value`; cand: numpy [n,6], n a multiple of 1024.
    Returns (search_value, search_count) CUDA fp32 [n]."""
    m = _mod("ref_tracker.cubin")
    n = cand.shape[0]
    value = torch.zeros(n, device="cuda"); count = torch.zeros(n, device="cuda")
    dummy = torch.zeros(1, device="cuda")                     # weight_vol / depth_map are never read by the kernel
    keep = [_dev(search_size), _dev(R), _dev(T), _dev(cand), _dev(K),
            _dev([vol_dim[0], vol_dim[1], vol_dim[2], vol_origin[0], vol_origin[1], vol_origin[2], voxel_size, n, level,
                  int(H / level), int(W / level), level_index, H, W, 0, 0, 0])]
    m.launch("compute_tsdf_value", (int(n / (32 * 32)), int(H / level), int(W / level)), (32 * 32, 1, 1),
             [tsdf.data_ptr(), dummy.data_ptr(), dummy.data_ptr(), vertex.data_ptr(), value.data_ptr(), count.data_ptr(),
              keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(), keep[3].data_ptr(), keep[5].data_ptr(), keep[4].data_ptr(),
              normal.data_ptr()])
    torch.cuda.synchronize()
    return value, count
