"""ctypes binding of the C-ABI in include/rf_abi.h (librf_b200.so).

There is no CPU fallback: if the shared object is missing this module raises at first use.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "librf_b200.so")
RF_MAX_LEVELS = 16


class RfError(RuntimeError):
    pass


class GridDesc(C.Structure):
    _fields_ = [
        ("n_levels", C.c_int32), ("n_features", C.c_int32), ("is_hash", C.c_int32), ("_pad", C.c_int32),
        ("scale", C.c_float * RF_MAX_LEVELS),
        ("resolution", C.c_uint32 * RF_MAX_LEVELS),
        ("size", C.c_uint32 * RF_MAX_LEVELS),
        ("offset", C.c_uint32 * (RF_MAX_LEVELS + 1)),
    ]

    @property
    def n_params(self) -> int:
        return int(self.offset[self.n_levels]) * int(self.n_features)

    @property
    def n_output_dims(self) -> int:
        return int(self.n_levels) * int(self.n_features)


class RayCfg(C.Structure):
    _fields_ = [
        ("range_d", C.c_float), ("n_range_d", C.c_int32), ("n_samples_d", C.c_int32),
        ("near_", C.c_float), ("far_", C.c_float), ("perturb", C.c_int32),
        ("c_trunc", C.c_float), ("trunc", C.c_float), ("clamp_mode", C.c_int32), ("clamp_thr", C.c_float),
        ("sc_factor", C.c_float), ("depth_trunc", C.c_float), ("rgb_missing", C.c_float),
        ("hidden", C.c_int32), ("n_bins", C.c_int32), ("geo_feat", C.c_int32),
        ("mlp_precision", C.c_int32), ("_pad", C.c_int32),
        ("n_rays_total", C.c_int64),
        ("bbox", C.c_double * 6),
    ]


class RayParams(C.Structure):
    _fields_ = [("hash_params", C.c_void_p), ("gbv_params", C.c_void_p), ("w_sdf0", C.c_void_p),
                ("w_sdf1", C.c_void_p), ("w_col0", C.c_void_p), ("w_col1", C.c_void_p)]


class RayGrads(C.Structure):
    _fields_ = [("g_hash", C.c_void_p), ("g_w_sdf0", C.c_void_p), ("g_w_sdf1", C.c_void_p),
                ("g_w_col0", C.c_void_p), ("g_w_col1", C.c_void_p), ("g_rays_o", C.c_void_p),
                ("g_rays_d", C.c_void_p)]


_lib = None


def lib() -> C.CDLL:
    """Load librf_b200.so (built by remixfusion_b200._build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RfError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the mapping hot path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.rf_last_error.restype = C.c_char_p
        _lib.rf_version.restype = C.c_int
        _lib.rf_ray_workspace_floats.restype = C.c_int64
        _lib.rf_ray_scratch_floats.restype = C.c_int64
        _lib.rf_point_workspace_floats.restype = C.c_int64
        _lib.rf_track_fitness_scratch_floats.restype = C.c_int64
        _lib.rf_track_random_optimization_scratch_floats.restype = C.c_int64
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().rf_last_error().decode(errors="replace")
        raise RfError(f"{what} failed with code {rc}: {msg}")


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t) -> C.c_void_p:
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise RfError("expected a CUDA tensor (the hot path has no CPU implementation)")
    if not t.is_contiguous():
        raise RfError("expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


_pixel_lambda_cache = {}


def pixel_lambda(K, H, W, device):
    """Cached per-camera 1/lambda image for the TSDF kernels (rf_tsdf_pixel_lambda): depends on the intrinsics only."""
    Kf = np.ascontiguousarray(np.asarray(K, dtype=np.float32).reshape(-1))
    key = (Kf.tobytes(), int(H), int(W), str(device))
    t = _pixel_lambda_cache.get(key)
    if t is None:
        t = torch.empty(int(H) * int(W), dtype=torch.float32, device=device)
        check(lib().rf_tsdf_pixel_lambda(Kf.ctypes.data_as(C.POINTER(C.c_float)), C.c_int(int(H)), C.c_int(int(W)), dptr(t),
                                         stream_ptr()), "rf_tsdf_pixel_lambda")
        if len(_pixel_lambda_cache) > 8:
            _pixel_lambda_cache.clear()
        _pixel_lambda_cache[key] = t
    return t


# The far plane of the TSDF row clip needs the frame's largest depth: a pre-pass of two tiny launches (~6 us).  It pays off
# when sweeping the part of the volume behind the farthest surface costs more than that — large volumes (BS3D-scale GBV, the
# shipped 1 cm Replica volume); below this many owned voxels the integrate calls run without it.
FAR_PLANE_MIN_VOXELS = 1 << 26


def depth_max(depth):
    """Largest depth of a frame as a device scalar (rf_tsdf_depth_max): the far plane of the TSDF row clip."""
    out = torch.empty(1, dtype=torch.float32, device=depth.device)
    check(lib().rf_tsdf_depth_max(dptr(depth), C.c_int64(depth.numel()), dptr(out), stream_ptr()), "rf_tsdf_depth_max")
    return out


def farr(values, n=None):
    a = np.ascontiguousarray(np.asarray(values, dtype=np.float32).reshape(-1))
    if n is not None and a.size != n:
        raise RfError(f"expected {n} floats, got {a.size}")
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def fptr(a: np.ndarray):
    """Host float32 array (contiguous; kept alive by the caller) -> const float*."""
    if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
        raise RfError("expected a contiguous float32 numpy array")
    return a.ctypes.data_as(C.POINTER(C.c_float))


def require_f32_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32:
        raise RfError(f"{name}: expected a float32 CUDA tensor")
    return t.contiguous()
