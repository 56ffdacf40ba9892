"""TEST INFRASTRUCTURE ONLY — ctypes front-end of oracle/tsdf_oracle.c (the CPU restatement of the reference's
TSDF kernels, model/Volume.py:196-336 and mp_slam/mapper.py:37-158).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None


def build() -> str:
    src = os.path.join(HERE, "tsdf_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def integrate_local(tsdf, weight, color, vol_dim, vol_origin, voxel_size, K, c2w, depth, packed_bgr, trunc_margin,
                    obs_weight=1.0, weight_clamp=1, reintegrate=0, old_bnd=None, x0=0, x1=None, touched=None,
                    threads=1):
    """In place on the float32 numpy arrays tsdf/weight/color [dx*dy*dz].  Returns (n_touched, n_band)."""
    dx, dy, dz = (int(v) for v in vol_dim)
    x1 = dx if x1 is None else x1
    H, W = depth.shape
    for a in (tsdf, weight, color):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    _o, o_p = _f(vol_origin)
    _k, k_p = _f(np.asarray(K).reshape(-1))
    _c, c_p = _f(np.asarray(c2w).reshape(-1))
    _d, d_p = _f(depth)
    _p, p_p = _f(packed_bgr)
    _b, b_p = _f(np.zeros(6) if old_bnd is None else np.asarray(old_bnd).reshape(-1))
    t_p = touched.ctypes.data_as(C.POINTER(C.c_uint8)) if touched is not None else None
    fn = lib().oracle_integrate_local

    def run(lo, hi):
        counts = (C.c_uint64 * 2)(0, 0)
        fn(tsdf.ctypes.data_as(C.POINTER(C.c_float)), weight.ctypes.data_as(C.POINTER(C.c_float)),
           color.ctypes.data_as(C.POINTER(C.c_float)), dx, dy, dz, o_p, C.c_float(voxel_size), k_p, c_p, d_p, p_p,
           H, W, C.c_float(trunc_margin), C.c_float(obs_weight), int(weight_clamp), int(reintegrate), b_p,
           int(lo), int(hi), counts, t_p)
        return counts[0], counts[1]

    return _sharded(run, x0, x1, threads)


def _sharded(run, lo, hi, threads):
    if threads <= 1 or hi - lo < 2:
        return run(lo, hi)
    cuts = np.linspace(lo, hi, min(threads * 4, hi - lo) + 1).astype(int)
    with ThreadPoolExecutor(max_workers=threads) as ex:      # ctypes releases the GIL during the call
        res = list(ex.map(lambda ab: run(ab[0], ab[1]), zip(cuts[:-1], cuts[1:])))
    return sum(r[0] for r in res), sum(r[1] for r in res)


def pack_bgr(rgb_0_255):
    rgb = np.ascontiguousarray(rgb_0_255, dtype=np.float32)
    out = np.empty(rgb.shape[:-1], dtype=np.float32)
    lib().oracle_pack_bgr(rgb.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data_as(C.POINTER(C.c_float)),
                          C.c_int64(out.size))
    return out


def integrate_global(trgb, wgt, R, box, K, c2w, depth, rgb, trunc_margin, obs_weight=1.0, z0=0, z1=None,
                     touched=None, threads=1):
    """In place on trgb [R^3*4], wgt [R^3] float32 numpy.  Returns n_touched."""
    z1 = R if z1 is None else z1
    H, W = depth.shape
    assert trgb.dtype == np.float32 and wgt.dtype == np.float32
    _b, b_p = _f(np.asarray(box).reshape(-1))
    _k, k_p = _f(np.asarray(K).reshape(-1))
    _c, c_p = _f(np.asarray(c2w).reshape(-1))
    _d, d_p = _f(depth)
    _r, r_p = _f(rgb)
    t_p = touched.ctypes.data_as(C.POINTER(C.c_uint8)) if touched is not None else None
    fn = lib().oracle_integrate_global

    def run(lo, hi):
        counts = (C.c_uint64 * 2)(0, 0)
        fn(trgb.ctypes.data_as(C.POINTER(C.c_float)), wgt.ctypes.data_as(C.POINTER(C.c_float)), int(R), b_p, k_p, c_p,
           d_p, r_p, H, W, C.c_float(trunc_margin), C.c_float(obs_weight), int(lo), int(hi), counts, t_p)
        return counts[0], counts[1]

    return _sharded(run, z0, z1, threads)[0]


def clear_global(trgb):
    lib().oracle_clear_global(trgb.ctypes.data_as(C.POINTER(C.c_float)), C.c_int64(trgb.size // 4))


def clear_local(tsdf, weight, color):
    lib().oracle_clear_local(tsdf.ctypes.data_as(C.POINTER(C.c_float)), weight.ctypes.data_as(C.POINTER(C.c_float)),
                             color.ctypes.data_as(C.POINTER(C.c_float)), C.c_int64(tsdf.size))


def recenter(old_tsdf, old_weight, old_color, old_dim, old_origin, new_dim, new_origin, voxel_size, threads=8):
    """model/Volume.py:128-194 (swap_rot_trans) on numpy fp32 arrays: returns the three new arrays."""
    L = lib()
    n = int(np.prod(new_dim))
    out = [np.empty(n, np.float32) for _ in range(3)]
    olds = [np.ascontiguousarray(a, np.float32) for a in (old_tsdf, old_weight, old_color)]
    dim = np.ascontiguousarray(new_dim, np.int32); odim = np.ascontiguousarray(old_dim, np.int32)
    org = np.ascontiguousarray(new_origin, np.float32); oorg = np.ascontiguousarray(old_origin, np.float32)
    P = lambda a, t: a.ctypes.data_as(C.POINTER(t))

    def run(lo, hi):
        L.oracle_recenter(P(out[0], C.c_float), P(out[1], C.c_float), P(out[2], C.c_float),
                          P(olds[0], C.c_float), P(olds[1], C.c_float), P(olds[2], C.c_float),
                          P(dim, C.c_int32), P(org, C.c_float), P(odim, C.c_int32), P(oorg, C.c_float),
                          C.c_float(voxel_size), C.c_int64(lo), C.c_int64(hi))
    cuts = np.linspace(0, n, max(1, min(threads * 4, n)) + 1).astype(np.int64)
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:          # ctypes releases the GIL during the call
        list(ex.map(lambda ab: run(int(ab[0]), int(ab[1])), zip(cuts[:-1], cuts[1:])))
    return out
