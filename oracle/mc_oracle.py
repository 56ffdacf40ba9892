"""TEST INFRASTRUCTURE ONLY.  The reference's own C++ marching cubes (thirdparty/NumpyMarchingCubes/marching_cubes/src/
marching_cubes.cpp, the `mcubes.marching_cubes` of utils.py:169) as a CPU oracle: oracle/build_ref.py compiles it from
/root/reference into oracle/_ref/libmc_ref.so (which travels to the GPU box); this module calls it through ctypes."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libmc_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def marching_cubes(volume: np.ndarray, isovalue: float, truncation: float):
    """(vertices [V,3] float64 in voxel units, faces [F,3] uint64) exactly as the reference's extension returns them."""
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB)
        _lib.mc_ref.restype = C.c_long
    vol = np.ascontiguousarray(volume, dtype=np.float64)
    assert vol.ndim == 3
    v = C.POINTER(C.c_double)(); f = C.POINTER(C.c_ulong)(); nf = C.c_long()
    nv = _lib.mc_ref(vol.ctypes.data_as(C.POINTER(C.c_double)), *[C.c_long(int(d)) for d in vol.shape], C.c_double(isovalue), C.c_double(truncation),
                     C.byref(v), C.byref(f), C.byref(nf))
    V = np.ctypeslib.as_array(v, (max(nv, 1), 3))[:nv].copy() if nv else np.zeros((0, 3))
    F = np.ctypeslib.as_array(f, (max(nf.value, 1), 3))[:nf.value].copy() if nf.value else np.zeros((0, 3), dtype=np.uint64)
    _lib.mc_free(v); _lib.mc_free(f)
    return V, F
