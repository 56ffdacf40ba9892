// TEST INFRASTRUCTURE ONLY.  Lets the reference's thirdparty/NumpyMarchingCubes/marching_cubes/src/marching_cubes.cpp be
// compiled WHERE IT LIES without Python / NumPy headers (its own wrapper targets an old NumPy C-API and no longer builds):
// oracle/build_ref.py compiles that file with `-D_EXTMODULE_H -include mc_ref_shim.h`, which skips the body of the
// reference's pyarraymodule.h and supplies the three names its marching_cubes.h needs — an array handle, the index type and
// the element getter — over a plain C-contiguous [X][Y][Z] double volume.
#pragma once
#include <cassert>
struct PyArrayObject { const double* data; long sx, sy, sz; };
typedef long npy_intp;
template <class T> T PyArray_SafeGet(const PyArrayObject* a, const npy_intp* c) { return (T)a->data[(c[0] * a->sy + c[1]) * a->sz + c[2]]; }
