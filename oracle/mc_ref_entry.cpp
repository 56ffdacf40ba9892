// TEST INFRASTRUCTURE ONLY.  C entry point around the reference's `marching_cubes()` (declared in its marching_cubes.h, defined
// in its marching_cubes.cpp; both are compiled from /root/reference by oracle/build_ref.py into oracle/_ref/libmc_ref.so).
#include "marching_cubes.h"
#include <cstdlib>
#include <cstring>

extern "C" long mc_ref(const double* vol, long sx, long sy, long sz, double iso, double trunc, double** verts, unsigned long** faces, long* n_faces) {
    PyArrayObject arr{vol, sx, sy, sz};
    npy_accessor acc(&arr, {sx, sy, sz});
    std::vector<double> v; std::vector<unsigned long> f;
    marching_cubes(acc, iso, trunc, v, f);
    *verts = (double*)malloc(v.size() * sizeof(double) + 8); memcpy(*verts, v.data(), v.size() * sizeof(double));
    *faces = (unsigned long*)malloc(f.size() * sizeof(unsigned long) + 8); memcpy(*faces, f.data(), f.size() * sizeof(unsigned long));
    *n_faces = (long)f.size() / 3;
    return (long)v.size() / 3;
}
extern "C" void mc_free(void* p) { free(p); }
