"""TEST INFRASTRUCTURE ONLY.  Build the *literal reference kernels* into oracle/_ref/ (git-ignored).

The reference keeps its TSDF kernels as CUDA-C strings handed to ``pycuda.compiler.SourceModule``
(model/Volume.py:127-611, mp_slam/mapper.py:36-185).  PyCUDA wraps such a string in ``extern "C" { }``
and runs ``nvcc --cubin -arch sm_XX`` with default flags.  This script does exactly that, without
PyCUDA: it reads the strings *where they lie* under /root/reference, writes them to a temporary
directory outside the repo, and cross-compiles them for sm_100a.  Only the resulting cubins land in
oracle/_ref/ — no reference source is copied into the repository.

The cubins travel to the GPU box with the repo snapshot (``.gitignore`` lists oracle/_ref/, ``.gpurunignore``
does not), where oracle/ref_kernels.py loads them with the CUDA driver API and launches them on torch
tensors: that is the reference itself, run beside the product kernels, and it is what pins parity for
Stage 1 (tests/test_tsdf_gpu.py; golden vectors in tests/golden/ are outputs of these cubins).
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("RF_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")

SOURCES = {
    # cubin name -> reference file holding the SourceModule string
    "ref_local_volume.cubin": "model/Volume.py",
    "ref_global_volume.cubin": "mp_slam/mapper.py",
    "ref_tracker.cubin": "model/ROtracker.py",          # compute_tsdf_value / compute_vertex / compute_normal (:141-400)
}
# model/ROtracker.py:398 passes no_extern_c=True (the string carries its own extern "C" block and includes curand)
NO_EXTERN_C = {"ref_tracker.cubin"}


MC_SRC = "thirdparty/NumpyMarchingCubes/marching_cubes/src"
MC_LIB = "libmc_ref.so"


def build_marching_cubes(verbose: bool = False) -> None:
    """The reference's C++ marching cubes (utils.py:169 `mcubes.marching_cubes`) as a CPU oracle: its marching_cubes.cpp is
    compiled where it lies, with oracle/mc_ref_shim.h standing in for the NumPy accessor and oracle/mc_ref_entry.cpp as the C
    entry point (the reference's own Cython / NumPy wrapper does not build against NumPy 2)."""
    src = os.path.join(REF_ROOT, MC_SRC)
    dst = os.path.join(OUT_DIR, MC_LIB)
    deps = [os.path.join(src, "marching_cubes.cpp"), os.path.join(HERE, "mc_ref_shim.h"), os.path.join(HERE, "mc_ref_entry.cpp")]
    if os.path.exists(dst) and all(os.path.getmtime(dst) >= os.path.getmtime(d) for d in deps):
        return
    cmd = ["g++", "-shared", "-fPIC", "-O2", "-std=c++14", "-w", "-D_EXTMODULE_H", "-include", os.path.join(HERE, "mc_ref_shim.h"),
           "-I" + src, os.path.join(HERE, "mc_ref_entry.cpp"), os.path.join(src, "marching_cubes.cpp"), "-o", dst]
    if verbose:
        print("[build_ref]", " ".join(cmd))
    subprocess.run(cmd, check=True)


def build(verbose: bool = False) -> bool:
    """Returns True when the cubins (and the marching-cubes oracle) are present after the call (built now or earlier)."""
    if not os.path.isdir(REF_ROOT):
        ok = all(os.path.exists(os.path.join(OUT_DIR, k)) for k in list(SOURCES) + [MC_LIB])
        if verbose:
            print(f"[build_ref] {REF_ROOT} absent; prebuilt oracle binaries present: {ok}")
        return ok
    os.makedirs(OUT_DIR, exist_ok=True)
    build_marching_cubes(verbose)
    for cubin, rel in SOURCES.items():
        dst = os.path.join(OUT_DIR, cubin)
        src_path = os.path.join(REF_ROOT, rel)
        if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src_path):
            continue
        text = open(src_path, "r", encoding="utf-8").read()
        m = re.search(r'SourceModule\("""(.*?)"""', text, re.S)
        if m is None:
            raise RuntimeError(f"no SourceModule string found in {src_path}")
        with tempfile.TemporaryDirectory(prefix="rf_ref_") as tmp:
            cu = os.path.join(tmp, "kernel.cu")
            with open(cu, "w") as f:
                if cubin in NO_EXTERN_C:
                    f.write(m.group(1))
                else:
                    f.write('extern "C" {\n' + m.group(1) + "\n}\n")   # what PyCUDA's SourceModule does
            cmd = ["nvcc", "--cubin", "-arch=sm_100a", "-w", "-o", dst, cu]   # nvcc defaults, as PyCUDA
            if verbose:
                print("[build_ref]", " ".join(cmd))
            subprocess.run(cmd, check=True)
    return True


if __name__ == "__main__":
    sys.exit(0 if build(verbose=True) else 1)
