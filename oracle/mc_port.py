"""TEST INFRASTRUCTURE ONLY.  NumPy restatement of the dual-grid marching cubes of the reference's
thirdparty/NumpyMarchingCubes/marching_cubes/src/marching_cubes.cpp (trilerp :94-118, vertexInterp :120-137,
extract_isosurface_at_position :139-244, run_marching_cubes_internal :424-438), in the SAME formulation the device kernels of
remixfusion_b200/csrc/marching_cubes.cu use — and reading the case table and the corner / edge numbering constants OUT OF
THAT .cu FILE, so that the CPU suite checks the very constants the kernels are compiled with against the reference's compiled
C++ (oracle/_ref/libmc_ref.so) and its golden outputs (tests/golden/mc_golden.npz)."""
from __future__ import annotations

import os
import re

import numpy as np

CU = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "remixfusion_b200", "csrc", "marching_cubes.cu")


def kernel_constants():
    src = open(CU).read()
    tri = [int(x, 16) for x in re.findall(r"0x([0-9a-f]{16})ull", re.search(r"kTriTable\[256\] = \{(.*?)\};", src, re.S).group(1))]
    assert len(tri) == 256
    ints = lambda name: [int(x) for x in re.findall(r"-?\d+", re.search(name + r"(?:\[\d+\])+ = \{(.*?)\};", src, re.S).group(1))]
    return tri, ints("kEdgeA"), ints("kEdgeB"), ints("kCornerBit"), np.array(ints("kCornerOff")).reshape(8, 3)


def triangle_soup(volume, isovalue=0.0, truncation=3.0, thresh=10.0):
    """[T,3,3] float32 triangle soup in the reference's order (cells i, j, k with k fastest; table order inside a cell)."""
    tri, ea, eb, cbit, coff = kernel_constants()
    vol = np.asarray(volume, dtype=np.float32)
    X, Y, Z = vol.shape
    f32 = np.float32
    corner = np.full((X + 1, Y + 1, Z + 1), np.nan, dtype=np.float32)
    if X > 1 and Y > 1 and Z > 1:
        ok = np.ones((X - 1, Y - 1, Z - 1), bool); acc = np.zeros((X - 1, Y - 1, Z - 1), np.float32)
        for ox, oy, oz in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 1)):     # trilerp's order
            d = vol[ox:X - 1 + ox, oy:Y - 1 + oy, oz:Z - 1 + oz]
            with np.errstate(invalid="ignore"):
                ok &= (d != -np.inf) & (np.abs(d) < f32(truncation))
            acc = (acc + f32(0.125) * d).astype(np.float32)
        corner[1:X, 1:Y, 1:Z] = np.where(ok, acc, np.nan)
    out = []
    iso = f32(isovalue)
    for i in range(X):
        for j in range(Y):
            for k in range(Z):
                d = np.array([corner[i + o[0], j + o[1], k + o[2]] for o in coff], dtype=np.float32)
                if np.isnan(d).any():
                    continue
                cube = sum(cbit[c] for c in range(8) if d[c] < iso)
                bad = np.any(np.abs(d) > thresh)
                prod = d[:, None] * d[None, :]
                bad |= np.any(np.where(prod < 0, np.abs(d)[:, None] + np.abs(d)[None, :], np.abs(d[:, None] - d[None, :])) > thresh)
                if bad:
                    continue
                row = tri[cube]
                es = []
                while len(es) < 15 and ((row >> (4 * len(es))) & 0xF) != 0xF:
                    es.append((row >> (4 * len(es))) & 0xF)
                if not es or sum(1 << e for e in set(es)) == 255:
                    continue
                pos = np.array([i, j, k], dtype=np.float32)
                verts = []
                for e in es:
                    a, b = ea[e], eb[e]
                    pa = pos + np.where(coff[a] > 0, f32(0.5), f32(-0.5)).astype(np.float32)
                    pb = pos + np.where(coff[b] > 0, f32(0.5), f32(-0.5)).astype(np.float32)
                    d1, d2 = d[a], d[b]
                    if abs(iso - d1) < f32(0.00001):
                        verts.append(pa)
                    elif abs(iso - d2) < f32(0.00001):
                        verts.append(pb)
                    elif abs(d1 - d2) < f32(0.00001):
                        verts.append(pa)
                    else:
                        mu = f32(f32(iso - d1) / f32(d2 - d1))
                        verts.append((pa + (mu * (pb - pa)).astype(np.float32)).astype(np.float32))
                out.append(np.stack(verts).reshape(-1, 3, 3))
    return np.concatenate(out, 0) if out else np.zeros((0, 3, 3), np.float32)
