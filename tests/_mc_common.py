"""Seeded test volumes and the triangle-set comparison shared by the marching-cubes tests."""
import numpy as np


def mc_volumes():
    """name -> (volume float32 [X,Y,Z], isovalue, truncation)."""
    rng = np.random.default_rng(0)
    vols = {}
    g = np.stack(np.meshgrid(np.arange(24), np.arange(19), np.arange(29), indexing="ij"), -1).astype(np.float64)
    sphere = np.linalg.norm(g - np.array([11.3, 9.1, 13.7]), axis=-1) - 6.2
    vols["sphere"] = (sphere.astype(np.float32), 0.0, 3.0)                      # truncation clips the far field: |d| >= 3 is invalid
    # two blobs + smooth noise, scaled like a truncated SDF in [-1, 1], with an unobserved (NaN) region and a -inf patch
    g2 = np.stack(np.meshgrid(np.arange(33), np.arange(27), np.arange(21), indexing="ij"), -1).astype(np.float64)
    f = np.minimum(np.linalg.norm(g2 - [10, 12, 9], axis=-1) - 5.5, np.linalg.norm(g2 - [22, 13, 11], axis=-1) - 6.5)
    f += 0.6 * np.sin(g2[..., 0] * 0.9) * np.cos(g2[..., 1] * 0.7 + g2[..., 2] * 0.5)
    f = np.clip(f / 3.0, -1.0, 1.0).astype(np.float32)
    f[25:, :, :6] = np.nan
    f[:4, 20:, :] = -np.inf
    vols["blobs"] = (f, 0.0, 3.0)
    vols["blobs_iso"] = (f, 0.13, 0.9)                                          # non-zero level, truncation inside the value range
    # exact hits: integer-valued field so that corner averages land on the iso level (vertexInterp short-cuts)
    q = (np.round(np.linalg.norm(g - [12, 9, 14], axis=-1)) - 6.0).astype(np.float32)
    vols["exact_hits"] = (q, 0.0, 3.0)
    vols["noise"] = ((rng.random((14, 15, 13)) * 2 - 1).astype(np.float32), 0.0, 3.0)     # every case of the table
    vols["empty"] = (np.full((9, 8, 7), 0.7, np.float32), 0.0, 3.0)
    return vols


def canonical_triangles(V, F, min_area=1e-9):
    """Triangle soup of an indexed mesh as a sorted [T, 9] array (vertices of a triangle sorted lexicographically, triangles
    sorted), slivers dropped: a representation independent of vertex numbering and of the orientation-preserving rotation."""
    V = np.asarray(V, dtype=np.float64); F = np.asarray(F).astype(np.int64)
    if F.size == 0:
        return np.zeros((0, 9))
    T = V[F]                                                                     # [T,3,3]
    area = 0.5 * np.linalg.norm(np.cross(T[:, 1] - T[:, 0], T[:, 2] - T[:, 0]), axis=1)
    T = T[area > min_area]
    order = np.lexsort((T[..., 2], T[..., 1], T[..., 0]), axis=1) if False else None
    rows = []
    for t in T:
        idx = sorted(range(3), key=lambda i: (round(t[i, 0], 4), round(t[i, 1], 4), round(t[i, 2], 4)))
        rows.append(t[idx].reshape(-1))
    A = np.asarray(rows)
    key = np.round(A, 3)
    return A[np.lexsort(key.T[::-1])]


def same_surface(Va, Fa, Vb, Fb, atol=3e-5):
    A, B = canonical_triangles(Va, Fa), canonical_triangles(Vb, Fb)
    if A.shape != B.shape:
        return False, f"{A.shape[0]} vs {B.shape[0]} triangles"
    if A.size == 0:
        return True, "empty"
    err = float(np.abs(A - B).max())
    return err <= atol, f"max vertex deviation {err:.2e} over {A.shape[0]} triangles"
