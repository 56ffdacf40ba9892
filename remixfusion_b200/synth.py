"""Seeded analytic RGB-D scene (datasets are not available offline; SURVEY.md §8d "Synthetic inputs").

A box room with three spheres and one slab inside; exact ray-cast z-depth (metres, fp32) for a pin-hole
camera in the OpenCV convention the reference uses (datasets/utils.py:24-56: directions ((i-cx)/fx,
(j-cy)/fy, 1), so the ray parameter *is* the z-depth), smooth procedural colour, and a seeded fraction of
zeroed pixels to exercise the invalid-depth paths.  Pure numpy: this is input generation, not the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Scene:
    room_lo: np.ndarray
    room_hi: np.ndarray
    spheres: list = field(default_factory=list)      # (centre[3], radius)
    slab_lo: np.ndarray | None = None
    slab_hi: np.ndarray | None = None


def make_scene(bound, seed: int = 0) -> Scene:
    """Room slightly inside `bound` ([[x0,x1],[y0,y1],[z0,z1]]), 3 spheres + 1 slab placed from the seed."""
    rng = np.random.default_rng(seed)
    b = np.asarray(bound, dtype=np.float64)
    ext = b[:, 1] - b[:, 0]
    lo = b[:, 0] + 0.04 * ext
    hi = b[:, 1] - 0.04 * ext
    spheres = []
    for _ in range(3):
        c = lo + (0.2 + 0.6 * rng.random(3)) * (hi - lo)
        r = 0.08 * ext.min() * (0.6 + 0.8 * rng.random())
        spheres.append((c, float(r)))
    sc = lo + (0.25 + 0.5 * rng.random(3)) * (hi - lo)
    sh = np.array([0.12, 0.03, 0.10]) * ext
    return Scene(lo, hi, spheres, sc - sh, sc + sh)


def intrinsics(fx, fy, cx, cy) -> np.ndarray:
    return np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64)


def look_at(eye, target, up=(0.0, 0.0, 1.0)) -> np.ndarray:
    """c2w (4x4, float64) for an OpenCV camera (x right, y down, z forward) at `eye` looking at `target`."""
    eye = np.asarray(eye, dtype=np.float64)
    f = np.asarray(target, dtype=np.float64) - eye
    f /= np.linalg.norm(f)
    upv = np.asarray(up, dtype=np.float64)
    r = np.cross(f, upv)
    if np.linalg.norm(r) < 1e-8:
        r = np.cross(f, np.array([0.0, 1.0, 0.0]))
    r /= np.linalg.norm(r)
    d = np.cross(f, r)
    c2w = np.eye(4)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = r, d, f, eye
    return c2w


def loop_trajectory(scene: Scene, n_frames: int, radius_frac: float = 0.18, seed: int = 0) -> np.ndarray:
    """Camera loop around the room centre looking outward-ish (poses [n,4,4], float64)."""
    c = 0.5 * (scene.room_lo + scene.room_hi)
    ext = scene.room_hi - scene.room_lo
    rad = radius_frac * min(ext[0], ext[1])
    poses = []
    for i in range(n_frames):
        a = 2 * np.pi * i / max(n_frames, 1)
        eye = c + np.array([rad * np.cos(a), rad * np.sin(a), 0.05 * ext[2] * np.sin(3 * a)])
        tgt = c + np.array([3 * rad * np.cos(a + 0.9), 3 * rad * np.sin(a + 0.9), -0.1 * ext[2]])
        poses.append(look_at(eye, tgt))
    return np.stack(poses)


def _box_interval(o, d, lo, hi):
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        t0 = (lo - o) * inv
        t1 = (hi - o) * inv
    tmin = np.minimum(t0, t1)
    tmax = np.maximum(t0, t1)
    return tmin.max(axis=-1), tmax.min(axis=-1)


def render_frame(scene: Scene, K: np.ndarray, H: int, W: int, c2w: np.ndarray, invalid_frac: float = 0.02,
                 seed: int = 0, max_depth: float | None = None):
    """Returns depth [H,W] float32 (0 = invalid) and rgb [H,W,3] float32 in [0,1]."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    jj, ii = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    dirs = np.stack([(ii - cx) / fx, (jj - cy) / fy, np.ones_like(ii)], -1).reshape(-1, 3)
    d = dirs @ c2w[:3, :3].T
    o = np.broadcast_to(c2w[:3, 3], d.shape)
    # room: camera is inside, the hit is the exit of the box interval
    _, t_room = _box_interval(o, d, scene.room_lo, scene.room_hi)
    t = np.where(t_room > 0, t_room, np.inf)
    for c, r in scene.spheres:
        oc = o - c
        a = (d * d).sum(-1)
        bq = (oc * d).sum(-1)
        cq = (oc * oc).sum(-1) - r * r
        disc = bq * bq - a * cq
        with np.errstate(invalid="ignore"):
            ts = (-bq - np.sqrt(disc)) / a
        ts = np.where((disc > 0) & (ts > 1e-3), ts, np.inf)
        t = np.minimum(t, ts)
    if scene.slab_lo is not None:
        tn, tf = _box_interval(o, d, scene.slab_lo, scene.slab_hi)
        ts = np.where((tn < tf) & (tn > 1e-3), tn, np.inf)
        t = np.minimum(t, ts)
    hit = o + d * np.where(np.isfinite(t), t, 0.0)[:, None]
    rgb = 0.5 + 0.5 * np.sin(hit * np.array([1.7, 2.3, 2.9]) + np.array([0.3, 1.1, 2.0]))
    rgb = 0.15 + 0.7 * rgb
    depth = np.where(np.isfinite(t), t, 0.0)
    if max_depth is not None:
        depth = np.where(depth > max_depth, 0.0, depth)
    if invalid_frac > 0:
        rng = np.random.default_rng(seed + 7919)
        depth = np.where(rng.random(depth.shape) < invalid_frac, 0.0, depth)
    return depth.reshape(H, W).astype(np.float32), rgb.reshape(H, W, 3).astype(np.float32)


def camera_dirs(K: np.ndarray, H: int, W: int) -> np.ndarray:
    """Per-pixel camera-frame ray directions [H,W,3] float32 (datasets/utils.py:24-56, 'OpenCV')."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    jj, ii = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    return np.stack([(ii - np.float32(cx)) / np.float32(fx), (jj - np.float32(cy)) / np.float32(fy),
                     np.ones_like(ii)], -1).astype(np.float32)


# Named configurations (BASELINE.json configs / SURVEY.md §8d)
REPLICA_BOUND = [[-1.0, 7.0], [-1.3, 3.7], [-1.7, 1.4]]          # configs/Replica/room0.yaml:3
REPLICA_CAM = dict(H=680, W=1200, fx=600.0, fy=600.0, cx=599.5, cy=339.5)   # configs/Replica/replica.yaml:86-97
CFG1_CAM = dict(H=480, W=640, fx=525.0, fy=525.0, cx=319.5, cy=239.5)
