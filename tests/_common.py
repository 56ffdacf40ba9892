"""Shared test helpers: seeded synthetic frames and volume set-ups."""
import numpy as np

from remixfusion_b200 import synth


def frame(cam, bound, eye, target, seed=0, invalid_frac=0.02):
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    scene = synth.make_scene(bound, seed)
    c2w = synth.look_at(eye, target)
    depth, rgb = synth.render_frame(scene, K, cam["H"], cam["W"], c2w, invalid_frac=invalid_frac, seed=seed)
    return K, c2w, depth, rgb


def small_cam(scale=4):
    c = synth.CFG1_CAM
    return dict(H=c["H"] // scale, W=c["W"] // scale, fx=c["fx"] / scale, fy=c["fy"] / scale,
                cx=(c["cx"] + 0.5) / scale - 0.5, cy=(c["cy"] + 0.5) / scale - 0.5)


def random_pose(rng, bound):
    b = np.asarray(bound, dtype=np.float64)
    lo, hi = b[:, 0], b[:, 1]
    eye = lo + (0.3 + 0.4 * rng.random(3)) * (hi - lo)
    tgt = lo + rng.random(3) * (hi - lo)
    c2w = synth.look_at(eye, tgt, up=rng.standard_normal(3))
    return c2w
