"""Generates tests/golden/keyframe_golden.npz by running the REFERENCE's own KeyFrameDatabase (model/keyframe.py, imported
from /root/reference — only possible in the build container) on seeded synthetic keyframes, recording the indices its
``random.sample`` calls drew so that the product can be driven with the same indices.  Run:  python tests/golden/make_keyframe_golden.py"""
import os
import random
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from model.keyframe import KeyFrameDatabase          # noqa: E402

H, W, NUM_KF, KEEP, BS = 12, 16, 4, 20, 37
cfg = {"cam": {"depth_trunc": 3.0}}
rng = np.random.default_rng(0)
drawn = []
_orig = random.sample


def _rec(pop, k):
    out = _orig(pop, k)
    drawn.append(np.asarray(out, dtype=np.int64))
    return out


random.sample = _rec
random.seed(0)
db = KeyFrameDatabase(cfg, H, W, NUM_KF, KEEP, "cpu")
out = {}
for i, (fid, filt) in enumerate([(0, False), (5, True), (10, True)]):
    batch = {"direction": torch.from_numpy(rng.standard_normal((1, H, W, 3)).astype(np.float32)),
             "rgb": torch.from_numpy(rng.random((1, H, W, 3)).astype(np.float32)),
             "depth": torch.from_numpy((rng.random((1, H, W)) * 4.0 * (rng.random((1, H, W)) > 0.2)).astype(np.float32)),
             "frame_id": fid}
    for k in ("direction", "rgb", "depth"):
        out[f"kf{i}_{k}"] = batch[k].numpy().copy()
    out[f"kf{i}_frame_id"] = np.int64(fid); out[f"kf{i}_filter"] = np.int64(filt)
    n0 = len(drawn)
    db.add_keyframe(batch, filter_depth=filt)
    out[f"kf{i}_idxs"] = drawn[n0]
    out[f"kf{i}_rays"] = db.rays[i].numpy().copy()
n0 = len(drawn)
rays, fids = db.sample_global_rays(BS)
out["sample_idxs"] = drawn[n0]; out["sample_rays"] = rays.numpy().copy(); out["sample_frame_ids"] = fids.numpy().copy()
out["meta"] = np.asarray([H, W, NUM_KF, KEEP, BS], dtype=np.int64)
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "keyframe_golden.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})
