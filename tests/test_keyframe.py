"""N3: the GPU-resident keyframe store against the reference's own KeyFrameDatabase (golden vectors made by
tests/golden/make_keyframe_golden.py from model/keyframe.py): driven with the indices the reference drew, it must store
and return exactly the same rays and frame ids (random and 'filter_depth' keyframes; the reference's fall-through branch for
too few valid depths indexes out of range in the reference itself and is not exercised)."""
import os

import numpy as np
import torch

from remixfusion_b200.ray_store import KeyFrameDatabase, _distinct

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "keyframe_golden.npz"))


def _run(device):
    H, W, NUM_KF, KEEP, BS = (int(v) for v in G["meta"])
    db = KeyFrameDatabase({"cam": {"depth_trunc": 3.0}}, H, W, NUM_KF, KEEP, device)
    for i in range(3):
        batch = {k: torch.from_numpy(G[f"kf{i}_{k}"]) for k in ("direction", "rgb", "depth")}
        batch["frame_id"] = int(G[f"kf{i}_frame_id"])
        db.add_keyframe(batch, filter_depth=bool(G[f"kf{i}_filter"]), idxs=G[f"kf{i}_idxs"])
        assert np.array_equal(db.rays[i].cpu().numpy(), G[f"kf{i}_rays"]), f"keyframe {i}"
    rays, fids = db.sample_global_rays(BS, idxs=G["sample_idxs"])
    assert np.array_equal(rays.cpu().numpy(), G["sample_rays"]) and np.array_equal(fids.cpu().numpy(), G["sample_frame_ids"])
    # device-side draws: distinct, in range, right count
    rays2, fids2 = db.sample_global_rays(BS)
    assert rays2.shape == (BS, 7) and fids2.shape == (BS,)
    idx = _distinct(3 * KEEP, BS, db.device)
    assert len(torch.unique(idx)) == BS and int(idx.min()) >= 0 and int(idx.max()) < 3 * KEEP
    return db


def test_keyframe_store_matches_reference_cpu():
    _run("cpu")


import pytest  # noqa: E402


@pytest.mark.gpu
def test_keyframe_store_matches_reference_gpu(cuda):
    db = _run(cuda)
    assert db.rays.is_cuda
