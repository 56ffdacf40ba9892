"""Generates tests/golden/tsdf_golden.npz by running the LITERAL reference kernels (oracle/_ref cubins, built
from /root/reference by oracle/build_ref.py) on a B200:

    gpurun -- 'python tests/golden/make_tsdf_golden.py gpurun_out/tsdf_golden.npz'

Inputs (frames, poses) are stored in the file so that the CPU suite does not depend on bit-reproducible
scene generation.  Outputs are stored as sha256 digests of the full fp32 arrays plus a strided sample of the
touched voxels (index, values), which is what tests/test_tsdf_oracle.py checks the C oracle against.
"""
import hashlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_kernels, tsdf_oracle as O          # noqa: E402
from remixfusion_b200 import synth                        # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float32).tobytes()).hexdigest()


def main(out):
    assert ref_kernels.available(), "needs a GPU and oracle/_ref/*.cubin"
    dev = torch.device("cuda:0")
    c = synth.CFG1_CAM
    s = 8
    cam = dict(H=c["H"] // s, W=c["W"] // s, fx=c["fx"] / s, fy=c["fy"] / s, cx=(c["cx"] + .5) / s - .5, cy=(c["cy"] + .5) / s - .5)
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    bound = [[-3.0, 3.0], [-3.0, 3.0], [-2.0, 2.0]]
    scene = synth.make_scene(bound, 3)
    rng = np.random.default_rng(42)
    poses, depths, rgbs = [], [], []
    for i in range(3):
        eye = (rng.random(3) - 0.5) * np.array([1.5, 1.5, 0.8])
        tgt = (rng.random(3) - 0.5) * np.array([5.0, 5.0, 2.5])
        c2w = synth.look_at(eye, tgt, up=rng.standard_normal(3))
        d, r = synth.render_frame(scene, K, cam["H"], cam["W"], c2w, seed=i)
        poses.append(c2w.astype(np.float32)); depths.append(d); rgbs.append(r)
    res = dict(K=K.astype(np.float32), poses=np.stack(poses), depths=np.stack(depths), rgbs=np.stack(rgbs),
               bound=np.asarray(bound, np.float32))
    # ---- local volume: 120 x 120 x 80 @ 0.05, origin (-3,-3,-2), trunc 0.06, weight clamp on
    dims = np.array([120, 120, 80]); origin = np.array([-3.0, -3.0, -2.0], np.float32); voxel = 0.05; trunc = 0.06
    n = int(dims.prod())
    vol = [torch.ones(n + 64, device=dev), torch.zeros(n + 64, device=dev), torch.zeros(n + 64, device=dev)]
    for c2w, d, r in zip(poses, depths, rgbs):
        packed = O.pack_bgr(np.floor(r * 255.0).astype(np.float32))
        ref_kernels.ref_integrate_local(vol[0], vol[1], vol[2], dims, origin, voxel, K, c2w, torch.from_numpy(d).to(dev),
                                        torch.from_numpy(packed).to(dev), trunc, obs_weight=1.0, weight_clamp=1.0)
    t, w, col = (v[:n].cpu().numpy() for v in vol)
    idx = np.nonzero(w > 0)[0]
    samp = idx[:: max(1, idx.size // 4096)]
    res.update(local_dims=dims, local_origin=origin, local_voxel=np.float32(voxel), local_trunc=np.float32(trunc),
               local_sha=np.array([sha(t), sha(w), sha(col)]), local_n_touched=np.int64(idx.size),
               local_idx=samp, local_tsdf=t[samp], local_weight=w[samp], local_color=col[samp])
    # ---- global volume: R = 64 over `bound`, trunc 0.1
    R = 64
    trgb = torch.zeros(4 * R ** 3 + 64, device=dev); gw = torch.zeros(R ** 3 + 64, device=dev)
    ref_kernels.ref_clear_global(trgb, R)
    box = [v for ax in bound for v in ax]
    for c2w, d, r in zip(poses, depths, rgbs):
        ref_kernels.ref_integrate_global(trgb, gw, R, box, K, c2w, torch.from_numpy(d).to(dev), torch.from_numpy(r).to(dev), 0.1, 1.0)
    gt = trgb[:4 * R ** 3].cpu().numpy(); gww = gw[:R ** 3].cpu().numpy()
    idx = np.nonzero(gww > 0)[0]
    samp = idx[:: max(1, idx.size // 4096)]
    res.update(global_R=np.int64(R), global_trunc=np.float32(0.1), global_sha=np.array([sha(gt), sha(gww)]),
               global_n_touched=np.int64(idx.size), global_idx=samp, global_trgb=gt.reshape(-1, 4)[samp], global_w=gww[samp])
    np.savez_compressed(out, **res)
    print("wrote", out, {k: (v.shape if hasattr(v, "shape") else v) for k, v in res.items() if "sha" not in k})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tsdf_golden.npz")
