"""Generate tests/golden/track_golden.npz on a GPU box: outputs of the LITERAL reference tracker kernels
(oracle/_ref/ref_tracker.cubin, compiled by oracle/build_ref.py from the strings in model/ROtracker.py:141-400) on a small
seeded case that the NumPy oracle replays on the CPU (tests/test_track_oracle.py).

    python tests/golden/make_track_golden.py gpurun_out/track_golden.npz      # then copy to tests/golden/
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_kernels as RK          # noqa: E402


def main(dst):
    g = np.random.default_rng(11)
    H, W = 96, 128
    K = np.array([[110.0, 0, 63.5], [0, 110.0, 47.5], [0, 0, 1]], np.float32)
    # a tilted plane ~2 m in front of the camera with a bump, some invalid pixels
    v, u = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    depth = (2.0 + 0.002 * u - 0.003 * v + 0.15 * np.exp(-((u - 70) ** 2 + (v - 40) ** 2) / 300.0)).astype(np.float32)
    depth[g.random((H, W)) < 0.03] = 0
    depth[:, :5] = 7.5                                           # beyond cut_dist
    dims = (72, 64, 80); voxel = 0.05; origin = np.array([-2.0, -1.0, 0.0], np.float32)   # integer origin (int truncation inert)
    tsdf = np.clip(g.normal(0, 0.6, int(np.prod(dims))), -1, 1).astype(np.float32)
    th = 0.1
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]], np.float32)
    T = np.array([0.1, 0.4, 0.3], np.float32)
    n = 1024
    cand = (g.random((n, 6)).astype(np.float32) * 2 - 1); cand[0] = 0
    ss = np.array([0.03, 0.02, 0.04, 0.01, 0.015, 0.02], np.float32)
    out = dict(depth=depth, K=K, tsdf=tsdf, dims=np.array(dims), origin=origin, voxel=np.float32(voxel), R=R, T=T, cand=cand, ss=ss)
    d = torch.from_numpy(depth).cuda(); t = torch.from_numpy(tsdf).cuda()
    for tag, seed, sr in (("a", 4242, 3.0), ("b", 17, 0.5)):
        vert, nrm = RK.ref_track_vertex_normal(d, K, 6.0, 0.06, seed, sr)
        vert_np = vert.cpu().numpy().reshape(H, W, 4)
        # the per-row sample the kernel drew: gt_tsdf = -sample wherever |sample| <= 1 (the clamp of :328-334 hides the rest)
        out[f"{tag}_vertex"] = vert_np; out[f"{tag}_normal"] = nrm.cpu().numpy().reshape(H, W, 3)
        out[f"{tag}_seed"] = np.int64(seed); out[f"{tag}_sample_range"] = np.float32(sr)
        # the per-row samples: curand is NVIDIA's generator and is not restated in the oracle, so the golden file carries
        # the draws.  They are read from the product's row_sample buffer and accepted only if the vertex map they
        # produce is bit-identical to the literal kernel's (which seeds one stream per pixel, subsequence = row).
        import ctypes as C
        from remixfusion_b200 import abi
        rs = torch.zeros(H, device="cuda"); pv = torch.zeros(H * W * 4, device="cuda"); pn = torch.zeros(H * W * 3, device="cuda")
        Kf = np.ascontiguousarray(K.reshape(-1))
        abi.check(abi.lib().rf_track_vertex_normal(abi.dptr(d), H, W, abi.fptr(Kf), C.c_float(6.0), C.c_float(0.06), C.c_int(seed),
                                                   C.c_float(sr), abi.dptr(rs), abi.dptr(pv), abi.dptr(pn), abi.stream_ptr()), "vertex")
        assert torch.equal(pv, vert) and torch.equal(pn, nrm), "row samples not validated"
        out[f"{tag}_row_sample"] = rs.cpu().numpy()
        for level, li in ((8, 3), (4, 1)):
            val, cnt = RK.ref_track_fitness(t, dims, origin, voxel, vert, nrm, H, W, K, R, T, cand, ss, level, li)
            out[f"{tag}_value_l{level}"] = val.cpu().numpy(); out[f"{tag}_count_l{level}"] = cnt.cpu().numpy()
    np.savez_compressed(dst, **out)
    print("wrote", dst, {k: getattr(v, "shape", None) for k, v in out.items() if k.startswith("a_")})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/track_golden.npz")
