import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from oracle import ref_kernels as RK
from remixfusion_b200 import configs, synth
from remixfusion_b200.volume import moving_volume
from remixfusion_b200.tracker import ROSearch
cuda = torch.device("cuda:0")
cfg = configs.replica(); cam = cfg["cam"]; H, W = cam["H"], cam["W"]
K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
sc = synth.make_scene(cfg["mapping"]["bound"], 0)
c2w = synth.loop_trajectory(sc, 200)[3].astype(np.float32)
depth, rgb = synth.render_frame(sc, K, H, W, c2w, seed=3)
vol = moving_volume(cfg, None, c2w, device=cuda)
vol.integrate(np.floor(rgb * 255.0).astype(np.float32), depth, K, c2w, None, 1.0, 0.0)
s = ROSearch(vol, H, W, 6.0, cfg["volume"]["trunc"], 3.0)
s.init_depth_vertex(depth, K, seed_num=1234)
n, level, li = 10240, 32, 5
g = np.random.default_rng(n)
R = c2w[:3, :3].copy(); T = c2w[:3, 3].copy()
cand = (g.random((n, 6)).astype(np.float32) * 2 - 1); cand[0] = 0
ss = np.array([0.02, 0.02, 0.02, 0.01, 0.01, 0.01], np.float32)
s.current_global_R, s.current_global_T, s.transform_candidate, s.search_size = R, T, cand, ss
def ref(normal=None):
    return RK.ref_track_fitness(vol.tsdf_vol_gpu, vol.vol_dim, vol.vol_origin, vol.voxel_size, s.depth_vertex_gpu,
                                s.normal_vertex_gpu if normal is None else normal, H, W, K, R, T, cand, ss, level, li)
_, val, cnt = s.evaluate_tsdf(0, level, n, K, li, as_numpy=False)
v1, c1 = ref(); v2, c2 = ref()
print("ref vs ref max diff", float((v1 - v2).abs().max()), "prod vs ref", float((val - v1).abs().max()))
bad = torch.nonzero((val - v1).abs() > 1e-2).view(-1).tolist()
print("bad candidates", bad, [(float(val[b]), float(v1[b]), float(cnt[b]), float(c1[b])) for b in bad])
full = s.normal_vertex_gpu.clone()
ph, pw = H // level, W // level
for p in range(ph * pw):
    pi = (p // pw) * level + li; pj = (p % pw) * level + li; i = pi * W + pj
    if float(full[3*i:3*i+3].abs().sum()) == 0: continue
    one = torch.zeros_like(full); one[3*i:3*i+3] = full[3*i:3*i+3]
    s.normal_vertex_gpu = one
    _, a, ca = s.evaluate_tsdf(0, level, n, K, li, as_numpy=False)
    b, cb = ref(one)
    d = (a - b).abs()
    if float(d.max()) > 0 or not torch.equal(ca, cb):
        j = int(d.argmax())
        print("pixel", pi, pj, "cand", j, float(a[j]), float(b[j]), float(ca[j]), float(cb[j]), "vertex", s.depth_vertex_gpu[4*i:4*i+4].tolist())
print("done")
