"""Kernel-only timing of the tracker kernels, product vs the literal reference kernel (DESIGN.md §7, N2 second half).

    python profiles/time_track.py        (GPU box; oracle/_ref cubins required)
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
ev = lambda: torch.cuda.Event(enable_timing=True)
def timeit(fn, it=20):
    fn(); torch.cuda.synchronize(); a, b = ev(), ev(); a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it * 1e3
print("vertex+normal us:", timeit(lambda: s.init_depth_vertex(torch.from_numpy(depth).cuda(), K, seed_num=5)))
d = torch.from_numpy(depth).cuda()
print("ref vertex+normal us (incl. its host uploads):", timeit(lambda: RK.ref_track_vertex_normal(d, K, 6.0, s.truncation, 5, 3.0), 5))
R = c2w[:3, :3].copy(); T = c2w[:3, 3].copy()
ss = np.array([0.02, 0.02, 0.02, 0.01, 0.01, 0.01], np.float32)
for n, level, li in [(10240, 32, 5), (3072, 16, 10), (1024, 8, 1)]:
    cand = (np.random.default_rng(n).random((n, 6)).astype(np.float32) * 2 - 1)
    s.current_global_R, s.current_global_T, s.transform_candidate, s.search_size = R, T, cand, ss
    t_p = timeit(lambda: s.evaluate_tsdf(0, level, n, K, li, as_numpy=False))
    # reference kernel only (device buffers prepared once)
    m = RK._mod("ref_tracker.cubin")
    value = torch.zeros(n, device="cuda"); count = torch.zeros(n, device="cuda"); dummy = torch.zeros(1, device="cuda")
    keep = [RK._dev(ss), RK._dev(R), RK._dev(T), RK._dev(cand), RK._dev(K),
            RK._dev([vol.vol_dim[0], vol.vol_dim[1], vol.vol_dim[2], vol.vol_origin[0], vol.vol_origin[1], vol.vol_origin[2], vol.voxel_size, n, level,
                     int(H / level), int(W / level), li, H, W, 0, 0, 0])]
    def ref():
        m.launch("compute_tsdf_value", (int(n / 1024), int(H / level), int(W / level)), (1024, 1, 1),
                 [vol.tsdf_vol_gpu.data_ptr(), dummy.data_ptr(), dummy.data_ptr(), s.depth_vertex_gpu.data_ptr(), value.data_ptr(), count.data_ptr(),
                  keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(), keep[3].data_ptr(), keep[5].data_ptr(), keep[4].data_ptr(), s.normal_vertex_gpu.data_ptr()])
    t_r = timeit(ref)
    import ctypes as C
    from remixfusion_b200 import abi
    L = abi.lib(); L.rf_profile_enable(1)
    s.evaluate_tsdf(0, level, n, K, li, as_numpy=False); buf = (C.c_float * 64)(); L.rf_profile_read(buf); L.rf_profile_enable(0)
    print(f"n={n} level={level}: product call {t_p:.1f} us (kernel {buf[14]*1e3:.1f} us), reference kernel {t_r:.1f} us, pairs {n*(H//level)*(W//level)/1e6:.1f} M")

# ---- the whole search loop: host-driven (the reference's structure: 20 x [fitness, read-back, cal_transform, policy in NumPy]) vs
# rf_track_random_optimization (20 iterations enqueued back to back, one read-back per frame); reference-sized PST tables
import time
from oracle import track_oracle as TO
def pst(seed, sizes=(10240, 3072, 1024), tables=7):
    g = np.random.default_rng(seed); out = []
    for n_ in sizes:
        a = np.clip(g.normal(0.0, 0.35, size=(tables, n_, 6)), -0.99, 0.99).astype(np.float32) * np.float32(0.57); a[:, 0, :] = 0.0; out.append(a)
    return out
ro = dict(init_size=0.02, scaling_coefficient=0.09, particle_iter_lens=20, PST_size=[10240, 3072, 1024], fix_level_index=False, count_search=200, iterative_scale=True)
from remixfusion_b200.tracker import ROSearch
a_dev = ROSearch(vol, H, W, 6.0, s.truncation, 3.0); a_dev.configure_search(ro, pst(1))
a_host = ROSearch(vol, H, W, 6.0, s.truncation, 3.0); a_host.configure_search(ro, pst(1))
def wall(fn, it=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / it * 1e3
t_dev = wall(lambda: a_dev.random_optimization(0, c2w, None, depth, K, seed_num=5))
t_host = wall(lambda: TO.random_optimization(a_host, 0, c2w, depth, K, seed_num=5))
print(f"search loop per frame (20 iterations, wall clock incl. vertex / normal maps): device loop {t_dev:.2f} ms, host-driven loop {t_host:.2f} ms")
