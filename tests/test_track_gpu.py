"""N2 (second half): tracker vertex / normal maps and candidate fitness vs the LITERAL reference kernels
(oracle/_ref/ref_tracker.cubin, built from the strings in model/ROtracker.py:141-400 by oracle/build_ref.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene(cuda, rf_lib):
    from oracle import ref_kernels as RK
    RK.require()                                   # mandatory on the GPU box: fail, never skip
    from remixfusion_b200 import configs, synth
    from remixfusion_b200.volume import moving_volume
    cfg = configs.replica()
    cam = cfg["cam"]; H, W = cam["H"], cam["W"]
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    sc = synth.make_scene(cfg["mapping"]["bound"], 0)
    traj = synth.loop_trajectory(sc, 200)
    c2w = traj[3].astype(np.float32)
    depth, rgb = synth.render_frame(sc, K, H, W, c2w, seed=3)
    vol = moving_volume(cfg, None, c2w, device=cuda)
    vol.integrate(np.floor(rgb * 255.0).astype(np.float32), depth, K, c2w, None, 1.0, 0.0)
    return dict(cfg=cfg, H=H, W=W, K=K, c2w=c2w, depth=depth, vol=vol, RK=RK)


def _search(scene, seed=1234, sample_range=3.0):
    from remixfusion_b200.tracker import ROSearch
    s = ROSearch(scene["vol"], scene["H"], scene["W"], cut_dist=6.0, truncation=scene["cfg"]["volume"]["trunc"], sample_range=sample_range)
    s.init_depth_vertex(scene["depth"], scene["K"], seed_num=seed)
    s.init_normal()
    return s


@pytest.mark.parametrize("seed,sample_range", [(1234, 3.0), (999983, 0.5), (7, 1.0)])
def test_vertex_normal_bit_exact(scene, seed, sample_range):
    s = _search(scene, seed, sample_range)
    d = torch.from_numpy(scene["depth"]).cuda()
    v_ref, n_ref = scene["RK"].ref_track_vertex_normal(d, scene["K"], 6.0, s.truncation, seed, sample_range)
    assert torch.equal(s.depth_vertex_gpu, v_ref)
    assert torch.equal(s.normal_vertex_gpu, n_ref)
    assert float(s.depth_vertex_gpu.abs().sum()) > 0 and float(s.normal_vertex_gpu.abs().sum()) > 0


def _pose_inputs(scene, n, g):
    c2w = scene["c2w"]
    R = c2w[:3, :3].copy(); T = c2w[:3, 3].copy()
    cand = (g.random((n, 6)).astype(np.float32) * 2 - 1)
    cand[0] = 0                                             # candidate 0 = the current pose (cal_transform's origin_tsdf)
    ss = np.array([0.02, 0.02, 0.02, 0.01, 0.01, 0.01], np.float32)
    return R, T, cand, ss


def _ref_fitness(scene, s, R, T, cand, ss, level, level_index, normal=None):
    v = scene["vol"]
    return scene["RK"].ref_track_fitness(v.tsdf_vol_gpu, v.vol_dim, v.vol_origin, v.voxel_size, s.depth_vertex_gpu,
                                         s.normal_vertex_gpu if normal is None else normal, scene["H"], scene["W"], scene["K"],
                                         R, T, cand, ss, level, level_index)


def test_fitness_terms_bit_exact(scene):
    """One valid pixel at a time: every candidate's sum is then a single term, so the comparison with the reference kernel
    is bit for bit (the reference's atomics make multi-term sums order-dependent)."""
    s = _search(scene)
    g = np.random.default_rng(5)
    R, T, cand, ss = _pose_inputs(scene, 1024, g)
    s.current_global_R, s.current_global_T, s.transform_candidate, s.search_size = R, T, cand, ss
    H, W = scene["H"], scene["W"]
    full = s.normal_vertex_gpu.clone()
    valid = torch.nonzero(full.view(-1, 3).abs().sum(1) > 0).view(-1)
    level = 8
    picks = []
    for i in valid[torch.randperm(valid.numel(), generator=torch.Generator().manual_seed(3))[:4000].cuda()].tolist():
        pi, pj = divmod(i, W)
        if pi % level == 3 and pj % level == 3 and pi < (H // level) * level and pj < (W // level) * level:
            picks.append(i)
    assert len(picks) >= 20
    hits = 0
    for i in picks[:40]:
        one = torch.zeros_like(full); one[3 * i:3 * i + 3] = full[3 * i:3 * i + 3]
        s.normal_vertex_gpu = one
        _, val, cnt = s.evaluate_tsdf(0, level, 1024, scene["K"], 3, as_numpy=False)
        v_ref, c_ref = _ref_fitness(scene, s, R, T, cand, ss, level, 3, normal=one)
        assert torch.equal(cnt, c_ref)
        assert torch.equal(val, v_ref), f"pixel {i}: max diff {float((val - v_ref).abs().max()):.3e}"
        hits += int(cnt.sum())
    assert hits > 1000
    # every valid pixel of the coarsest pyramid level (the three (pixel, candidate) pairs of this frame whose vertex lands
    # within an ulp of a voxel boundary — where a differently contracted multiply-add picks the neighbouring voxel — are here)
    level, li = 32, 5
    swept = 0
    for p in range((H // level) * (W // level)):
        i = ((p // (W // level)) * level + li) * W + (p % (W // level)) * level + li
        if float(full[3 * i:3 * i + 3].abs().sum()) == 0:
            continue
        one = torch.zeros_like(full); one[3 * i:3 * i + 3] = full[3 * i:3 * i + 3]
        s.normal_vertex_gpu = one
        _, val, cnt = s.evaluate_tsdf(0, level, 1024, scene["K"], li, as_numpy=False)
        v_ref, c_ref = _ref_fitness(scene, s, R, T, cand, ss, level, li, normal=one)
        assert torch.equal(cnt, c_ref) and torch.equal(val, v_ref), f"pixel {i}"
        swept += 1
    assert swept > 300
    s.normal_vertex_gpu = full


@pytest.mark.parametrize("n,level,level_index", [(10240, 32, 5), (3072, 16, 10), (1024, 8, 1)])
def test_fitness_matches_reference(scene, n, level, level_index):
    """The reference's PST sizes and pyramid levels on the 1200x680 frame: hit counts identical, sums within fp32
    summation-order noise (the reference adds with atomics in arbitrary order); two runs of the product are bit-identical."""
    s = _search(scene)
    g = np.random.default_rng(n)
    R, T, cand, ss = _pose_inputs(scene, n, g)
    s.current_global_R, s.current_global_T, s.transform_candidate, s.search_size = R, T, cand, ss
    norm, val, cnt = s.evaluate_tsdf(0, level, n, scene["K"], level_index, as_numpy=False)
    v_ref, c_ref = _ref_fitness(scene, s, R, T, cand, ss, level, level_index)
    assert torch.equal(cnt, c_ref) and float(cnt.sum()) > 0
    np.testing.assert_allclose(val.cpu().numpy(), v_ref.cpu().numpy(), rtol=2e-5, atol=1e-5)
    _, val2, cnt2 = s.evaluate_tsdf(0, level, n, scene["K"], level_index, as_numpy=False)
    assert torch.equal(val, val2) and torch.equal(cnt, cnt2)
    # the current pose (candidate 0) fits the volume it was integrated from better than the average perturbed pose
    assert float(norm[0]) <= float(norm[1:].mean())


@pytest.mark.parametrize("n,count_search,mode", [(10240, 30, "half"), (3072, 2000, "half"), (1024, 100000, "few"), (1024, 30, "none"), (2048, 1, "half")])
def test_cal_transform_matches_host_loop(cuda, rf_lib, n, count_search, mode):
    """rf_track_cal_transform vs the statement-by-statement restatement of the reference's Python loop
    (oracle/track_oracle.py: cal_transform; model/ROtracker.py:606-714)."""
    from oracle import track_oracle as TO
    from remixfusion_b200.tracker import ROSearch
    g = np.random.default_rng(n + count_search)
    cand = (g.random((n, 6)).astype(np.float32) * 2 - 1); cand[0] = 0
    ss = np.array([0.02, 0.03, 0.01, 0.01, 0.02, 0.015], np.float32)
    count = np.floor(g.random(n) * 700 + 1).astype(np.float32)
    fit = (0.2 + 0.1 * g.random(n)).astype(np.float32)
    if mode == "few":
        fit[1:] += 0.2; fit[g.integers(1, n, 7)] = 0.1
    if mode == "none":
        fit[0] = 0.05
    value = (fit * count).astype(np.float32)

    class MV:                                               # only the attributes ROSearch.__init__ reads
        tsdf_vol_gpu = torch.zeros(8, device=cuda)
    s = ROSearch(MV, 8, 8, 6.0, 0.06, 3.0)
    s.transform_candidate, s.search_size, s.count_search = cand, ss, count_search
    s._cand_dev = torch.from_numpy(cand).to(cuda)
    s._last = (torch.from_numpy(value).to(cuda), torch.from_numpy(count).to(cuda), n)
    ok, min_tsdf, mt = s.cal_transform()
    sv = (value / (count + np.float32(1e-6))).astype(np.float32)          # evaluate_tsdf :604
    ok_ref, min_ref, mt_ref = TO.cal_transform(sv, cand, ss, count_search)
    assert ok == ok_ref and (mode != "none" or not ok)
    assert abs(min_tsdf - min_ref) <= 1e-6 * abs(min_ref) + 1e-9
    np.testing.assert_allclose(mt, mt_ref, rtol=2e-6, atol=1e-9)


def _make_pst(seed, sizes=(2048, 1024, 1024), tables=7, zero_class=None):
    """Synthetic particle tables (the reference loads them from .tiff files, absent offline): candidate 0 is the identity
    (cal_transform compares everything with it), the others are Gaussian in the unit cube like the reference's PSTs."""
    g = np.random.default_rng(seed)
    out = []
    for n in sizes:
        a = np.clip(g.normal(0.0, 0.35, size=(tables, n, 6)), -0.99, 0.99).astype(np.float32) * np.float32(0.57)
        a[:, 0, :] = 0.0
        out.append(a)
    if zero_class is not None:
        out[zero_class][...] = 0.0          # every candidate of this class is the identity: no candidate beats candidate 0 -> failure branch
    return out


def _configured(scene, seed, zero_class=None):
    from remixfusion_b200.tracker import ROSearch
    s = ROSearch(scene["vol"], scene["H"], scene["W"], cut_dist=6.0, truncation=scene["cfg"]["volume"]["trunc"], sample_range=3.0)
    ro = dict(init_size=0.02, scaling_coefficient=0.09, particle_iter_lens=20, PST_size=[2048, 1024, 1024], fix_level_index=False,
              count_search=200, iterative_scale=True)
    s.configure_search(ro, _make_pst(seed, zero_class=zero_class))
    return s


@pytest.mark.parametrize("offset,zero_class", [(0.0, None), (0.015, None), (0.01, 1)])
def test_device_search_loop_matches_host_loop(scene, offset, zero_class):
    """rf_track_random_optimization (20 iterations on the device, one read-back) against the reference's loop restated in
    oracle/track_oracle.py driving the step methods (fitness + cal_transform kernels, policy in NumPy): same success sequence,
    pose and search sizes to float32 rounding (the policy scalars are float64 on the device, mixed float32 / float64 in NumPy).
    Two frames in a row with inherit=True exercise the state carried between frames; with the class-1 tables zeroed every second
    evaluation fails (no candidate beats the identity), which exercises the reset / failure branches of the policy."""
    from oracle import track_oracle as TO
    dev, host = _configured(scene, 11, zero_class), _configured(scene, 11, zero_class)
    pose0 = scene["c2w"].copy(); pose0[:3, 3] += np.float32(offset)                       # start off the true pose
    for frame in range(2):
        got = dev.random_optimization(frame, pose0, None, scene["depth"], scene["K"], beta=0.9, inherit=frame > 0, seed_num=77 + frame)
        want, flags = TO.random_optimization(host, frame, pose0, scene["depth"], scene["K"], beta=0.9, inherit=frame > 0, seed_num=77 + frame)
        mask = sum(1 << i for i, f in enumerate(flags) if f)
        assert dev.success_mask == mask, (frame, bin(dev.success_mask), bin(mask))
        assert any(flags) and (zero_class is None or not all(flags))
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(dev.search_size, host.search_size, rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(dev.previous_search_size, host.previous_search_size, rtol=1e-4, atol=1e-7)
        assert dev.previous_frame_success == host.previous_frame_success
        pose0 = got.copy()


def test_device_search_loop_argument_checks(scene):
    import ctypes as C
    from remixfusion_b200 import abi
    s = _configured(scene, 3)
    bad = (C.c_int * 20)(*([0] * 20))
    assert abi.lib().rf_track_random_optimization_scratch_floats(bad, bad, scene["H"], scene["W"]) == 0
    from remixfusion_b200.tracker import ROSearch
    s2 = ROSearch(scene["vol"], scene["H"], scene["W"], 6.0, 0.1, 3.0)
    with pytest.raises(abi.RfError, match="configure_search"):
        s2.random_optimization(0, scene["c2w"], None, scene["depth"], scene["K"])
    with pytest.raises(abi.RfError, match="multiple of 1024"):
        s2.configure_search(dict(init_size=0.02, scaling_coefficient=0.09, particle_iter_lens=20, PST_size=[1000, 1024, 1024],
                                 fix_level_index=False, count_search=200, iterative_scale=True), _make_pst(1, sizes=(1000, 1024, 1024)))
