"""Stage-1 parity (GPU): librf_b200 TSDF kernels vs the C oracle (oracle/tsdf_oracle.c) and vs the literal
reference kernels (oracle/_ref cubins built from /root/reference by oracle/build_ref.py), bit-exact.

Bar (BASELINE.json north_star): touched-voxel sets bit-exact; TSDF/weights within 1e-5 relative.  We hold the
stronger bar: every output array bit-identical (np.array_equal on the raw fp32 bits).
"""
import numpy as np
import pytest
import torch

from oracle import ref_kernels, tsdf_oracle as O
from remixfusion_b200 import synth
from remixfusion_b200.global_volume import MapVolume
from remixfusion_b200.volume import moving_volume
from tests import _common as T

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _cfg(voxel, lens, trunc=0.06, clamp=1.0):
    return {"volume": {"voxel_size": voxel, "trunc": trunc, "version": "center", "weight_clamp": clamp,
                       "x_config": {"fix": 0, "len": lens[0]}, "y_config": {"fix": 0, "len": lens[1]},
                       "z_config": {"fix": 0, "len": lens[2]}},
            "training": {"trunc": 0.05}}


class _Model:
    def __init__(self, R, dev):
        self.GBV = type("E", (), {})()
        self.GBW = type("E", (), {})()
        self.GBV.params = torch.zeros(4 * R ** 3 + 64, device=dev)     # padded: the reference's clean_tsdf writes voxel N
        self.GBW.params = torch.zeros(R ** 3 + 64, device=dev)


def _run_local(cuda, rf_lib, cam, lens, voxel, poses_frames, clamp=1.0, trunc=0.06, obs=1.0, with_ref=True):
    cfg = _cfg(voxel, lens, trunc, clamp)
    init = np.eye(4)
    mv = moving_volume(cfg, None, init, device=cuda)
    dims = mv.vol_dim
    n = int(np.prod(dims))
    tsdf = np.ones(n, np.float32); w = np.zeros(n, np.float32); col = np.zeros(n, np.float32)
    ref = None
    if with_ref:
        ref_kernels.require()                      # the literal reference kernel is part of the check, never skipped
        ref = [torch.ones(n + 64, device=cuda), torch.zeros(n + 64, device=cuda), torch.zeros(n + 64, device=cuda)]
    n_touched_total = 0
    for K, c2w, depth, rgb in poses_frames:
        rgb255 = np.floor(rgb * 255.0).astype(np.float32)            # model/ROtracker.py:82
        mv.integrate(rgb255, depth, K, c2w, None, obs_weight=obs)
        packed = O.pack_bgr(rgb255)
        nt, nb = O.integrate_local(tsdf, w, col, dims, mv.vol_origin, mv.voxel_size, K, c2w, depth, packed, trunc,
                                   obs_weight=obs, weight_clamp=int(clamp == 1.0), threads=8)
        cnt = mv.count_touched(torch.from_numpy(depth).to(cuda), K, c2w)
        assert cnt == (nt, nb)
        n_touched_total += nt
        if ref is not None:
            ref_kernels.ref_integrate_local(ref[0], ref[1], ref[2], dims, mv.vol_origin, mv.voxel_size, K, c2w,
                                            torch.from_numpy(depth).to(cuda), torch.from_numpy(packed).to(cuda),
                                            trunc, obs_weight=obs, weight_clamp=clamp)
    got = [mv.tsdf_vol_gpu.cpu().numpy(), mv.weight_vol_gpu.cpu().numpy(), mv.color_vol_gpu.cpu().numpy()]
    for g, o, name in zip(got, (tsdf, w, col), ("tsdf", "weight", "color")):
        assert np.array_equal(_bits(g), _bits(o)), f"{name}: product != C oracle ({(g != o).sum()} voxels)"
    if ref is not None:
        for g, r, name in zip(got, ref, ("tsdf", "weight", "color")):
            rr = r[:n].cpu().numpy()
            assert np.array_equal(_bits(g), _bits(rr)), f"{name}: product != reference kernel ({(g != rr).sum()} voxels)"
    assert n_touched_total > 0
    return n_touched_total


@pytest.mark.parametrize("far_plane", [False, True])
def test_local_small_multi_frame(cuda, rf_lib, monkeypatch, far_plane):
    """far_plane=True forces the far plane of the row clip (rf_tsdf_depth_max; normally only on volumes of >= 2^26 voxels):
    the sweep then stops behind the farthest surface, and every output must still equal the reference kernel's bit for bit."""
    from remixfusion_b200 import abi
    monkeypatch.setattr(abi, "FAR_PLANE_MIN_VOXELS", 0 if far_plane else 1 << 62)
    cam = T.small_cam(4)
    bound = [[-3, 3], [-3, 3], [-2, 2]]
    rng = np.random.default_rng(1)
    frames = []
    for i in range(4):
        c2w = T.random_pose(rng, [[-1, 1], [-1, 1], [-0.5, 0.5]])
        K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        depth, rgb = synth.render_frame(synth.make_scene(bound, 3), K, cam["H"], cam["W"], c2w, seed=i)
        frames.append((K, c2w, depth, rgb))
    _run_local(cuda, rf_lib, cam, (3, 3, 2), 0.05, frames)


@pytest.mark.parametrize("far_plane", [False, True])
def test_local_cfg1_256cube(cuda, rf_lib, monkeypatch, far_plane):
    """BASELINE config 1: one 640x480 frame into a 256^3 volume (voxel 6/256)."""
    from remixfusion_b200 import abi
    monkeypatch.setattr(abi, "FAR_PLANE_MIN_VOXELS", 0 if far_plane else 1 << 62)
    cam = synth.CFG1_CAM
    K, c2w, depth, rgb = T.frame(cam, [[-3, 3], [-3, 3], [-3, 3]], [0.2, 0.1, 0.3], [2.5, 1.0, 0.2])
    nt = _run_local(cuda, rf_lib, cam, (3, 3, 3), 6.0 / 256, [(K, c2w, depth, rgb)])
    assert nt > 100000


def test_local_fp32_decode_quirk(cuda, rf_lib):
    """> 2^24 voxels: the reference's fp32 index decode mis-places slab-tail voxels — reproduced."""
    cam = T.small_cam(2)
    bound = [[-4, 4], [-4, 4], [-3, 3]]
    frames = []
    rng = np.random.default_rng(5)
    for i in range(2):
        c2w = T.random_pose(rng, [[-1, 1], [-1, 1], [-0.5, 0.5]])
        K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
        depth, rgb = synth.render_frame(synth.make_scene(bound, 2), K, cam["H"], cam["W"], c2w, seed=i)
        frames.append((K, c2w, depth, rgb))
    _run_local(cuda, rf_lib, cam, (4, 4, 3), 0.02, frames)       # 400 x 400 x 300 = 48 M voxels (cfg 2 local volume)


def test_local_no_clamp_and_deintegrate(cuda, rf_lib):
    cam = T.small_cam(4)
    K, c2w, depth, rgb = T.frame(cam, [[-3, 3], [-3, 3], [-2, 2]], [0.0, 0.2, 0.1], [2.0, -1.0, 0.0])
    _run_local(cuda, rf_lib, cam, (3, 3, 2), 0.05, [(K, c2w, depth, rgb)] * 3, clamp=0.0)


def _special_frame(cam):
    """Axis-aligned camera at the origin, piecewise-constant depth with holes, colour with exactly black regions."""
    H, W = cam["H"], cam["W"]
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    c2w = np.eye(4)
    depth = np.full((H, W), 1.5, np.float32)
    depth[:, W // 2:] = 1.0
    depth[: H // 8] = 0.0
    rgb = np.zeros((H, W, 3), np.float32)                      # left third black: every colour numerator is exactly 0
    rgb[:, W // 3: 2 * W // 3] = [0.0, 0.5, 1.0]               # one channel 0
    rgb[:, 2 * W // 3:] = np.random.default_rng(3).random((H, W - 2 * W // 3, 3)).astype(np.float32)
    return K, c2w, depth, rgb


def test_division_special_operands(cuda, rf_lib):
    """The kernels take their IEEE quotients from a shared reciprocal when every operand lies within 2^+-40 and fall back to
    div.rn otherwise (csrc/tsdf_integrate.cu: div_fast / div_ok).  This frame drives the fall-backs: voxel coordinates that are
    exact binary fractions put X = 0 and Y = 0 exactly on two planes of voxels, black pixels make the colour numerators exactly
    0 on the first and on the repeated integrations; results must still equal the reference kernel's bit for bit."""
    cam = T.small_cam(4)
    fr = _special_frame(cam)
    nt = _run_local(cuda, rf_lib, cam, (2, 2, 2), 0.0625, [fr] * 3, trunc=0.125)
    assert nt > 5000
    n = _run_global(cuda, rf_lib, cam, 64, [[-2.0, 2.0], [-2.0, 2.0], [-2.0, 2.0]], [fr] * 3, trunc=0.125)
    assert n > 1000


def _run_global(cuda, rf_lib, cam, R, bound, frames, trunc=0.1, obs=1.0):
    cfg = {"globalV": {"base_resolution": R}, "mapping": {"bound": bound}, "training": {"c_trunc": trunc}}
    model = _Model(R, cuda)
    K0 = frames[0][0]
    mvol = MapVolume(cfg, model, K0)
    mvol.init_mapvolume()
    trgb = np.zeros(4 * R ** 3, np.float32); O.clear_global(trgb); gw = np.zeros(R ** 3, np.float32)
    box = [v for ax in bound for v in ax]
    ref_kernels.require()
    ref = _Model(R, cuda)
    ref_kernels.ref_clear_global(ref.GBV.params, R)
    for K, c2w, depth, rgb in frames:
        batch = {"rgb": torch.from_numpy(rgb)[None], "depth": torch.from_numpy(depth)[None]}
        pose = torch.from_numpy(c2w).float().to(cuda)
        cnt = mvol.count_touched(batch["depth"], pose, obs)
        mvol.integrate_kf(batch, pose, obs)
        nt = O.integrate_global(trgb, gw, R, box, K, c2w, depth, rgb, trunc, obs_weight=obs, threads=8)
        assert cnt == nt
        if ref is not None:
            ref_kernels.ref_integrate_global(ref.GBV.params, ref.GBW.params, R, box, K, c2w.astype(np.float32),
                                             torch.from_numpy(depth).to(cuda), torch.from_numpy(rgb).to(cuda), trunc, obs)
    g_t = model.GBV.params[:4 * R ** 3].cpu().numpy(); g_w = model.GBW.params[:R ** 3].cpu().numpy()
    assert np.array_equal(_bits(g_t), _bits(trgb)), f"GBV: product != C oracle ({(g_t != trgb).sum()})"
    assert np.array_equal(_bits(g_w), _bits(gw)), "GBW: product != C oracle"
    if ref is not None:
        r_t = ref.GBV.params[:4 * R ** 3].cpu().numpy(); r_w = ref.GBW.params[:R ** 3].cpu().numpy()
        assert np.array_equal(_bits(g_t), _bits(r_t)), f"GBV: product != reference kernel ({(g_t != r_t).sum()})"
        assert np.array_equal(_bits(g_w), _bits(r_w)), "GBW: product != reference kernel"
    assert (gw > 0).sum() > 0
    return int((gw > 0).sum())


def test_global_replica_200(cuda, rf_lib):
    """GBV R=200 over the Replica room0 bound with the 1200x680 camera (cfg 2), 3 keyframes."""
    cam = synth.REPLICA_CAM
    bound = synth.REPLICA_BOUND
    scene = synth.make_scene(bound, 0)
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    frames = []
    for i, c2w in enumerate(synth.loop_trajectory(scene, 3)):
        depth, rgb = synth.render_frame(scene, K, cam["H"], cam["W"], c2w, seed=i)
        frames.append((K, c2w, depth, rgb))
    n = _run_global(cuda, rf_lib, cam, 200, bound, frames)
    assert n > 100000


@pytest.mark.parametrize("far_plane", [False, True])
def test_global_small_random_poses(cuda, rf_lib, monkeypatch, far_plane):
    from remixfusion_b200 import abi
    monkeypatch.setattr(abi, "FAR_PLANE_MIN_VOXELS", 0 if far_plane else 1 << 62)
    _global_small_random_poses(cuda, rf_lib)


def _global_small_random_poses(cuda, rf_lib):
    cam = T.small_cam(4)
    bound = [[-2.0, 2.5], [-1.5, 2.0], [-1.0, 1.7]]
    rng = np.random.default_rng(11)
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    scene = synth.make_scene(bound, 4)
    frames = []
    for i in range(6):
        c2w = T.random_pose(rng, bound)
        depth, rgb = synth.render_frame(scene, K, cam["H"], cam["W"], c2w, seed=i)
        frames.append((K, c2w, depth, rgb))
    _run_global(cuda, rf_lib, cam, 96, bound, frames)


def test_global_deintegrate(cuda, rf_lib):
    cam = T.small_cam(4)
    bound = [[-2.0, 2.5], [-1.5, 2.0], [-1.0, 1.7]]
    K, c2w, depth, rgb = T.frame(cam, bound, [0.1, 0.2, 0.3], [2.0, 1.0, 0.0])
    cfg = {"globalV": {"base_resolution": 64}, "mapping": {"bound": bound}, "training": {"c_trunc": 0.1}}
    model = _Model(64, cuda)
    mvol = MapVolume(cfg, model, K)
    mvol.init_mapvolume()
    batch = {"rgb": torch.from_numpy(rgb), "depth": torch.from_numpy(depth)}
    mvol.integrate_kf(batch, torch.from_numpy(c2w).float(), 1.0)
    assert (model.GBW.params > 0).any()
    mvol.integrate_kf(batch, torch.from_numpy(c2w).float(), -1.0)        # mp_slam/mapper.py:126-133
    assert float(model.GBW.params.abs().sum()) == 0.0
    t = model.GBV.params[:4 * 64 ** 3].view(-1, 4)
    assert bool((t[:, 0] == 1).all()) and float(t[:, 1:].abs().sum()) == 0.0


def test_slab_sharding_is_bit_identical(cuda, rf_lib):
    """z-slab / x-slab sharding (multi-GPU partitioning): union of slabs == whole volume, bitwise."""
    cam = T.small_cam(4)
    bound = [[-2.0, 2.5], [-1.5, 2.0], [-1.0, 1.7]]
    K, c2w, depth, rgb = T.frame(cam, bound, [0.1, 0.2, 0.3], [2.0, 1.0, 0.0])
    R = 64
    cfg = {"globalV": {"base_resolution": R}, "mapping": {"bound": bound}, "training": {"c_trunc": 0.1}}
    whole = _Model(R, cuda)
    mv = MapVolume(cfg, whole, K); mv.init_mapvolume()
    batch = {"rgb": torch.from_numpy(rgb), "depth": torch.from_numpy(depth)}
    pose = torch.from_numpy(c2w).float()
    mv.integrate_kf(batch, pose)
    parts_t, parts_w = [], []
    for (z0, z1) in [(0, 20), (20, 41), (41, 64)]:
        m = type("M", (), {})()
        m.GBV = type("E", (), {})(); m.GBW = type("E", (), {})()
        m.GBV.params = torch.zeros(4 * (z1 - z0) * R * R, device=cuda)
        m.GBW.params = torch.zeros((z1 - z0) * R * R, device=cuda)
        s = MapVolume(cfg, m, K, z_slab=(z0, z1)); s.init_mapvolume(); s.integrate_kf(batch, pose)
        parts_t.append(m.GBV.params); parts_w.append(m.GBW.params)
    assert torch.equal(torch.cat(parts_t), whole.GBV.params[:4 * R ** 3])
    assert torch.equal(torch.cat(parts_w), whole.GBW.params[:R ** 3])
    # local volume, x-slabs
    lcfg = _cfg(0.05, (3, 3, 2))
    full = moving_volume(lcfg, None, np.eye(4), device=cuda)
    rgb255 = np.floor(rgb * 255).astype(np.float32)
    full.integrate(rgb255, depth, K, c2w, None)
    dx = int(full.vol_dim[0])
    outs = []
    for (x0, x1) in [(0, 50), (50, 51), (51, dx)]:
        p = moving_volume(lcfg, None, np.eye(4), device=cuda, x_slab=(x0, x1))
        p.integrate(rgb255, depth, K, c2w, None)
        outs.append((p.tsdf_vol_gpu, p.weight_vol_gpu, p.color_vol_gpu))
    for i, whole_arr in enumerate((full.tsdf_vol_gpu, full.weight_vol_gpu, full.color_vol_gpu)):
        assert torch.equal(torch.cat([o[i] for o in outs]), whole_arr)


def test_abi_errors(cuda, rf_lib):
    import ctypes as C
    rc = rf_lib.rf_tsdf_clear_global(C.c_void_p(0), C.c_int64(8), C.c_void_p(0))
    assert rc == -1 and b"NULL" in rf_lib.rf_last_error()
    t = torch.zeros(16, device=cuda)
    rc = rf_lib.rf_tsdf_clear_global(C.c_void_p(t.data_ptr() + 4), C.c_int64(1), C.c_void_p(0))
    assert rc == -3


def test_pixel_lambda_image_is_bit_identical_to_inline(cuda, rf_lib, monkeypatch):
    """The hoisted per-pixel 1/lambda image (rf_tsdf_pixel_lambda) and the inline per-voxel evaluation (NULL image) must
    give the same bits: same operations, same order (model/Volume.py:280-283)."""
    from remixfusion_b200 import abi
    cam = T.small_cam(2)
    K, c2w, depth, rgb = T.frame(cam, [[-3, 3], [-3, 3], [-2, 2]], eye=(0.2, -0.3, 0.1), target=(2.0, 1.0, 0.0), seed=5)
    rgb255 = np.floor(rgb * 255.0).astype(np.float32)
    outs = []
    for use_image in (True, False):
        if not use_image:
            monkeypatch.setattr(abi, "pixel_lambda", lambda *a, **k: None)
        mv = moving_volume(_cfg(0.04, (3, 3, 2)), None, np.eye(4), device=cuda)
        mv.integrate(rgb255, depth, K, c2w, None)
        m = _Model(48, cuda)
        gv = MapVolume({"globalV": {"base_resolution": 48}, "mapping": {"bound": [[-3, 3], [-3, 3], [-2, 2]]},
                        "training": {"c_trunc": 0.1}}, m, K)
        gv.init_mapvolume()
        gv.integrate_kf({"rgb": torch.from_numpy(rgb), "depth": torch.from_numpy(depth)}, torch.from_numpy(c2w).float(), 1.0)
        outs.append([t.cpu().numpy() for t in (mv.tsdf_vol_gpu, mv.weight_vol_gpu, mv.color_vol_gpu, m.GBV.params, m.GBW.params)])
    assert float(np.abs(outs[0][1]).sum()) > 0 and float(np.abs(outs[0][4]).sum()) > 0
    for a, b in zip(*outs):
        assert np.array_equal(_bits(a), _bits(b))


@pytest.mark.parametrize("lens,voxel,move", [((1, 1, 1), 0.04, (1, 0, 0)), ((1, 1, 1), 0.04, (0, -1, 1)), ((1, 1, 1), 0.04, (3, 0, 0)),
                                              ((4, 4, 3), 0.02, (1, -1, 0))])
def test_recenter_matches_oracle_and_reference_kernel(cuda, rf_lib, lens, voxel, move):
    """N2: moving_volume.update_tsdf_swap_rot_trans vs the C oracle vs the literal reference kernel (swap_rot_trans of the
    extracted cubin), bit for bit; the last case is the 400 x 400 x 300 volume (48 M voxels: fp32 index-decode quirk)."""
    mv = moving_volume(_cfg(voxel, lens), None, np.eye(4), device=cuda)
    n = int(np.prod(mv.vol_dim))
    g = torch.Generator(device="cuda").manual_seed(1)
    with torch.no_grad():
        mv.tsdf_vol_gpu.copy_(torch.rand(n, device=cuda, generator=g) * 2 - 1)
        mv.weight_vol_gpu.copy_(torch.floor(torch.rand(n, device=cuda, generator=g) * 40))
        mv.color_vol_gpu.copy_(torch.floor(torch.rand(n, device=cuda, generator=g) * 16777215))
    old = [t.clone() for t in (mv.tsdf_vol_gpu, mv.weight_vol_gpu, mv.color_vol_gpu)]
    old_bnds = mv.vol_bnds.copy(); old_dim = mv.vol_dim.copy(); old_origin = mv.vol_origin.copy()
    new_bnds = old_bnds + np.asarray(move, dtype=np.float64)[:, None]
    mv.copy_volume()
    mv.update_tsdf_swap_rot_trans(new_bnds.copy(), old_bnds.copy())
    got = [t.cpu().numpy() for t in (mv.tsdf_vol_gpu, mv.weight_vol_gpu, mv.color_vol_gpu)]
    exp = O.recenter(*[t.cpu().numpy() for t in old], old_dim, old_origin, mv.vol_dim, mv.vol_origin, mv.voxel_size, threads=8)
    for a, b, name in zip(got, exp, ("tsdf", "weight", "color")):
        assert np.array_equal(_bits(a), _bits(b)), f"{name}: product != C oracle ({(a != b).sum()} voxels)"
    kept = float((got[1] != 0).mean())
    assert (kept > 0) == all(abs(m) < 2 * l for m, l in zip(move, lens))
    ref_kernels.require()
    if True:
        new = [torch.full((n + 64,), 7.0, device=cuda) for _ in range(3)]
        oldp = [torch.cat([t, t.new_zeros(64)]) for t in old]
        ref_kernels.ref_recenter(new, oldp, mv.vol_dim, mv.vol_origin, old_dim, old_origin, mv.voxel_size)
        for a, r, name in zip(got, new, ("tsdf", "weight", "color")):
            rr = r[:n].cpu().numpy()
            assert np.array_equal(_bits(a), _bits(rr)), f"{name}: product != reference kernel ({(a != rr).sum()} voxels)"
    # a second move re-uses the ping-pong arrays; moving back restores the surviving region
    mv.update_tsdf_swap_rot_trans(old_bnds.copy(), new_bnds.copy())
    back = mv.weight_vol_gpu.cpu().numpy()
    assert np.array_equal(back[back != 0], old[1].cpu().numpy()[back != 0])


def test_back_to_back_integrate_behind_a_busy_stream(cuda, rf_lib):
    """moving_volume.integrate() stages host frames through pinned buffers with asynchronous copies.  Six different frames
    are queued back to back while the stream is still busy with earlier work (twice the number of staging slots, so slots
    are reused while their copies may still be in flight): the result must equal integrating them one at a time with a
    synchronisation in between (the reference's copies were synchronous, model/Volume.py:733-749)."""
    cam = dict(H=120, W=160, fx=131.0, fy=131.0, cx=79.5, cy=59.5)
    K = synth.intrinsics(cam["fx"], cam["fy"], cam["cx"], cam["cy"])
    scene = synth.make_scene([[-3, 3], [-3, 3], [-2, 2]], 1)
    poses = synth.loop_trajectory(scene, 12)
    frames = []
    for i in range(6):
        d, c = synth.render_frame(scene, K, cam["H"], cam["W"], poses[i], seed=i)
        frames.append((poses[i], d * (1.0 + 0.05 * i), np.floor(c * 255.0).astype(np.float32)))
    cfg = _cfg(0.05, (3, 3, 2))
    a = moving_volume(cfg, None, np.eye(4), device=cuda)
    b = moving_volume(cfg, None, np.eye(4), device=cuda)
    for c2w, d, c in frames:                                   # one at a time
        b.integrate(c, d, K, c2w, None)
        torch.cuda.synchronize()
    torch.cuda._sleep(int(4e8))                                # ~0.2 s of queued work ahead of the first copy
    for c2w, d, c in frames:                                   # back to back, nothing waits on the host side by itself
        a.integrate(c, d, K, c2w, None)
    torch.cuda.synchronize()
    for x, y, name in ((a.tsdf_vol_gpu, b.tsdf_vol_gpu, "tsdf"), (a.weight_vol_gpu, b.weight_vol_gpu, "weight"), (a.color_vol_gpu, b.color_vol_gpu, "color")):
        assert torch.equal(x, y), f"{name}: back-to-back integrate() differs from the synchronised sequence"
    assert float(a.weight_vol_gpu.sum()) > 0
