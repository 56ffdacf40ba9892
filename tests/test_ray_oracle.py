"""CPU suite: the Stage-2 restatement (oracle/ray_oracle.py) against golden vectors produced by the reference's
own model/scene_rep.py + model/decoder.py + model/utils.py (tests/golden/ray_golden.npz, generated in the build
container by tests/golden/make_ray_golden.py), forward and backward; plus known-answer tests for the
tiny-cuda-nn stand-in (SURVEY.md §8c "self-consistency KATs")."""
import numpy as np
import pytest
import torch

from oracle import ref_import, tcnn_standin
from tests import _ray_common as R

G = np.load(R.GOLDEN)
TOL = dict(rtol=2e-5, atol=1e-7)


@pytest.mark.parametrize("name,clamp,ray_grads", [("A", False, False), ("B", True, True)])
def test_oracle_training_matches_reference_golden(name, clamp, ray_grads):
    cfg, orc = R.oracle_from_golden(G, name)
    ro, rd, tc, td = R.inputs_from_golden(G, ray_grads)
    ret = orc.mapping(ro, rd, tc, td, clamp=clamp, u=torch.from_numpy(G[f"{name}_u"]))
    for k in ("rgb_res_loss", "depth_res_loss", "sdf_res_loss", "fs_res_loss", "rgb_res", "depth_res"):
        np.testing.assert_allclose(ret[k].detach().numpy(), G[f"{name}_{k}"], err_msg=k, **TOL)
    loss = orc.total_loss(ret)
    loss.backward()
    np.testing.assert_allclose(loss.item(), G[f"{name}_loss"], **TOL)
    got = {"g_hash": orc.embed_res_fn.params.grad, "g_w_sdf0": orc.w_sdf0.grad, "g_w_sdf1": orc.w_sdf1.grad,
           "g_w_col0": orc.w_col0.grad, "g_w_col1": orc.w_col1.grad}
    if ray_grads:
        got.update(g_rays_o=ro.grad, g_rays_d=rd.grad)
    for k, v in got.items():
        ref = G[f"{name}_{k}"]
        scale = np.abs(ref).max()
        np.testing.assert_allclose(v.numpy(), ref, rtol=1e-4, atol=1e-5 * scale, err_msg=k)


def test_oracle_eval_matches_reference_golden():
    cfg, orc = R.oracle_from_golden(G, "C", requires_grad=False)
    orc.training = False
    ro, rd, tc, td = R.inputs_from_golden(G)
    with torch.no_grad():
        ret = orc.mapping(ro, rd, tc, td, clamp=False, u=None)
    for k in ("rgb_res_map", "depth_res_map", "z_vals", "raw"):
        np.testing.assert_allclose(ret[k].numpy(), G[f"C_{k}"], err_msg=k, **TOL)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_reference_code_live():
    """Same comparison against the reference's code run live (second seed, ScanNet-style 21+96 sampling)."""
    cfg = R.base_config(hash_size=8, R=16)
    cfg["training"].update(n_range_d=21, n_samples_d=96, rgb_missing=0.0)
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
    ref = ref_import.make_reference_model(cfg, bb)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        ref.embed_res_fn.params.copy_((torch.rand(ref.embed_res_fn.params.shape, generator=g) - 0.5) * 0.2)
        ref.GBV.params.copy_(torch.rand(ref.GBV.params.shape, generator=g) * 2 - 1)
    n = 40
    ro = torch.tensor([3.0, 1.2, 0.0]).repeat(n, 1) + 0.1 * torch.randn(n, 3, generator=g)
    rd = torch.randn(n, 3, generator=g); rd = rd / rd.norm(dim=1, keepdim=True)
    td = torch.rand(n, 1, generator=g) * 3; td[::7] = 0.0
    tc = torch.rand(n, 3, generator=g)
    ref.train()
    torch.manual_seed(3); u = torch.rand(n, 117); torch.manual_seed(3)
    r_ref = ref.mapping(ro, rd, tc, td)
    h = R.hash_standin(cfg); gb = R.gbv_standin(cfg)
    with torch.no_grad():
        h.params.copy_(ref.embed_res_fn.params); gb.params.copy_(ref.GBV.params)
    from oracle.ray_oracle import RayOracle
    orc = RayOracle(cfg, bb, h, gb, ref.decoder_res.sdf_net.model[0].weight.detach(), ref.decoder_res.sdf_net.model[2].weight.detach(),
                    ref.decoder_res.color_net.model[0].weight.detach(), ref.decoder_res.color_net.model[2].weight.detach())
    r_orc = orc.mapping(ro, rd, tc, td, u=u)
    for k in ("rgb_res_loss", "depth_res_loss", "sdf_res_loss", "fs_res_loss", "rgb_res", "depth_res"):
        np.testing.assert_allclose(r_orc[k].detach().numpy(), r_ref[k].detach().numpy(), err_msg=k, **TOL)


# ---- known-answer tests for the tiny-cuda-nn stand-in (parity unpinned upstream) ----
def test_level_table_matches_survey_cfg3():
    pls = np.exp2(np.log2(512 / 16) / 15)
    s, res, size, off = tcnn_standin.grid_levels(16, 2, True, 19, 16, pls)
    assert res == [16, 21, 26, 33, 41, 51, 65, 81, 102, 129, 162, 204, 257, 323, 407, 513]
    assert off[-1] == 5261688
    assert all(size[l] == (res[l] ** 3 + 7) // 8 * 8 for l in range(7)) and all(sz == 2 ** 19 for sz in size[7:])


def test_hash_index_kats():
    # hashed level: index of cell (0,0,0) -> 0, (1,0,0) -> 1, (0,1,0) -> 2654435761 % T
    T = 2 ** 12
    scale, res = np.float32(300.0), 301
    def idx_of(cell):
        x = torch.tensor([[(c + 0.25 - 0.5) / float(scale) for c in cell]], dtype=torch.float32)
        i, w = tcnn_standin.grid_indices(x, scale, res, T, True)
        return int(i[0, 0])
    assert idx_of((0, 0, 0)) == 0 and idx_of((1, 0, 0)) == 1
    assert idx_of((0, 1, 0)) == 2654435761 % T and idx_of((0, 0, 1)) == 805459861 % T
    assert idx_of((3, 5, 7)) == ((3 * 1) ^ ((5 * 2654435761) & 0xFFFFFFFF) ^ ((7 * 805459861) & 0xFFFFFFFF)) % T


def test_dense_grid_hits_vertices_and_constant_volume():
    Rr = 8
    g = tcnn_standin.GridStandIn(1, 4, False, 0, Rr, 1)
    with torch.no_grad():
        g.params.copy_(torch.arange(g.params.numel(), dtype=torch.float32))
    v = np.array([3, 5, 2])
    x = torch.tensor((v - 0.5) / (Rr - 1), dtype=torch.float32)[None]        # vertex v sampled at (v-0.5)/(R-1)
    out = g(x)[0]
    lin = v[0] + v[1] * Rr + v[2] * Rr * Rr
    np.testing.assert_allclose(out.detach().numpy(), np.arange(4) + 4 * lin, rtol=2e-6)
    with torch.no_grad():
        g.params.fill_(0.37)
    np.testing.assert_allclose(g(torch.rand(50, 3) * 0.8 + 0.1).detach().numpy(), 0.37, rtol=1e-5)


def test_oneblob_rows_sum_to_one_and_wraps():
    ob = tcnn_standin.OneBlobStandIn(16)
    x = torch.rand(100, 3)
    o = ob(x).reshape(100, 3, 16)
    np.testing.assert_allclose(o.sum(-1).numpy(), 1.0, atol=1e-5)
    # wrap-around: the three shifted kernels make x = 0 and x = 1 the same point
    np.testing.assert_allclose(ob(torch.zeros(1, 3)).numpy(), ob(torch.ones(1, 3)).numpy(), atol=2e-6)
    assert (o >= -1e-6).all()


def test_level_scale_rounding_sensitivity_on_the_bench_table():
    """R6 is 'parity unpinned': the stand-in derives each level's scale with a correctly rounded exp2, upstream tiny-cuda-nn with
    the device's exp2f (<= 2 ulp).  This QUANTIFIES what such a difference can do on the bench table (Replica: 16 levels, base 16,
    2 cm finest cells, T = 2^16): with every level's scale moved by +-2 ulp, the fraction of (position, level) pairs whose cell —
    and therefore hash index set — changes, and the largest change of an interpolated feature, on 2 * 10^5 seeded positions."""
    import numpy as np
    import torch
    from oracle.tcnn_standin import GridStandIn
    from remixfusion_b200 import configs
    cfg = configs.replica()
    from oracle.ray_oracle import hash_standin
    ref = hash_standin(cfg)
    x = torch.rand(200000, 3, generator=torch.Generator().manual_seed(5))
    base_idx = ref.level_indices(x)[:, :, 0]
    with torch.no_grad():
        base_feat = ref(x)
    worst_flip, worst_feat = 0.0, 0.0
    for ulps in (-2, 2):
        moved = hash_standin(cfg)
        moved.params.data.copy_(ref.params.data)
        # level 0 is exact in both (exp2(0) = 1: scale 15.0, the one level where an ulp would change the resolution); the others move
        moved.scale = [s if l == 0 else np.nextafter(np.nextafter(s, np.float32(np.inf * ulps)), np.float32(np.inf * ulps))
                       for l, s in enumerate(ref.scale)]
        assert all(int(np.ceil(s)) + 1 == r for s, r in zip(moved.scale, ref.res)), "a 2-ulp scale change must not change a level's resolution"
        idx = moved.level_indices(x)[:, :, 0]
        with torch.no_grad():
            feat = moved(x)
        worst_flip = max(worst_flip, float((idx != base_idx).float().mean()))
        worst_feat = max(worst_feat, float((feat - base_feat).abs().max() / base_feat.abs().max()))
    # observed: 4e-5 of the (position, level) pairs land in the neighbouring cell (positions within 2 ulp of a cell face; the
    # interpolant is continuous there).  The feature change comes from the fractional position itself: 2 ulp of a scale of 399 move it
    # by 1e-4 cells, i.e. 2.6e-4 of the feature scale on this UNCORRELATED random table (a trained table is smooth across a cell) —
    # inside the 1e-3 bar of the path, and the reason R6 is reported as unpinned rather than as exact
    print(f"2-ulp level scales: cell flips {worst_flip:.2e} of (position, level) pairs, feature change {worst_feat:.2e} of the feature scale")
    assert worst_flip < 1e-4, worst_flip
    assert worst_feat < 1e-3, worst_feat


def test_fast_sdf_weight_formula_error_bound():
    """csrc/ray_query.cu: sdf_weight computes sigmoid(a) sigmoid(-a) (model/scene_rep.py:116) as t / (1 + t)^2 with t = exp(-|a|) from
    ex2.approx (2 ulp, argument rounded once) and rcp.approx (1 ulp).  A float32 model of that evaluation with the approximations'
    worst-case errors injected stays within the bound the kernel's comment states (8e-7 + 6e-8 |a|, relative), against the float64
    value of the reference formula."""
    import numpy as np
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.uniform(-30, 30, 400000), rng.normal(0, 2, 400000)]).astype(np.float32)
    a64 = a.astype(np.float64)
    truth = 1.0 / (1.0 + np.exp(-a64)) / (1.0 + np.exp(a64))
    sig_truth = 1.0 / (1.0 + np.exp(-a64))
    bound = 8e-7 + 6e-8 * np.abs(a64)
    for d_t in (-2, 0, 2):
        for d_r in (-1, 0, 1):
            arg = (-np.abs(a) * np.float32(1.4426950408889634)).astype(np.float32)          # x * log2(e), rounded once
            t = np.exp2(arg.astype(np.float64)).astype(np.float32)
            t = (t.view(np.int32) + d_t).view(np.float32)                                   # ex2.approx: 2 ulp
            inv = (np.float32(1.0) / (np.float32(1.0) + t)).astype(np.float32)
            inv = (inv.view(np.int32) + d_r).view(np.float32)                               # rcp.approx: 1 ulp
            ti = (t * inv).astype(np.float32)
            e = (ti * inv).astype(np.float32)
            sg = np.where(a >= 0, inv, ti)
            rel_e = np.abs(e.astype(np.float64) - truth) / truth
            rel_s = np.abs(sg.astype(np.float64) - sig_truth) / sig_truth
            assert np.all(rel_e <= bound), (d_t, d_r, float(np.max(rel_e - bound)))
            assert np.all(rel_s <= bound), (d_t, d_r, float(np.max(rel_s - bound)))
