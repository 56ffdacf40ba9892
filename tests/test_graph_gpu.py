"""The CUDA-graph-captured mapping iteration (remixfusion_b200.graph) must do what the eager iteration does: same losses
at every step and, after several optimisation steps, the same parameters.  Two eager runs do not agree bit for bit either
(fp32 reduction order in the atomics), and Adam with eps = 1e-15 turns a near-zero gradient into a step of +-lr whose sign
that noise can flip: parameters are compared with 1e-3 relative + 2e-5 absolute (0.2 % of one lr = 1e-2 step)."""
import numpy as np
import pytest
import torch

from remixfusion_b200 import configs
from remixfusion_b200.graph import GraphedMappingStep
from remixfusion_b200.optim import Adam
from remixfusion_b200.scene_rep import JointEncoding

pytestmark = pytest.mark.gpu


def _make(cuda, capturable):
    cfg = configs.replica(hash_size=12)
    cfg["training"]["perturb"] = 0                       # deterministic sampling: both runs see the same z_vals
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
    torch.manual_seed(3)
    m = JointEncoding(cfg, bb).to(cuda); m.train()
    with torch.no_grad():
        m.GBV.params.copy_((torch.rand_like(m.GBV.params) * 2 - 1) * 0.5)
    groups = [{"params": list(m.decoder_res.parameters()), "weight_decay": 1e-6, "lr": 1e-2},
              {"params": list(m.embed_res_fn.parameters()), "eps": 1e-15, "lr": 1e-2}]
    return cfg, m, Adam(groups, betas=(0.9, 0.99), capturable=capturable)


def test_graphed_iteration_matches_eager(cuda, rf_lib):
    n = 2048
    cfg_e, m_e, opt_e = _make(cuda, False)
    cfg_g, m_g, opt_g = _make(cuda, True)
    step = GraphedMappingStep(m_g, opt_g, n, lambda r: configs.total_loss(cfg_g, r), eager_steps=2)
    g = torch.Generator().manual_seed(0)
    b = torch.tensor(cfg_e["mapping"]["bound"])
    for it in range(8):
        ro = (b[:, 0] + (0.3 + 0.4 * torch.rand(n, 3, generator=g)) * (b[:, 1] - b[:, 0])).to(cuda)
        rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(cuda)
        tc = torch.rand(n, 3, generator=g).to(cuda); td = (0.3 + 2.5 * torch.rand(n, 1, generator=g)).to(cuda)
        ret = m_e.mapping(ro, rd, tc, td)
        loss_e = configs.total_loss(cfg_e, ret)
        loss_e.backward()
        opt_e.step(zero_grad=True)
        loss_g, _ = step(ro, rd, tc, td)
        assert abs(float(loss_g) - float(loss_e)) <= 1e-4 * abs(float(loss_e)), (it, float(loss_g), float(loss_e))
    assert step.graph is not None
    for pe, pg in zip(m_e.parameters(), m_g.parameters()):
        if pe.requires_grad:
            # Gradient noise is 1e-7 relative in BOTH paths (profiles/determinism_probe.py); for the few table entries whose
            # gradient nearly cancels, Adam with eps = 1e-15 normalises that noise into a visible fraction of one lr = 1e-2
            # step.  Which entries those are depends on the seed and on the kernels' rounding, so a handful of outliers
            # (< 0.1 % of the elements, each below one lr step) is part of the contract; everything else must agree.
            bad = (pg - pe).abs() > 2e-5 + 1e-3 * pe.abs()
            n_bad = int(bad.sum())
            assert n_bad <= max(1, pe.numel() // 1000), (n_bad, pe.numel())
            if n_bad:
                assert float((pg - pe).abs().max()) < 1e-2, float((pg - pe).abs().max())


def test_smoothness_matches_oracle_and_is_capturable(cuda, rf_lib):
    """The feature-grid smoothness term (mp_slam/slam.py:193-217): value and hash-table gradient vs the stand-in encoder on the
    CPU with the same two random draws; then an iteration whose loss includes it runs inside the captured graph (device-side
    draws) and keeps reducing the same kind of loss as the eager loop."""
    from oracle.ray_oracle import hash_standin
    from remixfusion_b200.losses import Smoothness, make_loss_fn
    cfg, m, opt = _make(cuda, True)
    cfg["training"].update(smooth_weight=0.001, smooth_pts=24, smooth_vox=0.1, smooth_margin=0.05)
    sm = Smoothness(m, 24, 0.1, 0.05)
    g = torch.Generator().manual_seed(5)
    r3 = torch.rand(3, generator=g); r1 = torch.rand((1, 1, 1, 3), generator=g)
    with torch.no_grad():
        m.embed_res_fn.params.copy_((torch.rand_like(m.embed_res_fn.params) * 2 - 1) * 0.05)
    m.embed_res_fn.params.grad = None
    val = sm(r3.to(cuda), r1.to(cuda))
    val.backward()
    val = float(val)            # drop the autograd graph: it holds the table's gradient accumulator, created on the default stream,
                                # and a capture on a side stream must not be made to synchronise with that stream
    # CPU restatement with the stand-in hash grid
    h = hash_standin(cfg)
    with torch.no_grad():
        h.params.copy_(m.embed_res_fn.params.cpu())
    bb = torch.tensor(cfg["mapping"]["bound"], dtype=torch.float64)
    vol = bb[:, 1] - bb[:, 0]
    off = r3.to(vol) * (vol - 23 * 0.1 - 2 * 0.05) + 0.05
    ax = torch.arange(0, 23)
    coords = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).float()
    pts = (coords.to(vol) + r1.to(vol)) * 0.1 + bb[:, 0] + off
    f = h(((pts - bb[:, 0]) / vol).reshape(-1, 3).float()).reshape(23, 23, 23, -1)
    ref = (torch.pow(f[1:] - f[:-1], 2).sum() + torch.pow(f[:, 1:] - f[:, :-1], 2).sum() + torch.pow(f[:, :, 1:] - f[:, :, :-1], 2).sum()) / 24 ** 3
    ref.backward()
    assert abs(float(val) - float(ref)) <= 1e-4 * abs(float(ref)), (float(val), float(ref))
    gc, gr = m.embed_res_fn.params.grad.cpu(), h.params.grad
    assert float((gc - gr).norm()) <= 1e-3 * float(gr.norm())
    # captured iteration with the smoothness term inside
    n = 2048
    m.embed_res_fn.params.grad = None
    step = GraphedMappingStep(m, opt, n, make_loss_fn(cfg, m), eager_steps=2)
    b = torch.tensor(cfg["mapping"]["bound"])
    losses = []
    for it in range(6):
        ro = (b[:, 0] + (0.3 + 0.4 * torch.rand(n, 3, generator=g)) * (b[:, 1] - b[:, 0])).to(cuda)
        rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(cuda)
        tc = torch.rand(n, 3, generator=g).to(cuda); td = (0.3 + 2.5 * torch.rand(n, 1, generator=g)).to(cuda)
        loss, _ = step(ro, rd, tc, td)
        losses.append(float(loss))
    assert step.graph is not None and all(np.isfinite(losses))


def test_graphed_ba_iteration_matches_eager(cuda, rf_lib):
    """ray_grads=True: the captured bundle-adjustment iteration (clamp=True, gradients w.r.t. the ray origins / directions handed
    back with the loss) against the eager one on the same batches: same losses, ray gradients to 1e-3 of their scale."""
    n = 2048
    cfg_e, m_e, opt_e = _make(cuda, False)
    cfg_g, m_g, opt_g = _make(cuda, True)
    step = GraphedMappingStep(m_g, opt_g, n, lambda r: configs.total_loss(cfg_g, r), eager_steps=2, ray_grads=True)
    g = torch.Generator().manual_seed(1)
    b = torch.tensor(cfg_e["mapping"]["bound"])
    for it in range(6):
        ro = (b[:, 0] + (0.3 + 0.4 * torch.rand(n, 3, generator=g)) * (b[:, 1] - b[:, 0])).to(cuda)
        rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(cuda)
        tc = torch.rand(n, 3, generator=g).to(cuda); td = (0.3 + 2.5 * torch.rand(n, 1, generator=g)).to(cuda)
        ro_e = ro.clone().requires_grad_(True); rd_e = rd.clone().requires_grad_(True)
        loss_e = configs.total_loss(cfg_e, m_e.mapping(ro_e, rd_e, tc, td, clamp=True))
        loss_e.backward()
        opt_e.step(zero_grad=True)
        loss_g, out = step(ro, rd, tc, td)
        assert abs(float(loss_g) - float(loss_e.detach())) <= 1e-4 * abs(float(loss_e.detach())), (it, float(loss_g), float(loss_e.detach()))
        for got, want in ((out["g_rays_o"], ro_e.grad), (out["g_rays_d"], rd_e.grad)):
            scale = float(want.abs().max())
            assert scale > 0 and float((got - want).abs().max()) <= 1e-3 * scale, (it, float((got - want).abs().max()), scale)
    assert step.graph is not None
