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
            torch.testing.assert_close(pg, pe, rtol=1e-3, atol=2e-5)
