"""N1: fused Adam vs torch.optim.Adam — the optimiser the reference constructs (mp_slam/slam.py:271-286) —
run on the CPU in fp32 on the same parameters and gradients.  Tolerance 2e-6 relative + 3e-8 absolute (half an ulp of the
largest parameters, a few ulps of the lr = 1e-2 steps accumulated over six iterations: torch's vectorised CPU kernels contract / order the same operations differently; a parameter that
lands near zero after the update only has that absolute accuracy on either side)."""
import pytest
import torch

from remixfusion_b200.optim import Adam

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [3, 1024, 100003])
def test_fused_adam_matches_torch(cuda, rf_lib, n):
    g = torch.Generator().manual_seed(n)
    dec0 = (torch.rand(n, generator=g) - 0.5); tab0 = (torch.rand(2 * n + 1, generator=g) - 0.5) * 1e-2
    ref_p = [dec0.clone().requires_grad_(True), tab0.clone().requires_grad_(True)]
    got_p = [dec0.clone().to(cuda).requires_grad_(True), tab0.clone().to(cuda).requires_grad_(True)]
    groups = lambda ps: [{"params": [ps[0]], "weight_decay": 1e-6, "lr": 1e-2}, {"params": [ps[1]], "eps": 1e-15, "lr": 1e-2}]
    ref = torch.optim.Adam(groups(ref_p), betas=(0.9, 0.99))
    got = Adam(groups(got_p), betas=(0.9, 0.99))
    for it in range(6):
        for rp, gp in zip(ref_p, got_p):
            gr = torch.randn(rp.shape, generator=g) * (10.0 ** -(it % 3))
            gr[::7] = 0.0                                   # entries without gradient still move (dense semantics)
            rp.grad = gr.clone(); gp.grad = gr.clone().to(cuda)
        ref.step(); got.step(zero_grad=(it % 2 == 1))
        for rp, gp in zip(ref_p, got_p):
            torch.testing.assert_close(gp.detach().cpu(), rp.detach(), rtol=2e-6, atol=3e-8)
            if it % 2 == 1:
                assert float(gp.grad.abs().sum()) == 0.0
    for rp, gp in zip(ref_p, got_p):
        # moments: gradients are O(1), so an ulp of the accumulators is ~1e-7 (torch's CPU lerp / addcmul use fused multiply-adds)
        torch.testing.assert_close(got.state[gp]["exp_avg"].cpu(), ref.state[rp]["exp_avg"], rtol=1e-5, atol=2e-7)
        torch.testing.assert_close(got.state[gp]["exp_avg_sq"].cpu(), ref.state[rp]["exp_avg_sq"], rtol=1e-5, atol=2e-7)
    assert set(got.state_dict()["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}


def test_sharded_adam_single_rank_is_the_fused_adam(cuda, rf_lib):
    """dist.ShardedAdam with one rank (no collectives): parameters / gradients re-homed into the flat buffers, one rf_adam_step
    per group segment — bit-identical to remixfusion_b200.optim.Adam on the same gradients, gradients cleared by the step."""
    from remixfusion_b200.dist import ShardedAdam
    g = torch.Generator().manual_seed(5)
    shapes = [(1001,), (7, 5), (3, 7)]
    init = [torch.randn(s, generator=g) for s in shapes]
    a = [torch.nn.Parameter(t.clone().to(cuda)) for t in init]
    b = [torch.nn.Parameter(t.clone().to(cuda)) for t in init]
    groups = lambda ps: [{"params": [ps[1], ps[2]], "weight_decay": 1e-6, "lr": 1e-2}, {"params": [ps[0]], "eps": 1e-15, "lr": 1e-2}]
    sh = ShardedAdam(groups(a), betas=(0.9, 0.99))
    ref = Adam(groups(b), betas=(0.9, 0.99))
    for it in range(5):
        for pa, pb in zip(a, b):
            gr = torch.randn(pa.shape, generator=g).to(cuda)
            pa.grad.add_(gr)                                  # the views accumulate like autograd does
            pb.grad = gr.clone()
        sh.step(); ref.step()
        for pa, pb in zip(a, b):
            assert torch.equal(pa.detach(), pb.detach())
            assert float(pa.grad.abs().sum()) == 0.0 and pa.grad._base is sh.gflat
