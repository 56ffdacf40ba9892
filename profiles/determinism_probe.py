"""Run-to-run spread of the mapping gradients on identical inputs (the atomics make them order-dependent at the fp32 rounding
level; anything larger is a race).  Usage: python profiles/determinism_probe.py [lib.so]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import remixfusion_b200.abi as abi
if len(sys.argv) > 1:
    abi.LIB_PATH = os.path.abspath(sys.argv[1])
from remixfusion_b200 import configs
from remixfusion_b200.scene_rep import JointEncoding

dev = torch.device("cuda", 0)
for hidden, hs, n in ((32, 12, 2048), (32, 16, 65536), (64, 14, 8192)):
    cfg = configs.replica(hidden=hidden, hash_size=hs)
    cfg["training"]["perturb"] = 0
    bb = torch.from_numpy(np.array(cfg["mapping"]["bound"])).to(torch.float64)
    torch.manual_seed(3)
    m = JointEncoding(cfg, bb).to(dev); m.train()
    with torch.no_grad():
        m.GBV.params.copy_((torch.rand_like(m.GBV.params) * 2 - 1) * 0.5)
        m.embed_res_fn.params.copy_((torch.rand_like(m.embed_res_fn.params) * 2 - 1) * 1e-2)
    g = torch.Generator().manual_seed(0)
    b = torch.tensor(cfg["mapping"]["bound"])
    ro = (b[:, 0] + (0.3 + 0.4 * torch.rand(n, 3, generator=g)) * (b[:, 1] - b[:, 0])).to(dev)
    rd = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1).to(dev)
    tc = torch.rand(n, 3, generator=g).to(dev); td = (0.3 + 2.5 * torch.rand(n, 1, generator=g)).to(dev)
    params = [p for p in m.parameters() if p.requires_grad]
    names = [k for k, p in m.named_parameters() if p.requires_grad]
    ref = None
    worst = [0.0] * len(params); nd = [0] * len(params)
    for it in range(12):
        for p in params: p.grad = None
        loss = configs.total_loss(cfg, m.mapping(ro, rd, tc, td)); loss.backward()
        if ref is None:
            keep = [i for i, p in enumerate(params) if p.grad is not None]
            params = [params[i] for i in keep]; names = [names[i] for i in keep]
            worst = [0.0] * len(params); nd = [0] * len(params)
        gs = [p.grad.detach().clone() for p in params]
        if ref is None:
            ref = gs; continue
        for i, (a, r) in enumerate(zip(gs, ref)):
            d = (a - r).abs()
            worst[i] = max(worst[i], float(d.max() / (r.abs().max() + 1e-30)))
            nd[i] = max(nd[i], int((d > 0).sum()))
    print(f"hidden {hidden} hash 2^{hs} rays {n}:")
    for nm, w, c, p in zip(names, worst, nd, params):
        print(f"   {nm:40s} numel {p.numel():9d}  max|dg|/max|g| {w:.3e}  differing {c}")
