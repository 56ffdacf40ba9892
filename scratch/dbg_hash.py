import sys, numpy as np, torch
sys.path.insert(0, '.')
from oracle import tcnn_standin
from remixfusion_b200.encodings import GridEncoding
cuda = torch.device('cuda')
g = torch.Generator().manual_seed(0)
x = torch.cat([torch.rand(500, 3, generator=g), torch.rand(200, 3, generator=g) * 3 - 1,
               torch.tensor([[0., 0., 0.], [1., 1., 1.], [0.5, 0.999999, 1e-7]])])
for hash_size, res in ((10, 400), (14, 512), (19, 2048)):
    pls = np.exp2(np.log2(res / 16) / 15)
    enc = GridEncoding(16, 2, 16, pls, hash_size, True, cuda)
    ref = tcnn_standin.GridStandIn(16, 2, True, hash_size, 16, pls)
    with torch.no_grad(): enc.params.fill_(1.0)
    for l in range(16):
        i, w = tcnn_standin.grid_indices(x, ref.scale[l], ref.res[l], ref.size[l], True)
        # per-sample check: scatter one sample at a time at this level is expensive; instead compare per-level sets
    out = enc(x.to(cuda)); out.sum().backward()
    grad = enc.params.grad.view(-1, 2).cpu()
    for l in range(16):
        i, w = tcnn_standin.grid_indices(x, ref.scale[l], ref.res[l], ref.size[l], True)
        exp = torch.zeros(ref.size[l], 2, dtype=torch.float64)
        exp.index_add_(0, i.reshape(-1), w.reshape(-1, 1).double().expand(-1, 2))
        got = grad[ref.offset[l]:ref.offset[l + 1]].double()
        bad = (exp - got).abs().max(1).values > 1e-4
        if bad.any():
            b = bad.nonzero()[:5, 0]
            print(hash_size, 'level', l, 'res', ref.res[l], 'size', ref.size[l], 'bad', int(bad.sum()), [(int(k), exp[k, 0].item(), got[k, 0].item()) for k in b])
            # which samples hit the first bad entry
            k = int(b[0]); s = (i == k).any(1).nonzero()[:3, 0]
            print('   samples', s.tolist(), x[s].tolist())
print('done')
