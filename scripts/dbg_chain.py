import torch, torch.nn.functional as F, sys
sys.path.insert(0, '.')
from shapeformer_b200 import ops
def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale
cuda = torch.device('cuda:0')
for (M, N, K) in [(33, 1024, 4096), (64, 1024, 2048), (4, 256, 1536), (64, 128, 4096)]:
    x, W, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=0.05), rnd(N, seed=3)
    g, be = rnd(K, seed=5, scale=0.2) + 1, rnd(K, seed=6, scale=0.1)
    xs = x * 3.0 + 0.5
    ref = F.layer_norm(xs.double(), (K,), g.double(), be.double(), 1e-5) @ W.double().t() + b.double()
    out = ops.linear_chain(xs.to(cuda), W.to(cuda), b.to(cuda), ln=(g.to(cuda), be.to(cuda))).cpu()
    nan = torch.isnan(out)
    print(M, N, K, 'nan rows', nan.any(1).nonzero().flatten().tolist()[:20], 'nan cols', int(nan.any(0).sum()),
          'err', (out.double() - ref)[~nan].abs().max().item())
    ws = ops._chain_ws[cuda]
        pieces = (K + 127) // 128
    stv = ws[256:256 + M * pieces * 8].view(torch.float32).view(M, pieces, 2).cpu()
    xm = xs.view(M, pieces, -1)
    print('  mean err', (stv[..., 0] - xm.mean(-1)).abs().max().item(), 'M2 err', (stv[..., 1] - ((xm - xm.mean(-1, keepdim=True)) ** 2).sum(-1)).abs().max().item(), 'nan stats', int(torch.isnan(stv).sum()))
