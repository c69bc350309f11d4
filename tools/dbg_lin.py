"""Bring-up check of the tcgen05 Linear at shapes that need column groups: max relative error of y, dx, dW, db against torch CPU."""
import sys, torch
sys.path.insert(0, '/root/repo')
import torch.nn.functional as F
from rbm_b200 import ops
for (M,N,K) in [(300,1024,256),(300,16,64),(300,64,1024),(5000,48,96),(300,512,128)]:
    torch.manual_seed(M+N)
    x, w, b = torch.randn(M, K), torch.randn(N, K) * 0.2, torch.randn(N)
    dy = torch.randn(M, N)
    xg, wg, bg = (t.cuda().requires_grad_(True) for t in (x, w, b))
    y = ops.linear(xg, wg, bg)
    y.backward(dy.cuda())
    xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
    ref = F.linear(xr, wr, br); ref.backward(dy)
    e = lambda a, r: ((a.detach().cpu() - r).abs().max() / r.abs().max()).item()
    print((M,N,K), 'y %.2e dx %.2e dw %.2e db %.2e' % (e(y, ref), e(xg.grad, xr.grad), e(wg.grad, wr.grad), e(bg.grad, br.grad)))
    bad = ((y.detach().cpu()-ref).abs() > 1e-3).nonzero()
    if len(bad): print('  bad y entries', len(bad), 'cols', sorted(set(bad[:,1].tolist()))[:20], 'rows', sorted(set(bad[:,0].tolist()))[:10])
