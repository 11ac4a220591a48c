"""GPU diagnostics for the planar small-channel conv kernels vs torch CPU fp32."""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmcnet_b200 import ops


def rel(a, b):
    return float((a.double().cpu() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def case(N, Cin, Cout, H, W, ks, stride, which):
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, ks, ks, generator=g) * 0.2
    b = torch.randn(Cout, generator=g)
    pad = ks // 2
    xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    y = F.leaky_relu(F.conv2d(xr, wr, br, stride, pad), 0.1)
    dy = torch.randn(y.shape, generator=g)
    Ho, Wo = y.shape[-2:]
    pre = F.conv2d(x, w, b, stride, pad)
    dpre = dy * torch.where(pre > 0, torch.ones_like(pre), torch.full_like(pre, 0.1))
    y.backward(dy)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    out = {}
    if 'fwd' in which:
        yo = torch.full((N, Cout, Ho, Wo), float('nan'), device='cuda')
        ops.conv_fwd(xd, Cin * H * W, Cin, H, W, wd, bd, Cout, ks, stride, yo, Cout * Ho * Wo, N, slope=0.1)
        torch.cuda.synchronize()
        out['fwd'] = rel(yo, y.detach())
    dpd = dpre.cuda()
    if 'wgrad' in which:
        dW = torch.zeros_like(wd); dB = torch.zeros_like(bd)
        ops.conv_wgrad(xd, Cin * H * W, Cin, H, W, dpd, Cout * Ho * Wo, Cout, ks, stride, dW, dB, N)
        torch.cuda.synchronize()
        out['wgrad'] = rel(dW, wr.grad); out['bgrad'] = rel(dB, br.grad)
    if 'dgrad' in which:
        dX = torch.full((N, Cin, H, W), float('nan'), device='cuda')
        ops.conv_dgrad(dpd, Cout * Ho * Wo, Cout, wd, Cin, Cin, ks, stride, dX, Cin * H * W, H, W, N)
        torch.cuda.synchronize()
        out['dgrad'] = rel(dX, xr.grad)
        if ks == 3 and stride == 1:
            wT = torch.zeros(Cin * Cout * 9, device='cuda')
            ops.weight_flip(wd, Cout, Cin, Cin, wT)
            dX2 = torch.full((N, Cin, H, W), float('nan'), device='cuda')
            ops.conv_fwd(dpd, Cout * Ho * Wo, Cout, H, W, wT, None, Cin, 3, 1, dX2, Cin * H * W, N)
            torch.cuda.synchronize()
            out['dgrad_as_fwd'] = rel(dX2, xr.grad)
    print('N=%d Cin=%d Cout=%d %dx%d k%d s%d:' % (N, Cin, Cout, H, W, ks, stride),
          ' '.join('%s %.2e' % kv for kv in out.items()), flush=True)
    return max(out.values())


if __name__ == '__main__':
    which = sys.argv[1].split(',') if len(sys.argv) > 1 else ['fwd', 'wgrad', 'dgrad']
    worst = 0.0
    for args in [(2, 33, 2, 64, 64, 3, 1), (2, 5, 8, 64, 96, 3, 1), (1, 21, 6, 40, 36, 3, 1),
                 (2, 27, 4, 32, 64, 3, 1), (2, 16, 16, 56, 56, 3, 1), (2, 2, 16, 64, 64, 3, 2),
                 (2, 32, 64, 28, 28, 3, 2), (2, 2, 64, 64, 64, 7, 2), (1, 64, 64, 28, 28, 3, 1)]:
        worst = max(worst, case(*args, which))
    print('DIAG_CONV_DONE worst=%.3e' % worst)
