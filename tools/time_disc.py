"""Times the planar conv kernels on every Discriminator3 layer shape at the bench size (384 frames)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmcnet_b200 import ops

M = 384
FMA_PEAK = 148 * 128 * 1.95e9
LAYERS = [(2, 16, 224, 2), (16, 16, 112, 1), (16, 32, 112, 2), (32, 32, 56, 1), (32, 64, 56, 2),
          (64, 64, 28, 1), (64, 128, 28, 2), (128, 128, 14, 1)]


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    tot = [0.0, 0.0, 0.0]
    for cin, cout, h, st in LAYERS:
        ho = h // st
        x = torch.randn(M, cin, h, h, device='cuda')
        w = torch.randn(cout, cin, 3, 3, device='cuda') * 0.1
        b = torch.randn(cout, device='cuda')
        y = torch.empty(M, cout, ho, ho, device='cuda')
        dy = torch.randn_like(y)
        dx = torch.empty_like(x)
        dW, dB = torch.zeros_like(w), torch.zeros_like(b)
        fma = cin * cout * 9 * M * ho * ho
        tf = timeit(lambda: ops.conv_fwd(x, cin * h * h, cin, h, h, w, b, cout, 3, st, y, cout * ho * ho, M, slope=0.2))
        td = timeit(lambda: ops.conv_dgrad(dy, cout * ho * ho, cout, w, cin, cin, 3, st, dx, cin * h * h, h, h, M))
        tw = timeit(lambda: ops.conv_wgrad(x, cin * h * h, cin, h, h, dy, cout * ho * ho, cout, 3, st, dW, dB, M))
        tot[0] += tf; tot[1] += td; tot[2] += tw
        print('%3d->%3d %3dpx s%d  fwd %7.1f us (%4.1f%%)  dgrad %7.1f us (%4.1f%%)  wgrad %7.1f us (%4.1f%%)' % (
            cin, cout, h, st, tf * 1e6, 100 * fma / tf / FMA_PEAK, td * 1e6, 100 * fma / td / FMA_PEAK,
            tw * 1e6, 100 * fma / tw / FMA_PEAK), flush=True)
    print('total (s1 layers run twice in D3): fwd %.2f dgrad %.2f wgrad %.2f ms' % tuple(t * 1e3 for t in tot))


if __name__ == '__main__':
    main()
