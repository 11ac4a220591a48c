"""Times the generator's conv layers one by one at the bench size (192 frames of 224x224) and
prints each against the fp32-FMA and HBM ceilings measured on this GPU class."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmcnet_b200 import ops

N, H, W = 192, 224, 224
FMA_PEAK = 148 * 128 * 1.95e9          # FMA/s
HBM = 7.0e12
LAYERS = [(5, 8), (13, 8), (21, 6), (27, 4), (31, 2), (33, 2)]


def timeit(fn, reps=int(os.environ.get('REPS', '5'))):
    fn(); torch.cuda.synchronize()
    if reps == 0:
        return 1.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    which = sys.argv[1].split(',') if len(sys.argv) > 1 else ['fwd', 'wgrad']
    buf = torch.randn(N, 33, H, W, device='cuda')
    flat = buf.view(-1)
    tot = {}
    for cin, cout in LAYERS:
        w = torch.randn(cout, cin, 3, 3, device='cuda') * 0.1
        b = torch.randn(cout, device='cuda')
        out = torch.empty(N, cout, H, W, device='cuda')
        fma = cin * cout * 9 * N * H * W
        byt = (cin + cout) * N * H * W * 4
        src = flat[(33 - cin) * H * W:]
        line = '%2d->%d ' % (cin, cout)
        if 'fwd' in which:
            t = timeit(lambda: ops.conv_fwd(src, 33 * H * W, cin, H, W, w, b, cout, 3, 1, out, cout * H * W, N, slope=0.1))
            tot['fwd'] = tot.get('fwd', 0) + t
            line += ' fwd %6.1f us (FMA %4.1f%%, HBM %4.1f%%)' % (t * 1e6, 100 * fma / t / FMA_PEAK, 100 * byt / t / HBM)
        if 'wgrad' in which:
            dW, dB = torch.zeros_like(w), torch.zeros_like(b)
            t = timeit(lambda: ops.conv_wgrad(src, 33 * H * W, cin, H, W, out, cout * H * W, cout, 3, 1, dW, dB, N))
            tot['wgrad'] = tot.get('wgrad', 0) + t
            line += ' wgrad %6.1f us (FMA %4.1f%%, HBM %4.1f%%)' % (t * 1e6, 100 * fma / t / FMA_PEAK, 100 * byt / t / HBM)
        print(line, flush=True)
    print('total', ' '.join('%s %.3f ms' % (k, v * 1e3) for k, v in tot.items()))


if __name__ == '__main__':
    main()
