"""Times the classifier-stem kernels at the bench size (192 frames): 7x7/2 conv forward and weight
gradient, BN+ReLU+maxpool forward, and its backward; prints achieved HBM GB/s / FMA fraction."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmcnet_b200 import ops

N, H, W, C = 192, 224, 224, 64
FMA_PEAK = 148 * 128 * 1.95e9


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def main():
    Ho = Wo = 112
    Hq = Wq = 56
    x = torch.randn(N, 2, H, W, device='cuda')
    w = torch.randn(C, 2, 7, 7, device='cuda') * 0.1
    Y = torch.empty(N, C, Ho, Wo, device='cuda')
    t = timeit(lambda: ops.conv_fwd(x, 2 * H * W, 2, H, W, w, None, C, 7, 2, Y, C * Ho * Wo, N, slope=1.0))
    fma = 2 * 49 * C * N * Ho * Wo
    print('stem conv fwd   %7.1f us  FMA %4.1f%%  write %5.0f GB/s' % (t * 1e6, 100 * fma / t / FMA_PEAK, Y.numel() * 4 / t / 1e9))
    wb = torch.zeros(128 * 128, dtype=torch.bfloat16, device='cuda')
    Y2 = torch.empty_like(Y)
    t = timeit(lambda: ops.stem_conv_tc_fwd(x, 2 * H * W, H, W, w, wb, Y2, C * Ho * Wo, N))
    print('stem conv fwd TC %6.1f us  write %5.0f GB/s  max|diff| vs simt %.2e' % (t * 1e6, Y.numel() * 4 / t / 1e9, float((Y2 - Y).abs().max())))
    dW = torch.zeros_like(w)
    dZ = torch.randn_like(Y)
    t = timeit(lambda: ops.conv_wgrad(x, 2 * H * W, 2, H, W, dZ, C * Ho * Wo, C, 7, 2, dW, None, N))
    print('stem conv wgrad %7.1f us  FMA %4.1f%%  read  %5.0f GB/s' % (t * 1e6, 100 * fma / t / FMA_PEAK, dZ.numel() * 4 / t / 1e9))
    ws = torch.empty(ops.stem_wgrad_workspace_floats(), device='cuda')
    dW2 = torch.zeros_like(w)
    t = timeit(lambda: ops.stem_conv_tc_wgrad(x, 2 * H * W, H, W, dZ, C * Ho * Wo, dW2, ws, N))
    print('stem conv wgrad TC %5.1f us  read  %5.0f GB/s' % (t * 1e6, dZ.numel() * 4 / t / 1e9))
    scale = torch.rand(C, device='cuda') + 0.5
    shift = torch.randn(C, device='cuda') * 0.1
    a_hi = torch.zeros(N, Hq + 2, Wq + 2, C, dtype=torch.bfloat16, device='cuda')
    a_lo = torch.zeros_like(a_hi)
    idx = torch.zeros(N, Hq, Wq, C, dtype=torch.uint8, device='cuda')
    t = timeit(lambda: ops.stem_pool_fwd(Y, scale, shift, N, C, Ho, Wo, a_hi, a_lo, idx))
    byt = Y.numel() * 4 + N * Hq * Wq * C * 5
    print('stem pool fwd   %7.1f us  %5.0f GB/s' % (t * 1e6, byt / t / 1e9))
    g_a = torch.randn(N, Hq + 2, Wq + 2, C, device='cuda')
    g_b = torch.randn_like(g_a)
    t = timeit(lambda: ops.stem_pool_bwd(g_a, g_b, idx, Y, scale, shift, N, C, Ho, Wo, dZ))
    byt = Y.numel() * 8 + g_a.numel() * 8 + idx.numel()
    print('stem pool bwd   %7.1f us  %5.0f GB/s' % (t * 1e6, byt / t / 1e9))


if __name__ == '__main__':
    main()
