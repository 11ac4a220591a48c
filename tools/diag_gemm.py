"""GPU diagnostics for the tcgen05 tap GEMM / wgrad kernels (run under gpurun).

Compares the tensor-core kernels and their CUDA-core twins against an fp64
torch reference and prints error statistics plus a coarse map of where errors
sit (row / column blocks), which is what one needs to debug descriptor,
swizzle or barrier mistakes without a local GPU.
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmcnet_b200 import ops  # noqa: E402


def split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def ref_tap_gemm(A, B, M, shift, phase, bsel, Hp, Wp):
    # A [phases][rows][K] fp32, B [slices][N][K]
    rows = A.shape[1]
    out = torch.zeros(M, B.shape[1], dtype=torch.float64, device=A.device)
    q = torch.arange(M, device=A.device)
    for s, p, b in zip(shift, phase, bsel):
        idx = q + s
        ok = (idx >= 0) & (idx < rows)
        a = A[p].double()[idx.clamp(0, rows - 1)] * ok[:, None]
        out += a @ B[b].double().t()
    if Hp:
        wp = q % Wp
        hp = (q // Wp) % Hp
        keep = (wp >= 1) & (wp <= Wp - 2) & (hp >= 1) & (hp <= Hp - 2)
        out *= keep[:, None]
    return out


def report(name, got, ref):
    err = (got.double() - ref).abs()
    scale = ref.abs().max().item() + 1e-30
    print('  %-6s max_abs_err %.3e  rel_to_max %.3e  mean_err %.3e  nan %d' % (
        name, err.max().item(), err.max().item() / scale, err.mean().item(),
        int(torch.isnan(got).sum())))
    bad = err > 1e-3 * scale
    if bad.any():
        rb = bad.any(1).nonzero().flatten()
        cb = bad.any(0).nonzero().flatten()
        print('    bad rows %d (first %s ...) bad cols %d (first %s ...)' % (
            rb.numel(), rb[:12].tolist(), cb.numel(), cb[:12].tolist()))
        r0 = rb[0].item()
        print('    row %d got %s' % (r0, got[r0, :8].tolist()))
        print('    row %d ref %s' % (r0, ref[r0, :8].tolist()))
    return err.max().item() / scale


def case_tap(M, K, N, ntaps, phases, Hp, Wp, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    rows = M
    A = torch.randn(phases, rows, K, device='cuda', generator=g)
    B = torch.randn(ntaps, N, K, device='cuda', generator=g) * 0.1
    Ah, Al = split(A)
    Bh, Bl = split(B)
    if ntaps == 9:
        shift = [(r - 1) * Wp + (s - 1) for r in range(3) for s in range(3)]
    else:
        shift = [0] * ntaps
    phase = [t % phases for t in range(ntaps)]
    bsel = list(range(ntaps))
    ref = ref_tap_gemm(A, B, M, shift, phase, bsel, Hp, Wp)
    print('tap_gemm M=%d K=%d N=%d taps=%d phases=%d Hp=%d Wp=%d' % (M, K, N, ntaps, phases, Hp, Wp))
    worst = 0.0
    for eng in ('simt', 'tc'):
        D = torch.full((M, N), float('nan'), device='cuda')
        try:
            ops.tap_gemm(Ah, Al, Bh, Bl, D, a_phases=phases, a_rows=rows, K=K, b_slices=ntaps, N=N,
                         M=M, ldD=N, Hp=Hp, Wp=Wp, shift=shift, phase=phase, bsel=bsel, engine=eng)
            torch.cuda.synchronize()
            worst = max(worst, report(eng, D, ref))
        except Exception as e:  # noqa: BLE001
            print('  %s FAILED: %s' % (eng, e))
            worst = float('inf')
    return worst


def case_wgrad(P, Cout, Cin, ntaps, phases, Wp, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    G = torch.randn(P, Cout, device='cuda', generator=g)
    X = torch.randn(phases, P, Cin, device='cuda', generator=g)
    Gh, Gl = split(G)
    Xh, Xl = split(X)
    if ntaps == 9:
        shift = [(r - 1) * Wp + (s - 1) for r in range(3) for s in range(3)]
    else:
        shift = [0] * ntaps
    phase = [t % phases for t in range(ntaps)]
    bsel = list(range(ntaps))
    ref = torch.zeros(ntaps, Cout, Cin, dtype=torch.float64, device='cuda')
    q = torch.arange(P, device='cuda')
    for t in range(ntaps):
        idx = q + shift[t]
        ok = (idx >= 0) & (idx < P)
        x = X[phase[t]].double()[idx.clamp(0, P - 1)] * ok[:, None]
        ref[t] = G.double().t() @ x
    print('wgrad P=%d Cout=%d Cin=%d taps=%d phases=%d' % (P, Cout, Cin, ntaps, phases))
    worst = 0.0
    for eng in ('simt', 'tc'):
        dW = torch.zeros(ntaps, Cout, Cin, device='cuda')
        try:
            ops.wgrad_gemm(Gh, Gl, Xh, Xl, dW, P=P, Cout=Cout, x_phases=phases, Cin=Cin, shift=shift,
                           phase=phase, bsel=bsel, engine=eng)
            torch.cuda.synchronize()
            worst = max(worst, report(eng, dW.view(ntaps * Cout, Cin), ref.view(ntaps * Cout, Cin)))
        except Exception as e:  # noqa: BLE001
            print('  %s FAILED: %s' % (eng, e))
            worst = float('inf')
    return worst


def bench_tap(frames, H, C, Cout, reps=5):
    Hp = Wp = H + 2
    P = frames * Hp * Wp
    A = torch.randn(1, P, C, device='cuda')
    B = torch.randn(9, Cout, C, device='cuda') * 0.05
    Ah, Al = split(A)
    Bh, Bl = split(B)
    D = torch.empty(P, Cout, device='cuda')
    shift = [(r - 1) * Wp + (s - 1) for r in range(3) for s in range(3)]
    if os.environ.get('ZERO_SHIFT') == '1':
        shift = [0, 8, 16, 24, 32, 40, 48, 56, 64]       # atom-aligned row offsets
    if os.environ.get('ZERO_SHIFT') == '2':
        shift = [0, 1, 2, 3, 4, 5, 6, 7, 9]              # small unaligned window
    kw = dict(a_phases=1, a_rows=P, K=C, b_slices=9, N=Cout, M=P, ldD=Cout, Hp=Hp, Wp=Wp,
              shift=shift, phase=[0] * 9, bsel=list(range(9)))
    for eng in ('tc',):
        for _ in range(2):
            ops.tap_gemm(Ah, Al, Bh, Bl, D, engine=eng, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(reps):
            ops.tap_gemm(Ah, Al, Bh, Bl, D, engine=eng, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = 2.0 * frames * H * H * 9 * C * Cout
        print('bench %s frames=%d H=%d C=%d Cout=%d: %.3f ms  %.1f TFLOP/s (algorithmic)' % (
            eng, frames, H, C, Cout, ms, flops / ms / 1e9))
    G = torch.randn(P, Cout, device='cuda')
    Gh, Gl = split(G)
    dW = torch.zeros(9, Cout, C, device='cuda')
    kw2 = dict(P=P, Cout=Cout, x_phases=1, Cin=C, shift=shift, phase=[0] * 9, bsel=list(range(9)))
    for _ in range(2):
        ops.wgrad_gemm(Gh, Gl, Ah, Al, dW, engine='tc', **kw2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        ops.wgrad_gemm(Gh, Gl, Ah, Al, dW, engine='tc', **kw2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('bench wgrad frames=%d H=%d C=%d Cout=%d: %.3f ms  %.1f TFLOP/s (algorithmic)' % (
        frames, H, C, Cout, ms, flops / ms / 1e9))


def main():
    print(torch.cuda.get_device_name(0))
    t0 = time.time()
    worst = 0.0
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if what in ('all', 'tap'):
        # plain GEMM first (1 tap, no mask), then shifted taps with the border mask
        worst = max(worst, case_tap(256, 64, 64, 1, 1, 0, 0))
        worst = max(worst, case_tap(1000, 128, 128, 1, 1, 0, 0))
        worst = max(worst, case_tap(384, 256, 32, 1, 1, 0, 0))
        worst = max(worst, case_tap(2 * 10 * 10, 64, 64, 9, 1, 10, 10))
        worst = max(worst, case_tap(3 * 16 * 16, 128, 256, 9, 1, 16, 16))
        worst = max(worst, case_tap(3 * 9 * 9, 64, 128, 9, 4, 9, 9))
    if what in ('all', 'wgrad'):
        worst = max(worst, case_wgrad(512, 128, 64, 1, 1, 10))
        worst = max(worst, case_wgrad(2 * 10 * 10, 64, 64, 9, 1, 10))
        worst = max(worst, case_wgrad(3 * 16 * 16 + 7, 256, 128, 9, 1, 16))
        worst = max(worst, case_wgrad(5000, 128, 128, 9, 4, 30))
    print('WORST rel err %.3e   (%.1fs)' % (worst, time.time() - t0))
    if worst < 1e-3 and what in ('all', 'bench'):
        bench_tap(192, 56, 64, 64)
        bench_tap(192, 28, 128, 128)
        bench_tap(192, 14, 256, 256)
        bench_tap(192, 7, 512, 512)
    print('DIAG_GEMM_DONE worst=%.3e' % worst)


if __name__ == '__main__':
    main()
