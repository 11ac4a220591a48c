"""Per-kernel-family time of one I3D train step (CUDA events around every launch, eager):
python tools/time_i3d.py [clips] -> ms per step per C-ABI entry point, with the tap-GEMM / wgrad calls split by
tap count, and the issued GEMM FLOPs."""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmcnet_b200 import ops
from dmcnet_b200.i3d_engine import I3DEngine
from dmcnet_b200.i3d_model import build_i3d_state
from dmcnet_b200.i3d_trainer import I3DHParams, I3DTrainStep


class Timer:
    def __init__(self):
        self.ev = []

    def __call__(self, name, args):
        import contextlib

        @contextlib.contextmanager
        def cm():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            yield
            e1.record()
            v = lambda a: a.value if hasattr(a, 'value') else a
            key, fl = name.replace('dmc_', ''), 0.0
            if name == 'dmc_tc_tap_gemm_ex':
                K, N, M, nt = v(args[4]), v(args[8]), v(args[10]), v(args[14])
                key += '[%dtap%s]' % (nt, ',bw' if v(args[19]) else '')
                fl = 2.0 * M * N * K * nt
            elif name == 'dmc_tc_wgrad_ex':
                P, Cout, Cin, nt = v(args[3]), v(args[4]), v(args[8]), v(args[10])
                key += '[%dtap]' % nt
                fl = 2.0 * P * Cout * Cin * nt
            shape = ''
            if name == 'dmc_tc_tap_gemm_ex':
                shape = 'M=%d N=%d K=%d ldD=%d lda=%d%s%s' % (M, N, K, v(args[11]), v(args[2]),
                                                            ' stats' if v(args[17]) else '', ' gb' if v(args[21]) else '')
            elif name == 'dmc_tc_wgrad_ex':
                shape = 'P=%d Cout=%d Cin=%d' % (P, Cout, Cin)
            elif name in ('dmc_maxpool3d_fwd', 'dmc_maxpool3d_bwd'):
                shape = 'C=%d thw=%s k=%s s=%s' % (v(args[3]), list(args[4]), list(args[6]), list(args[7]))
            self.ev.append((key, e0, e1, fl, shape))
        return cm()


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    eng = I3DEngine(51, B, 16)
    eng.load_state(build_i3d_state(51, 'DenseNetTiny', seed=1))
    tr = I3DTrainStep(eng, I3DHParams(epoch_thre=0))
    g = torch.Generator().manual_seed(0)
    data = torch.empty(B, 7, 16, 224, 224).normal_(generator=g).cuda()
    target = torch.randint(0, 51, (B,), generator=g).cuda()
    mask = eng.draw_dropout_mask(0.5, g)
    for _ in range(2):
        tr.step(data, target, dropout_mask=mask, metrics=False)
    t = Timer()
    ops.set_call_hook(t)
    steps = 2
    for _ in range(steps):
        tr.step(data, target, dropout_mask=mask, metrics=False)
    ops.set_call_hook(None)
    torch.cuda.synchronize()
    ms, fl, cnt = defaultdict(float), defaultdict(float), defaultdict(int)
    calls = defaultdict(float)
    for k, e0, e1, f, shape in t.ev:
        if shape:
            calls[(k, shape)] += e0.elapsed_time(e1) / steps
    for k, e0, e1, f, shape in t.ev:
        ms[k] += e0.elapsed_time(e1) / steps
        fl[k] += f / steps
        cnt[k] += 1
    tot = sum(ms.values())
    print('I3D train step B=%d: %.2f ms summed over %d launches' % (B, tot, len(t.ev) // steps))
    for k in sorted(ms, key=lambda k: -ms[k]):
        extra = '  %.0f TFLOP/s issued' % (fl[k] / ms[k] / 1e9) if fl[k] else ''
        print('  %-34s %8.3f ms  %5.1f %%  x%d%s' % (k, ms[k], 100 * ms[k] / tot, cnt[k] // steps, extra))
    ncall = defaultdict(int)
    for k, e0, e1, f, shape in t.ev:
        if shape:
            ncall[(k, shape)] += 1
    print('slowest GEMM shapes (ms per step, all calls of the shape):')
    for (k, shape), v in sorted(calls.items(), key=lambda kv: -kv[1])[:48]:
        print('  %-28s %-58s %7.3f ms  x%d' % (k, shape, v, ncall[(k, shape)] // steps))


if __name__ == '__main__':
    main()
