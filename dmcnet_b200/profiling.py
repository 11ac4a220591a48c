"""Live per-kernel-family timing of a train step (CUDA events on the launch
stream) and the roofline figures bench.py reports.

Algorithmic work: a tap GEMM launch does
2 * valid_pixels * N * K * ntaps FLOP where valid_pixels excludes the padding
ring; the tensor-pipe peak is the measured cuBLAS bf16 figure in
MEASURED_PEAKS.json (the bf16x3 split issues 3 MMAs per algorithmic MAC, so the
ceiling of `frac` for these kernels is 1/3).
"""
from __future__ import annotations

import contextlib
import json
import os
from collections import defaultdict

import torch

from . import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return {'hbm_gbs': d['hbm_gbs'], 'tensor_tflops': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'source': 'measured (MEASURED_PEAKS.json: copy GB/s, cuBLAS bf16 sustained)'}
    return {'hbm_gbs': 6650.0, 'tensor_tflops': 1590.0, 'source': 'fallback (B200_PROFILING.md)'}


class FamilyTimer:
    def __init__(self):
        self.events = []            # (name, e0, e1, flops, bytes)

    @contextlib.contextmanager
    def __call__(self, name, args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        self.events.append((name, e0, e1) + _work(name, args))

    def totals(self):
        torch.cuda.synchronize()
        ms, fl, by, cnt = defaultdict(float), defaultdict(float), defaultdict(float), defaultdict(int)
        for name, e0, e1, f, b in self.events:
            ms[name] += e0.elapsed_time(e1)
            fl[name] += f
            by[name] += b
            cnt[name] += 1
        return ms, fl, by, cnt


def _v(a):
    return a.value if hasattr(a, 'value') else a


def _interior_fraction(Hp, Wp):
    ph = ops.pad_hi()
    return ((Hp - 1 - ph) * (Wp - 1 - ph)) / float(Hp * Wp) if Hp else 1.0


def _flops(name, args):
    """Algorithmic FLOPs of one call (0 for bandwidth kernels)."""
    return _work(name, args)[0]


def _work(name, args):
    """(algorithmic FLOPs, algorithmic HBM bytes) of one call; bytes = every operand touched once."""
    if name in ('dmc_tc_tap_gemm', 'dmc_simt_tap_gemm', 'dmc_tc_tap_gemm_act'):
        K, N, M, Hp, Wp, ntaps = _v(args[4]), _v(args[8]), _v(args[10]), _v(args[12]), _v(args[13]), _v(args[14])
        valid = M * _interior_fraction(Hp, Wp)
        return 2.0 * valid * N * K * ntaps, M * (K * 4.0 + N * 4.0)
    if name == 'dmc_tc_tap_gemm_ring':
        K, N, M, Hp, Wp, R, ntaps = (_v(args[i]) for i in (4, 8, 10, 12, 13, 14, 15))
        valid = M * ((Hp - R) * (Wp - R)) / float(Hp * Wp)
        return 2.0 * valid * N * K * ntaps, M * (K * 4.0 + N * 4.0)
    if name in ('dmc_tc_wgrad', 'dmc_simt_wgrad'):
        P, Cout, Cin, ntaps = _v(args[2]), _v(args[3]), _v(args[7]), _v(args[9])
        # ring rows of dY are zero: only interior pixels are algorithmic work.  The call carries no
        # geometry, so the fraction comes from the square frame the row count implies (set by the hook)
        return 2.0 * P * _WGRAD_INTERIOR.get(int(P), 1.0) * Cout * Cin * ntaps, P * (Cout + Cin) * 4.0
    if name in ('dmc_conv_fwd',):
        Cin, H, W, Cout, ks, stride, N = (_v(args[i]) for i in (2, 3, 4, 7, 8, 9, 17))
        ho, wo = H // stride, W // stride
        return 2.0 * N * ho * wo * Cin * Cout * ks * ks, 4.0 * N * (Cin * H * W + Cout * ho * wo)
    if name in ('dmc_conv_wgrad',):
        Cin, H, W, Cout, ks, stride, N = (_v(args[i]) for i in (2, 3, 4, 7, 8, 9, 12))
        ho, wo = H // stride, W // stride
        return 2.0 * N * ho * wo * Cin * Cout * ks * ks, 4.0 * N * (Cin * H * W + Cout * ho * wo)
    if name in ('dmc_conv_dgrad',):
        Cout, cic, ks, stride, H, W, N = (_v(args[i]) for i in (2, 5, 6, 7, 10, 11, 13))
        ho, wo = H // stride, W // stride
        return 2.0 * N * ho * wo * cic * Cout * ks * ks, 4.0 * N * (cic * H * W + Cout * ho * wo)
    if name in ('dmc_conv3x3_dgrad_fused',):
        Cy, H, W, Cx, N = (_v(args[i]) for i in (2, 3, 4, 6, 14))
        return 2.0 * N * H * W * Cy * Cx * 9, 4.0 * N * H * W * (Cy + 2 * Cx)
    if name in ('dmc_conv3x3_taps2',):
        Cin, H, W, Cout, N = (_v(args[i]) for i in (2, 3, 4, 7, 13))
        return 2.0 * N * H * W * Cin * Cout * 4, 4.0 * N * H * W * (Cin + Cout)
    return 0.0, 0.0


_WGRAD_INTERIOR = {}          # rows P of a pixel-major tensor -> interior fraction (filled by register_geometry)


def register_geometry(frames, H, W):
    """Tell the FLOP model which padded row counts belong to which frame geometry, so that the
    weight-gradient count excludes the zero ring like the tap-GEMM count does."""
    Hp, Wp = ops.padded(H), ops.padded(W)
    _WGRAD_INTERIOR[int(frames * Hp * Wp)] = _interior_fraction(Hp, Wp)


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the newest committed ncu --set full capture
    (profiles/rNN_tap_gemm_traffic.json), or None."""
    d = os.path.join(ROOT, 'profiles')
    cands = sorted(f for f in os.listdir(d) if f.endswith('_tap_gemm_traffic.json')) if os.path.isdir(d) else []
    if cands:
        with open(os.path.join(d, cands[-1])) as f:
            return json.load(f).get('dram_bytes_per_launch_avg')
    return None


def fma_peak_tflops():
    """fp32 FMA peak: 128 FMA/clk/SM (tools/ubench/pipes.cu) x 148 SMs x max SM clock."""
    mhz = 1965.0
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        with open(p) as f:
            mhz = float(json.load(f).get('sm_max_mhz', mhz))
    return 148 * 128 * 2 * mhz * 1e6 / 1e12


def measure_roofline(resident_step, per, steps, trainer):
    """Run `steps` instrumented eager steps; return the dominant family's roofline
    and the per-family time breakdown (ms per train step)."""
    peaks = measured_peaks()
    was_graph, was_overlap = trainer.use_graph, getattr(trainer, 'overlap', False)
    trainer.use_graph = False
    trainer.overlap = False                          # every kernel timed alone, on one stream
    timer = FamilyTimer()
    eng = trainer.eng
    for blk in getattr(eng, 'blocks', []):
        register_geometry(eng.N, blk['geo'].H, blk['geo'].W)
    resident_step()                                   # eager warm-up outside the hook
    if per == 2:
        resident_step()
    ops.set_call_hook(timer)
    try:
        for _ in range(steps * per):
            resident_step()
    finally:
        ops.set_call_hook(None)
        trainer.use_graph = was_graph
        trainer.overlap = was_overlap
    ms, fl, by, cnt = timer.totals()
    if os.environ.get('DMC_DUMP_CALLS'):            # per-call times of one family, in launch order
        fam_dump = os.environ['DMC_DUMP_CALLS']
        import sys
        for name, e0, e1, f, _b in timer.events:
            if fam_dump in name:
                print('CALL %s %.1f us %.1f GFLOP' % (name, e0.elapsed_time(e1) * 1e3, f / 1e9), file=sys.stderr)
    denom = float(steps * per)
    breakdown = {k.replace('dmc_', ''): round(v / denom, 4) for k, v in sorted(ms.items(), key=lambda kv: -kv[1])}
    top = max(ms, key=lambda k: ms[k])
    # GEMM-space FLOPs: for the discriminator / ContextNetwork plans N and K are the PADDED column counts
    # (zero-padded channels and the structural zeros of the space-to-depth weight slices are issued MMAs,
    # not algorithmic work): those families are reported under other_families with that caveat, and the
    # headline roofline stays on the classifier's tap GEMM, whose N, K are the true channel counts.
    gemm = 'dmc_tc_tap_gemm'
    fam = gemm if gemm in ms else top
    out = {'breakdown': breakdown, 'top_family': top.replace('dmc_', '')}
    t_ms = ms[fam]
    achieved = fl[fam] / (t_ms * 1e-3) / 1e12 if t_ms > 0 else 0.0
    out['roofline'] = {
        'kernel': fam.replace('dmc_', '') + '_kernel', 'bound': 'tensor', 'achieved': achieved,
        'peak': peaks['tensor_tflops'], 'unit': 'TFLOP/s', 'frac': achieved / peaks['tensor_tflops'],
        'traffic': measured_traffic(), 'launches_per_step': cnt[fam] / denom, 'avg_launch_ms': t_ms / max(cnt[fam], 1),
        'share_of_step': t_ms / max(sum(ms.values()), 1e-9),
        'peak_source': peaks['source'],
        'note': 'algorithmic FLOPs (padding ring excluded); bf16x3 split issues 3 MMAs per MAC -> ceiling 1/3; '
                'traffic = DRAM bytes per launch (bytes) from the committed ncu capture',
    }
    # the CUDA-core conv families of the generator / discriminator (north star: report achieved HBM
    # GB/s; they are FMA-bound, so the fp32-FMA-pipe fraction is given alongside) and the wgrad GEMM
    others = {}
    fma = fma_peak_tflops()
    for k in ms:
        if fl[k] > 0 and k != fam:
            tf = fl[k] / (ms[k] * 1e-3) / 1e12
            o = {'tflops': tf, 'ms_per_step': ms[k] / denom,
                 'algorithmic_gbs': by[k] / (ms[k] * 1e-3) / 1e9,
                 'hbm_frac': by[k] / (ms[k] * 1e-3) / 1e9 / peaks['hbm_gbs']}
            if 'tc_' in k:
                o['tensor_frac'] = tf / peaks['tensor_tflops']
            else:
                o['fma_frac'] = tf / fma
            others[k.replace('dmc_', '')] = o
    out['roofline']['other_families'] = others
    return out
