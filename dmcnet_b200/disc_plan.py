"""Host-side plan of the discriminators on the tensor-core path (csrc/disc_pm.cu, csrc/gemm_tc.cu).

For every block of ``Discriminator`` / ``Discriminator2`` / ``3`` / ``5``
(code/dmcnet_GAN/model.py:254-438) this module decides the layout ("form") of its input and output
maps and builds the integer tables that tell the gather kernels which OIHW weight sits at which
position of the GEMM operand:

  form 's2d4'  the 2-channel 224x224 input as a 56x56 grid of 4x4 pixel blocks:
               column (a*4 + b)*2 + c  =  x[c][4i + a][4j + b]            (32 of 64 columns used)
  form 's2d'   a C <= 16 channel map at 2G x 2G as a G x G grid of 2x2 blocks:
               column (a*2 + b)*C + c  =  x[c][2i + a][2j + b]            (4C of 64 columns used)
  form 'pm'    a C-channel map at its own resolution, zero-padded to 64 / 128 columns

A conv between two forms is a tap GEMM  out[q][n] = sum_t in[phase_t][q + shift_t][:] . Wg[t][n][:]
over grid pixels q; ``gmap[t][n][k]`` is the flat OIHW index of the weight at that position or -1.
Nothing here touches the device; the tables are plain numpy and are checked on the CPU against
``torch.nn.functional.conv2d`` (tests/test_disc_plan.py).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

R_INV = 4            # a true weight appears in at most 4 GEMM positions (the 4 output phases of 's2d')


def pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def supported(arch_d: str, height: int, width: int) -> bool:
    """Whether the plan can express this discriminator: the four 16/32/64/128-channel ones, at sizes that
    are multiples of 16 (Discriminator4's 8/16/32 channels have no space-to-depth form here)."""
    return arch_d in ('Discriminator', 'Discriminator2', 'Discriminator3', 'Discriminator5') \
        and height % 16 == 0 and width % 16 == 0


def preferred(arch_d: str, height: int, width: int) -> bool:
    """Whether the engine picks the tensor-core plan by default.  Discriminator5 (20 BatchNorm'd blocks) is
    left on the fp32 planar kernels: each BatchNorm backward amplifies the 2^-17 operand rounding of the
    bf16 hi/lo GEMMs, and through five blocks per stage the gradients of its early blocks were measured
    3e-2 .. 2e-1 away from the float64 oracle (1e-2 for Discriminator3); `disc_engine='tc'` still forces it."""
    return supported(arch_d, height, width) and arch_d != 'Discriminator5'


def _split2(v: int) -> Tuple[int, int]:
    """full-resolution offset a + r - 1 in {-1, 0, 1, 2} -> (grid offset, phase) on a 2x2 grid"""
    return {-1: (-1, 1), 0: (0, 0), 1: (0, 1), 2: (1, 0)}[v]


def _split4(v: int) -> Tuple[int, int]:
    """full-resolution offset 2a + r - 1 in {-1 .. 3} -> (grid offset, phase) on a 4x4 grid"""
    return (-1, 3) if v == -1 else (0, v)


def layer_plan(kind: str, cin: int, cout: int) -> Dict[str, object]:
    """kind: 'S4' s2d4 -> s2d, stride 2 | 'S1' s2d -> s2d, stride 1 | 'S2' s2d -> pm, stride 2 |
    'P1' pm -> pm, stride 1 | 'P2' pm -> pm, stride 2 (phase-split input, taps of engine._taps_s2).
    Returns offsets (list of (di, dj) grid offsets, one per GEMM tap; P2: kernel taps r*3+s), Kp, Np,
    gmap [T][Np][Kp], cmap [Np] (true output channel per column), kform / nform."""
    oihw = lambda co, ci, r, s: ((co * cin + ci) * 3 + r) * 3 + s
    if kind == 'S4':
        assert 4 * cout <= 64 and 16 * cin <= 64
        offs = [(di, dj) for di in (-1, 0) for dj in (-1, 0)]
        Kp, Np = 64, 64
        g = -np.ones((len(offs), Np, Kp), np.int32)
        for a in range(2):
            for b in range(2):
                for r in range(3):
                    for s in range(3):
                        (di, a4), (dj, b4) = _split4(2 * a + r - 1), _split4(2 * b + s - 1)
                        t = offs.index((di, dj))
                        for co in range(cout):
                            for ci in range(cin):
                                g[t, (a * 2 + b) * cout + co, (a4 * 4 + b4) * cin + ci] = oihw(co, ci, r, s)
        cmap = np.array([j % cout if j < 4 * cout else -1 for j in range(Np)], np.int32)
    elif kind == 'S1':
        assert 4 * cout <= 64 and 4 * cin <= 64
        offs = [(di, dj) for di in (-1, 0, 1) for dj in (-1, 0, 1)]
        Kp, Np = 64, 64
        g = -np.ones((9, Np, Kp), np.int32)
        for a in range(2):
            for b in range(2):
                for r in range(3):
                    for s in range(3):
                        (di, a2), (dj, b2) = _split2(a + r - 1), _split2(b + s - 1)
                        t = offs.index((di, dj))
                        for co in range(cout):
                            for ci in range(cin):
                                g[t, (a * 2 + b) * cout + co, (a2 * 2 + b2) * cin + ci] = oihw(co, ci, r, s)
        cmap = np.array([j % cout if j < 4 * cout else -1 for j in range(Np)], np.int32)
    elif kind == 'S2':
        assert 4 * cin <= 64
        offs = [(di, dj) for di in (-1, 0) for dj in (-1, 0)]
        Kp, Np = 64, pad64(cout)
        g = -np.ones((4, Np, Kp), np.int32)
        for r in range(3):
            for s in range(3):
                (di, a2), (dj, b2) = _split2(r - 1), _split2(s - 1)
                t = offs.index((di, dj))
                for co in range(cout):
                    for ci in range(cin):
                        g[t, co, (a2 * 2 + b2) * cin + ci] = oihw(co, ci, r, s)
        cmap = np.array([j if j < cout else -1 for j in range(Np)], np.int32)
    elif kind in ('P1', 'P2'):
        offs = [(r - 1, s - 1) for r in range(3) for s in range(3)]
        Kp, Np = pad64(cin), pad64(cout)
        g = -np.ones((9, Np, Kp), np.int32)
        co, ci = np.meshgrid(np.arange(cout), np.arange(cin), indexing='ij')
        for r in range(3):
            for s in range(3):
                g[r * 3 + s, :cout, :cin] = ((co * cin + ci) * 3 + r) * 3 + s
        cmap = np.array([j if j < cout else -1 for j in range(Np)], np.int32)
    else:
        raise ValueError(kind)
    # inverse tables: OIHW element -> its (<= 4) GEMM positions; true channel -> its columns
    n_w = cout * cin * 9
    inv = -np.ones((n_w, R_INV), np.int32)
    fill = np.zeros(n_w, np.int32)
    flat = g.reshape(-1)
    pos = np.nonzero(flat >= 0)[0]
    for p in pos:
        e = flat[p]
        inv[e, fill[e]] = p
        fill[e] += 1
    assert fill.min() >= 1 and fill.max() <= R_INV
    binv = -np.ones((cout, R_INV), np.int32)
    bfill = np.zeros(cout, np.int32)
    for j, c in enumerate(cmap):
        if c >= 0:
            binv[c, bfill[c]] = j
            bfill[c] += 1
    return {'kind': kind, 'offsets': offs, 'Kp': Kp, 'Np': Np, 'gmap': g, 'cmap': cmap, 'inv': inv,
            'binv': binv, 'cin': cin, 'cout': cout}


def plan(blocks: List[Tuple[str, int, int, int, bool]], height: int, width: int) -> List[Dict[str, object]]:
    """Per block of ``engine.disc_blocks(arch_d)``: kind, grid size (Hg, Wg) of its OUTPUT map, true
    output size (Ho, Wo) and the tables of ``layer_plan``."""
    out = []
    h, w = height, width
    form = 's2d4'
    for name, cin, cout, stride, bn in blocks:
        ho, wo = h // stride, w // stride
        if form == 's2d4':
            assert stride == 2 and cout <= 16
            kind, nform, grid = 'S4', 's2d', (ho // 2, wo // 2)
        elif form == 's2d' and stride == 1:
            kind, nform, grid = 'S1', 's2d', (ho // 2, wo // 2)
        elif form == 's2d':
            kind, nform, grid = 'S2', 'pm', (ho, wo)
        elif stride == 1:
            kind, nform, grid = 'P1', 'pm', (ho, wo)
        else:
            kind, nform, grid = 'P2', 'pm', (ho, wo)
        lp = layer_plan(kind, cin, cout)
        lp.update(name=name, stride=stride, bn=bn, form_in=form, form_out=nform, grid=grid, out_hw=(ho, wo))
        out.append(lp)
        form, h, w = nform, ho, wo
    return out
