"""CPU: the index tables of dmcnet_b200/disc_plan.py.  A numpy model of the tap GEMM
(out[q] = sum_t in[phase_t][q + shift_t] . Wg[t]^T on ring-padded pixel-major maps, exactly what
csrc/gemm_tc.cu computes) with the tables' operands must reproduce torch's conv2d forward, input
gradient and weight gradient for every layer kind, through the space-to-depth forms."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dmcnet_b200 import disc_plan as DP
from dmcnet_b200.engine import _taps_s2, disc_blocks


def ring(x_nhwc):
    """[M,H,W,C] -> shared-ring pixel-major [M,H+1,W+1,C] (row 0 / column 0 zero)."""
    m, h, w, c = x_nhwc.shape
    out = np.zeros((m, h + 1, w + 1, c), x_nhwc.dtype)
    out[:, 1:, 1:] = x_nhwc
    return out


def to_s2d4(x):        # [M,2,H,W] -> [M,H/4,W/4,64]
    m, c, h, w = x.shape
    o = np.zeros((m, h // 4, w // 4, 64), x.dtype)
    for a in range(4):
        for b in range(4):
            for ch in range(c):
                o[..., (a * 4 + b) * c + ch] = x[:, ch, a::4, b::4]
    return o


def from_s2d4(o, c):
    m, hg, wg, _ = o.shape
    x = np.zeros((m, c, hg * 4, wg * 4), o.dtype)
    for a in range(4):
        for b in range(4):
            for ch in range(c):
                x[:, ch, a::4, b::4] = o[..., (a * 4 + b) * c + ch]
    return x


def to_s2d(x):         # [M,C,2G,2G] -> [M,G,G,64]
    m, c, h, w = x.shape
    o = np.zeros((m, h // 2, w // 2, 64), x.dtype)
    for a in range(2):
        for b in range(2):
            o[..., (a * 2 + b) * c:(a * 2 + b + 1) * c] = x[:, :, a::2, b::2].transpose(0, 2, 3, 1)
    return o


def from_s2d(o, c):
    m, g, _, _ = o.shape
    x = np.zeros((m, c, 2 * g, 2 * g), o.dtype)
    for a in range(2):
        for b in range(2):
            x[:, :, a::2, b::2] = o[..., (a * 2 + b) * c:(a * 2 + b + 1) * c].transpose(0, 3, 1, 2)
    return x


def to_pm(x, cp):
    m, c, h, w = x.shape
    o = np.zeros((m, h, w, cp), x.dtype)
    o[..., :c] = x.transpose(0, 2, 3, 1)
    return o


def from_pm(o, c):
    return o[..., :c].transpose(0, 3, 1, 2)


FORM_TO = {'s2d4': lambda x, cp: to_s2d4(x), 's2d': lambda x, cp: to_s2d(x), 'pm': to_pm}
FORM_FROM = {'s2d4': from_s2d4, 's2d': from_s2d, 'pm': from_pm}


def shifted(flat, s):
    """rows q -> flat[q + s], zero outside [0, P) (TMA out-of-bounds fill)."""
    out = np.zeros_like(flat)
    P = flat.shape[0]
    lo, hi = max(0, -s), min(P, P - s)
    if hi > lo:
        out[lo:hi] = flat[lo + s:hi + s]
    return out


def interior_mask(m, hg, wg):
    k = np.zeros((m, hg + 1, wg + 1, 1))
    k[:, 1:, 1:] = 1
    return k.reshape(-1, 1)


@pytest.mark.parametrize('kind,cin,cout,hw', [('S4', 2, 16, 32), ('S1', 16, 16, 16), ('S2', 16, 32, 16),
                                              ('P1', 32, 32, 8), ('P2', 32, 64, 8), ('P1', 128, 128, 4)])
def test_tables_reproduce_conv_forward_and_both_gradients(kind, cin, cout, hw):
    rng = np.random.default_rng(0)
    M = 2
    stride = 2 if kind in ('S4', 'S2', 'P2') else 1
    x = torch.tensor(rng.standard_normal((M, cin, hw, hw)), requires_grad=True)
    w = torch.tensor(rng.standard_normal((cout, cin, 3, 3)) * 0.2, requires_grad=True)
    y = F.conv2d(x, w, None, stride, 1)
    dy = torch.tensor(rng.standard_normal(tuple(y.shape)))
    gx, gw = torch.autograd.grad(y, (x, w), dy)
    lp = DP.layer_plan(kind, cin, cout)
    fin = {'S4': 's2d4', 'S1': 's2d', 'S2': 's2d', 'P1': 'pm', 'P2': 'pm'}[kind]
    fout = {'S4': 's2d', 'S1': 's2d', 'S2': 'pm', 'P1': 'pm', 'P2': 'pm'}[kind]
    Kp, Np, g = lp['Kp'], lp['Np'], lp['gmap']
    wf = np.concatenate((w.detach().numpy().reshape(-1), [0.0]))
    Wg = wf[g]                                                      # [T][Np][Kp], -1 -> the appended zero
    ho = hw // stride
    hg = ho // 2 if fout == 's2d' else ho                           # output grid
    Wp = hg + 1
    keep = interior_mask(M, hg, hg)
    # ---- operands
    if kind == 'P2':
        xin = to_pm(x.detach().numpy(), Kp)
        phases = [ring(xin[:, ph::2, pw::2]).reshape(-1, Kp) for ph in (0, 1) for pw in (0, 1)]
        shift, phase, bsel = _taps_s2(Wp)
    else:
        phases = [ring(FORM_TO[fin](x.detach().numpy(), Kp)).reshape(-1, Kp)]
        shift = [di * Wp + dj for di, dj in lp['offsets']]
        phase, bsel = [0] * len(shift), list(range(len(shift)))
    # ---- forward
    out = sum(shifted(phases[phase[t]], shift[t]) @ Wg[bsel[t]].T for t in range(len(shift))) * keep
    got = FORM_FROM[fout](out.reshape(M, hg + 1, hg + 1, Np)[:, 1:, 1:], cout)
    np.testing.assert_allclose(got, y.detach().numpy(), atol=1e-10)
    if Np > lp['cmap'].max() + 1 and fout == 'pm':
        assert np.abs(out[:, cout:]).max() == 0                     # padding columns stay zero
    # ---- gradient operand in the output form
    G = ring(FORM_TO[fout](dy.numpy(), Np)).reshape(-1, Np)
    # weight gradient in GEMM space, gathered back through the inverse table
    dWg = np.stack([G.T @ shifted(phases[phase[t]], shift[t]) for t in range(len(shift))])   # [T][Np][Kp]
    order = np.argsort(bsel)
    dWg_by_slice = np.zeros_like(Wg)
    for t in range(len(shift)):
        dWg_by_slice[bsel[t]] += dWg[t]
    flat = np.concatenate((dWg_by_slice.reshape(-1), [0.0]))
    dW = flat[lp['inv']].sum(1).reshape(cout, cin, 3, 3)
    np.testing.assert_allclose(dW, gw.numpy(), atol=1e-9)
    # data gradient: dIn[ph][q'] = sum_{t: phase_t = ph} G[q' - shift_t] . Wg[bsel_t]
    dins = []
    for ph in range(len(phases)):
        acc = np.zeros((G.shape[0], Kp))
        for t in range(len(shift)):
            if phase[t] == ph:
                acc += shifted(G, -shift[t]) @ Wg[bsel[t]]
        dins.append(acc * keep)
    if kind == 'P2':
        full = np.zeros((M, hw, hw, Kp))
        for i, (ph, pw) in enumerate((a, b) for a in (0, 1) for b in (0, 1)):
            full[:, ph::2, pw::2] = dins[i].reshape(M, hg + 1, hg + 1, Kp)[:, 1:, 1:]
        got_dx = from_pm(full, cin)
    else:
        got_dx = FORM_FROM[fin](dins[0].reshape(M, hg + 1, hg + 1, Kp)[:, 1:, 1:], cin)
    np.testing.assert_allclose(got_dx, gx.numpy(), atol=1e-9)
    # folding tables
    for c in range(cout):
        cols = [j for j in lp['binv'][c] if j >= 0]
        assert cols and all(lp['cmap'][j] == c for j in cols)


@pytest.mark.parametrize('arch_d', ['Discriminator', 'Discriminator2', 'Discriminator3', 'Discriminator5'])
def test_whole_discriminator_plans_chain(arch_d):
    p = DP.plan(disc_blocks(arch_d), 224, 224)
    assert p[0]['kind'] == 'S4' and p[0]['grid'] == (56, 56)
    form = 's2d4'
    for lp in p:
        assert lp['form_in'] == form
        form = lp['form_out']
    assert p[-1]['form_out'] == 'pm' and p[-1]['grid'] == (14, 14) and p[-1]['Np'] == 128
    assert not DP.supported('Discriminator4', 224, 224) and DP.supported(arch_d, 224, 224)
