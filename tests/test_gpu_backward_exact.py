"""GPU: the classifier BACKWARD of the product against autograd of the oracle, with the forward
state made identical.

End-to-end gradient comparisons (test_gpu_parity.py) are limited by the discontinuity of the
ResNet backward in the forward activations: tests/test_grad_sensitivity.py shows that perturbing
the ORACLE's own conv operands at the fp32 ulp level already moves its per-tensor gradients by
5e-3 (ReLU / max-pool switch flips, square-root law in the perturbation size), so no forward that is
not bit-identical can be held to 1e-3 there.  This test removes that effect instead of widening the
bar: the oracle's forward intermediates (raw conv outputs, activations, batch statistics, max-pool
argmax, logits) are loaded into the engine's saved-state buffers, then ONLY the product's backward
runs -- every switch is the oracle's, and what remains is the arithmetic of the backward kernels
(bf16x3 tensor-core dgrad / wgrad GEMMs with fused BatchNorm-backward epilogues, stem, heads) over
the whole network at once.  Bar: per-tensor relative L2 <= 1e-3 (north-star tolerance; measured
values are recorded in profiles/r02_backward_exact.json).
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import dmc_oracle as O          # noqa: E402  (checker only)

if torch.cuda.is_available():
    from dmcnet_b200 import ops
    from dmcnet_b200.engine import DmcEngine
    from dmcnet_b200.trainer import HParams, loss_grad_scales


def rel2(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def pixel(x):      # [N,C,H,W] -> padded pixel-major [N*Hp*Wp, C] (cuda fp32)
    n, c, h, w = x.shape
    p = torch.zeros(n, ops.padded(h), ops.padded(w), c, device='cuda')
    p[:, 1:h + 1, 1:w + 1, :] = x.cuda().permute(0, 2, 3, 1)
    return p.reshape(-1, c).contiguous()


def split(x):
    hi = x.to(torch.bfloat16)
    return hi.contiguous(), (x - hi.float()).to(torch.bfloat16).contiguous()


def bn_stats(y, gamma, beta, eps=1e-5):
    yd = y.double()
    mean = yd.mean((0, 2, 3))
    var = yd.var((0, 2, 3), unbiased=False)
    invstd = (var + eps).rsqrt()
    scale = gamma.double() * invstd
    return [t.float().cuda().contiguous() for t in (mean, invstd, scale, beta.double() - mean * scale)]


def load_unit(u, y, gamma, beta):
    u.Y.copy_(pixel(y))
    mean, invstd, scale, shift = bn_stats(y, gamma, beta)
    u.mean.copy_(mean); u.invstd.copy_(invstd); u.scale.copy_(scale); u.shift.copy_(shift)


def load_forward_state(eng, cap, st, n):
    """Oracle intermediates -> the buffers DmcEngine._cls_backward reads."""
    eng.stem_Y.copy_(cap['conv1'].cuda())
    mean, invstd, scale, shift = bn_stats(cap['conv1'], st['base_model.bn1.weight'].detach(),
                                          st['base_model.bn1.bias'].detach())
    for k, v in (('mean', mean), ('invstd', invstd), ('scale', scale), ('shift', shift)):
        eng.stem[k].copy_(v)
    # arg-max byte r*3+s of the 3x3/2 window (csrc/stem.cu) from ATen's flat input index
    idx = cap['pool_idx']                                  # [N,64,56,56] -> h*W + w of the 112x112 map
    W2 = eng.W // 2
    hq = torch.arange(idx.shape[2]).view(1, 1, -1, 1)
    wq = torch.arange(idx.shape[3]).view(1, 1, 1, -1)
    r = idx // W2 - (2 * hq - 1)
    s = idx % W2 - (2 * wq - 1)
    assert int(r.min()) >= 0 and int(r.max()) <= 2 and int(s.min()) >= 0 and int(s.max()) <= 2
    eng.pool_idx.copy_((r * 3 + s).permute(0, 2, 3, 1).contiguous().view(-1).to(torch.uint8).cuda())
    hi, lo = split(pixel(cap['pool']))
    eng.A0_hi.copy_(hi); eng.A0_lo.copy_(lo)
    x_hi, x_lo = eng.A0_hi, eng.A0_lo
    for blk in eng.blocks:
        q, gi = blk['name'], blk['geo_in']
        g = lambda k: st[k].detach()
        if 'ds' in blk:
            ops.phase_split(x_hi, x_lo, n, gi.H, gi.W, blk['cin'], blk['xp_hi'], blk['xp_lo'])
            load_unit(blk['ds'], cap[q + '.downsample.0'], g(q + '.downsample.1.weight'), g(q + '.downsample.1.bias'))
        load_unit(blk['c1'], cap[q + '.conv1'], g(q + '.bn1.weight'), g(q + '.bn1.bias'))
        hi, lo = split(pixel(cap[q + '.act1']))
        blk['c1'].act_hi.copy_(hi); blk['c1'].act_lo.copy_(lo)
        load_unit(blk['c2'], cap[q + '.conv2'], g(q + '.bn2.weight'), g(q + '.bn2.bias'))
        hi, lo = split(pixel(cap[q + '.out']))
        blk['c2'].act_hi.copy_(hi); blk['c2'].act_lo.copy_(lo)
        x_hi, x_lo = blk['c2'].act_hi, blk['c2'].act_lo
    eng.pooled.copy_(cap['pooled'].cuda())


RESULTS = {}


@pytest.mark.parametrize('batch', [2, 16, 64])
def test_classifier_backward_with_oracle_forward_state(batch):
    num_class, S = 51, 3
    n = batch * S
    sd = O.build_state(num_class, None, seed=1)
    flow, mv, res, target = O.make_inputs(batch, S, num_class, seed=0)
    eng = DmcEngine(num_class, S, n)
    eng.load_state(sd)
    eng.forward(mv.cuda(), res.cuda(), train=True)            # generator + weight operands + buffers
    x0 = eng.gen_flow.detach().cpu().clone().requires_grad_(True)
    # oracle: ResNet-18 forward / backward from the SAME classifier input
    st = {k: (v.clone().requires_grad_(True) if not O.is_buffer(k) else v.clone()) for k, v in sd.items()}
    cap = {}
    logits = O.resnet18_forward(st, x0, True, capture=cap)
    out = logits.view(batch, S, num_class).mean(1)
    hp = HParams()
    (F.cross_entropy(out, target) * hp.lr_cls).backward()
    # product: backward only, on the oracle's forward state
    load_forward_state(eng, cap, st, n)
    eng.logits.copy_(logits.detach().cuda())
    cons = torch.zeros(batch, num_class, device='cuda')
    stats = torch.zeros(4, device='cuda')
    sc = loss_grad_scales(hp, batch, 1, n, eng.H, eng.W)
    ops.ce_head(eng.logits, batch, S, num_class, target.cuda(), sc['cls'], cons, eng.d_logits, stats)
    eng.zero_grads()
    ops.memset_zero(eng.dD)
    eng.backward(n, cls=True, cls_wgrad=True, gen_grad=False, cls_to_gen=True)
    torch.cuda.synchronize()
    errs = {}
    for k in eng.specs:
        if k.startswith('base_model'):
            errs[k] = rel2(eng.grad_view(k), st[k].grad)
    errs['d_gen_flow'] = rel2(eng.dD[:, 0:2], x0.grad)
    worst = max(errs, key=errs.get)
    RESULTS[batch] = {'median': float(np.median(list(errs.values()))), 'worst': errs[worst], 'worst_key': worst}
    if os.path.isdir('gpurun_out'):
        with open('gpurun_out/r02_backward_exact.json', 'w') as f:
            json.dump({'per_batch': RESULTS, 'last_per_tensor': errs}, f, indent=1)
    assert errs[worst] < 1e-3, (worst, errs[worst])
    assert RESULTS[batch]['median'] < 2e-4, RESULTS[batch]
