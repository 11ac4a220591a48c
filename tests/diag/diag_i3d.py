"""Per-stage activation errors and per-tensor gradient errors (relative L2, in module order) of the I3D
engine against the oracle at B=1 -- run on a GPU box: python tests/diag/diag_i3d.py [seed]."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dmcnet_b200 import ops
from dmcnet_b200.i3d_engine import I3DEngine
from oracle import i3d_oracle as O


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    sd = O.build_state(51, 'DenseNetTiny', seed=1)
    beta = float(os.environ.get('I3D_BN_BIAS', '0'))
    if beta:                      # push every pre-activation far from the ReLU switch: no switch noise
        for k in sd:
            if k.endswith('batch3d.bias'):
                sd[k].fill_(beta)
    eng = I3DEngine(51, B, 16)
    eng.load_state(sd)
    data, target = O.make_inputs(B, 16, 51, seed=0)
    st = {k: (v.clone() if O.is_buffer(k) else v.clone().requires_grad_(True)) for k, v in sd.items()}
    rec = {}
    logits_o, flow_o = O.i3d_forward(st, data[:, :5], train=True, record=rec)
    (F.cross_entropy(logits_o, target) + F.mse_loss(flow_o, data[:, 5:7])).backward()
    # the same step in float64: how far is the fp32 oracle itself from the exact gradient?
    st64 = {k: (v.clone().double() if v.is_floating_point() else v.clone()) for k, v in sd.items()}
    for k in st64:
        if not O.is_buffer(k):
            st64[k].requires_grad_(True)
    l64, f64 = O.i3d_forward(st64, data[:, :5].double(), train=True)
    (F.cross_entropy(l64, target) + F.mse_loss(f64, data[:, 5:7].double())).backward()
    dev = eng.device
    eng.zero_grads()
    eng.forward_data(data.cuda(), train=True)
    ops.ce_head(eng.logits, B, 1, 51, target.cuda(), 1.0 / B, torch.zeros(B, 51, device=dev), eng.d_logits,
                torch.zeros(4, device=dev))
    numel = eng.N * 2 * 224 * 224
    ops.mse_head(eng.gen_flow, eng.in_flow, numel, 2.0 / numel, eng.dD, torch.zeros(1, dtype=torch.float64, device=dev),
                 frame_elems=2 * 224 * 224, dgen_ns=eng.dD.shape[1] * 224 * 224)
    eng.backward(eng.N, cls=True, cls_wgrad=True, gen_grad=True, cls_to_gen=True)
    torch.cuda.synchronize()
    print('logits', rel_l2(eng.logits, logits_o))
    stages = [('conv3d_1a_7x7', eng.m_stem, list(range(64))), ('conv3d_2b_1x1', eng.m_2b, list(range(64))),
              ('conv3d_2c_3x3', eng.m_2c, list(range(192)))]
    for i, M in enumerate(eng.mixed):
        cols = eng.mixed[i + 1]['b0'].in_cols if i + 1 < len(eng.mixed) else list(range(1024))
        stages.append((M['name'], M['m_cat'], cols))
    for name, m, cols in stages:
        geo = m['geo']
        a = (m['hi'].float() + m['lo'].float()).view(geo.P, m['width'])[:, cols]
        a = a.reshape(B, geo.Tp, geo.Hp, geo.Wp, -1)[:, 1:1 + geo.T, 1:, 1:].permute(0, 4, 1, 2, 3)
        flips = float(((a.cpu() > 0) != (rec[name] > 0)).float().mean())
        print('act  %-16s rel %.2e  relu-switch flips %.2e' % (name, rel_l2(a, rec[name]), flips))
    e1 = sorted(rel_l2(eng.grad_view(k), st[k].grad) for k in eng.specs)
    e2 = sorted(rel_l2(eng.grad_view(k), st64[k].grad) for k in eng.specs)
    e3 = sorted(rel_l2(st[k].grad, st64[k].grad) for k in eng.specs)
    print('SUMMARY B=%d: engine vs oracle median %.2e worst %.2e | engine vs fp64 median %.2e worst %.2e | '
          'fp32 oracle vs fp64 median %.2e worst %.2e' % (B, e1[len(e1) // 2], e1[-1], e2[len(e2) // 2], e2[-1],
                                                          e3[len(e3) // 2], e3[-1]))
    for k in eng.specs:
        print('grad %-44s vs oracle %.2e | engine vs fp64 %.2e | fp32 oracle vs fp64 %.2e | |g| %.2e'
              % (k, rel_l2(eng.grad_view(k), st[k].grad), rel_l2(eng.grad_view(k), st64[k].grad),
                 rel_l2(st[k].grad, st64[k].grad), float(st[k].grad.norm())))


if __name__ == '__main__':
    main()
