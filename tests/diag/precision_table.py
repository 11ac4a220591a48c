"""Measured error of cheaper tensor-core operand formats for the ResNet-18 convs (VERDICT r01 item 4,
SURVEY section 7 "decide per layer from the measured error").  CPU experiment on the oracle: the
operands of the selected convs are rounded the way the candidate MMA scheme would see them
(oracle.OPERAND_HOOK, straight-through gradients), everything else stays exact fp32; reported are the
relative error of the consensus logits (north-star bar 1e-3) and the per-tensor relative L2 change of
the classifier gradients (median / worst).  Candidates per k-step, cost in bf16-rate MMAs:

  bf16x3   hi*hi + lo*hi + hi*lo   3 MMAs   (shipped)            operands ~2^-17
  tf32     one kind::tf32 MMA      2        (half-rate)          operands 2^-11 (10-bit mantissa)
  bf16x2a  hi*hi + lo*hi           2        activations split, weights bf16 (2^-9)
  bf16x2w  hi*hi + hi*lo           2        weights split, activations bf16 (2^-9)
  bf16x1   hi*hi                   1        both bf16

    python tests/diag/precision_table.py [batch]      -> profiles/r02_precision_table.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dmc_oracle as O          # noqa: E402  (diagnostic: the oracle is the subject)


def r_bf16(d):
    return d.to(torch.bfloat16).float()


def r_hilo(d):
    hi = r_bf16(d)
    return hi + r_bf16(d - hi)


def r_tf32(d):
    i = d.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF                                   # round to nearest, 10-bit mantissa
    return i.view(torch.float32)


SCHEMES = {'bf16x3': (r_hilo, r_hilo), 'tf32': (r_tf32, r_tf32), 'bf16x2a': (r_hilo, r_bf16),
           'bf16x2w': (r_bf16, r_hilo), 'bf16x1': (r_bf16, r_bf16)}

# conv index in forward order -> stage name (stem, layer1 .. layer4); 20 convs
STAGE_OF = ['stem'] + ['layer1'] * 4 + ['layer2'] * 5 + ['layer3'] * 5 + ['layer4'] * 5


class Hook:
    """OPERAND_HOOK is called as hook(activation), hook(weight) per conv, in forward order."""

    def __init__(self, scheme, stages):
        self.act, self.wgt = SCHEMES[scheme]
        self.base_a, self.base_w = SCHEMES['bf16x3']
        self.stages, self.calls = stages, 0

    def __call__(self, x):
        conv, is_w = (self.calls // 2) % 20, self.calls % 2
        self.calls += 1
        sel = STAGE_OF[conv] in self.stages
        f = (self.wgt if is_w else self.act) if sel else (self.base_w if is_w else self.base_a)
        d = x.detach()
        return x + (f(d) - d)


def run(batch, hook):
    sd = O.build_state(51, None, seed=1)
    flow, mv, res, target = O.make_inputs(batch, 3, 51, seed=0)
    tr = O.OracleTrainer(sd, O.HParams())
    O.OPERAND_HOOK = hook
    try:
        tr.step(flow, mv, res, target, apply=False)
    finally:
        O.OPERAND_HOOK = None
    return tr.grads(), tr.last_output


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    torch.set_num_threads(os.cpu_count() or 1)
    g0, out0 = run(batch, None)
    rows = []
    groups = [('all', {'stem', 'layer1', 'layer2', 'layer3', 'layer4'}), ('layer1', {'layer1'}),
              ('layer2', {'layer2'}), ('layer3', {'layer3'}), ('layer4', {'layer4'}),
              ('layer3+4', {'layer3', 'layer4'})]
    for scheme in ('bf16x3', 'tf32', 'bf16x2a', 'bf16x2w', 'bf16x1'):
        for gname, stages in groups:
            if scheme == 'bf16x3' and gname != 'all':
                continue
            g1, out1 = run(batch, Hook(scheme, stages))
            e = [float((g1[k].double() - g0[k].double()).norm() / g0[k].double().norm())
                 for k in g0 if k.startswith('base_model') and float(g0[k].abs().max()) > 0]
            row = {'scheme': scheme, 'layers': gname,
                   'logits_rel_err': float((out1 - out0).abs().max() / out0.abs().max()),
                   'argmax_same': bool(torch.equal(out1.argmax(1), out0.argmax(1))),
                   'grad_rel_l2_median': float(np.median(e)), 'grad_rel_l2_worst': float(max(e))}
            rows.append(row)
            print('%-8s %-9s logits %.2e  grads median %.2e worst %.2e' % (
                scheme, gname, row['logits_rel_err'], row['grad_rel_l2_median'], row['grad_rel_l2_worst']), flush=True)
    out = {'batch': batch, 'note': 'rows other than "all" apply the scheme to the named stage only; every other conv '
                                   'keeps the shipped bf16x3 operands', 'rows': rows}
    with open(os.path.join(ROOT, 'profiles', 'r02_precision_table.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
