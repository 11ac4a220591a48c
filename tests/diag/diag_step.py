"""GPU diagnostics: FusedTrainStep vs the CPU oracle, per-tensor error report.

    python tests/diag/diag_step.py [dmcnet|gan_d3|gan_d] [batch] [tc|simt]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dmc_oracle as O          # noqa: E402  (checker only)
from dmcnet_b200.engine import DmcEngine    # noqa: E402
from dmcnet_b200.trainer import FusedTrainStep, HParams  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rel2(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else 'dmcnet'
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    gemm = sys.argv[3] if len(sys.argv) > 3 else 'tc'
    num_class, arch_d = {'dmcnet': (51, None), 'gan_d3': (101, 'Discriminator3'),
                         'gan_d': (51, 'Discriminator')}[which]
    gan = arch_d is not None
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.build_state(num_class, arch_d, seed=1)
    flow, mv, res, target = O.make_inputs(batch, 3, num_class, seed=0)
    hp_o = O.HParams()
    tr_o = O.OracleTrainer(sd, hp_o, gan=gan, arch_d=arch_d)
    eng = DmcEngine(num_class, 3, batch * 3, gan=gan, arch_d=arch_d, gemm_engine=gemm)
    eng.load_state(sd)
    tr = FusedTrainStep(eng, HParams(), batch)
    assert list(eng.state_keys()) == list(sd.keys()), 'state_dict key order differs'
    worst = 0.0
    for it in range(2):
        masks = None
        if gan:
            torch.manual_seed(100 + it)
            m = batch * 3 * (2 if it % 2 == 0 else 1)
            masks = O.draw_dropout_masks(arch_d, m)
        t0 = time.time()
        mo = tr_o.step(flow, mv, res, target, masks=masks)
        t1 = time.time()
        mg = tr.step(flow.cuda(), mv.cuda(), res.cuda(), target.cuda(), masks=masks)
        torch.cuda.synchronize()
        print('--- step %d (%s)  oracle %.1fs  ours %.2fs' % (it, which, t1 - t0, time.time() - t1))
        for k in sorted(mo):
            print('  %-9s oracle %.6f  ours %.6f' % (k, mo[k], mg.get(k, float('nan'))))
        print('  gen_flow rel err %.3e' % rel(eng.gen_flow, tr_o.last_gen_flow))
        print('  consensus rel err %.3e' % rel(tr.consensus, tr_o.last_output))
        if gan:
            print('  validity rel err %.3e' % rel(eng.validity[:tr_o.last_validity.shape[0]], tr_o.last_validity))
        og = tr_o.grads()
        stepped = tr._step_groups('full' if not gan else ('D' if it % 2 == 0 else 'G'))
        rows = []
        for k in eng.specs:
            if not any(k.startswith(g) for g in stepped):
                continue
            rows.append((rel(eng.grad_view(k), og[k]), k, rel2(eng.grad_view(k), og[k])))
        if os.environ.get('DIAG_ALL'):
            for e, k, sc in rows:
                print('    ALL %.3e  %-50s (L2 rel %.3e)' % (e, k, sc))
        rows.sort(reverse=True)
        print('  worst grads:')
        for e, k, s in rows[:8]:
            print('    %.3e  %-50s (L2 rel %.3e)' % (e, k, s))
        print('  worst L2-rel grad: %.3e   median max-rel %.3e' % (max(r[2] for r in rows), rows[len(rows) // 2][0]))
        worst = max(worst, rows[0][0])
        osd = tr_o.state_dict()
        gsd = eng.state_dict()
        srows = sorted(((rel(gsd[k].float(), osd[k].float()), k) for k in osd), reverse=True)
        print('  worst state:')
        for e, k in srows[:6]:
            print('    %.3e  %s' % (e, k))
        worst = max(worst, srows[0][0])
    print('DIAG_STEP_DONE %s worst=%.3e' % (which, worst))


if __name__ == '__main__':
    main()
