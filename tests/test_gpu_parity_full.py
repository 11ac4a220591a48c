"""GPU: the product against the CPU oracle at the BASELINE configurations themselves (B=64), the
gradient error as a function of the batch size, and a 20-step training trajectory.

Bars
  * forward outputs, losses, running statistics: 1e-3 relative (north star), argmax / prec@k exact;
  * gradients that do not pass through the ResNet-18 backward (generator in dmcnet, discriminator
    in D-steps): 1e-4;
  * gradients downstream of the ResNet-18 backward: set from the measured B=64 figure, see
    GRAD_BARS -- tests/test_grad_sensitivity.py shows the oracle's OWN gradients move by 5e-3 under an
    ulp-level perturbation of its conv operands and by 1.2e-2 under the 2^-17 operand rounding of the
    tensor-core split, independent of the batch size (switch flips: the flipped fraction, not the
    count, sets the relative L2), and tests/test_gpu_backward_exact.py shows the product's backward
    agrees with autograd to ~1e-5 once the forward state is identical.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import dmc_oracle as O          # noqa: E402  (checker only)

if torch.cuda.is_available():
    from dmcnet_b200.engine import DmcEngine
    from dmcnet_b200.trainer import FusedTrainStep, HParams

# per-tensor relative L2 of gradients downstream of the ResNet-18 backward: (median, worst)
GRAD_BARS = (3e-2, 5e-2)     # measured at B=64 on B200: 1.41e-2 / 1.89e-2 (oracle's own figure: 1.2e-2 / 1.8e-2)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rel2(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def _record(name, payload):
    if os.path.isdir('gpurun_out'):
        path = 'gpurun_out/r02_parity_full.json'
        data = {}
        if os.path.isfile(path):
            with open(path) as f:
                data = json.load(f)
        data[name] = payload
        with open(path, 'w') as f:
            json.dump(data, f, indent=1)


def _pair(num_class, arch_d, batch):
    gan = arch_d is not None
    sd = O.build_state(num_class, arch_d, seed=1)
    ref = O.OracleTrainer(sd, O.HParams(), gan=gan, arch_d=arch_d)
    eng = DmcEngine(num_class, 3, batch * 3, gan=gan, arch_d=arch_d)
    eng.load_state(sd)
    tr = FusedTrainStep(eng, HParams(), batch)
    return ref, eng, tr


def _forward_checks(ref, eng, tr, mo, mg, gan):
    for k in mo:
        if k in ('prec1', 'prec5', 'acc_adv'):
            assert mg[k] == pytest.approx(mo[k], abs=1e-9), k
        else:
            assert mg[k] == pytest.approx(mo[k], rel=1e-3, abs=1e-6), k
    errs = {'gen_flow': rel(eng.gen_flow, ref.last_gen_flow), 'consensus': rel(tr.consensus, ref.last_output)}
    assert torch.equal(tr.consensus.argmax(1).cpu(), ref.last_output.argmax(1))
    if gan:
        m = ref.last_validity.shape[0]
        errs['validity'] = rel(eng.validity[:m], ref.last_validity)
        assert torch.equal(eng.validity[:m].argmax(1).cpu(), ref.last_validity.argmax(1))
    for k, e in errs.items():
        assert e < 1e-3, (k, e)
    return errs


def _grad_errors(ref, eng, prefix):
    og = ref.grads()
    return {k: rel2(eng.grad_view(k), og[k]) for k in eng.specs
            if k.startswith(prefix) and float(og[k].abs().max()) > 0}


def _disc_grads_fp64(ref, eng, arch_d, masks, n, ref_flow):
    """d(lr_adv_d * CE(validity, [fake, real]))/d(discriminator parameters) from the oracle's
    functions evaluated in float64 on the D input the product saw (first fake, then real)."""
    import torch.nn.functional as F
    st0 = ref.pre_step_state
    st = {}
    for k, v in st0.items():
        if k.startswith('discriminator'):
            t = v.double() if v.is_floating_point() else v.clone()
            st[k] = t.requires_grad_(True) if (v.is_floating_point() and not O.is_buffer(k)) else t
    x = torch.cat((eng.gen_flow.detach(), ref_flow.reshape(-1, 2, eng.H, eng.W).cuda()), 0).double().cpu()   # [fake | real]
    validity = O.disc_forward(st, x, arch_d, True, [m.double() for m in masks])
    tgt = torch.cat((torch.zeros(n, dtype=torch.int64), torch.ones(n, dtype=torch.int64)))
    (F.cross_entropy(validity, tgt) * O.HParams().lr_adv_d).backward()
    return {k: t.grad for k, t in st.items() if t.requires_grad}


def test_config2_b64_train_step_vs_oracle():
    """BASELINE config 2 at full size: dmcnet train step, B=64 x 3 segments, 51 classes."""
    B = 64
    ref, eng, tr = _pair(51, None, B)
    flow, mv, res, target = O.make_inputs(B, 3, 51, seed=0)
    mo = ref.step(flow, mv, res, target)
    mg = tr.step(flow.cuda(), mv.cuda(), res.cuda(), target.cuda())
    errs = _forward_checks(ref, eng, tr, mo, mg, False)
    gen = _grad_errors(ref, eng, 'gen_flow_model')
    assert max(gen.values()) < 1e-4, gen                           # MSE path only (classifier input detached)
    cls = _grad_errors(ref, eng, 'base_model')
    med, worst = float(np.median(list(cls.values()))), max(cls.values())
    _record('config2_b64', {'forward': errs, 'gen_grad_worst': max(gen.values()),
                            'cls_grad_median': med, 'cls_grad_worst': worst, 'metrics': mg})
    assert med < GRAD_BARS[0] and worst < GRAD_BARS[1], (med, worst)
    osd, gsd = ref.state_dict(), eng.state_dict()
    for k in osd:                                                   # post-Adam state
        if k.endswith(('running_mean', 'running_var')) or k.startswith('gen_flow_model'):
            assert rel(gsd[k].float(), osd[k].float()) < 1e-3, k
        elif k.endswith('num_batches_tracked'):
            assert int(gsd[k]) == int(osd[k])


def test_config3_b64_gan_d_and_g_step_vs_oracle():
    """BASELINE config 3 at full size: dmcnet_GAN, Discriminator3, B=64, 101 classes; D-step then G-step."""
    B, arch_d = 64, 'Discriminator3'
    ref, eng, tr = _pair(101, arch_d, B)
    flow, mv, res, target = O.make_inputs(B, 3, 101, seed=0)
    rec = {}
    for it in range(2):
        torch.manual_seed(100 + it)
        masks = O.draw_dropout_masks(arch_d, B * 3 * (2 if it == 0 else 1))
        ref.pre_step_state = ref.state_dict()
        mo = ref.step(flow, mv, res, target, masks=masks, apply=False)
        mg = tr.step(flow.cuda(), mv.cuda(), res.cuda(), target.cuda(), masks=masks, apply=False)
        errs = _forward_checks(ref, eng, tr, mo, mg, True)
        if it == 0:       # D-step: discriminator gradients do not depend on the ResNet backward
            d = _grad_errors(ref, eng, 'discriminator')
            # 1e-2 against both evaluations of the oracle: at 384 frames the fp32 oracle is itself 1e-3 ..
            # 2.7e-3 away from its float64 evaluation upstream of block_2 (torch's CPU BatchNorm backward
            # sums 1.2 M fp32 terms per channel), and the tensor-core plan carries 4e-3 .. 5e-3 there
            # (2^-17 operand rounding amplified by block_3's BatchNorm backward; tests/test_gpu_disc_tc.py)
            assert max(d.values()) < 1e-2, d
            d64 = _disc_grads_fp64(ref, eng, arch_d, masks, B * 3, flow)
            e64 = {k: rel2(eng.grad_view(k), d64[k]) for k in d64}
            assert max(e64.values()) < 1e-2, e64
            rec['D'] = {'forward': errs, 'disc_grad_worst_vs_fp32_oracle': max(d.values()),
                        'disc_grad_worst_vs_fp64_oracle': max(e64.values()), 'metrics': mg}
        else:             # G-step: generator gradient = MSE + adversarial (through D) + CE (through ResNet-18)
            g = _grad_errors(ref, eng, 'gen_flow_model')
            med, worst = float(np.median(list(g.values()))), max(g.values())
            rec['G'] = {'forward': errs, 'gen_grad_median': med, 'gen_grad_worst': worst, 'metrics': mg}
            assert med < GRAD_BARS[0] and worst < GRAD_BARS[1], (med, worst)
    _record('config3_b64', rec)


def test_gradient_error_vs_batch():
    """Per-tensor relative L2 error of the classifier gradients against the oracle at B = 2, 8, 64
    (dmcnet step).  Recorded in profiles/r02_parity_full.json; the bar is the same at every size
    because the effect (switch flips under a ~1e-5 forward difference) is a FRACTION of the
    activations, see tests/test_grad_sensitivity.py."""
    table = {}
    for B in (2, 8, 64):
        ref, eng, tr = _pair(51, None, B)
        flow, mv, res, target = O.make_inputs(B, 3, 51, seed=0)
        ref.step(flow, mv, res, target, apply=False)
        tr.step(flow.cuda(), mv.cuda(), res.cuda(), target.cuda(), apply=False)
        cls = _grad_errors(ref, eng, 'base_model')
        table[B] = {'median': float(np.median(list(cls.values()))), 'worst': max(cls.values()),
                    'worst_key': max(cls, key=cls.get), 'logits': rel(tr.consensus, ref.last_output)}
        del eng, tr
        torch.cuda.empty_cache()
    _record('grad_error_vs_batch', table)
    for B, t in table.items():
        assert t['median'] < GRAD_BARS[0] and t['worst'] < GRAD_BARS[1], (B, t)


def test_twenty_step_trajectory_vs_oracle():
    """20 Adam steps on a fixed batch (B=2), product vs oracle: losses, consensus argmax and every
    parameter / running statistic -- a slow drift of the loose gradient path would show here."""
    B, steps = 2, 20
    ref, eng, tr = _pair(51, None, B)
    flow, mv, res, target = O.make_inputs(B, 3, 51, seed=0)
    fc, mc, rc, tc = flow.cuda(), mv.cuda(), res.cuda(), target.cuda()
    hist = []
    for it in range(steps):
        mo = ref.step(flow, mv, res, target)
        mg = tr.step(fc, mc, rc, tc)
        hist.append({k: (mo[k], mg[k]) for k in ('loss', 'loss_cls', 'loss_mse')})
        for k in ('loss_cls', 'loss_mse', 'loss'):
            assert mg[k] == pytest.approx(mo[k], rel=2e-3, abs=1e-6), (it, k)
        # top-k can only be compared while the oracle's own decision is not a near-tie: late in the
        # trajectory two logits of one clip come within the accumulated difference of the two runs
        lo = ref.last_output
        top = lo.topk(6, 1).values
        margin = float((top[:, :-1] - top[:, 1:]).min() / lo.abs().max())
        dist = rel(tr.consensus, lo)
        if margin > 3 * dist:
            assert mg['prec1'] == mo['prec1'] and mg['prec5'] == mo['prec5'], (it, margin, dist)
        # the two runs drift apart as the ~1e-2 gradient difference passes through Adam step after step
        # (6 frames per BatchNorm batch); recorded per step in profiles/r02_parity_full.json
        assert dist < (1e-3 if it == 0 else 1e-1), (it, dist)
        hist[-1]['consensus_rel_err'] = dist
    osd, gsd = ref.state_dict(), eng.state_dict()
    init = O.build_state(51, None, seed=1)
    hp = HParams()
    worst_w, worst_travel = {}, {}
    for k in osd:
        if k.endswith('num_batches_tracked'):
            assert int(gsd[k]) == steps == int(osd[k])
            continue
        a, b = gsd[k].double().cpu(), osd[k].double().cpu()
        worst_w[k] = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        if not O.is_buffer(k):
            # Adam(eps 1e-3) moves an element by at most ~lr per step whatever the gradient size, so
            # elements whose gradient is near zero follow the gradient NOISE: measure the distance
            # between the two trajectories in units of the distance travelled, lr * steps
            lr = hp.lr * (hp.lr_cls_mult if k.startswith('base_model') else hp.lr_mse_mult)
            worst_travel[k] = float((a - b).abs().max() / (lr * steps))
    nz = {k: v for k, v in worst_w.items() if float(init[k].double().abs().max()) > 0}      # not zero-initialised
    bad, far = max(nz, key=nz.get), max(worst_travel, key=worst_travel.get)
    _record('trajectory_20_steps', {'history': hist, 'worst_state_rel': nz[bad], 'worst_key': bad,
                                    'worst_fraction_of_travel': worst_travel[far], 'worst_travel_key': far})
    assert nz[bad] < 1e-2, (bad, nz[bad])                       # weights, BN weights, running statistics
    assert worst_travel[far] < 0.5, (far, worst_travel[far])   # measured 0.18 (a zero-initialised BN bias)
    assert hist[-1]['loss'][1] < hist[0]['loss'][1]


def test_config4_b512_on_one_gpu_equals_b64_on_a_tiled_batch():
    """BASELINE config 4 at N = 1 (the strong-scaling base: all 512 clips, 3072 discriminator frames, on one
    GPU) cannot be compared with the CPU oracle in reasonable time, so it is held to a size-independent
    property instead: a batch made of 8 copies of a 64-clip batch (Dropout2d masks tiled alike) has the same
    BatchNorm statistics, hence the same losses, logits and -- the losses being batch means -- the same
    gradients as the 64-clip batch.  Catches 32-bit index / grid-size errors at the largest shapes."""
    arch_d, nc, b, reps = 'Discriminator3', 51, 64, 8
    sd = O.build_state(nc, arch_d, seed=1)
    flow, mv, res, target = O.make_inputs(b, 3, nc, seed=0)
    torch.manual_seed(5)
    masks = [O.draw_dropout_masks(arch_d, b * 3 * 2), O.draw_dropout_masks(arch_d, b * 3)]
    out = {}
    for B in (b, b * reps):
        r = B // b
        eng = DmcEngine(nc, 3, B * 3, gan=True, arch_d=arch_d)
        eng.load_state(sd)
        tr = FusedTrainStep(eng, HParams(), B)
        f, m, rs, t = (x.repeat((r,) + (1,) * (x.dim() - 1)).cuda() for x in (flow, mv, res, target))
        rec = []
        for it in range(2):                                   # D-step, then G-step, same parameters
            n = b * 3
            if it == 0:    # D input = [fake frames | real frames]: tile each half
                mk = [torch.cat((k[:n].repeat(r, 1), k[n:].repeat(r, 1))) for k in masks[0]]
            else:
                mk = [k.repeat(r, 1) for k in masks[1]]
            mt = tr.step(f, m, rs, t, masks=mk, apply=False)
            rec.append((mt, tr.consensus[:b].clone().cpu(), eng.grads.clone().cpu()))
        out[B] = rec
        del eng, tr
        torch.cuda.empty_cache()
    for it in range(2):
        (m1, c1, g1), (m8, c8, g8) = out[b][it], out[b * reps][it]
        for k in m1:
            assert m8[k] == pytest.approx(m1[k], rel=1e-4, abs=1e-6), (it, k)
        assert rel(c8, c1) < 1e-4, it
        # gradients: identical up to the summation order of the batch statistics, i.e. up to the switch
        # sensitivity of tests/test_grad_sensitivity.py (5e-3 for one ulp)
        assert rel2(g8, g1) < 3e-2, it
    _record('config4_b512_vs_b64_tiled', {'loss_D': out[b * reps][0][0]['loss'], 'loss_G': out[b * reps][1][0]['loss'],
                                           'grad_rel_l2_D': rel2(out[b * reps][0][2], out[b][0][2]),
                                           'grad_rel_l2_G': rel2(out[b * reps][1][2], out[b][1][2])})
