"""Generate tests/golden/*.npz from the REFERENCE's own model.py (build container only).

    python tests/golden/make_golden.py

The reference ``Model`` (imported unmodified from /root/reference, see
oracle/ref_loader.py) is driven by the step bodies restated from
code/dmcnet/train.py:221-266 and code/dmcnet_GAN/train.py:237-372, with the
optimizer wiring of train.py:121-142 / GAN :122-153.  Inputs and weights are
seed-generated (oracle.make_inputs seed 0, constructor under manual_seed(1),
Dropout2d draws under manual_seed(100+it)), so only the *outputs* are stored:
scalars in full, tensors as digests (oracle/digest.py).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import dmc_oracle as O                       # noqa: E402
from oracle import ref_loader as R                       # noqa: E402
from oracle.digest import digest                         # noqa: E402
from oracle.pin_against_reference import ref_optimizers  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def build(variant, num_class, arch_d, arch_estimator='DenseNetTiny'):
    torch.manual_seed(1)
    kw = dict(base_model='resnet18', arch_estimator=arch_estimator, gen_flow_or_delta=1, use_databn=0)
    if variant == 'dmcnet_GAN':
        kw['arch_d'] = arch_d
    return R.build_reference_model(variant, num_class, 3, 'mv', **kw)


def golden_infer():
    """BASELINE config 1: single clip, eval forward, 51 classes (test.py:139-151)."""
    ref = build('dmcnet', 51, None).eval()
    flow, mv, res, target = O.make_inputs(1, 3, 51, seed=0)
    with torch.no_grad():
        base_out, gen_flow = ref(mv, res)
    scores = base_out.view(-1, 3, 51).mean(1)
    np.savez(os.path.join(HERE, 'infer_cfg1.npz'),
             base_out=base_out.numpy(), scores=scores.numpy(),
             argmax=scores.argmax(1).numpy(), gen_flow=digest(gen_flow),
             init_state=np.stack([digest(v.float()) for v in ref.state_dict().values()]))


def golden_train(variant, num_class, arch_d, batch, name, steps=2, arch_estimator='DenseNetTiny'):
    gan = variant == 'dmcnet_GAN'
    ref = build(variant, num_class, arch_d, arch_estimator).train()
    hp = O.HParams()
    flow, mv, res, target = O.make_inputs(batch, 3, num_class, seed=0)
    fl = flow.view((-1,) + tuple(mv.shape[-3:]))
    opt_cls, opt_gf, opt_d = ref_optimizers(ref, hp, gan)
    out = {'keys': np.array(list(ref.state_dict().keys())),
           'param_keys': np.array([k for k, _ in ref.named_parameters()])}
    for it in range(steps):
        torch.manual_seed(100 + it)
        scal = {}
        if not gan:
            output, gen_flow = ref(mv, res)
            output = output.view((-1, 3) + tuple(output.shape[1:])).mean(1)
            loss_cls = F.cross_entropy(output, target)
            loss_mse = F.mse_loss(gen_flow, fl)
            loss = loss_cls * hp.lr_cls + loss_mse * hp.lr_mse
            steppers = (opt_cls, opt_gf)
            scal['loss_mse'] = float(loss_mse.detach())
        else:
            valid = torch.ones(batch * 3, dtype=torch.int64)
            fake = torch.zeros_like(valid)
            if it % 2 == 0:
                output, validity, gen_flow = ref(mv, res, fl)
                output = output.view((-1, 3) + tuple(output.shape[1:])).mean(1)
                loss_cls = F.cross_entropy(output, target)
                loss_adv = F.cross_entropy(validity, torch.cat((fake, valid), 0))
                loss = loss_cls * hp.lr_cls + loss_adv * hp.lr_adv_d
                steppers = (opt_cls, opt_d)
            else:
                output, validity, gen_flow = ref(mv, res)
                output = output.view((-1, 3) + tuple(output.shape[1:])).mean(1)
                loss_cls = F.cross_entropy(output, target)
                loss_adv = F.cross_entropy(validity, valid)
                loss_mse = F.mse_loss(gen_flow, fl)
                loss = loss_cls * hp.lr_cls + loss_adv * hp.lr_adv_g + loss_mse * hp.lr_mse
                steppers = (opt_gf,)
                scal['loss_mse'] = float(loss_mse.detach())
            scal['loss_adv'] = float(loss_adv.detach())
            out['s%d_validity' % it] = validity.detach().numpy()
        for op in (opt_cls, opt_gf, opt_d):
            if op is not None:
                op.zero_grad(set_to_none=False)
        loss.backward()
        scal['loss'] = float(loss.detach())
        scal['loss_cls'] = float(loss_cls.detach())
        p1, p5 = O.accuracy(output.detach(), target, topk=(1, 5))
        scal['prec1'], scal['prec5'] = p1, p5
        for k, v in scal.items():
            out['s%d_%s' % (it, k)] = np.float64(v)
        out['s%d_output' % it] = output.detach().numpy()
        out['s%d_gen_flow' % it] = digest(gen_flow)
        out['s%d_grads' % it] = np.stack([
            digest(p.grad if p.grad is not None else torch.zeros_like(p))
            for _, p in ref.named_parameters()])
        for op in steppers:
            op.step()
        out['s%d_state' % it] = np.stack([digest(v.float()) for v in ref.state_dict().values()])
    np.savez(os.path.join(HERE, name), **out)


def main():
    assert R.reference_available(), 'needs /root/reference'
    torch.set_num_threads(os.cpu_count() or 1)
    golden_infer()
    golden_train('dmcnet', 51, None, 2, 'train_dmcnet_b2.npz')
    golden_train('dmcnet_GAN', 101, 'Discriminator3', 2, 'train_gan_d3_b2.npz')
    golden_train('dmcnet_GAN', 51, 'Discriminator', 1, 'train_gan_d_b1.npz')
    # the default generator: fixture for the kernels of the next round
    golden_train('dmcnet', 51, None, 1, 'train_context_b1.npz', arch_estimator='ContextNetwork')
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
