"""Pin the oracle restatement against the reference's own model.py.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference):

    python oracle/pin_against_reference.py

For dmcnet (code/dmcnet/model.py) and dmcnet_GAN (code/dmcnet_GAN/model.py) it
checks, on the same seeds, that
  * ``build_state`` reproduces the reference constructor's state_dict bit for bit,
  * eval- and train-mode forwards agree bit for bit (incl. Dropout2d draws and
    BatchNorm running-stat updates),
  * gradients and the post-Adam state of two consecutive restated steps agree
    with the same losses/optimizers driven through the reference ``Model``
    (optimizer wiring as code/dmcnet/train.py:121-142, GAN :122-153).
Exit code 0 = pinned.  ``tests/test_oracle_pin.py`` runs the same checks when
/root/reference is present and is skipped otherwise.
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import dmc_oracle as O                      # noqa: E402
from oracle import ref_loader as R                      # noqa: E402


def _maxdiff(a, b):
    return float((a.double() - b.double()).abs().max()) if a.numel() else 0.0


def check_state(variant, num_class, arch_d, arch_estimator='DenseNetTiny', att=0, ds=0):
    torch.manual_seed(1)
    kw = dict(base_model='resnet18', arch_estimator=arch_estimator, gen_flow_or_delta=1, use_databn=0,
              att=att, gen_flow_ds_factor=ds)
    if variant == 'dmcnet_GAN':
        kw['arch_d'] = arch_d
    ref = R.build_reference_model(variant, num_class, 3, 'mv', **kw)
    sd_ref = ref.state_dict()
    sd = O.build_state(num_class, arch_d if variant == 'dmcnet_GAN' else None, seed=1,
                       arch_estimator=arch_estimator, att=att, gen_flow_ds_factor=ds)
    assert list(sd.keys()) == list(sd_ref.keys()), (set(sd) ^ set(sd_ref))
    for k in sd:
        assert sd[k].shape == sd_ref[k].shape and torch.equal(sd[k], sd_ref[k]), k
    return ref, sd


def ref_optimizers(ref, hp, gan):
    groups = {'base_model': [], 'gen_flow_model': [], 'discriminator': []}
    mults = {'base_model': hp.lr_cls_mult, 'gen_flow_model': hp.lr_mse_mult, 'discriminator': hp.lr_d_mult}
    for key, value in dict(ref.named_parameters()).items():
        for tag in groups:
            if tag in key:
                groups[tag].append({'params': value, 'lr': hp.lr * mults[tag],
                                    'weight_decay': hp.weight_decay * (0.0 if 'bias' in key else 1.0)})
    mk = lambda g: torch.optim.Adam(g, eps=hp.eps)
    return mk(groups['base_model']), mk(groups['gen_flow_model']), (mk(groups['discriminator']) if gan else None)


def pin(variant, num_class, arch_d, batch=2, verbose=True, arch_estimator='DenseNetTiny', att=0, ds=0):
    """arch_estimator / att / ds: the generator choice (--arch_estimator, --att, --gen_flow_ds_factor);
    with att == 1 the flow criterion weights both sides by the attention map (train.py:244-247)."""
    gan = variant == 'dmcnet_GAN'
    ref, sd = check_state(variant, num_class, arch_d, arch_estimator, att, ds)
    gen_kw = dict(arch_estimator=arch_estimator, att=att, gen_flow_ds_factor=ds)
    use_att = att == 1 and arch_estimator == 'ContextNetwork'

    def flow_loss(gen_flow, fl, rest):
        if use_att:
            return F.mse_loss(rest[0] * gen_flow, rest[0] * fl)
        return F.mse_loss(gen_flow, fl)
    hp = O.HParams()
    flow, mv, res, target = O.make_inputs(batch, 3, num_class, seed=0)
    worst = 0.0

    # eval forward
    ref.eval()
    with torch.no_grad():
        r = ref(mv, res)
        o = O.model_forward({k: v.clone() for k, v in sd.items()}, mv, res, gan=gan, arch_d=arch_d, train=False,
                            **gen_kw)
    assert len(r) == len(o)
    for a, b in zip(r, o):
        worst = max(worst, _maxdiff(a, b))
        assert torch.equal(a, b), 'eval forward differs'

    if use_att:
        # The reference's ContextNetworkAtt cannot be trained on torch >= 1.x: predict_att ends in
        # LeakyReLU(inplace) -> ReLU(inplace) (code/dmcnet/model.py:94-97) and autograd refuses the
        # second in-place write ("modified by an inplace operation").  Only its forwards can be
        # pinned: eval above, train mode (batch statistics + running-stat update) here.
        ref.train()
        st = {k: v.clone() for k, v in sd.items()}
        with torch.no_grad():
            r = ref(mv, res)
            o = O.model_forward(st, mv, res, train=True, **gen_kw) if not gan else None   # GAN: Dropout2d draws
        if o is not None:
            for a, b in zip(r, o):
                assert torch.equal(a, b), 'train forward differs'
            rsd = ref.state_dict()
            for k in rsd:
                assert torch.equal(rsd[k], st[k]), ('running stats', k)
        if verbose:
            print('pinned %-11s C=%d arch_d=%s %s att=%d ds=%d  forwards only (reference backward raises)'
                  % (variant, num_class, arch_d, arch_estimator, att, ds))
        return worst

    # two train steps (GAN: D-step then G-step)
    ref.train()
    tr = O.OracleTrainer(sd, hp, gan=gan, arch_d=arch_d, **gen_kw)
    opt_cls, opt_gf, opt_d = ref_optimizers(ref, hp, gan)
    for it in range(2):
        torch.manual_seed(100 + it)
        fl = flow.view((-1,) + tuple(mv.shape[-3:]))
        if not gan:
            output, gen_flow, *rest = ref(mv, res)
            output = output.view((-1, 3) + tuple(output.shape[1:])).mean(1)
            loss = F.cross_entropy(output, target) * hp.lr_cls + flow_loss(gen_flow, fl, rest) * hp.lr_mse
            steppers = (opt_cls, opt_gf)
        else:
            valid = torch.ones(batch * 3, dtype=torch.int64)
            fake = torch.zeros_like(valid)
            if it % 2 == 0:
                output, validity, gen_flow, *rest = ref(mv, res, fl)
                output = output.view((-1, 3) + tuple(output.shape[1:])).mean(1)
                loss = F.cross_entropy(output, target) * hp.lr_cls + \
                    F.cross_entropy(validity, torch.cat((fake, valid), 0)) * hp.lr_adv_d
                steppers = (opt_cls, opt_d)
            else:
                output, validity, gen_flow, *rest = ref(mv, res)
                output = output.view((-1, 3) + tuple(output.shape[1:])).mean(1)
                loss = F.cross_entropy(output, target) * hp.lr_cls + \
                    F.cross_entropy(validity, valid) * hp.lr_adv_g + flow_loss(gen_flow, fl, rest) * hp.lr_mse
                steppers = (opt_gf,)
        for op in (opt_cls, opt_gf, opt_d):
            if op is not None:
                op.zero_grad(set_to_none=False)
        loss.backward()
        ref_grads = {k: v.grad.clone() for k, v in ref.named_parameters()}
        for op in steppers:
            op.step()

        torch.manual_seed(100 + it)                     # same Dropout2d draws
        m = tr.step(flow, mv, res, target)
        assert abs(m['loss'] - float(loss.detach())) <= 1e-6 * max(1.0, abs(float(loss.detach()))), m["loss"]
        og = tr.grads()
        for k in ref_grads:
            d = _maxdiff(ref_grads[k], og[k])
            worst = max(worst, d)
            assert d <= 1e-6 * (1.0 + float(ref_grads[k].abs().max())), ('grad', k, d)
        osd, rsd = tr.state_dict(), ref.state_dict()
        for k in rsd:
            d = _maxdiff(rsd[k].float(), osd[k].float())
            worst = max(worst, d)
            assert d <= 1e-6 * (1.0 + float(rsd[k].float().abs().max())), ('state', k, d)
    if verbose:
        print('pinned %-11s C=%d arch_d=%s %s att=%d ds=%d  worst |diff| = %.3g'
              % (variant, num_class, arch_d, arch_estimator, att, ds, worst))
    return worst


def main():
    assert R.reference_available(), 'needs /root/reference'
    torch.set_num_threads(os.cpu_count() or 1)
    pin('dmcnet', 51, None)
    pin('dmcnet_GAN', 101, 'Discriminator3')
    pin('dmcnet_GAN', 51, 'Discriminator', batch=1)
    for arch_d in ('Discriminator2', 'Discriminator4', 'Discriminator5'):
        check_state('dmcnet_GAN', 51, arch_d)
        print('state pinned dmcnet_GAN arch_d=%s' % arch_d)
    # every other generator choice: state, forwards, gradients, post-Adam state
    pin('dmcnet', 51, None, batch=1, arch_estimator='DenseNetSmall')
    pin('dmcnet', 51, None, batch=1, arch_estimator='DenseNetTinyEarlyFusionSum')
    pin('dmcnet', 51, None, batch=1, arch_estimator='DenseNetTinyEarlyFusionStack')
    pin('dmcnet', 51, None, batch=1, arch_estimator='ContextNetwork')
    pin('dmcnet', 51, None, batch=1, arch_estimator='ContextNetwork', att=1)
    pin('dmcnet', 51, None, batch=1, arch_estimator='ContextNetwork', att=1, ds=4)
    pin('dmcnet_GAN', 51, 'Discriminator', batch=1, arch_estimator='ContextNetwork')
    pin('dmcnet', 51, None, batch=1, arch_estimator='DenseNetTiny', ds=4)
    print('ORACLE PINNED against /root/reference model.py')


if __name__ == '__main__':
    main()
