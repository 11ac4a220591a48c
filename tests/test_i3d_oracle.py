"""CPU checks of the I3D oracle (oracle/i3d_oracle.py) and of the host logic around the I3D engine: the
reference-generated fixture, the pin against the live reference (build container only), the drop-in module's
state_dict, parameter grouping and the two-stage learning-rate rule."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import i3d_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'i3d_b1.npz')


def test_oracle_reproduces_the_reference_generated_fixture():
    gold = np.load(GOLD)
    sd = O.build_state(51, 'DenseNetTiny', seed=1)
    data, target = O.make_inputs(1, 16, 51, seed=0)
    st = {k: (v.clone() if O.is_buffer(k) else v.clone().requires_grad_(True)) for k, v in sd.items()}
    logits, flow = O.i3d_forward(st, data[:, :5], train=True)
    loss = F.cross_entropy(logits, target) + F.mse_loss(flow, data[:, 5:7])
    loss.backward()
    assert np.allclose(logits.detach().numpy(), gold['logits'], rtol=0, atol=1e-6)
    assert abs(float(loss) - float(gold['loss'])) < 1e-6
    assert abs(float(flow.double().sum()) - float(gold['flow_sum'])) < 1e-3 * abs(float(gold['flow_abs']))
    for k in ('classifier.weight', 'conv3d_1a_7x7.conv3d.weight', 'mixed_4c.branch_2.1.batch3d.bias',
              'gen_flow_model.conv_0.0.weight'):
        assert abs(float(st[k].grad.double().norm()) - float(gold['gnorm/' + k])) < 1e-4 * float(gold['gnorm/' + k]), k
    for k in ('conv3d_1a_7x7.batch3d.running_mean', 'mixed_5c.branch_3.1.batch3d.running_var'):
        assert np.allclose(st[k].numpy(), gold['buf/' + k], rtol=1e-6, atol=1e-8)


def test_oracle_pinned_against_the_live_reference():
    from oracle import pin_i3d
    if not pin_i3d.reference_available():
        pytest.skip('/root/reference is only present in the build container')
    assert pin_i3d.run(write=False) == 187


def test_dropin_module_state_dict_and_groups():
    from dmcnet_b200.i3d_model import I3D, build_i3d_state
    from dmcnet_b200.i3d_trainer import lr_mult_rule, param_group_of
    a, b = build_i3d_state(51, 'DenseNetTiny', seed=1), O.build_state(51, 'DenseNetTiny', seed=1)
    assert list(a.keys()) == list(b.keys()) and all(torch.equal(a[k], b[k]) for k in a)
    plain = I3D(101, modality='flow')
    assert not hasattr(plain, 'gen_flow_model') and plain.classifier.weight.shape == (101, 400)
    keys = [k for k in b if not O.is_buffer(k)]
    gf, base, new = O.param_groups(keys)
    for k in keys:
        assert param_group_of(k) == ('gf' if k in gf else 'new' if k in new else 'base')
    assert len(gf) == 12 and len(new) == 4 and len(base) == 171
    # train/model.py:268-283
    assert lr_mult_rule(0.2, 0, 1) == 0.0 and lr_mult_rule(0.2, 1, 1) == 0.2
    assert lr_mult_rule(0.5, 0, 1) == 0.0 and lr_mult_rule(0.5, 3, 1) == 1.0
    assert lr_mult_rule(1.0, 0, 1) == 1.0 and lr_mult_rule(0.5, 0, 0) == 1.0
    for a_, b_ in ((0.2, 0), (0.5, 2), (1.0, 1)):
        assert lr_mult_rule(a_, b_, 2) == O.lr_mult_rule(a_, b_, 2)
    with pytest.raises(RuntimeError):
        plain(torch.zeros(1, 2, 16, 224, 224))          # CPU tensor: there is no CPU path


def test_oracle_adversarial_stages_keep_the_other_groups_gradients():
    """fit() zeroes only the optimizers a stage steps (train/model.py:387-446): after a D stage the generator's
    gradient is still in .grad (and becomes part of the G step), after the G stage the classifier's and the
    discriminator's are."""
    from oracle import dmc_oracle as O2
    sd = O.build_state(51, 'DenseNetTiny', seed=1, arch_d='Discriminator')
    assert len(sd) == 383
    tr = O.I3DOracleTrainer(sd, O.I3DHParams(optim='sgd', epoch_thre=0, dropout=0.0, adv=1.0), arch_d='Discriminator')
    tr.set_epoch(1)
    g = torch.Generator().manual_seed(3)
    data, target = O.make_inputs(1, 16, 51, seed=3)
    m = tr.step(data, target, disc_masks=O2.draw_dropout_masks('Discriminator', 32, g))
    assert m['stage'] == 'D' and m['stepped'] and 0.3 < m['loss_adv'] < 2.0
    gr = tr.grads()
    assert all(k.startswith('gen_flow_model') for k in gr) and len(gr) == 12
    before = {k: v.clone() for k, v in tr.state_dict().items()}
    m = tr.step(data, target, disc_masks=O2.draw_dropout_masks('Discriminator', 32, g))
    assert m['stage'] == 'G' and m['stepped']
    gr = tr.grads()
    assert not any(k.startswith('gen_flow_model') for k in gr) and 'classifier.weight' in gr \
        and 'discriminator.adv_layer.weight' in gr
    after = tr.state_dict()
    assert not torch.equal(after['gen_flow_model.conv_0.0.weight'], before['gen_flow_model.conv_0.0.weight'])
    assert torch.equal(after['classifier.weight'], before['classifier.weight'])
    assert torch.equal(after['discriminator.adv_layer.weight'], before['discriminator.adv_layer.weight'])
