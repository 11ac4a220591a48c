"""CPU: the oracle restatement reproduces the committed reference outputs
(tests/golden/*.npz, produced by tests/golden/make_golden.py from the
reference's own model.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import dmc_oracle as O
from oracle.digest import digest_close

RTOL = 1e-5     # oracle vs reference on the same torch build: ~bit-identical


def test_infer_cfg1(golden_dir):
    g = np.load(os.path.join(golden_dir, 'infer_cfg1.npz'))
    sd = O.build_state(51, None, seed=1)
    for i, v in enumerate(sd.values()):
        digest_close(g['init_state'][i], v.float(), RTOL, 'init %d' % i)
    flow, mv, res, target = O.make_inputs(1, 3, 51, seed=0)
    with torch.no_grad():
        base_out, gen_flow = O.model_forward(sd, mv, res, train=False)
    np.testing.assert_allclose(base_out.numpy(), g['base_out'], rtol=RTOL, atol=1e-6)
    scores = O.infer_video_scores(sd, mv, res, 3)
    np.testing.assert_allclose(scores.numpy(), g['scores'], rtol=RTOL, atol=1e-6)
    assert np.array_equal(scores.argmax(1).numpy(), g['argmax'])
    digest_close(g['gen_flow'], gen_flow, RTOL, 'gen_flow')


@pytest.mark.parametrize('name,num_class,arch_d,batch,arch_estimator', [
    ('train_dmcnet_b2.npz', 51, None, 2, 'DenseNetTiny'),
    ('train_gan_d3_b2.npz', 101, 'Discriminator3', 2, 'DenseNetTiny'),
    ('train_gan_d_b1.npz', 51, 'Discriminator', 1, 'DenseNetTiny'),
    ('train_context_b1.npz', 51, None, 1, 'ContextNetwork'),
])
def test_train_steps(golden_dir, name, num_class, arch_d, batch, arch_estimator):
    g = np.load(os.path.join(golden_dir, name))
    gan = arch_d is not None
    sd = O.build_state(num_class, arch_d, seed=1, arch_estimator=arch_estimator)
    assert list(sd.keys()) == list(g['keys'])
    tr = O.OracleTrainer(sd, O.HParams(), gan=gan, arch_d=arch_d, arch_estimator=arch_estimator)
    flow, mv, res, target = O.make_inputs(batch, 3, num_class, seed=0)
    pkeys = list(g['param_keys'])
    for it in range(2):
        torch.manual_seed(100 + it)
        m = tr.step(flow, mv, res, target)
        for k in ('loss', 'loss_cls', 'loss_mse', 'loss_adv', 'prec1', 'prec5'):
            gk = 's%d_%s' % (it, k)
            if gk in g:
                assert m[k] == pytest.approx(float(g[gk]), rel=RTOL, abs=1e-7), gk
        np.testing.assert_allclose(tr.last_output.numpy(), g['s%d_output' % it], rtol=RTOL, atol=1e-6)
        if gan:
            np.testing.assert_allclose(tr.last_validity.numpy(), g['s%d_validity' % it], rtol=RTOL, atol=1e-6)
        digest_close(g['s%d_gen_flow' % it], tr.last_gen_flow, RTOL, 'gen_flow')
        grads = tr.grads()
        for i, k in enumerate(pkeys):
            digest_close(g['s%d_grads' % it][i], grads[k], RTOL, 'grad ' + k)
        for i, (k, v) in enumerate(tr.state_dict().items()):
            digest_close(g['s%d_state' % it][i], v.float(), RTOL, 'state ' + k)
