"""Pin oracle/i3d_oracle.py against the reference's own I3D (code/dmcnet_I3D/network/i3d.py) imported from
/root/reference -- build container only (TEST INFRASTRUCTURE).  Checks, bit for bit: state_dict keys, shapes
and seeded random init; logits and generated flow of I3D.forward(node='flow+logit') in train mode; every
parameter gradient of CE + MSE; the BatchNorm running statistics after the forward.  `python -m
oracle.pin_i3d --write` also writes the reference's outputs to tests/golden/i3d_b1.npz."""
import os
import sys
import warnings

import torch
import torch.nn.functional as F

REF = '/root/reference/code/dmcnet_I3D/network'


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF, 'i3d.py'))


def load_reference():
    import importlib.util
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        sys.path.insert(0, REF)
        try:
            spec = importlib.util.spec_from_file_location('_ref_i3d', os.path.join(REF, 'i3d.py'))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        finally:
            sys.path.remove(REF)
    return mod


def run(write: bool = False, clip_len: int = 16, num_class: int = 51):
    from oracle import i3d_oracle as O
    ref = load_reference()
    torch.manual_seed(1)
    net = ref.I3D(num_class, modality='flow+mp4', dropout_prob=0, arch_estimator='DenseNetTiny')
    net.train()
    sd_ref = net.state_dict()
    sd = O.build_state(num_class, 'DenseNetTiny', seed=1)
    assert list(sd.keys()) == list(sd_ref.keys()), 'state_dict keys / order differ'
    for k in sd:
        assert sd[k].shape == sd_ref[k].shape and torch.equal(sd[k], sd_ref[k]), k
    data, target = O.make_inputs(1, clip_len, num_class, seed=0)
    # reference
    logits_r, flow_r = net(data[:, :5], node='flow+logit')
    loss_r = F.cross_entropy(logits_r, target) + F.mse_loss(flow_r, data[:, 5:7])
    loss_r.backward()
    g_ref = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    # restatement
    st = {k: (v.clone() if O.is_buffer(k) else v.clone().requires_grad_(True)) for k, v in sd.items()}
    logits, flow = O.i3d_forward(st, data[:, :5], arch_estimator='DenseNetTiny', train=True)
    loss = F.cross_entropy(logits, target) + F.mse_loss(flow, data[:, 5:7])
    loss.backward()
    assert torch.equal(logits, logits_r) and torch.equal(flow, flow_r), 'forward differs'
    for k, g in g_ref.items():
        assert torch.equal(st[k].grad, g), 'gradient of %s differs' % k
    after = net.state_dict()
    for k in sd:
        if O.is_buffer(k):
            assert torch.equal(st[k], after[k]), k
    # eval forward
    net.eval()
    with torch.no_grad():
        le_r, _ = net(data[:, :5], node='flow+logit')
        le, _ = O.i3d_forward(st, data[:, :5], arch_estimator='DenseNetTiny', train=False)
    assert torch.equal(le, le_r), 'eval forward differs'
    # with a discriminator: construction order / init, and I3D.forward(x, node='D')
    from oracle.dmc_oracle import disc_forward
    torch.manual_seed(1)
    netd = ref.I3D(num_class, modality='flow+mp4', dropout_prob=0, arch_estimator='DenseNetTiny', arch_d='Discriminator')
    sdd = O.build_state(num_class, 'DenseNetTiny', seed=1, arch_d='Discriminator')
    assert list(sdd.keys()) == list(netd.state_dict().keys())
    for k, v in netd.state_dict().items():
        assert torch.equal(sdd[k], v), k
    netd.eval()
    xs = torch.randn(3, 2, 224, 224)
    with torch.no_grad():
        assert torch.equal(netd(xs, node='D'), disc_forward({k: v.clone() for k, v in sdd.items()}, xs, 'Discriminator', False))
    if write:
        import numpy as np
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'i3d_b1.npz')
        rec = {'logits': logits_r.detach().numpy(), 'logits_eval': le_r.numpy(), 'loss': np.float64(float(loss_r)),
               'flow_sum': np.float64(float(flow_r.double().sum())), 'flow_abs': np.float64(float(flow_r.double().abs().sum()))}
        for k, g in g_ref.items():
            rec['gnorm/' + k] = np.float64(float(g.double().norm()))
            rec['gsum/' + k] = np.float64(float(g.double().sum()))
        for k in ('conv3d_1a_7x7.batch3d.running_mean', 'mixed_5c.branch_3.1.batch3d.running_var'):
            rec['buf/' + k] = after[k].numpy()
        np.savez_compressed(out, **rec)
        print('wrote', out)
    return len(g_ref)


if __name__ == '__main__':
    n = run(write='--write' in sys.argv)
    print('i3d oracle pinned: state, forward, %d gradients, running statistics bit-identical' % n)
