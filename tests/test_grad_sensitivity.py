"""CPU: how far do the gradients of the hot path move when the FORWARD is not bit-identical?

The ResNet-18 backward is a discontinuous function of the forward activations (ReLU and max-pool
switches); train-mode BatchNorm turns one flipped switch into a per-channel offset that the weight
gradient sums coherently.  The product's forward carries ~2^-17 relative operand error (bf16 hi/lo
split on the tensor cores).  Here the ORACLE's own ResNet convs are given operands rounded the same
way (oracle.OPERAND_HOOK, straight-through) and its per-tensor gradients are compared with the
exact-fp32 oracle's, at several batch sizes.  This is the evidence behind the end-to-end gradient
bars of tests/test_gpu_parity*.py: they are set from what ANY 2^-17-accurate forward does to the
reference's own gradient.  Measured here (8 host cores, torch 2.11 CPU):

    perturbation (relative, per conv operand)   median / worst per-tensor rel. L2 of base_model grads
    2^-14 uniform                               4.9e-2 / 7.0e-2
    2^-17 uniform                               1.6e-2 / 2.3e-2      (bf16 hi/lo rounding: 1.2e-2 / 1.8e-2)
    2^-20 uniform                               4.5e-3 / 6.5e-3
    2^-23 uniform  (one fp32 ulp)               5.0e-3 / 7.5e-3      <- floor: fp32 re-association

i.e. a square-root law (a flipped FRACTION of switches ~ eps gives a relative L2 ~ sqrt(eps)) down to
a floor of 5e-3 that even an fp32 summation-order change produces; and the figure does not depend on
the batch size (B = 2, 4, 8, 16: 1.16e-2, 1.06e-2, 1.03e-2, 1.06e-2) because the flipped fraction,
not the count, sets it.  The north star's 1e-3 is therefore a bar for forward outputs; gradients
behind the ResNet backward are held to 1e-3 with the forward state made identical
(tests/test_gpu_backward_exact.py) and to the sensitivity figure end to end.
"""
import numpy as np
import pytest
import torch

from oracle import dmc_oracle as O


def hilo_round(x):
    """x -> bf16(x) + bf16(x - bf16(x)) (|error| <= 2^-17 |x|), gradient passed straight through."""
    d = x.detach()
    hi = d.to(torch.bfloat16).float()
    lo = (d - hi).to(torch.bfloat16).float()
    return x + ((hi + lo) - d)


def per_tensor_rel_l2(ga, gb, prefix):
    out = {}
    for k in ga:
        if k.startswith(prefix) and float(gb[k].abs().max()) > 0:
            out[k] = float((ga[k].double() - gb[k].double()).norm() / gb[k].double().norm())
    return out


def oracle_grads(batch, hook, gan=False, steps=1, num_class=51, seed=0):
    arch_d = 'Discriminator' if gan else None
    sd = O.build_state(num_class, arch_d, seed=1)
    flow, mv, res, target = O.make_inputs(batch, 3, num_class, seed=seed)
    tr = O.OracleTrainer(sd, O.HParams(), gan=gan, arch_d=arch_d)
    O.OPERAND_HOOK = hook
    try:
        for it in range(steps):
            masks = None
            if gan:
                torch.manual_seed(100 + it)
                masks = O.draw_dropout_masks(arch_d, batch * 3 * (2 if it % 2 == 0 else 1))
            m = tr.step(flow, mv, res, target, masks=masks, apply=(it < steps - 1))
    finally:
        O.OPERAND_HOOK = None
    return tr.grads(), m, tr.last_output


def sensitivity(batch, gan=False):
    """median / worst per-tensor relative L2 change of the gradients downstream of the ResNet backward."""
    steps = 2 if gan else 1                       # GAN: the G-step is where the generator sees ResNet's dgrad
    g0, m0, out0 = oracle_grads(batch, None, gan, steps)
    g1, m1, out1 = oracle_grads(batch, hilo_round, gan, steps)
    prefix = 'gen_flow_model' if gan else 'base_model'
    e = per_tensor_rel_l2(g1, g0, prefix)
    fwd = float((out1 - out0).abs().max() / out0.abs().max())
    return float(np.median(list(e.values()))), max(e.values()), fwd


def test_operand_hook_off_is_the_pinned_oracle():
    a, _, oa = oracle_grads(1, None)
    b, _, ob = oracle_grads(1, lambda t: t)
    assert torch.equal(oa, ob)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def uniform_noise(eps, seed=5):
    g = torch.Generator().manual_seed(seed)

    def hook(x):
        d = x.detach()
        return x + d * eps * (torch.rand(d.shape, generator=g) * 2 - 1)
    return hook


def _cls_change(batch, hook):
    g0, _, out0 = oracle_grads(batch, None)
    g1, _, out1 = oracle_grads(batch, hook)
    e = per_tensor_rel_l2(g1, g0, 'base_model')
    return float(np.median(list(e.values()))), max(e.values()), float((out1 - out0).abs().max() / out0.abs().max())


def test_rounded_forward_moves_resnet_gradients_far_more_than_outputs():
    """B=2: a forward perturbed by 2^-17 keeps the logits to ~1e-5 but moves per-tensor classifier
    gradients by ~1e-2 -- the 1e-3 bar of the north star is a FORWARD bar; no implementation that is
    not bit-identical in the forward can meet it on these gradients."""
    med, worst, fwd = sensitivity(2)
    assert fwd < 1e-4
    assert med > 3e-3                         # the effect is real ...
    assert med < 5e-2 and worst < 1.2e-1      # ... and inside the bar the GPU parity tests use


def test_even_one_ulp_of_operand_noise_moves_gradients_by_more_than_1e3():
    """fp32 re-association noise (2^-23 relative on the conv operands) already exceeds 1e-3, and a
    64x larger perturbation (2^-17) costs only ~sqrt(64)/2: the switch-flip square-root law."""
    med_ulp, worst_ulp, fwd_ulp = _cls_change(2, uniform_noise(2.0 ** -23))
    med_17, worst_17, _ = _cls_change(2, uniform_noise(2.0 ** -17))
    assert fwd_ulp < 5e-6
    assert med_ulp > 1e-3, med_ulp
    assert med_17 < 8 * med_ulp, (med_17, med_ulp)          # far from linear (64x)


def test_gradient_sensitivity_does_not_depend_on_batch():
    med2, worst2, _ = sensitivity(2)
    med8, worst8, _ = sensitivity(8)
    assert 0.5 < med8 / med2 < 2.0, (med2, med8)
    assert worst8 < 1.2e-1 and worst2 < 1.2e-1


# ------------------------------------------------------------------ discriminator: operand rounding
def _disc_grads(quantised, m=4, arch_d='Discriminator3'):
    """Discriminator3 forward / backward in float64 from the oracle's tables; quantised=True rounds the
    operands of every conv (activation, weight, incoming gradient) to bf16 hi + lo, i.e. what a
    tensor-core GEMM on split operands sees, and nothing else."""
    import torch.nn.functional as F

    def hilo64(d):
        d32 = d.float()
        hi = d32.to(torch.bfloat16).float()
        return (hi + (d32 - hi).to(torch.bfloat16).float()).double()

    class QConv(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, b, stride):
            xq, wq = hilo64(x), hilo64(w)
            ctx.save_for_backward(xq, wq)
            ctx.stride = stride
            return F.conv2d(xq, wq, b, stride, 1)

        @staticmethod
        def backward(ctx, g):
            xq, wq = ctx.saved_tensors
            gq = hilo64(g)
            return (torch.nn.grad.conv2d_input(xq.shape, wq, gq, ctx.stride, 1),
                    torch.nn.grad.conv2d_weight(xq, wq.shape, gq, ctx.stride, 1), g.sum((0, 2, 3)), None)

    sd = O.build_state(51, arch_d, seed=1)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(m, 2, 224, 224, generator=g).double()
    dval = torch.randn(m, 2, generator=g).double()
    torch.manual_seed(7)
    masks = [t.double() for t in O.draw_dropout_masks(arch_d, m)]
    st = {k: (v.double().clone().requires_grad_(True) if not O.is_buffer(k) else v.double())
          for k, v in sd.items() if k.startswith('discriminator') and v.is_floating_point()}
    h = x
    for i, (name, _, _, stride, bn) in enumerate(O.disc_blocks(arch_d)):
        p = 'discriminator.discriminator_block_%s' % name
        if quantised:
            h = QConv.apply(h, st[p + '.0.weight'], st[p + '.0.bias'], stride)
        else:
            h = F.conv2d(h, st[p + '.0.weight'], st[p + '.0.bias'], stride, 1)
        h = F.leaky_relu(h, 0.2) * masks[i].view(m, -1, 1, 1)
        if bn:
            h = F.batch_norm(h, None, None, st[p + '.3.weight'], st[p + '.3.bias'], True, 0.1, 0.8)
    v = F.linear(h.reshape(m, -1), st['discriminator.adv_layer.weight'], st['discriminator.adv_layer.bias'])
    v.backward(dval)
    return {k: t.grad for k, t in st.items() if t.requires_grad}


def test_discriminator_backward_amplifies_operand_rounding_at_one_batchnorm():
    """Why the tensor-core discriminator plan is held to 1e-2 on the gradients of its early blocks:
    rounding the conv operands of the float64 oracle to bf16 hi+lo (2^-17) changes the gradients
    downstream of block_3's BatchNorm by ~1e-5 and everything upstream of it by ~5e-3."""
    exact, rounded = _disc_grads(False), _disc_grads(True)
    err = {k: float((exact[k] - rounded[k]).norm() / exact[k].norm()) for k in exact if k.endswith('.0.weight')}
    late = [v for k, v in err.items() if 'block_3_' in k or 'block_4' in k]
    early = [v for k, v in err.items() if 'block_1' in k or 'block_2' in k or k.endswith('block_3.0.weight')]
    assert max(late) < 1e-4, err
    assert 1e-3 < min(early) and max(early) < 1e-2, err
