"""CPU oracle for the I3D classifier behind the DMC generator  --  TEST INFRASTRUCTURE ONLY.

Plain torch-CPU fp32 restatement of ``I3D.forward`` (code/dmcnet_I3D/network/i3d.py:435-533) on a
functional ``state_dict`` and of one iteration of ``model.fit`` (code/dmcnet_I3D/train/model.py:286-446:
CE + MSE [+ the adversarial CE of ``--adv``], the classifier's / generator's / discriminator's optimizers,
the alternating D / G stages, gradient accumulation over ``iter_size`` batches, the two-stage learning-rate
rule of ``adjust_learning_rate`` :268-283).  Only
``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU-baseline legs may import it.

Parity pin: the reference has no tests or golden vectors for this path; this restatement is pinned
bit-exactly against the reference's own ``i3d.py`` imported from /root/reference in the build container
(``oracle/pin_i3d.py``: state_dict keys / shapes, logits, generated flow and every gradient at B=1), and
reference-generated outputs are committed under ``tests/golden/i3d_b1.npz``.

All ``file:line`` citations are relative to /root/reference.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

MIXED = [('mixed_3b', 192, [64, 96, 128, 16, 32, 32]), ('mixed_3c', 256, [128, 128, 192, 32, 96, 64]),
         ('mixed_4b', 480, [192, 96, 208, 16, 48, 64]), ('mixed_4c', 512, [160, 112, 224, 24, 64, 64]),
         ('mixed_4d', 512, [128, 128, 256, 24, 64, 64]), ('mixed_4e', 512, [112, 144, 288, 32, 64, 64]),
         ('mixed_4f', 528, [256, 160, 320, 32, 128, 128]), ('mixed_5b', 832, [256, 160, 320, 32, 128, 128]),
         ('mixed_5c', 832, [384, 192, 384, 48, 128, 128])]
DENSE_GROWTH = {'DenseNetTiny': (8, 8, 6, 4, 2), 'DenseNetSmall': (32, 32, 24, 16, 8),
                'DenseNet': (128, 128, 96, 64, 32)}


def tf_same_pad(kernel: Sequence[int], stride: Sequence[int]) -> Tuple[int, ...]:
    """get_padding_shape, i3d.py:299-315: per dimension pad_along = max(k - s, 0), front = pad_along // 2,
    back = the rest; returned in the order ConstantPad3d receives it there (H, W, then depth)."""
    out = []
    for k, s in zip(kernel, stride):
        pad = max(k - s, 0)
        out += [pad // 2, pad - pad // 2]
    return tuple(out[2:] + out[:2])


def build_state(num_class: int, arch_estimator: Optional[str] = 'DenseNetTiny', seed: Optional[int] = 1,
                arch_d: Optional[str] = None) -> "OrderedDict[str, Tensor]":
    """state_dict of ``I3D(num_class, 'flow+mp4', arch_estimator=...)`` with the constructors' random
    init, in the reference's construction order (generator first, i3d.py:457-465)."""
    from torch import nn
    if seed is not None:
        torch.manual_seed(seed)
    mods: "OrderedDict[str, nn.Module]" = OrderedDict()
    if arch_estimator is not None:
        cin = 5
        for k, g in enumerate(DENSE_GROWTH[arch_estimator]):
            mods['gen_flow_model.conv_%d.0' % k] = nn.Conv2d(cin, g, 3, 1, 1, bias=True)
            cin += g
        mods['gen_flow_model.predict_flow'] = nn.Conv2d(cin, 2, 3, 1, 1, bias=True)
    if arch_d is not None:                                     # i3d.py:466-476, blocks :112-137 (same as dmcnet_GAN)
        from oracle.dmc_oracle import disc_blocks, disc_fc_in
        for name, ci, co, stride, bn in disc_blocks(arch_d):
            if bn:
                nn.Conv2d(ci, co, 3, stride, 1)              # the bn-less block built first and discarded
            mods['discriminator.discriminator_block_%s.0' % name] = nn.Conv2d(ci, co, 3, stride, 1)
            if bn:
                mods['discriminator.discriminator_block_%s.3' % name] = nn.BatchNorm2d(co, 0.8)
        mods['discriminator.adv_layer'] = nn.Linear(disc_fc_in(arch_d), 2)

    def unit(name, cin, cout, k, stride=1):
        pad = tf_same_pad((k,) * 3, (stride,) * 3)
        simple = all(p == pad[0] for p in pad)
        mods[name + '.conv3d'] = nn.Conv3d(cin, cout, (k,) * 3, stride=(stride,) * 3,
                                           padding=(pad[0] if simple else 0), bias=False)
        mods[name + '.batch3d'] = nn.BatchNorm3d(cout)

    unit('conv3d_1a_7x7', 2, 64, 7, 2)
    unit('conv3d_2b_1x1', 64, 64, 1)
    unit('conv3d_2c_3x3', 64, 192, 3)
    for name, cin, oc in MIXED:
        unit(name + '.branch_0', cin, oc[0], 1)
        unit(name + '.branch_1.0', cin, oc[1], 1)
        unit(name + '.branch_1.1', oc[1], oc[2], 3)
        unit(name + '.branch_2.0', cin, oc[3], 1)
        unit(name + '.branch_2.1', oc[3], oc[4], 3)
        unit(name + '.branch_3.1', cin, oc[5], 1)
    mods['conv3d_0c_1x1.conv3d'] = nn.Conv3d(1024, 400, (1, 1, 1), bias=True)
    mods['classifier'] = nn.Linear(400, num_class)
    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            sd[name + '.' + k] = v.detach().clone()
    return sd


def is_buffer(key: str) -> bool:
    return key.endswith(('running_mean', 'running_var', 'num_batches_tracked'))


def _unit(st: Dict[str, Tensor], name: str, x: Tensor, k: int, train: bool, stride: int = 1) -> Tensor:
    """Unit3Dpy.forward, i3d.py:341-356 (use_bn, relu)."""
    pad = tf_same_pad((k,) * 3, (stride,) * 3)
    if all(p == pad[0] for p in pad):
        x = F.conv3d(x, st[name + '.conv3d.weight'], None, stride, pad[0])
    else:
        x = F.conv3d(F.pad(x, pad, 'constant', 0.0), st[name + '.conv3d.weight'], None, stride, 0)
    b = name + '.batch3d'
    if train:
        st[b + '.num_batches_tracked'] += 1
    x = F.batch_norm(x, st[b + '.running_mean'], st[b + '.running_var'], st[b + '.weight'], st[b + '.bias'],
                     train, 0.1, 1e-5)
    return F.relu(x)


def _pool(x: Tensor, kernel, stride) -> Tensor:
    """MaxPool3dTFPadding, i3d.py:375-388: ConstantPad3d zeros, then MaxPool3d(ceil_mode=True)."""
    return F.max_pool3d(F.pad(x, tf_same_pad(kernel, stride), 'constant', 0.0), kernel, stride, ceil_mode=True)


def _mixed(st, name: str, x: Tensor, train: bool) -> Tensor:
    """Mixed.forward, i3d.py:425-432."""
    o0 = _unit(st, name + '.branch_0', x, 1, train)
    o1 = _unit(st, name + '.branch_1.1', _unit(st, name + '.branch_1.0', x, 1, train), 3, train)
    o2 = _unit(st, name + '.branch_2.1', _unit(st, name + '.branch_2.0', x, 1, train), 3, train)
    o3 = _unit(st, name + '.branch_3.1', _pool(x, (3, 3, 3), (1, 1, 1)), 1, train)
    return torch.cat((o0, o1, o2, o3), 1)


def gen_forward(st: Dict[str, Tensor], x: Tensor, arch_estimator: str) -> Tensor:
    """EstimatorDenseNet*.forward (identical copy of the 2-D ones, i3d.py:33-107)."""
    for k in range(len(DENSE_GROWTH[arch_estimator])):
        p = 'gen_flow_model.conv_%d.0' % k
        x = torch.cat((F.leaky_relu(F.conv2d(x, st[p + '.weight'], st[p + '.bias'], 1, 1), 0.1), x), 1)
    p = 'gen_flow_model.predict_flow'
    return F.conv2d(x, st[p + '.weight'], st[p + '.bias'], 1, 1)


def i3d_forward(st: Dict[str, Tensor], inp: Tensor, *, arch_estimator: Optional[str] = 'DenseNetTiny',
                train: bool = True, dropout_mask: Optional[Tensor] = None, detach: bool = False,
                record: Optional[Dict[str, Tensor]] = None):
    """I3D.forward(inp, node='flow+logit', detach), i3d.py:497-533.  inp [B, 5, T, H, W] (or [B, 2, T, H, W]
    without an estimator).  dropout_mask [B, 400] replaces nn.Dropout's own draw (None: no dropout).
    Returns (logits [B, num_class], flow [B, 2, T, H, W])."""
    if arch_estimator is not None:
        b, c, t, h, w = inp.shape
        y = gen_forward(st, torch.reshape(torch.transpose(inp, 1, 2), (-1, c, h, w)), arch_estimator)
        inp = torch.transpose(torch.reshape(y, (b, t, 2, h, w)), 1, 2)
    rec = (lambda k, v: record.__setitem__(k, v.detach())) if record is not None else (lambda k, v: None)
    out = _unit(st, 'conv3d_1a_7x7', inp.detach() if detach else inp, 7, train, 2)
    rec('conv3d_1a_7x7', out)
    out = _pool(out, (1, 3, 3), (1, 2, 2))
    rec('pool_2a', out)
    out = _unit(st, 'conv3d_2b_1x1', out, 1, train)
    rec('conv3d_2b_1x1', out)
    out = _unit(st, 'conv3d_2c_3x3', out, 3, train)
    rec('conv3d_2c_3x3', out)
    out = _pool(out, (1, 3, 3), (1, 2, 2))
    for name, _, _ in MIXED:
        if name == 'mixed_4b':
            out = _pool(out, (3, 3, 3), (2, 2, 2))
        elif name == 'mixed_5b':
            out = _pool(out, (2, 2, 2), (2, 2, 2))
        out = _mixed(st, name, out, train)
        rec(name, out)
    out = F.avg_pool3d(out, (2, 7, 7), (1, 1, 1))
    out = F.conv3d(out, st['conv3d_0c_1x1.conv3d.weight'], st['conv3d_0c_1x1.conv3d.bias'])
    out = out.squeeze(3).squeeze(3).mean(2)                      # Unit3Dpy squeeze / mean, i3d.py:352-356
    if dropout_mask is not None:
        out = out * dropout_mask
    out = F.linear(out, st['classifier.weight'], st['classifier.bias'])
    return out, inp


def make_inputs(batch: int, clip_len: int, num_class: int, seed: int = 0, hw: int = 224):
    """Synthetic sample [B, 7, T, H, W] (mv 2 | residual 3 | flow 2) with the uint8 value model of SURVEY.md
    section 8(d) and integer targets."""
    g = torch.Generator().manual_seed(seed)

    def u8(ch, sigma, scale):
        v = torch.clamp(torch.round(128.0 + sigma * torch.randn((batch, ch, clip_len, hw, hw), generator=g)), 0, 255)
        return (v / 255.0 - 0.5) / scale
    data = torch.cat((u8(2, 25.0, 0.226), u8(3, 20.0, 0.226), u8(2, 30.0, 0.226)), 1).contiguous()
    target = torch.randint(0, num_class, (batch,), generator=g)
    return data, target


class I3DHParams:
    """The knobs of train_model.py:20-29 / train_hmdb51.py that reach the step."""

    def __init__(self, optim: str = 'sgd', lr_base: float = 0.005, lr_base2: float = 0.002, weight_decay: float = 1e-4,
                 iter_size: int = 1, epoch_thre: int = 1, fine_tune: bool = True, detach: bool = False,
                 dropout: float = 0.5, adv: float = 0.0, lr_d: Optional[float] = None):
        self.optim, self.lr_base, self.lr_base2, self.weight_decay = optim, lr_base, lr_base2, weight_decay
        self.iter_size, self.epoch_thre, self.fine_tune, self.detach, self.dropout = iter_size, epoch_thre, fine_tune, detach, dropout
        self.adv, self.lr_d = adv, (lr_base if lr_d is None else lr_d)


def param_groups(keys: Sequence[str]):
    """train_model.py:62-86 with modality 'flow+mp4': generator | I3D base layers | new layers."""
    gf = [k for k in keys if k.startswith('gen_flow_model')]
    new = [k for k in keys if k.startswith('conv3d_0c_1x1') or k.startswith('classifier')]
    base = [k for k in keys if k not in gf and k not in new and not k.startswith('discriminator')]
    return gf, base, new


def lr_mult_rule(lr_mult: float, epoch: int, epoch_thre: int) -> float:
    """model.adjust_learning_rate, train/model.py:268-283: the main convolutional part (lr_mult 0.2 or 0.5)
    is frozen during stage one; 0.5 becomes 1.0 afterwards."""
    if lr_mult in (0.2, 0.5):
        if epoch_thre > 0 and epoch + 1 <= epoch_thre:
            return 0.0
        if lr_mult == 0.5:
            return 1.0
    return lr_mult


class I3DOracleTrainer:
    """``model.fit``, one call = one batch of the epoch.  Without a discriminator (hp.adv == 0, optimizer_3 is
    None) every batch is a classifier + generator iteration; with one, batches alternate in runs of
    ``iter_size``: D stage (loss CE + adv * adversarial CE; steps ``optimizer`` and ``optimizer_3``), G stage
    (loss [0 in epoch 0] * CE + MSE + adv * the SAME adversarial CE; steps ``optimizer_mse`` only).  Gradients
    are those of autograd's .grad accumulation: a stage zeroes only the optimizers it steps, so the other
    groups' gradients carry over into their next step -- reference behaviour, train/model.py:344-446."""

    def __init__(self, state: Dict[str, Tensor], hp: I3DHParams, arch_estimator: str = 'DenseNetTiny',
                 arch_d: Optional[str] = None):
        self.hp, self.arch, self.arch_d = hp, arch_estimator, arch_d
        self.st: "OrderedDict[str, Tensor]" = OrderedDict()
        for k, v in state.items():
            self.st[k] = v.detach().clone() if is_buffer(k) else v.detach().clone().requires_grad_(True)
        keys = [k for k in self.st if not is_buffer(k)]
        gf, base, new = param_groups(keys)
        dk = [k for k in keys if k.startswith('discriminator')]
        self.keys = {'gf': gf, 'base': base, 'new': new, 'd': dk}
        lr_mul = 0.2 if hp.fine_tune else 0.5                           # train_model.py:100-105
        self.lr_mul = lr_mul
        P = lambda ks: [self.st[k] for k in ks]

        def make(lr, params_groups, eps=1e-8):
            if hp.optim == 'adam':
                return torch.optim.Adam(params_groups, lr=lr, weight_decay=hp.weight_decay, eps=eps)
            return torch.optim.SGD(params_groups, lr=lr, momentum=0.9, weight_decay=hp.weight_decay, nesterov=True)
        grp = lambda: [{'params': P(base), 'lr_mult': lr_mul}, {'params': P(new), 'lr_mult': 1.0}]
        # stage one / stage two optimizers (train_model.py:122-176); the generator's stage-two Adam has eps 1e-3
        self.opt = [make(hp.lr_base, grp()), make(hp.lr_base2, grp())]
        self.opt_mse = [make(hp.lr_base, [{'params': P(gf)}]), make(hp.lr_base2, [{'params': P(gf)}], eps=1e-3)]
        self.opt_d = None
        if hp.adv > 0:
            assert arch_d is not None
            self.opt_d = torch.optim.Adam(P(dk), lr=hp.lr_base, weight_decay=hp.weight_decay, eps=1e-3)   # :143-149
        self.i = 0
        self.i_batch = 0
        self.epoch = 0

    def set_epoch(self, epoch: int):
        self.epoch, self.i_batch = epoch, 0

    def _adjust(self, optimizer, lr, epoch=0, epoch_thre=0):
        for g in optimizer.param_groups:
            g['lr'] = lr * lr_mult_rule(g.get('lr_mult', 1.0), epoch, epoch_thre)

    def _forward(self, data, target, dropout_mask, disc_masks):
        logits, flow = i3d_forward(self.st, data[:, :5], arch_estimator=self.arch, train=True,
                                   dropout_mask=dropout_mask)
        loss = F.cross_entropy(logits, target)
        mse = F.mse_loss(flow, data[:, 5:7])
        adv = None
        if self.opt_d is not None:                                       # static_model.forward, :148-160
            from oracle.dmc_oracle import disc_forward
            b, _, t, h, w = flow.shape
            x = torch.cat((torch.reshape(torch.transpose(flow, 1, 2), (-1, 2, h, w)),
                           torch.reshape(torch.transpose(data[:, 5:7], 1, 2), (-1, 2, h, w))), 0)
            validity = disc_forward(self.st, x, self.arch_d, True, disc_masks)
            labels = torch.cat((torch.zeros(b * t, dtype=torch.int64), torch.ones(b * t, dtype=torch.int64)))
            adv = F.cross_entropy(validity, labels)
            self.last_validity = validity.detach()
        self.last_logits, self.last_flow = logits.detach(), flow.detach()
        return logits, loss, mse, adv

    def _div(self, opts):
        if self.hp.iter_size != 1:
            for o in opts:
                for g in o.param_groups:
                    for p in g['params']:
                        if p.grad is not None:
                            p.grad /= self.hp.iter_size

    def step(self, data: Tensor, target: Tensor, dropout_mask: Optional[Tensor] = None,
             disc_masks: Optional[Sequence[Tensor]] = None) -> Dict[str, float]:
        hp = self.hp
        stage2 = self.epoch >= hp.epoch_thre
        opt, opt_mse = self.opt[1 if stage2 else 0], self.opt_mse[1 if stage2 else 0]
        d_stage = self.opt_d is not None and self.i_batch % (2 * hp.iter_size) < hp.iter_size
        self.i_batch += 1
        logits, loss, mse, adv = self._forward(data, target, dropout_mask, disc_masks)
        if not stage2:
            lr = hp.lr_base
            lr1 = 0.0 if hp.detach else lr                               # :405-411
        else:
            lr = lr1 = hp.lr_base2
        stepped = False
        if d_stage:
            (loss + hp.adv * adv).backward()                             # :362-369
            self._adjust(opt, lr1, self.epoch, hp.epoch_thre)
            self._adjust(self.opt_d, hp.lr_d)
            self.i += 1
            if self.i % hp.iter_size == 0:
                self._div((opt, self.opt_d))
                opt.step(); opt.zero_grad()
                self.opt_d.step(); self.opt_d.zero_grad()
                self.i, stepped = 0, True
        else:
            if self.opt_d is None:
                (loss + mse).backward()                                  # :392-396
                self._adjust(opt, lr1, self.epoch, hp.epoch_thre)
            else:
                ((0.0 if self.epoch < 1 else 1.0) * loss + mse + hp.adv * adv).backward()    # :397-402
            self._adjust(opt_mse, lr)
            self.i += 1
            if self.i % hp.iter_size == 0:
                self._div(((opt, opt_mse) if self.opt_d is None else (opt_mse,)))
                if self.opt_d is None:
                    opt.step(); opt.zero_grad()
                opt_mse.step(); opt_mse.zero_grad()
                self.i, stepped = 0, True
        top = logits.detach().topk(min(5, logits.shape[1]), 1).indices
        out = {'loss_ce': float(loss.detach()), 'loss_mse': float(mse.detach()), 'stepped': stepped,
               'stage': 'D' if d_stage else 'G',
               'top1': float((top[:, :1] == target.view(-1, 1)).any(1).float().mean() * 100),
               'top5': float((top == target.view(-1, 1)).any(1).float().mean() * 100)}
        if adv is not None:
            out['loss_adv'] = float(adv.detach())
        return out

    def grads(self) -> Dict[str, Tensor]:
        return {k: v.grad.detach().clone() for k, v in self.st.items() if not is_buffer(k) and v.grad is not None}

    def state_dict(self) -> Dict[str, Tensor]:
        return OrderedDict((k, v.detach().clone()) for k, v in self.st.items())
