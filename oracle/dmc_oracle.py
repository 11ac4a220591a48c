"""CPU oracle for the DMC-Net training hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU fp32 restatement of the reference's algorithm
for the hot path (generator -> ResNet-18 [-> discriminator] -> losses ->
per-tensor Adam).  It is *not* part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the timed
CPU baseline.  Nothing under ``dmcnet_b200/`` imports it.

Parity pin: the reference repo has no tests, golden vectors or KATs for this
path, so the oracle is pinned against the reference's own
``model.py`` imported from /root/reference in the build container
(``oracle/pin_against_reference.py``; bit-identical state_dict init, forward
outputs and gradients), and the outputs of the *reference* are committed as
fixtures under ``tests/golden/`` (``tests/golden/make_golden.py``).

Third-party arithmetic: all conv / BN / loss / Adam maths is PyTorch's
(reference pins 0.3.1 / 0.4.0, README.md:28; the installed torch 2.11 CPU fp32
semantics are the oracle) and torchvision's ``resnet18`` topology
(code/dmcnet/model.py:305).

All ``file:line`` citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------
# architecture tables
# --------------------------------------------------------------------------

# EstimatorDenseNetTiny: code/dmcnet/model.py:172-194 (growth 8,8,6,4,2 then
# predict_flow -> 2).  Each conv sees cat(new_k-1, ..., new_0, mv, res).
GEN_TINY_GROWTH = (8, 8, 6, 4, 2)
GEN_CH_IN = 5
# EstimatorDenseNetSmall (:147-169) and EstimatorDenseNet (:122-144): same structure, wider layers
DENSE_GROWTH = {'DenseNetTiny': GEN_TINY_GROWTH, 'DenseNetSmall': (32, 32, 24, 16, 8),
                'DenseNet': (128, 128, 96, 64, 32)}

# ContextNetwork / ContextNetworkAtt: code/dmcnet/model.py:45-104.  (cout, dilation) per layer; the
# fifth dilation is 16 at full resolution and 1 when gen_flow_ds_factor != 0.
def context_layers(att: int, gen_flow_ds_factor: int) -> List[Tuple[int, int]]:
    d5 = 16 if gen_flow_ds_factor == 0 else 1
    layers = [(32, 1), (128, 2), (128, 4), (96, 8), (64, d5), (32, 1)]
    return layers if att else layers + [(2, 1)]


# ...TinyEarlyFusionSum / ...Stack: code/dmcnet/model.py:197-250
EARLY_FUSION = {'DenseNetTinyEarlyFusionSum': False, 'DenseNetTinyEarlyFusionStack': True}

# ResNet-18 = torchvision BasicBlock x (2,2,2,2), widths 64..512
RESNET18_STAGES = ((64, 1), (128, 2), (256, 2), (512, 2))

# discriminators: code/dmcnet_GAN/model.py:282-438.  (name_suffix, cin, cout, stride, bn)
def disc_blocks(arch_d: str) -> List[Tuple[str, int, int, int, bool]]:
    """Block list of a discriminator, in forward order.

    discriminator_block (stride 2): code/dmcnet_GAN/model.py:254-265
    discriminator_block2 (stride 1): code/dmcnet_GAN/model.py:268-279
    """
    if arch_d == 'Discriminator4':              # :369-385
        return [('1', 2, 8, 2, False), ('2', 8, 16, 2, True), ('3', 16, 32, 2, True)]
    extra = {'Discriminator': 0, 'Discriminator2': 1, 'Discriminator3': 2,
             'Discriminator5': 4}[arch_d]       # :282, :303, :332, :388
    blocks = []
    cin = 2
    for stage, cout in enumerate((16, 32, 64, 128), start=1):
        blocks.append((str(stage), cin, cout, 2, stage != 1))
        for j in range(extra):
            blocks.append(('%d_%d' % (stage, j + 2), cout, cout, 1, True))
        cin = cout
    return blocks


def disc_fc_in(arch_d: str) -> int:
    return 32 * 28 * 28 if arch_d == 'Discriminator4' else 128 * 14 * 14


# --------------------------------------------------------------------------
# parameter construction  (same RNG consumption order as the reference ctor)
# --------------------------------------------------------------------------

def build_state(num_class: int, arch_d: Optional[str] = None, seed: Optional[int] = 1,
                arch_estimator: str = 'DenseNetTiny', att: int = 0, gen_flow_ds_factor: int = 0
                ) -> "OrderedDict[str, Tensor]":
    """state_dict of ``Model(num_class, S, 'mv', 'resnet18', arch_estimator=
    'DenseNetTiny'[, arch_d=...], use_databn=0)`` with random init.

    Construction order follows ``Model.__init__``: ``_prepare_base_model``
    (torchvision resnet18 -> generator -> discriminator; dmcnet/model.py:301-327,
    dmcnet_GAN/model.py:495-530) then ``_prepare_tsn`` (new fc, new 2-channel
    conv1; dmcnet/model.py:283-294).  ``pretrained=True`` (model.py:305) cannot
    download here, so weights are the constructors' random init.
    """
    import torchvision
    from torch import nn
    if seed is not None:
        torch.manual_seed(seed)
    base = torchvision.models.resnet18(weights=None)
    gen = OrderedDict()
    cin = GEN_CH_IN
    if arch_estimator in DENSE_GROWTH:
        for k, g in enumerate(DENSE_GROWTH[arch_estimator]):
            gen['conv_%d.0' % k] = nn.Conv2d(cin, g, 3, 1, 1, bias=True)      # model.py:111-115
            cin += g
        gen['predict_flow'] = nn.Conv2d(cin, 2, 3, 1, 1, bias=True)           # model.py:118-119
    elif arch_estimator in EARLY_FUSION:                                      # model.py:197-250
        gen['conv_0_mv.0'] = nn.Conv2d(2, 8, 3, 1, 1, bias=True)
        gen['conv_0_r.0'] = nn.Conv2d(3, 8, 3, 1, 1, bias=True)
        cin = 16 if EARLY_FUSION[arch_estimator] else 8
        for k, g in zip((1, 2, 3, 4), (8, 6, 4, 2)):
            gen['conv_%d.0' % k] = nn.Conv2d(cin, g, 3, 1, 1, bias=True)
            cin += g
        gen['predict_flow'] = nn.Conv2d(cin, 2, 3, 1, 1, bias=True)
    elif arch_estimator == 'ContextNetwork':                                  # model.py:31-104
        def block(prefix, ci, co, d):
            gen[prefix + '.0'] = nn.Conv2d(ci, co, 3, 1, d, d, bias=False)
            gen[prefix + '.1'] = nn.BatchNorm2d(co)
        for i, (co, d) in enumerate(context_layers(att, gen_flow_ds_factor)):
            block('conv_context.%d' % i, cin, co, d)
            cin = co
        if att:
            block('predict_flow', 32, 2, 1)
            block('predict_att.0', 32, 2, 1)
    else:
        raise ValueError('unknown arch_estimator %r' % (arch_estimator,))
    disc = OrderedDict()
    if arch_d is not None:
        for name, ci, co, stride, bn in disc_blocks(arch_d):
            if bn:
                # discriminator_block builds (and discards) a bn-less block first,
                # GAN/model.py:255-258 -> one extra Conv2d init draw from the RNG.
                nn.Conv2d(ci, co, 3, stride, 1)
            disc['discriminator_block_%s.0' % name] = nn.Conv2d(ci, co, 3, stride, 1)
            if bn:
                disc['discriminator_block_%s.3' % name] = nn.BatchNorm2d(co, 0.8)
        disc['adv_layer'] = nn.Linear(disc_fc_in(arch_d), 2)
    base.fc = nn.Linear(base.fc.in_features, num_class)                     # model.py:285-286
    base.conv1 = nn.Conv2d(2, 64, kernel_size=7, stride=2, padding=3, bias=False)  # :289-294

    sd: "OrderedDict[str, Tensor]" = OrderedDict()
    for k, v in base.state_dict().items():
        sd['base_model.' + k] = v.detach().clone()
    for name, m in gen.items():
        for k, v in m.state_dict().items():
            sd['gen_flow_model.%s.%s' % (name, k)] = v.detach().clone()
    for name, m in disc.items():
        for k, v in m.state_dict().items():
            sd['discriminator.%s.%s' % (name, k)] = v.detach().clone()
    return sd


def is_buffer(key: str) -> bool:
    return key.endswith(('running_mean', 'running_var', 'num_batches_tracked'))


# --------------------------------------------------------------------------
# synthetic inputs  (CoviarDataSet sample layout)
# --------------------------------------------------------------------------

INPUT_STD = (0.229, 0.224, 0.225)       # code/dmcnet/dataset.py:109-112


def make_inputs(batch: int, segments: int, num_class: int, seed: int = 0, hw: int = 224
                ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """(input_flow[B,S,2,H,W], input_mv[B,S,2,H,W], input_residual[B,S,3,H,W], target[B]).

    uint8-like values normalised exactly as code/dmcnet/dataset.py:251-263:
    ``(v/255 - 0.5)/mean(std)`` for flow and mv, ``(v/255 - 0.5)/std_c`` for
    the residual.
    """
    g = torch.Generator().manual_seed(seed)
    def u8(shape, sigma):
        return torch.clamp(torch.round(128.0 + sigma * torch.randn(shape, generator=g)), 0, 255)
    mv = u8((batch, segments, 2, hw, hw), 25.0)
    res = u8((batch, segments, 3, hw, hw), 20.0)
    flow = u8((batch, segments, 2, hw, hw), 30.0)
    std = torch.tensor(INPUT_STD, dtype=torch.float32)
    mstd = torch.mean(std)
    mv = (mv / 255.0 - 0.5) / mstd
    flow = (flow / 255.0 - 0.5) / mstd
    res = (res / 255.0 - 0.5) / std.view(1, 1, 3, 1, 1)
    target = torch.randint(0, num_class, (batch,), generator=g)
    return flow.float(), mv.float(), res.float(), target


# --------------------------------------------------------------------------
# functional forward
# --------------------------------------------------------------------------

def _bn(x: Tensor, st: Dict[str, Tensor], prefix: str, train: bool, eps: float) -> Tensor:
    """nn.BatchNorm2d forward incl. running-stat update (momentum 0.1)."""
    if train:
        st[prefix + '.num_batches_tracked'] += 1
    return F.batch_norm(x, st[prefix + '.running_mean'], st[prefix + '.running_var'],
                        st[prefix + '.weight'], st[prefix + '.bias'], train, 0.1, eps)


def gen_tiny_forward(st: Dict[str, Tensor], x: Tensor) -> Tensor:
    """EstimatorDenseNetTiny.forward, code/dmcnet/model.py:186-194: new channels
    are *prepended*: x = cat(LeakyReLU_0.1(conv_k(x)), x)."""
    for k in range(len(GEN_TINY_GROWTH)):
        p = 'gen_flow_model.conv_%d.0' % k
        y = F.leaky_relu(F.conv2d(x, st[p + '.weight'], st[p + '.bias'], 1, 1), 0.1)
        x = torch.cat((y, x), 1)
    p = 'gen_flow_model.predict_flow'
    return F.conv2d(x, st[p + '.weight'], st[p + '.bias'], 1, 1)


def early_fusion_forward(st: Dict[str, Tensor], x: Tensor, stack: bool) -> Tensor:
    """EstimatorDenseNetTinyEarlyFusionSum / ...Stack.forward, code/dmcnet/model.py:212-222, :240-250."""
    q = 'gen_flow_model.'
    lr = lambda t: F.leaky_relu(t, 0.1)
    x_mv = lr(F.conv2d(x[:, :2], st[q + 'conv_0_mv.0.weight'], st[q + 'conv_0_mv.0.bias'], 1, 1))
    x_r = lr(F.conv2d(x[:, 2:], st[q + 'conv_0_r.0.weight'], st[q + 'conv_0_r.0.bias'], 1, 1))
    x = torch.cat((x_mv, x_r), 1) if stack else x_mv + x_r
    for k in (1, 2, 3, 4):
        p = q + 'conv_%d.0' % k
        x = torch.cat((lr(F.conv2d(x, st[p + '.weight'], st[p + '.bias'], 1, 1)), x), 1)
    return F.conv2d(x, st[q + 'predict_flow.weight'], st[q + 'predict_flow.bias'], 1, 1)


def context_forward(st: Dict[str, Tensor], x: Tensor, att: int, gen_flow_ds_factor: int, train: bool):
    """ContextNetwork.forward (code/dmcnet/model.py:69-71) / ContextNetworkAtt.forward (:99-102):
    every layer is dilated Conv3x3(bias=False) -> BatchNorm2d(eps 1e-5) -> LeakyReLU(0.1), the
    output layers included; the attention head adds a ReLU."""
    def block(x, prefix, d):
        x = F.conv2d(x, st[prefix + '.0.weight'], None, 1, d, d)
        return F.leaky_relu(_bn(x, st, prefix + '.1', train, 1e-5), 0.1)
    for i, (_, d) in enumerate(context_layers(att, gen_flow_ds_factor)):
        x = block(x, 'gen_flow_model.conv_context.%d' % i, d)
    if not att:
        return x
    return block(x, 'gen_flow_model.predict_flow', 1), F.relu(block(x, 'gen_flow_model.predict_att.0', 1))


# Sensitivity experiments (tests/test_grad_sensitivity.py): when set, every ResNet-18 conv sees
# OPERAND_HOOK(activation) and OPERAND_HOOK(weight) instead of the exact fp32 operands (straight-
# through: the hook's result carries the operand's gradient).  None = the pinned fp32 restatement.
OPERAND_HOOK = None


def _conv(x: Tensor, w: Tensor, stride: int, pad: int) -> Tensor:
    if OPERAND_HOOK is not None:
        x, w = OPERAND_HOOK(x), OPERAND_HOOK(w)
    return F.conv2d(x, w, None, stride, pad)


def resnet18_forward(st: Dict[str, Tensor], x: Tensor, train: bool, prefix: str = 'base_model',
                     capture: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """torchvision resnet18 forward with the 2-channel conv1 and num_class fc of
    code/dmcnet/model.py:283-294.  ``capture`` (tests only) receives the intermediate tensors:
    raw conv outputs ('<unit>.conv'), block activations and the max-pool argmax."""
    p = prefix
    cap = (lambda k, t: capture.__setitem__(k, t.detach())) if capture is not None else (lambda k, t: None)
    x = _conv(x, st[p + '.conv1.weight'], 2, 3)
    cap('conv1', x)
    x = F.relu(_bn(x, st, p + '.bn1', train, 1e-5))
    if capture is not None:
        cap('pool_idx', F.max_pool2d(x.detach(), 3, 2, 1, return_indices=True)[1])
    x = F.max_pool2d(x, 3, 2, 1)
    cap('pool', x)
    for li, (width, stride) in enumerate(RESNET18_STAGES, start=1):
        for b in range(2):
            q = '%s.layer%d.%d' % (p, li, b)
            s = stride if b == 0 else 1
            out = _conv(x, st[q + '.conv1.weight'], s, 1)
            cap(q + '.conv1', out)
            out = F.relu(_bn(out, st, q + '.bn1', train, 1e-5))
            cap(q + '.act1', out)
            out = _conv(out, st[q + '.conv2.weight'], 1, 1)
            cap(q + '.conv2', out)
            out = _bn(out, st, q + '.bn2', train, 1e-5)
            if (q + '.downsample.0.weight') in st:
                idt = _conv(x, st[q + '.downsample.0.weight'], s, 0)
                cap(q + '.downsample.0', idt)
                idt = _bn(idt, st, q + '.downsample.1', train, 1e-5)
            else:
                idt = x
            x = F.relu(out + idt)
            cap(q + '.out', x)
    x = F.adaptive_avg_pool2d(x, 1).flatten(1)
    cap('pooled', x)
    return F.linear(x, st[p + '.fc.weight'], st[p + '.fc.bias'])


def draw_dropout_masks(arch_d: str, m: int, generator: Optional[torch.Generator] = None
                       ) -> List[Tensor]:
    """Dropout2d(0.25) feature masks, one ``[M, C]`` tensor per block, drawn with
    the same ATen calls, shapes and order as ``F.dropout2d`` inside the
    reference's blocks (noise = empty([M,C,1,1]).bernoulli_(0.75).div_(0.75))."""
    masks = []
    for _, _, co, _, _ in disc_blocks(arch_d):
        noise = torch.empty(m, co, 1, 1).bernoulli_(0.75, generator=generator).div_(0.75)
        masks.append(noise.view(m, co))
    return masks


def disc_forward(st: Dict[str, Tensor], x: Tensor, arch_d: str, train: bool,
                 masks: Optional[Sequence[Tensor]] = None) -> Tensor:
    """Discriminator*.forward, code/dmcnet_GAN/model.py:254-438.  Block order is
    Conv(bias) -> LeakyReLU(0.2) -> Dropout2d(0.25) -> BatchNorm2d(eps=0.8)."""
    for i, (name, _, _, stride, bn) in enumerate(disc_blocks(arch_d)):
        p = 'discriminator.discriminator_block_%s' % name
        x = F.leaky_relu(F.conv2d(x, st[p + '.0.weight'], st[p + '.0.bias'], stride, 1), 0.2)
        if train:
            if masks is None:
                x = F.dropout2d(x, 0.25, True)
            else:
                x = x * masks[i].view(x.shape[0], x.shape[1], 1, 1)
        if bn:
            x = _bn(x, st, p + '.3', train, 0.8)
    x = x.reshape(x.shape[0], -1)
    return F.linear(x, st['discriminator.adv_layer.weight'], st['discriminator.adv_layer.bias'])


def model_forward(st: Dict[str, Tensor], input_mv: Tensor, input_residual: Tensor,
                  input_flow: Optional[Tensor] = None, *, gan: bool = False,
                  arch_d: Optional[str] = None, train: bool = True,
                  gen_flow_or_delta: int = 1, masks: Optional[Sequence[Tensor]] = None,
                  arch_estimator: str = 'DenseNetTiny', att: int = 0, gen_flow_ds_factor: int = 0):
    """Model.forward.  dmcnet: code/dmcnet/model.py:330-357 -> (base_out, gen_flow[, att_flow]);
    GAN: code/dmcnet_GAN/model.py:533-566 -> (base_out, validity, gen_flow[, att_flow]).
    ``att_flow`` is returned only for ContextNetwork with att == 1."""
    mv = input_mv.reshape((-1,) + tuple(input_mv.shape[-3:]))
    res = input_residual.reshape((-1,) + tuple(input_residual.shape[-3:]))
    if gen_flow_ds_factor != 0:                                              # model.py:335-337
        mv = F.avg_pool2d(mv, gen_flow_ds_factor, gen_flow_ds_factor)
        res = F.avg_pool2d(res, gen_flow_ds_factor, gen_flow_ds_factor)
    x = torch.cat((mv, res), 1)
    att_flow = None
    if arch_estimator in DENSE_GROWTH:
        gen_flow = gen_tiny_forward(st, x)
    elif arch_estimator in EARLY_FUSION:
        gen_flow = early_fusion_forward(st, x, EARLY_FUSION[arch_estimator])
    elif arch_estimator == 'ContextNetwork' and att == 1:
        gen_flow, att_flow = context_forward(st, x, 1, gen_flow_ds_factor, train)
    elif arch_estimator == 'ContextNetwork':
        gen_flow = context_forward(st, x, 0, gen_flow_ds_factor, train)
    else:
        raise AttributeError("'Model' object has no attribute 'gen_flow_model'")     # model.py:310-325
    if gen_flow_or_delta == 1:
        gen_flow = torch.add(gen_flow, mv)
    if gen_flow_ds_factor != 0:                                              # model.py:347-348 (a TILING)
        gen_flow = gen_flow.repeat(1, 1, gen_flow_ds_factor, gen_flow_ds_factor)
    extra = () if att_flow is None else (att_flow,)
    if not gan:
        base_out = resnet18_forward(st, gen_flow.detach(), train)           # model.py:352
        return (base_out, gen_flow) + extra
    if input_flow is not None:
        flow = input_flow.reshape((-1,) + tuple(input_flow.shape[-3:]))
        d_in = torch.cat((gen_flow, flow), 0)                                # "first fake then real"
    else:
        d_in = gen_flow
    base_out = resnet18_forward(st, gen_flow, train)                         # GAN/model.py:560
    validity = disc_forward(st, d_in, arch_d, train, masks)                  # :561
    return (base_out, validity, gen_flow) + extra


# --------------------------------------------------------------------------
# training step restatement
# --------------------------------------------------------------------------

@dataclass
class HParams:
    """Defaults = exp_my/hmdb51_gen_flow/split1/run.sh:12-35 and
    exp_my/hmdb51_gan/split1/run.sh:12-39."""
    lr: float = 0.01
    lr_cls: float = 1.0          # loss weights (code/dmcnet/train_options.py:69-73)
    lr_mse: float = 10.0
    lr_adv_g: float = 1.0
    lr_adv_d: float = 0.01
    lr_cls_mult: float = 0.01    # per-group lr multipliers (:74-75)
    lr_mse_mult: float = 1.0
    lr_d_mult: float = 1.0
    weight_decay: float = 1e-4
    lr_steps: Tuple[int, ...] = (20, 35, 45)
    lr_decay: float = 0.1
    num_segments: int = 3
    eps: float = 1e-3            # Adam eps, code/dmcnet/train.py:137,142
    betas: Tuple[float, float] = (0.9, 0.999)
    loss_mse: str = 'MSELoss'    # --loss-mse (train_options.py:71): MSELoss | SmoothL1Loss | L1


def flow_criterion(name: str):
    """criterion_mse of code/dmcnet/train.py:166-172 (GAN/train.py:179-185).  Any other
    string leaves criterion_mse undefined in the reference (NameError at the first use)."""
    if name == 'MSELoss':
        return torch.nn.MSELoss()
    if name == 'SmoothL1Loss':
        return torch.nn.SmoothL1Loss()
    if name == 'L1':
        return torch.nn.L1Loss()
    raise NameError("name 'criterion_mse' is not defined")


def accuracy(output: Tensor, target: Tensor, topk=(1,)) -> List[float]:
    """code/dmcnet/train.py:411-424."""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand(maxk, -1))
    return [float(correct[:k].reshape(-1).float().sum(0)) * 100.0 / target.size(0) for k in topk]


class OracleTrainer:
    """Line-for-line restatement of the step bodies
    code/dmcnet/train.py:221-266 (``gan=False``) and
    code/dmcnet_GAN/train.py:237-372 (``gan=True``), plus the optimizer
    wiring code/dmcnet/train.py:121-142 / code/dmcnet_GAN/train.py:122-153 and
    ``adjust_learning_rate`` (train.py:398-408)."""

    def __init__(self, state: Dict[str, Tensor], hp: HParams, *, gan: bool = False,
                 arch_d: Optional[str] = None, arch_estimator: str = 'DenseNetTiny', att: int = 0,
                 gen_flow_ds_factor: int = 0):
        self.hp, self.gan, self.arch_d = hp, gan, arch_d
        # generator choice (--arch_estimator / --att / --gen_flow_ds_factor, train.py:53-60)
        self.gen_kw = dict(arch_estimator=arch_estimator, att=att, gen_flow_ds_factor=gen_flow_ds_factor)
        self.att = int(att == 1 and arch_estimator == 'ContextNetwork')
        self.st: Dict[str, Tensor] = OrderedDict()
        for k, v in state.items():
            t = v.detach().clone()
            if not is_buffer(k):
                t.requires_grad_(True)
            self.st[k] = t
        groups = {'base_model': [], 'gen_flow_model': [], 'discriminator': []}
        mults = {'base_model': hp.lr_cls_mult, 'gen_flow_model': hp.lr_mse_mult,
                 'discriminator': hp.lr_d_mult}
        for key, value in self.st.items():
            if is_buffer(key):
                continue
            for tag in groups:
                if tag in key:                                               # train.py:125,129
                    decay_mult = 0.0 if 'bias' in key else 1.0
                    groups[tag].append({'params': value, 'lr': hp.lr, 'lr_mult': mults[tag],
                                        'decay_mult': decay_mult})
        mk = lambda g: torch.optim.Adam(g, weight_decay=hp.weight_decay, eps=hp.eps,
                                        betas=hp.betas)
        self.opt_cls = mk(groups['base_model'])
        self.opt_gf = mk(groups['gen_flow_model'])
        self.opt_d = mk(groups['discriminator']) if gan else None
        self.iteration = 0
        self.criterion_mse = flow_criterion(hp.loss_mse)
        self.set_epoch(0, epoch_thre=0)

    def set_epoch(self, epoch: int, epoch_thre: int = 0):
        """adjust_learning_rate for every optimizer.  dmcnet freezes the
        classifier while epoch < epoch_thre (train.py:177,183); the GAN script
        ignores epoch_thre (GAN/train.py:190)."""
        hp = self.hp
        self.freeze = (not self.gan) and epoch < epoch_thre
        decay = hp.lr_decay ** sum(epoch >= s for s in hp.lr_steps)
        for opt, frozen in ((self.opt_cls, self.freeze), (self.opt_gf, False), (self.opt_d, False)):
            if opt is None:
                continue
            lr, wd = hp.lr * decay, hp.weight_decay
            if frozen:
                lr, wd = 0.0, 0.0
            for g in opt.param_groups:
                g['lr'] = lr * g['lr_mult']
                g['weight_decay'] = wd * g['decay_mult']

    def _zero(self):
        for opt in (self.opt_cls, self.opt_gf, self.opt_d):
            if opt is not None:
                opt.zero_grad(set_to_none=False)

    def grads(self) -> Dict[str, Tensor]:
        return {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v))
                for k, v in self.st.items() if not is_buffer(k)}

    def state_dict(self) -> Dict[str, Tensor]:
        return OrderedDict((k, v.detach().clone()) for k, v in self.st.items())

    def _flow_loss(self, gen_flow: Tensor, flow: Tensor, att_flow) -> Tensor:
        """code/dmcnet/train.py:244-247 (GAN/train.py:349-352): with pixel attention both the
        generated and the target flow are weighted by the attention map."""
        if self.att:
            return self.criterion_mse(att_flow[0] * gen_flow, att_flow[0] * flow)
        return self.criterion_mse(gen_flow, flow)

    def step(self, input_flow: Tensor, input_mv: Tensor, input_residual: Tensor, target: Tensor,
             masks: Optional[Sequence[Tensor]] = None, apply: bool = True) -> Dict[str, float]:
        hp, S = self.hp, self.hp.num_segments
        ce = F.cross_entropy
        flow = input_flow.reshape((-1,) + tuple(input_mv.shape[-3:]))         # train.py:230
        out: Dict[str, float] = {}
        if not self.gan:
            output, gen_flow, *att_flow = model_forward(self.st, input_mv, input_residual, train=True,
                                                        **self.gen_kw)
            output = output.view((-1, S) + tuple(output.shape[1:])).mean(dim=1)   # :239-240
            loss_cls = ce(output, target)                                     # :241
            loss_mse = self._flow_loss(gen_flow, flow, att_flow)              # :244-247
            loss = loss_cls * hp.lr_cls + loss_mse * hp.lr_mse                # :248
            self._zero()
            if self.freeze:                                                   # :260-265
                (loss_mse * hp.lr_mse).backward()
            else:
                loss.backward()
                if apply:
                    self.opt_cls.step()
            if apply:
                self.opt_gf.step()                                            # :266
            out.update(loss_mse=float(loss_mse.detach()))
        else:
            valid = torch.ones(target.shape[0] * S, dtype=torch.int64)        # GAN/train.py:253-256
            fake = torch.zeros_like(valid)
            if self.iteration % 2 == 0:                                       # D-step :261-302
                output, validity, gen_flow, *att_flow = model_forward(
                    self.st, input_mv, input_residual, flow, gan=True, arch_d=self.arch_d,
                    train=True, masks=masks, **self.gen_kw)
                output = output.view((-1, S) + tuple(output.shape[1:])).mean(dim=1)
                loss_cls = ce(output, target)
                adv_t = torch.cat((fake, valid), 0)
                loss_adv = ce(validity, adv_t)                                # :274
                loss = loss_cls * hp.lr_cls + loss_adv * hp.lr_adv_d          # :278
                self._zero()
                loss.backward()
                if apply:
                    self.opt_cls.step()
                    self.opt_d.step()
                out.update(acc_adv=accuracy(validity.detach(), adv_t)[0])
            else:                                                             # G-step :331-371
                output, validity, gen_flow, *att_flow = model_forward(
                    self.st, input_mv, input_residual, None, gan=True, arch_d=self.arch_d,
                    train=True, masks=masks, **self.gen_kw)
                output = output.view((-1, S) + tuple(output.shape[1:])).mean(dim=1)
                loss_cls = ce(output, target)
                loss_adv = ce(validity, valid)                                # :346
                loss_mse = self._flow_loss(gen_flow, flow, att_flow)          # :349-352
                loss = loss_cls * hp.lr_cls + loss_adv * hp.lr_adv_g + loss_mse * hp.lr_mse
                self._zero()
                loss.backward()
                if apply:
                    self.opt_gf.step()
                out.update(loss_mse=float(loss_mse.detach()),
                           acc_adv=accuracy(validity.detach(), valid)[0])
            out.update(loss_adv=float(loss_adv.detach()))
            self.last_validity = validity.detach()
        prec1, prec5 = accuracy(output.detach(), target, topk=(1, 5))         # :250
        out.update(loss=float(loss.detach()), loss_cls=float(loss_cls.detach()), prec1=prec1, prec5=prec5)
        self.last_output = output.detach()
        self.last_gen_flow = gen_flow.detach()
        self.iteration += 1
        return out


def validate_batch(state: Dict[str, Tensor], hp: HParams, input_flow: Tensor, input_mv: Tensor,
                   input_residual: Tensor, target: Tensor, *, gan: bool = False,
                   arch_d: Optional[str] = None) -> Dict[str, float]:
    """One iteration of ``validate``: code/dmcnet/train.py:309-347 (``gan=False``) and
    code/dmcnet_GAN/train.py:417-459 (``gan=True``; it reports no total loss, the sum below is
    the G-step weighting of :355 for convenience)."""
    S = hp.num_segments
    st = {k: v.detach() for k, v in state.items()}
    flow = input_flow.reshape((-1,) + tuple(input_mv.shape[-3:]))
    crit = flow_criterion(hp.loss_mse)
    out: Dict[str, float] = {}
    with torch.no_grad():
        if not gan:
            output, gen_flow = model_forward(st, input_mv, input_residual, train=False)
        else:
            output, validity, gen_flow = model_forward(st, input_mv, input_residual, None, gan=True,
                                                       arch_d=arch_d, train=False)
        output = output.view((-1, S) + tuple(output.shape[1:])).mean(dim=1)
        loss_cls = F.cross_entropy(output, target)
        loss_mse = crit(gen_flow, flow)
        loss = loss_cls * hp.lr_cls + loss_mse * hp.lr_mse
        if gan:
            valid = torch.cat([target.clone().fill_(1)] * S, 0)                 # GAN/train.py:442
            loss_adv = F.cross_entropy(validity, valid)
            loss = loss + loss_adv * hp.lr_adv_g
            out.update(loss_adv=float(loss_adv), acc_adv=accuracy(validity, valid)[0])
    prec1, prec5 = accuracy(output, target, topk=(1, 5))
    out.update(loss=float(loss), loss_cls=float(loss_cls), loss_mse=float(loss_mse), prec1=prec1, prec5=prec5)
    return out


def infer_video_scores(state: Dict[str, Tensor], input_mv: Tensor, input_residual: Tensor,
                       segments: int) -> Tensor:
    """Inference flavour (BASELINE config 1): code/dmcnet/test.py:139-151 --
    eval-mode forward then mean of the logits over segments*crops."""
    st = {k: v.detach() for k, v in state.items()}
    with torch.no_grad():
        scores, _ = model_forward(st, input_mv, input_residual, train=False)
    return scores.view((-1, segments) + tuple(scores.shape[1:])).mean(dim=1)
