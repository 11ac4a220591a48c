"""Drop-in ``model`` module: the reference's public nn.Module surface on top of
the B200 engine.

Mirrors code/dmcnet/model.py:253-378 (``Model``) and
code/dmcnet_GAN/model.py:442-585 (GAN ``Model`` with ``arch_d`` and the optional
``input_flow`` argument): identical constructor signatures, attribute names
(``base_model``, ``gen_flow_model``, ``discriminator``, ``num_segments``),
``named_parameters`` / ``state_dict`` keys, ``crop_size`` / ``scale_size`` /
``get_augmentation``, and the same error behaviour (``ValueError`` for an
unknown base model; unknown ``arch_estimator`` / ``arch_d`` strings leave the
attribute undefined -> ``AttributeError`` at use).

``forward`` runs the hand-written kernels through ``DmcEngine`` and is wrapped in
a ``torch.autograd.Function`` so an unmodified reference-style training loop
(``loss.backward()``; torch optimizers) works: parameters are views into the
engine's flat bucket, gradients come from the engine's backward pass.  Only the
shipped recipe is native: ``base_model='resnet18'``, ``arch_estimator=
'DenseNetTiny'``, ``representation='mv'``; other generator / backbone choices
construct (so checkpoints load) but ``forward`` raises ``NotImplementedError`` --
there is no PyTorch fallback path.
"""
from __future__ import annotations

import contextlib
import io
from collections import OrderedDict
from typing import Optional

import torch
from torch import nn

from .engine import DmcEngine, disc_blocks, GEN_GROWTH, GEN_IN


# ---------------------------------------------------------------- generator / discriminator shells
def conv(in_planes, out_planes, kernel_size=3, stride=1, padding=1, dilation=1):
    """code/dmcnet/model.py:111-115."""
    return nn.Sequential(
        nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=padding,
                  dilation=dilation, bias=True),
        nn.LeakyReLU(0.1))


def predict_flow(in_planes):
    """code/dmcnet/model.py:118-119."""
    return nn.Conv2d(in_planes, 2, kernel_size=3, stride=1, padding=1, bias=True)


# growth of the dense estimators (new channels per layer): code/dmcnet/model.py:122-194
DENSE_GROWTH = {'DenseNetTiny': GEN_GROWTH, 'DenseNetSmall': (32, 32, 24, 16, 8),
                'DenseNet': (128, 128, 96, 64, 32)}


class _EstimatorShell(nn.Module):
    """Generators are parameter containers with the reference's names; the computation is done
    by the fused engine, never by torch modules (no PyTorch fallback)."""

    def forward(self, x):  # pragma: no cover
        raise NotImplementedError('%s runs inside dmcnet_b200.engine.DmcEngine' % type(self).__name__)


class EstimatorDense(_EstimatorShell):
    """EstimatorDenseNet / ...Small / ...Tiny (code/dmcnet/model.py:122-194): five dense-connected
    3x3 convs ``conv_k`` then ``predict_flow``; only the growth table differs."""

    def __init__(self, ch_in, growth=GEN_GROWTH):
        super().__init__()
        cin = ch_in
        for k, g in enumerate(growth):
            setattr(self, 'conv_%d' % k, conv(cin, g))
            cin += g
        self.predict_flow = predict_flow(cin)


class EstimatorDenseNetTiny(EstimatorDense):
    def __init__(self, ch_in):
        super().__init__(ch_in, DENSE_GROWTH['DenseNetTiny'])


class EstimatorDenseNetSmall(EstimatorDense):
    def __init__(self, ch_in):
        super().__init__(ch_in, DENSE_GROWTH['DenseNetSmall'])


class EstimatorDenseNet(EstimatorDense):
    def __init__(self, ch_in):
        super().__init__(ch_in, DENSE_GROWTH['DenseNet'])


class _EarlyFusion(_EstimatorShell):
    """...EarlyFusionSum / ...EarlyFusionStack (code/dmcnet/model.py:197-250): separate first convs
    for the motion vectors and the residual, summed (8 channels) or stacked (16)."""

    def __init__(self, stack):
        super().__init__()
        self.conv_0_mv = conv(2, 8)
        self.conv_0_r = conv(3, 8)
        dd = 16 if stack else 8
        for k, g in zip((1, 2, 3, 4), (8, 6, 4, 2)):
            setattr(self, 'conv_%d' % k, conv(dd, g))
            dd += g
        self.predict_flow = predict_flow(dd)


class EstimatorDenseNetTinyEarlyFusionSum(_EarlyFusion):
    def __init__(self, ch_in):                 # ch_in is unused by the reference too (model.py:198-201)
        super().__init__(stack=False)


class EstimatorDenseNetTinyEarlyFusionStack(_EarlyFusion):
    def __init__(self, ch_in):
        super().__init__(stack=True)


class Flatten(nn.Module):
    """code/dmcnet/model.py:20-28."""

    def forward(self, x):
        return x.view(x.size(0), -1)


def conv_dilation(batch_norm, in_planes, out_planes, kernel_size=3, stride=1, dilation=1):
    """code/dmcnet/model.py:31-42."""
    pad = ((kernel_size - 1) * dilation) // 2
    if batch_norm:
        return nn.Sequential(
            nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, dilation=dilation,
                      padding=pad, bias=False),
            nn.BatchNorm2d(out_planes), nn.LeakyReLU(0.1, inplace=True))
    return nn.Sequential(
        nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, dilation=dilation,
                  padding=pad, bias=True),
        nn.LeakyReLU(0.1, inplace=True))


class _Context(_EstimatorShell):
    """ContextNetwork / ContextNetworkAtt (code/dmcnet/model.py:45-104): dilated 3x3 conv + BN +
    LeakyReLU(0.1) stack 5-32-128-128-96-64-32(-2); dilation 16 of the fifth layer drops to 1 when
    the generator runs on a down-sampled input; ``att`` splits the head into flow and attention."""

    def __init__(self, ch_in, batch_norm, gen_flow_ds_factor, att):
        super().__init__()
        widths = (32, 128, 128, 96, 64, 32) + (() if att else (2,))
        dil = (1, 2, 4, 8, 16 if gen_flow_ds_factor == 0 else 1, 1, 1)
        layers, cin = [], ch_in
        for w, d in zip(widths, dil):
            layers.append(conv_dilation(batch_norm, cin, w, 3, 1, d))
            cin = w
        self.conv_context = nn.Sequential(*layers)
        if att:
            self.predict_flow = conv_dilation(batch_norm, 32, 2, 3, 1, 1)
            self.predict_att = nn.Sequential(conv_dilation(batch_norm, 32, 2, 3, 1, 1), nn.ReLU(inplace=True))


class ContextNetwork(_Context):
    def __init__(self, ch_in, batch_norm=True, gen_flow_ds_factor=0):
        super().__init__(ch_in, batch_norm, gen_flow_ds_factor, att=False)


class ContextNetworkAtt(_Context):
    def __init__(self, ch_in, batch_norm=True, gen_flow_ds_factor=0):
        super().__init__(ch_in, batch_norm, gen_flow_ds_factor, att=True)


def build_estimator(arch_estimator, ch_in, gen_flow_ds_factor=0, att=0):
    """The selection of code/dmcnet/model.py:310-325 (GAN: code/dmcnet_GAN/model.py:501-517).  The
    dmcnet variant tests ``arch_estimator is 'ContextNetwork'`` (identity, :311), which only holds
    for the default argument; equality is used here, as the GAN variant does.  An unknown string
    returns None: the attribute stays undefined, as in the reference."""
    if arch_estimator == 'ContextNetwork':
        if att == 0:
            return ContextNetwork(ch_in, True, gen_flow_ds_factor)
        if att == 1:
            return ContextNetworkAtt(ch_in, True, gen_flow_ds_factor)
        return None
    named = {'DenseNet': EstimatorDenseNet, 'DenseNetSmall': EstimatorDenseNetSmall,
             'DenseNetTiny': EstimatorDenseNetTiny,
             'DenseNetTinyEarlyFusionSum': EstimatorDenseNetTinyEarlyFusionSum,
             'DenseNetTinyEarlyFusionStack': EstimatorDenseNetTinyEarlyFusionStack}
    if arch_estimator in named:
        return named[arch_estimator](ch_in)
    return None


def _disc_block(in_filters, out_filters, stride, bn):
    """Conv(bias) -> LeakyReLU(0.2) -> Dropout2d(0.25) -> BatchNorm2d(out, eps=0.8)
    (code/dmcnet_GAN/model.py:254-279; the reference builds and discards a bn-less
    block first, which consumes one extra Conv2d init from the RNG)."""
    layers = [nn.Conv2d(in_filters, out_filters, 3, stride, 1), nn.LeakyReLU(0.2, inplace=True),
              nn.Dropout2d(0.25)]
    if bn:
        layers = [nn.Conv2d(in_filters, out_filters, 3, stride, 1), nn.LeakyReLU(0.2, inplace=True),
                  nn.Dropout2d(0.25), nn.BatchNorm2d(out_filters, 0.8)]
    return nn.Sequential(*layers)


def discriminator_block(in_filters, out_filters, bn=True):
    """Stride-2 block, code/dmcnet_GAN/model.py:254-265."""
    return _disc_block(in_filters, out_filters, 2, bn)


def discriminator_block2(in_filters, out_filters, bn=True):
    """Stride-1 block, code/dmcnet_GAN/model.py:268-279."""
    return _disc_block(in_filters, out_filters, 1, bn)


class _Discriminator(nn.Module):
    """Discriminator / Discriminator2..5 parameter container (GAN/model.py:282-438)."""

    def __init__(self, arch_d, height=224, width=224):
        super().__init__()
        for name, ci, co, stride, bn in disc_blocks(arch_d):
            setattr(self, 'discriminator_block_%s' % name, _disc_block(ci, co, stride, bn))
        fc_in = 32 * (height // 8) * (width // 8) if arch_d == 'Discriminator4' \
            else 128 * (height // 16) * (width // 16)
        self.adv_layer = nn.Linear(fc_in, 2)

    def forward(self, x):  # pragma: no cover
        raise NotImplementedError('the discriminator runs inside dmcnet_b200.engine.DmcEngine')


def _named_discriminator(arch_d, doc):
    def __init__(self, ch_in):
        if ch_in != 2:
            raise NotImplementedError('the discriminators take the 2-channel flow map (GAN/model.py:519-529)')
        _Discriminator.__init__(self, arch_d)
    return type(arch_d, (_Discriminator,), {'__init__': __init__, '__doc__': doc})


Discriminator = _named_discriminator('Discriminator', 'code/dmcnet_GAN/model.py:282-300')
Discriminator2 = _named_discriminator('Discriminator2', 'code/dmcnet_GAN/model.py:303-329')
Discriminator3 = _named_discriminator('Discriminator3', 'code/dmcnet_GAN/model.py:332-366')
Discriminator4 = _named_discriminator('Discriminator4', 'code/dmcnet_GAN/model.py:369-385')
Discriminator5 = _named_discriminator('Discriminator5', 'code/dmcnet_GAN/model.py:388-438')


DISCRIMINATORS = ('Discriminator', 'Discriminator2', 'Discriminator3', 'Discriminator4', 'Discriminator5')


# ---------------------------------------------------------------- autograd bridge
class _DmcFunction(torch.autograd.Function):
    """Model.forward as one autograd node: the engine's forward, and its backward
    for ``loss.backward()`` (code/dmcnet/train.py:264)."""

    @staticmethod
    def forward(ctx, model, input_mv, input_residual, input_flow, *params):
        eng = model._engine_for(input_mv)
        ctx.model, ctx.eng = model, eng
        ctx.n = input_mv.numel() // (2 * eng.H * eng.W)
        out = eng.forward(input_mv, input_residual, input_flow, train=model.training)
        return tuple(o.clone() for o in out)

    @staticmethod
    def backward(ctx, *grads):
        eng, model, n = ctx.eng, ctx.model, ctx.n
        from . import ops
        d_att = None
        if eng.att:
            grads, d_att = grads[:-1], grads[-1]
        if eng.gan:
            d_logits, d_validity, d_gen = grads
        else:
            (d_logits, d_gen), d_validity = grads, None
        z = lambda t, g: t.copy_(g) if g is not None else t.zero_()
        z(eng.d_logits[:n], d_logits)
        z(eng.d_gen_flow[:n], d_gen)
        if eng.att:
            z(eng.d_att[:n], d_att)
        if eng.gan:
            z(eng.d_validity[:eng._m], d_validity)
        eng.zero_grads()
        eng.backward(n, cls=True, cls_wgrad=True, gen_grad=True, cls_to_gen=eng.gan, disc=eng.gan,
                     disc_wgrad=True, disc_to_gen=eng.gan)
        pg = [eng.grad_view(k).clone() for k in model._param_keys]
        return (None, None, None, None) + tuple(pg)


# ---------------------------------------------------------------- Model
class DmcModel(nn.Module):
    """Shared implementation; the two drop-in ``Model`` classes below only fix the
    positional signature of their reference counterpart."""

    def __init__(self, num_class, num_segments, representation, base_model='resnet152', new_length=1,
                 use_databn=1, gen_flow_or_delta=0, gen_flow_ds_factor=0,
                 arch_estimator='ContextNetwork', arch_d=None, att=0):
        super().__init__()
        self._representation = representation
        self.num_segments = num_segments
        self.gen_flow_or_delta = gen_flow_or_delta
        self.gen_flow_ds_factor = gen_flow_ds_factor
        self.arch_estimator = arch_estimator
        self.arch_d = arch_d
        self.att = att
        self.new_length = new_length
        self.use_databn = use_databn
        self._num_class = num_class
        self._base_name = base_model
        print(("""
Initializing model:
    base model:         {}.
    input_representation:     {}.
    num_class:          {}.
    num_segments:       {}.
    new_length:       {}.
        """.format(base_model, self._representation, num_class, self.num_segments, self.new_length)))
        self._prepare_base_model(base_model)
        self._prepare_tsn(num_class)
        self._engines = {}
        self._param_keys = None

    # -- construction, in the reference's order (RNG parity with its constructor)
    def _prepare_base_model(self, base_model):
        import torchvision
        if 'resnet' in base_model:
            # pretrained=True in the reference (model.py:305); no network here -> random init
            self.base_model = getattr(torchvision.models, base_model)(weights=None)
            self._input_size = 224
        else:
            raise ValueError('Unknown base model: {}'.format(base_model))
        gen = build_estimator(self.arch_estimator, 5, self.gen_flow_ds_factor, self.att)
        if gen is not None:
            self.gen_flow_model = gen
        if self.arch_d is not None and self.arch_d in DISCRIMINATORS:
            self.discriminator = globals()[self.arch_d](2)

    def _prepare_tsn(self, num_class):
        feature_dim = self.base_model.fc.in_features
        self.base_model.fc = nn.Linear(feature_dim, num_class)
        if self.gen_flow_ds_factor != 0:                     # parameter-free (model.py:326-327)
            self.downsample = nn.AvgPool2d(self.gen_flow_ds_factor, stride=self.gen_flow_ds_factor)
        if self._representation in ('mv', 'flow'):
            self.base_model.conv1 = nn.Conv2d(2 * self.new_length, 64, kernel_size=(7, 7), stride=(2, 2),
                                              padding=(3, 3), bias=False)
            if self.use_databn == 1:                         # created, never called (model.py:295-299)
                self.data_bn = nn.BatchNorm2d(2)
        if self._representation == 'residual' and self.use_databn == 1:
            self.data_bn = nn.BatchNorm2d(3)

    # -- engine plumbing
    def _native_supported(self) -> bool:
        dense = self.arch_estimator in DENSE_GROWTH or self.arch_estimator in (
            'DenseNetTinyEarlyFusionSum', 'DenseNetTinyEarlyFusionStack')
        context = self.arch_estimator == 'ContextNetwork' and self.att in (0, 1)
        # (--att 1 with any other estimator unpacks ONE tensor into two in the reference, model.py:341-342)
        return (self._base_name == 'resnet18' and ((dense and self.att == 0) or context)
                and self._representation == 'mv' and self.new_length == 1)

    def _engine_for(self, input_mv) -> DmcEngine:
        if not hasattr(self, 'gen_flow_model'):
            raise AttributeError("'%s' object has no attribute 'gen_flow_model'" % type(self).__name__)
        if not self._native_supported():
            raise NotImplementedError(
                'dmcnet_b200 runs base_model=resnet18, representation=mv with arch_estimator=DenseNetTiny | '
                'DenseNetSmall | DenseNet | DenseNetTinyEarlyFusionSum | DenseNetTinyEarlyFusionStack (any '
                'gen_flow_ds_factor) or ContextNetwork (att 0 / 1, any gen_flow_ds_factor) natively; this '
                'configuration has no kernels (and there is no PyTorch fallback)')
        if not input_mv.is_cuda:
            raise RuntimeError('dmcnet_b200: inputs must be CUDA tensors (no CPU path exists)')
        H, W = input_mv.shape[-2], input_mv.shape[-1]
        n = input_mv.numel() // (2 * H * W)
        key = (n, H, W, input_mv.device.index)
        if key not in self._engines:
            gan = getattr(self, 'discriminator', None) is not None
            eng = DmcEngine(self._num_class, self.num_segments, n, gan=gan, arch_d=self.arch_d,
                            gen_flow_or_delta=self.gen_flow_or_delta, height=H, width=W,
                            device=input_mv.device,
                            gen_growth=DENSE_GROWTH.get(self.arch_estimator, DENSE_GROWTH['DenseNetTiny']),
                            arch_estimator=self.arch_estimator, gen_flow_ds_factor=self.gen_flow_ds_factor,
                            att=self.att)
            sd = {k: v for k, v in self.state_dict().items() if not k.startswith('data_bn')}
            eng.load_state(sd)
            # parameters and buffers become views of the engine's storage
            named_p = dict(self.named_parameters())
            for k in eng.specs:
                named_p[k].data = eng.param_view(k)
            mods = dict(self.named_modules())
            for k, b in eng.buffers.items():
                mod, _, leaf = k.rpartition('.')
                mods[mod]._buffers[leaf] = b
            self._param_keys = list(eng.specs.keys())
            self._engines = {key: eng}                     # one live engine (it owns the parameters)
        return self._engines[key]

    def forward(self, input_mv, input_residual, input_flow=None):
        eng = self._engine_for(input_mv)
        named_p = dict(self.named_parameters())
        params = [named_p[k] for k in self._param_keys]
        out = _DmcFunction.apply(self, input_mv.contiguous().float(), input_residual.contiguous().float(),
                                 None if input_flow is None else input_flow.contiguous().float(), *params)
        return out

    @property
    def crop_size(self):
        return self._input_size

    @property
    def scale_size(self):
        return self._input_size * 256 // 224

    def get_augmentation(self):
        """code/dmcnet/model.py:369-378 (needs the reference's transforms module on sys.path)."""
        import torchvision
        from transforms import GroupMultiScaleCrop, GroupRandomHorizontalFlip
        scales = [1, .875, .75] if self._representation in ['mv', 'residual', 'flow'] else [1, .875, .75, .66]
        print('Augmentation scales:', scales)
        return torchvision.transforms.Compose(
            [GroupMultiScaleCrop(self._input_size, scales),
             GroupRandomHorizontalFlip()])


class Model(DmcModel):
    """``from model import Model`` of code/dmcnet (model.py:253-255): no discriminator."""

    def __init__(self, num_class, num_segments, representation, base_model='resnet152', new_length=1,
                 use_databn=1, gen_flow_or_delta=0, gen_flow_ds_factor=0,
                 arch_estimator='ContextNetwork', att=0):
        super().__init__(num_class, num_segments, representation, base_model, new_length, use_databn,
                         gen_flow_or_delta, gen_flow_ds_factor, arch_estimator, None, att)

    def forward(self, input_mv, input_residual):
        return super().forward(input_mv, input_residual)


class GANModel(DmcModel):
    """``from model import Model`` of code/dmcnet_GAN (model.py:442-444)."""

    def __init__(self, num_class, num_segments, representation, base_model='resnet152', new_length=1,
                 use_databn=1, gen_flow_or_delta=0, gen_flow_ds_factor=0,
                 arch_estimator='ContextNetwork', arch_d='Discriminator', att=0):
        super().__init__(num_class, num_segments, representation, base_model, new_length, use_databn,
                         gen_flow_or_delta, gen_flow_ds_factor, arch_estimator, arch_d, att)


def build_state(num_class: int, arch_d: Optional[str] = None, seed: Optional[int] = 1,
                arch_estimator: str = 'DenseNetTiny') -> "OrderedDict[str, torch.Tensor]":
    """state_dict of a freshly constructed Model (random init under ``seed``)."""
    if seed is not None:
        torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = DmcModel(num_class, 3, 'mv', base_model='resnet18', arch_estimator=arch_estimator,
                     gen_flow_or_delta=1, use_databn=0, arch_d=arch_d)
    return OrderedDict((k, v.detach().clone()) for k, v in m.state_dict().items())
