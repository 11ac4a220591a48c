"""Drop-in ``I3D`` (code/dmcnet_I3D/network/i3d.py:435-533): same constructor signature, module tree,
``state_dict`` keys / shapes and seeded random init as the reference; ``forward(inp, node, detach)`` is one
autograd node running the I3DEngine (no PyTorch compute path: a CPU tensor or an unsupported configuration
raises).  ``from network.i3d import I3D`` keeps working through ``dropin/dmcnet_I3D/network/i3d.py``.
"""
from __future__ import annotations

import contextlib
import io
from collections import OrderedDict
from typing import Optional

import torch
from torch import nn

from .i3d_engine import GEN_TABLE, MIXED, I3DEngine
from .model import build_estimator


def get_padding_shape(filter_shape, stride):
    """i3d.py:299-315 (TF "SAME": pad_along = max(k - s, 0), the smaller half in front)."""
    shape = []
    for k, s in zip(filter_shape, stride):
        pad = max(k - s, 0)
        shape += [pad // 2, pad - pad // 2]
    return tuple(shape[2:] + shape[:2])


class Unit3Dpy(nn.Module):
    """Parameter shell of i3d.py:318-373 (Conv3d [+ BatchNorm3d]); the engine runs it."""

    def __init__(self, in_channels, out_channels, kernel_size=(1, 1, 1), stride=(1, 1, 1), activation='relu',
                 padding='SAME', use_bias=False, use_bn=True, squeeze=False, mean=False):
        super().__init__()
        if padding not in ('SAME', 'VALID'):
            raise ValueError('padding should be in [VALID|SAME] but got {}'.format(padding))
        pad = get_padding_shape(kernel_size, stride) if padding == 'SAME' else (0,) * 6
        simple = all(p == pad[0] for p in pad)
        self.conv3d = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride,
                                padding=(pad[0] if simple else 0), bias=use_bias)
        if use_bn:
            self.batch3d = nn.BatchNorm3d(out_channels)

    def forward(self, inp):
        raise NotImplementedError('Unit3Dpy runs inside I3D.forward (dmcnet_b200 has no per-module PyTorch path)')


class MaxPool3dTFPadding(nn.Module):
    def __init__(self, kernel_size, stride=None, padding='SAME'):
        super().__init__()
        self.kernel_size, self.stride = kernel_size, stride

    def forward(self, inp):
        raise NotImplementedError('MaxPool3dTFPadding runs inside I3D.forward')


class Mixed(nn.Module):
    """i3d.py:391-432."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.branch_0 = Unit3Dpy(in_channels, out_channels[0], kernel_size=(1, 1, 1))
        self.branch_1 = nn.Sequential(Unit3Dpy(in_channels, out_channels[1], kernel_size=(1, 1, 1)),
                                      Unit3Dpy(out_channels[1], out_channels[2], kernel_size=(3, 3, 3)))
        self.branch_2 = nn.Sequential(Unit3Dpy(in_channels, out_channels[3], kernel_size=(1, 1, 1)),
                                      Unit3Dpy(out_channels[3], out_channels[4], kernel_size=(3, 3, 3)))
        self.branch_3 = nn.Sequential(MaxPool3dTFPadding((3, 3, 3), (1, 1, 1)),
                                      Unit3Dpy(in_channels, out_channels[5], kernel_size=(1, 1, 1)))


class _I3DFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, inp, detach, *params):
        eng = model._engine_for(inp)
        ctx.model, ctx.eng, ctx.detach = model, eng, bool(detach)
        if model.training and model.dropout.p > 0:
            eng.set_dropout(model.dropout.p, eng.draw_dropout_mask(model.dropout.p))
        else:
            eng.set_dropout(0.0)
        if eng.has_gen:
            logits, flow = eng.forward_data(inp, train=model.training)
        else:
            eng._ensure_inputs()
            eng.in_mv.copy_(inp.transpose(1, 2).reshape(-1, 2, eng.H, eng.W))
            logits, _ = eng.forward_frames(train=model.training)
            flow = eng.in_mv
        B, T = eng.clips, eng.clip_len
        return logits.clone(), flow.view(B, T, 2, eng.H, eng.W).transpose(1, 2).clone()

    @staticmethod
    def backward(ctx, d_logits, d_flow):
        eng, model = ctx.eng, ctx.model
        eng.d_logits.copy_(d_logits) if d_logits is not None else eng.d_logits.zero_()
        if d_flow is not None:
            eng.d_gen_flow.copy_(d_flow.transpose(1, 2).reshape(-1, 2, eng.H, eng.W))
        else:
            eng.d_gen_flow.zero_()
        eng.zero_grads()
        eng.backward(eng.N, cls=True, cls_wgrad=True, gen_grad=eng.has_gen, cls_to_gen=eng.has_gen and not ctx.detach)
        return (None, None, None) + tuple(eng.grad_view(k).clone() for k in model._param_keys)


class I3D(nn.Module):
    def __init__(self, num_classes, modality='rgb', dropout_prob=0, arch_estimator=None, arch_d=None,
                 name='inception', **kwargs):
        super().__init__()
        self.name = name
        self.num_classes = num_classes
        if modality == 'rgb':
            in_channels = 3
        elif modality in ['flow', 'mv', 'flow+mp4']:
            in_channels = 2
        else:                                    # `elif modality == 'res' or 'I'` (i3d.py:451) is always true
            in_channels = 3
        self.modality = modality
        self.arch_estimator = arch_estimator
        if arch_estimator in GEN_TABLE:
            self.gen_flow_model = build_estimator(arch_estimator, 5, 0, 0)
        self.arch_d = arch_d
        if arch_d is not None:
            from . import model as M
            if arch_d in M.DISCRIMINATORS:
                self.discriminator = getattr(M, arch_d)(2)
        self.conv3d_1a_7x7 = Unit3Dpy(out_channels=64, in_channels=in_channels, kernel_size=(7, 7, 7),
                                      stride=(2, 2, 2), padding='SAME')
        self.maxPool3d_2a_3x3 = MaxPool3dTFPadding((1, 3, 3), (1, 2, 2))
        self.conv3d_2b_1x1 = Unit3Dpy(out_channels=64, in_channels=64, kernel_size=(1, 1, 1), padding='SAME')
        self.conv3d_2c_3x3 = Unit3Dpy(out_channels=192, in_channels=64, kernel_size=(3, 3, 3), padding='SAME')
        self.maxPool3d_3a_3x3 = MaxPool3dTFPadding((1, 3, 3), (1, 2, 2))
        for mname, cin, oc in MIXED:
            if mname == 'mixed_4b':
                self.maxPool3d_4a_3x3 = MaxPool3dTFPadding((3, 3, 3), (2, 2, 2))
            elif mname == 'mixed_5b':
                self.maxPool3d_5a_2x2 = MaxPool3dTFPadding((2, 2, 2), (2, 2, 2))
            setattr(self, mname, Mixed(cin, oc))
        self.avg_pool = nn.AvgPool3d((2, 7, 7), (1, 1, 1))
        self.dropout = nn.Dropout(dropout_prob)
        self.conv3d_0c_1x1 = Unit3Dpy(in_channels=1024, out_channels=400, kernel_size=(1, 1, 1), activation=None,
                                      use_bias=True, use_bn=False, squeeze=True, mean=True)
        self.classifier = nn.Linear(400, num_classes)
        self.softmax = nn.Softmax(1)
        self._in_channels = in_channels
        self._engine: Optional[I3DEngine] = None
        self._param_keys = None

    def _engine_for(self, inp) -> I3DEngine:
        if self._in_channels != 2:
            raise NotImplementedError('dmcnet_b200 runs I3D on two-channel stacks (modality flow | mv | flow+mp4); '
                                      'this configuration has no kernels')
        if not inp.is_cuda:
            raise RuntimeError('dmcnet_b200: inputs must be CUDA tensors (no CPU path exists)')
        B, T, H, W = int(inp.shape[0]), int(inp.shape[2]), int(inp.shape[3]), int(inp.shape[4])
        eng = self._engine
        if eng is None or (eng.clips, eng.clip_len) != (B, T):
            gen = self.arch_estimator if self.arch_estimator in GEN_TABLE else None
            arch_d = self.arch_d if getattr(self, 'discriminator', None) is not None else None
            new = I3DEngine(self.num_classes, B, T, arch_estimator=gen, arch_d=arch_d, height=H, width=W,
                            device=inp.device, share_from=eng)
            if eng is None:
                new.load_state(self.state_dict())
                named_p = dict(self.named_parameters())
                for k in new.specs:
                    named_p[k].data = new.param_view(k)
                mods = dict(self.named_modules())
                for k, b in new.buffers.items():
                    mod, _, leaf = k.rpartition('.')
                    mods[mod]._buffers[leaf] = b
                self._param_keys = list(new.specs.keys())
            self._engine = new
        return self._engine

    def forward(self, inp, node='logit', detach=False):
        if node == 'D':
            raise NotImplementedError('I3D(node="D") as a separate autograd call is not built: the adversarial stages '
                                      'run inside dmcnet_b200.i3d_trainer.I3DTrainStep (adv > 0)')
        self._engine_for(inp)
        named_p = dict(self.named_parameters())
        params = [named_p[k] for k in self._param_keys]
        out, flow = _I3DFunction.apply(self, inp.contiguous().float(), detach, *params)
        if node == 'flow+logit':
            return out, flow
        if node == 'gen_flow':
            return flow
        return out


def build_i3d_state(num_class: int, arch_estimator: Optional[str] = 'DenseNetTiny', seed: Optional[int] = 1
                    ) -> "OrderedDict[str, torch.Tensor]":
    """state_dict of a freshly constructed I3D(num_class, 'flow+mp4', arch_estimator=...) under `seed`."""
    if seed is not None:
        torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        m = I3D(num_class, modality='flow+mp4', arch_estimator=arch_estimator)
    return OrderedDict((k, v.detach().clone()) for k, v in m.state_dict().items())
