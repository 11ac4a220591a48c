"""CPU stand-ins that let the HOST logic of ``FusedTrainStep`` run without a GPU (tests only).

``SimEngine`` is a ``DmcEngine`` whose parameter table, flat buckets and public attributes are
built by the engine's own code on the CPU, but whose ``forward`` / ``backward`` are computed with
torch autograd on the oracle's functional model instead of kernels; ``patch_ops`` replaces the
handful of C-ABI calls the trainer issues directly (heads, Adam, memset) by torch expressions of
the documented kernel semantics (include/dmc_b200.h).  What is exercised for real: mode
selection, loss-gradient scales, dead-work flags, the Adam chunk / hyper-parameter tables, the
step counters, metrics assembly, validate_batch, checkpoint / resume / warm_start and the epoch
driver.  Nothing here is importable from the product.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

from dmcnet_b200 import ops
from dmcnet_b200.engine import DmcEngine
from oracle import dmc_oracle as O


class SimEngine(DmcEngine):
    def __init__(self, num_class, num_segments, frames, *, gan=False, arch_d=None, gen_flow_or_delta=1,
                 height=224, width=224, gen_growth=(8, 8, 6, 4, 2), share_from=None, arch_estimator=None,
                 att=0, gen_flow_ds_factor=0):
        self.device = torch.device('cpu')
        self._share_from = share_from
        # generator selection, as DmcEngine.__init__ (parameter table only; the simulated forward below
        # implements the dense family)
        self._arch_estimator = arch_estimator
        self.gen_arch = 'context' if arch_estimator == 'ContextNetwork' else 'dense'
        self.gen_fusion = {'DenseNetTinyEarlyFusionSum': 'sum', 'DenseNetTinyEarlyFusionStack': 'stack'}.get(
            arch_estimator or '')
        self.gen_ds = int(gen_flow_ds_factor)
        self.att = int(att == 1 and self.gen_arch == 'context')
        self.num_class, self.S, self.N = num_class, num_segments, frames
        self.gan, self.arch_d = gan, (arch_d if gan else None)
        self.gen_flow_or_delta = gen_flow_or_delta
        self.gen_growth = tuple(gen_growth)
        self.H, self.W = height, width
        self._build_param_table()
        N, H, W = frames, height, width
        self.gen_ctot = 5 + sum(self.gen_growth)
        self.dD = torch.zeros(N, 2 + self.gen_ctot - 5, H, W)
        self.d_gen_flow = self.dD[:, 0:2]
        self.gen_flow = torch.zeros(N, 2, H, W)
        self.logits, self.d_logits = torch.zeros(N, num_class), torch.zeros(N, num_class)
        self.validity, self.d_validity = torch.zeros(2 * N, 2), torch.zeros(2 * N, 2)
        self._masks, self._m, self._graph = None, N, None

    def sibling(self, frames):
        return SimEngine(self.num_class, self.S, frames, gan=self.gan, arch_d=self.arch_d,
                         gen_flow_or_delta=self.gen_flow_or_delta, height=self.H, width=self.W,
                         gen_growth=self.gen_growth, share_from=self)

    # -- kernels replaced by autograd on the oracle's functional model
    def forward(self, input_mv, input_residual, input_flow=None, *, train=True, masks=None, use_dropout=True):
        st = OrderedDict()
        for k in self.state_keys():
            st[k] = self.param_view(k).detach().clone().requires_grad_(True) if k in self.specs \
                else self.buffers[k]                        # buffers are updated in place, as BatchNorm does
        n = self.N
        mk = None
        if self.gan and train and use_dropout:
            if not (isinstance(masks, str) and masks == 'preloaded') and masks is not None:
                self.set_masks(masks, 2 * n if input_flow is not None else n)
            mk = self._masks
        with torch.enable_grad():
            out = O.model_forward(st, input_mv.reshape(-1, 2, self.H, self.W), input_residual.reshape(-1, 3, self.H, self.W),
                                  input_flow, gan=self.gan, arch_d=self.arch_d, train=train,
                                  gen_flow_or_delta=self.gen_flow_or_delta, masks=mk)
        self._graph = (st, out)
        self.logits[:n].copy_(out[0].detach())
        self.gen_flow[:n].copy_(out[-1].detach())
        if not self.gan:
            return self.logits[:n], self.gen_flow[:n]
        self._m = out[1].shape[0]
        self.validity[:self._m].copy_(out[1].detach())
        return self.logits[:n], self.validity[:self._m], self.gen_flow[:n]

    def set_masks(self, masks, m):
        self._masks = [mk.reshape(m, -1).to(torch.float32).clone() for mk in masks]

    def backward(self, n, *, cls=True, cls_wgrad=True, gen_grad=True, cls_to_gen=False, disc=False,
                 disc_wgrad=False, disc_to_gen=False):
        st, out = self._graph
        outs, gouts = [out[-1]], [self.d_gen_flow[:n].clone()]
        if cls:
            outs.append(out[0]); gouts.append(self.d_logits[:n].clone())
        if disc:
            outs.append(out[1]); gouts.append(self.d_validity[:self._m].clone())
        leaves = [st[k] for k in self.specs]
        grads = torch.autograd.grad(outs, leaves, gouts, allow_unused=True)
        want = {'base_model': cls and cls_wgrad, 'gen_flow_model': gen_grad, 'discriminator': disc and disc_wgrad}
        for k, g in zip(self.specs, grads):
            tag = next(t for t in want if k.startswith(t))
            if g is not None and want[tag]:
                self.g(k).add_(g.reshape(-1))


def patch_ops(monkeypatch):
    """torch expressions of the C-ABI calls FusedTrainStep issues itself."""

    def ce_head(logits, B, S, C, target, gscale, consensus, dlogits, out_stats):
        lg = logits[:B * S].detach().clone().requires_grad_(True)
        out = lg.view(B, S, C).mean(1)
        loss = F.cross_entropy(out, target[:B], reduction='sum')
        if consensus is not None:
            consensus[:B].copy_(out.detach())
        if dlogits is not None:
            dlogits[:B * S].copy_(torch.autograd.grad(loss, lg)[0] * gscale)
        _, pred = out.detach().topk(min(5, C), 1, True, True)
        hit = pred.eq(target[:B].view(-1, 1))
        out_stats[0], out_stats[1], out_stats[2] = float(loss.detach()), float(hit[:, :1].sum()), float(hit.sum())

    def flow_loss_head(kind, gen, flow, numel, gscale, dgen, loss_sum, frame_elems=None, dgen_ns=None):
        d = (gen.reshape(-1)[:numel] - flow.reshape(-1)[:numel]).double()
        if kind == 0:
            val, slope = d * d, 2 * d
        elif kind == 1:
            val, slope = torch.where(d.abs() < 1, 0.5 * d * d, d.abs() - 0.5), d.clamp(-1, 1)
        else:
            val, slope = d.abs(), d.sign()
        loss_sum[0] = float(val.sum())
        if dgen is not None:
            fe = frame_elems or numel
            ns = dgen_ns or fe
            frames = numel // fe
            dgen.view(-1)[:frames * ns].view(frames, ns)[:, :fe].copy_((gscale * slope).float().view(frames, fe))

    def mse_head(gen, flow, numel, gscale, dgen, loss_sum, frame_elems=None, dgen_ns=None):
        # dmc_mse_head's gscale already carries the factor 2 of d(d^2)
        flow_loss_head(0, gen, flow, numel, gscale / 2.0, dgen, loss_sum, frame_elems, dgen_ns)

    def adam_step(p, g, m, v, chunks, nchunks, hyper, step, beta1, beta2, eps, grad_scale=1.0):
        step += 1
        t = int(step[0])
        bc1, bc2 = 1.0 - beta1 ** t, 1.0 - beta2 ** t
        for off, cnt, ti, _ in chunks[:nchunks].tolist():
            lr, wd = float(hyper[2 * ti]), float(hyper[2 * ti + 1])
            sl = slice(off, off + cnt)
            gk = g[sl] * grad_scale + wd * p[sl]
            m[sl] = m[sl] + (gk - m[sl]) * (1 - beta1)
            v[sl] = beta2 * v[sl] + (1 - beta2) * gk * gk
            p[sl] = p[sl] - (lr / bc1) * (m[sl] / (v[sl].sqrt() / (bc2 ** 0.5) + eps))

    monkeypatch.setattr(ops, 'ce_head', ce_head)
    monkeypatch.setattr(ops, 'mse_head', mse_head)
    monkeypatch.setattr(ops, 'flow_loss_head', flow_loss_head)
    monkeypatch.setattr(ops, 'adam_step', adam_step)
    monkeypatch.setattr(ops, 'memset_zero', lambda t: t.zero_())
    monkeypatch.setattr(ops, 'launch_count', lambda: 0)
