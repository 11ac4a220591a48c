"""Device-side execution plan of the DMC-Net hot path.

``DmcEngine`` owns every device buffer (flat parameter / gradient / Adam-moment
buckets, activations, scratch) and sequences the C-ABI kernels for

  * the generator  EstimatorDenseNetTiny   (code/dmcnet/model.py:172-194)
  * the classifier ResNet-18 on the 2-channel DMC map (model.py:283-308)
  * the discriminator blocks               (code/dmcnet_GAN/model.py:254-438)
  * forward  = ``Model.forward``           (dmcnet/model.py:330-357, GAN :533-566)
  * backward = what autograd does in the reference for
    ``loss.backward()`` (dmcnet/train.py:264, GAN/train.py:300,369)

PyTorch is used for memory, streams and (optionally) CUDA-graph capture only;
no torch operator runs on the data path.  Layouts:

  planar   fp32 [N][C][H][W]                    generator, discriminator, stem conv
  pixel    [N][H+2][W+2][C] with a zero ring    classifier, as bf16 hi/lo planes
           (activations, GEMM operands) or fp32 (raw conv outputs, gradients)
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops

GEN_GROWTH = (8, 8, 6, 4, 2)          # code/dmcnet/model.py:175-183
GEN_IN = 5
BN_MOMENTUM = 0.1
# ContextNetwork (code/dmcnet/model.py:45-71): (cout, dilation) of its seven dilated 3x3 convs, each
# followed by BatchNorm2d and LeakyReLU(0.1) -- the last one (-> 2 channels) included
CONTEXT_LAYERS = ((32, 1), (128, 2), (128, 4), (96, 8), (64, 16), (32, 1), (2, 1))
CONTEXT_RING = 16                     # zero ring of the pixel-major layout >= the largest dilation


def disc_blocks(arch_d: str) -> List[Tuple[str, int, int, int, bool]]:
    """(suffix, cin, cout, stride, has_bn) per block; code/dmcnet_GAN/model.py:254-438."""
    if arch_d == 'Discriminator4':
        return [('1', 2, 8, 2, False), ('2', 8, 16, 2, True), ('3', 16, 32, 2, True)]
    extra = {'Discriminator': 0, 'Discriminator2': 1, 'Discriminator3': 2, 'Discriminator5': 4}[arch_d]
    blocks, cin = [], 2
    for stage, cout in enumerate((16, 32, 64, 128), start=1):
        blocks.append((str(stage), cin, cout, 2, stage != 1))
        for j in range(extra):
            blocks.append(('%d_%d' % (stage, j + 2), cout, cout, 1, True))
        cin = cout
    return blocks


class _Geo:
    """Padded pixel-major geometry of one classifier stage."""

    def __init__(self, frames: int, H: int, W: int):
        self.frames, self.H, self.W = frames, H, W
        self.Hp, self.Wp = ops.padded(H), ops.padded(W)     # shared zero ring: row 0 / column 0
        self.P = frames * self.Hp * self.Wp
        self.count = float(frames * H * W)


def _taps_s1(Wp: int):
    shift = [(r - 1) * Wp + (s - 1) for r in range(3) for s in range(3)]
    return shift, [0] * 9, list(range(9))


# stride-2 3x3: kernel row r reads input row 2*oh + r - 1 = phase (r+1)%2 at oh + (r==0 ? -1 : 0)
_S2 = {0: (1, -1), 1: (0, 0), 2: (1, 0)}


def _taps_s2(Wq: int):
    shift, phase, bsel = [], [], []
    for r in range(3):
        for s in range(3):
            (ph, dr), (pw, ds) = _S2[r], _S2[s]
            shift.append(dr * Wq + ds)
            phase.append(ph * 2 + pw)
            bsel.append(r * 3 + s)
    return shift, phase, bsel


class _ConvBN:
    """One conv + BatchNorm unit of the classifier (pixel-major, tensor cores)."""

    def __init__(self, eng: 'DmcEngine', name_conv: str, name_bn: str, cin: int, cout: int, ks: int,
                 stride: int, geo_out: _Geo):
        dev = eng.device
        self.name_conv, self.name_bn = name_conv, name_bn
        self.cin, self.cout, self.ks, self.stride, self.geo = cin, cout, ks, stride, geo_out
        self.taps = ks * ks
        bf = dict(dtype=torch.bfloat16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.W_hi = torch.zeros(self.taps, cout, cin, **bf)
        self.W_lo = torch.zeros(self.taps, cout, cin, **bf)
        self.Wt_hi = self.Wt_lo = None            # assigned by the owner (may be a shared stack)
        self.Y = torch.zeros(geo_out.P, cout, **f32)
        self.act_hi = torch.zeros(geo_out.P, cout, **bf)
        self.act_lo = torch.zeros(geo_out.P, cout, **bf)
        self.sums = eng._alloc_sums(cout)          # zeroed in one memset per forward
        self.sums2 = eng._alloc_sums(cout, bwd=True)   # zeroed in one memset per backward
        self.scale = torch.zeros(cout, **f32)
        self.shift = torch.zeros(cout, **f32)
        self.mean = torch.zeros(cout, **f32)
        self.invstd = torch.ones(cout, **f32)
        self.fbias = torch.zeros(cout, **f32)      # folded BatchNorm shift (eval-mode GEMM bias)


class DmcEngine:
    # defaults for subclasses that build only the parameter table (tests/sim_engine.py)
    gen_arch, gen_fusion, gen_ds, _arch_estimator, att, fold_bn = 'dense', None, 0, None, 0, True

    def __init__(self, num_class: int, num_segments: int, frames: int, *, gan: bool = False,
                 arch_d: Optional[str] = None, gen_flow_or_delta: int = 1, height: int = 224,
                 width: int = 224, device: Optional[torch.device] = None, gemm_engine: str = 'tc',
                 grad_bf16: bool = False, gen_growth: Sequence[int] = GEN_GROWTH,
                 share_from: Optional['DmcEngine'] = None, disc_engine: Optional[str] = None,
                 arch_estimator: Optional[str] = None, gen_flow_ds_factor: int = 0, att: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError('dmcnet_b200: a CUDA device is required (no CPU path exists)')
        if height % 32 or width % 32:
            raise ValueError('height/width must be multiples of 32')
        self.device = device or torch.device('cuda', torch.cuda.current_device())
        self.num_class, self.S, self.N = num_class, num_segments, frames
        self.gan, self.arch_d = gan, (arch_d if gan else None)
        self.gen_flow_or_delta = gen_flow_or_delta
        # channels added by each of the five dense layers: (8,8,6,4,2) = EstimatorDenseNetTiny, the
        # shipped recipe; (32,32,24,16,8) = ...Small and (128,128,96,64,32) = EstimatorDenseNet share
        # the structure (code/dmcnet/model.py:122-194) and the kernels
        self.gen_growth = tuple(int(g) for g in gen_growth)
        if len(self.gen_growth) != 5 or min(self.gen_growth) < 1:
            raise ValueError('gen_growth must list the widths of the five dense layers')
        self.H, self.W = height, width
        self.gemm_engine = gemm_engine
        # generator: the dense-concat family (growth table, CUDA-core kernels) or ContextNetwork, the
        # reference's default --arch_estimator (dilated convs as tcgen05 tap GEMMs)
        self._arch_estimator = arch_estimator
        self.gen_arch = 'context' if arch_estimator == 'ContextNetwork' else 'dense'
        if self.gen_arch == 'context' and gemm_engine != 'tc':
            raise ValueError('ContextNetwork runs on the tensor-core GEMM engine only')
        # EstimatorDenseNetTinyEarlyFusionSum / ...Stack (code/dmcnet/model.py:197-250): separate first convs
        # for the motion vectors and the residual, summed or concatenated, then four dense layers
        self.gen_fusion = {'DenseNetTinyEarlyFusionSum': 'sum', 'DenseNetTinyEarlyFusionStack': 'stack'}.get(
            arch_estimator or '')
        # --gen_flow_ds_factor f: AvgPool2d(f) in, f x f tiling out (model.py:326-327, :335-337, :347-348)
        self.gen_ds = int(gen_flow_ds_factor)
        if self.gen_ds < 0 or (self.gen_ds and (height % self.gen_ds or width % self.gen_ds)):
            raise ValueError('gen_flow_ds_factor must divide the frame size')
        # --att 1: ContextNetworkAtt (model.py:74-104) -- six context layers, then a flow head and an
        # attention head (+ReLU); the flow criterion is weighted by the attention map (train.py:244-247).
        # The reference ignores --att for every other estimator's constructor but still unpacks two
        # outputs, so att is only meaningful with ContextNetwork
        self.att = int(att == 1 and self.gen_arch == 'context')
        # eval-mode forwards (validate(), test.py scoring) fold BatchNorm into the conv operands: the GEMM
        # epilogue adds the folded shift and the residual, applies the activation and writes the next
        # GEMM's bf16 hi/lo operand directly -- no bn_apply pass, no fp32 conv output
        self.fold_bn = True
        # grad_bf16=True: the backward GEMMs (data and weight gradients of the ResNet convs) read
        # the incoming gradient dY rounded to bf16 (2 MMAs per k-step, no dY_lo plane; weights and
        # saved activations keep the full hi/lo split).  Measured on B200: 5% faster step, but the
        # per-layer rounding noise (~1e-3) random-walks to 7e-3 relative at the stem/generator
        # gradients, so it is opt-in; the default keeps every GEMM at the fp32-equivalent split.
        self.grad_bf16 = grad_bf16
        # share_from: alias another engine's parameter / gradient / Adam buckets and BatchNorm buffers
        # (a second execution plan for a different frame count over the SAME model state)
        self._share_from = share_from
        # discriminator execution plan: 'tc' = every conv as a tcgen05 tap GEMM on pixel-major /
        # space-to-depth operands (disc_plan.py, csrc/disc_pm.cu); 'planar' = the CUDA-core NCHW kernels
        # (Discriminator4, odd sizes, and the cross-check of the tensor-core plan)
        if disc_engine is None:
            from . import disc_plan as DP
            disc_engine = 'tc' if (gan and gemm_engine == 'tc' and DP.preferred(arch_d, height, width)) \
                else 'planar'
        self.disc_engine = disc_engine if gan else None
        self._build_param_table()
        if self.gen_arch == 'context':
            self._alloc_context()
        else:
            self._alloc_generator()
        self._alloc_classifier()
        if self.gan:
            if self.disc_engine == 'tc':
                self._alloc_discriminator_tc()
            else:
                self._alloc_discriminator()

    def sibling(self, frames: int) -> 'DmcEngine':
        """A second execution plan for `frames` frames over the SAME model state (shared parameter,
        gradient, Adam buckets and BatchNorm buffers) -- the short last batch of an epoch."""
        return DmcEngine(self.num_class, self.S, frames, gan=self.gan, arch_d=self.arch_d,
                         gen_flow_or_delta=self.gen_flow_or_delta, height=self.H, width=self.W,
                         device=self.device, gemm_engine=self.gemm_engine, grad_bf16=self.grad_bf16,
                         gen_growth=self.gen_growth, share_from=self, disc_engine=self.disc_engine,
                         arch_estimator=self._arch_estimator, gen_flow_ds_factor=self.gen_ds, att=self.att)

    def _alloc_sums(self, cout: int, bwd: bool = False) -> torch.Tensor:
        """[2][cout] double view inside one pool, so all BN statistics are zeroed by one memset."""
        if not hasattr(self, '_sums_pool'):
            self._sums_pool = torch.zeros(20 * 2 * 512, dtype=torch.float64, device=self.device)
            self._sums2_pool = torch.zeros(20 * 2 * 512, dtype=torch.float64, device=self.device)
            self._sums_used = 0
            self._sums2_used = 0
        n = 2 * cout
        pool, used = (self._sums2_pool, self._sums2_used) if bwd else (self._sums_pool, self._sums_used)
        if used + n > pool.numel():
            raise RuntimeError('BN statistics pool exhausted')
        v = pool[used:used + n].view(2, cout)
        if bwd:
            self._sums2_used += n
        else:
            self._sums_used += n
        return v

    # ------------------------------------------------------------------ parameters
    def _param_specs(self) -> "OrderedDict[str, Tuple[int, ...]]":
        C = self.num_class
        specs: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
        specs['base_model.conv1.weight'] = (64, 2, 7, 7)
        specs['base_model.bn1.weight'] = (64,)
        specs['base_model.bn1.bias'] = (64,)
        cin = 64
        for li, (width, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), start=1):
            for b in range(2):
                q = 'base_model.layer%d.%d' % (li, b)
                s = stride if b == 0 else 1
                specs[q + '.conv1.weight'] = (width, cin, 3, 3)
                specs[q + '.bn1.weight'] = (width,)
                specs[q + '.bn1.bias'] = (width,)
                specs[q + '.conv2.weight'] = (width, width, 3, 3)
                specs[q + '.bn2.weight'] = (width,)
                specs[q + '.bn2.bias'] = (width,)
                if b == 0 and (s != 1 or cin != width):
                    specs[q + '.downsample.0.weight'] = (width, cin, 1, 1)
                    specs[q + '.downsample.1.weight'] = (width,)
                    specs[q + '.downsample.1.bias'] = (width,)
                cin = width
        specs['base_model.fc.weight'] = (C, 512)
        specs['base_model.fc.bias'] = (C,)
        cin = GEN_IN
        if self.gen_arch == 'context':
            for name, ci, co, _ in self._context_layout():
                specs[name + '.0.weight'] = (co, ci, 3, 3)
                specs[name + '.1.weight'] = (co,)
                specs[name + '.1.bias'] = (co,)
        else:
            names, growth, base = self._dense_layout()
            if self.gen_fusion:
                specs['gen_flow_model.conv_0_mv.0.weight'] = (8, 2, 3, 3)
                specs['gen_flow_model.conv_0_mv.0.bias'] = (8,)
                specs['gen_flow_model.conv_0_r.0.weight'] = (8, 3, 3, 3)
                specs['gen_flow_model.conv_0_r.0.bias'] = (8,)
            cin = base
            for name, g in zip(names, growth):
                specs[name + '.weight'] = (g, cin, 3, 3)
                specs[name + '.bias'] = (g,)
                cin += g
            specs['gen_flow_model.predict_flow.weight'] = (2, cin, 3, 3)
            specs['gen_flow_model.predict_flow.bias'] = (2,)
        if self.gan:
            for name, ci, co, stride, bn in disc_blocks(self.arch_d):
                p = 'discriminator.discriminator_block_%s' % name
                specs[p + '.0.weight'] = (co, ci, 3, 3)
                specs[p + '.0.bias'] = (co,)
                if bn:
                    specs[p + '.3.weight'] = (co,)
                    specs[p + '.3.bias'] = (co,)
            fc_in = 32 * (self.H // 8) * (self.W // 8) if self.arch_d == 'Discriminator4' \
                else 128 * (self.H // 16) * (self.W // 16)
            specs['discriminator.adv_layer.weight'] = (2, fc_in)
            specs['discriminator.adv_layer.bias'] = (2,)
        return specs

    def _build_param_table(self):
        dev = self.device
        self.specs = self._param_specs()
        self.offsets: Dict[str, int] = {}
        self.group_range: Dict[str, Tuple[int, int]] = {}
        off = 0
        for tag in ('base_model', 'gen_flow_model', 'discriminator'):
            start = off
            for k, shp in self.specs.items():
                if not k.startswith(tag):
                    continue
                n = 1
                for d in shp:
                    n *= d
                self.offsets[k] = off
                off += (n + 63) // 64 * 64
            self.group_range[tag] = (start, off)
        self.total = off
        if getattr(self, '_share_from', None) is not None:
            o = self._share_from
            if list(o.specs.items()) != list(self.specs.items()):
                raise ValueError('share_from: the two engines describe different models')
            self.params, self.grads, self.exp_avg, self.exp_avg_sq = o.params, o.grads, o.exp_avg, o.exp_avg_sq
            self.buffers = o.buffers
            return
        z = lambda: torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.params, self.grads, self.exp_avg, self.exp_avg_sq = z(), z(), z(), z()
        # BatchNorm buffers (state_dict entries that are not parameters)
        self.buffers: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        for k, shp in self.specs.items():
            if k.endswith('.weight') and len(shp) == 1:        # a BatchNorm weight
                base = k[:-len('.weight')]
                self.buffers[base + '.running_mean'] = torch.zeros(shp, dtype=torch.float32, device=dev)
                self.buffers[base + '.running_var'] = torch.ones(shp, dtype=torch.float32, device=dev)
                self.buffers[base + '.num_batches_tracked'] = torch.zeros((), dtype=torch.int64, device=dev)

    def numel(self, key: str) -> int:
        n = 1
        for d in self.specs[key]:
            n *= d
        return n

    def p(self, key: str) -> torch.Tensor:
        o = self.offsets[key]
        return self.params[o:o + self.numel(key)]

    def g(self, key: str) -> torch.Tensor:
        o = self.offsets[key]
        return self.grads[o:o + self.numel(key)]

    def param_view(self, key: str) -> torch.Tensor:
        return self.p(key).view(self.specs[key])

    def grad_view(self, key: str) -> torch.Tensor:
        return self.g(key).view(self.specs[key])

    def state_keys(self) -> List[str]:
        """state_dict key order of the reference ``Model`` (torch module order)."""
        keys: List[str] = []
        for k in self.specs:
            keys.append(k)
            if k.endswith('.bias') and (k[:-len('.bias')] + '.running_mean') in self.buffers:
                base = k[:-len('.bias')]
                keys += [base + '.running_mean', base + '.running_var', base + '.num_batches_tracked']
        return keys

    def load_state(self, sd: Dict[str, torch.Tensor]) -> None:
        for k in self.specs:
            self.p(k).copy_(sd[k].detach().reshape(-1).to(self.device, torch.float32))
        for k, b in self.buffers.items():
            if k not in sd and k.endswith('.num_batches_tracked'):
                # checkpoints written by torch < 0.4.1 (the reference trained with 0.3.1, README.md:28)
                # have no such key; torch's own loader keeps the current value (BN version check)
                continue
            b.copy_(sd[k].detach().to(self.device))

    def state_dict(self) -> "OrderedDict[str, torch.Tensor]":
        out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        for k in self.state_keys():
            out[k] = (self.param_view(k) if k in self.specs else self.buffers[k]).detach().clone()
        return out

    # ------------------------------------------------------------------ allocation
    def _dense_layout(self):
        """(parameter prefixes of the dense layers, their growth, width of the base block).  The base is
        the raw input (mv | residual, 5 channels) for EstimatorDenseNet*, or the fused first-layer
        features for the EarlyFusion variants (8 summed / 16 stacked channels)."""
        if self.gen_fusion:
            return (['gen_flow_model.conv_%d.0' % k for k in (1, 2, 3, 4)], self.gen_growth[1:],
                    16 if self.gen_fusion == 'stack' else 8)
        return ['gen_flow_model.conv_%d.0' % k for k in range(5)], self.gen_growth, GEN_IN

    def _alloc_generator(self):
        dev, N, H, W = self.device, self.N, self.H, self.W
        f32 = dict(dtype=torch.float32, device=dev)
        f = self.gen_ds or 1
        self.gH, self.gW = H // f, W // f                       # resolution the estimator runs at
        gH, gW = self.gH, self.gW
        names, growth, base = self._dense_layout()
        self.gen_names, self.gen_dense, self.gen_base = names, tuple(growth), base
        self.gen_ctot = base + sum(growth)                      # 33 for DenseNetTiny
        self.gen_base_off = self.gen_ctot - base                # X channel of the base block (28)
        base_grad = self.gen_fusion is not None                 # the raw input needs no gradient
        self.X = torch.zeros(N, self.gen_ctot, gH, gW, **f32)
        # gradient buffer of the dense block: [d out (2) | d new_last .. d new_first | d base (fusion only)];
        # the gradient of one slice is a single convolution over ALL channels in front of it
        self.gD = torch.zeros(N, 2 + self.gen_base_off + (base if base_grad else 0), gH, gW, **f32)
        if self.gen_ds:
            self.dD = torch.zeros(N, 2, H, W, **f32)            # d loss / d gen_flow at frame resolution
            self.gen_small = torch.zeros(N, 2, gH, gW, **f32)   # estimator output (+ pooled mv)
            self.in_small = torch.zeros(N, GEN_IN, gH, gW, **f32)     # pooled mv | residual
        else:
            self.dD = self.gD
        self.d_gen_flow = self.dD[:, 0:2]                       # strided view (frame stride 30*H*W for Tiny)
        self.gen_flow = torch.zeros(N, 2, H, W, **f32)
        if self.gen_fusion == 'sum':
            self.ef = [torch.zeros(N, 8, gH, gW, **f32) for _ in range(3)]      # lrelu(conv_mv), lrelu(conv_r), scratch
        # channel offset of each layer's OUTPUT inside X: new channels are prepended
        outs, off = [], self.gen_base_off
        for g in growth:
            off -= g
            outs.append(off)
        self.gen_out_off = outs                                          # [20, 12, 6, 2, 0]
        self.gen_in_off = [o + g for o, g in zip(outs, growth)]          # [28, 20, 12, 6, 2]
        self.wflip = torch.zeros(128 * 128 * 9, **f32)                   # dgrad weight scratch
        # combined (flipped, transposed) dgrad weights of the slices, see _gen_backward
        L = len(growth)
        slices = [(growth[k], outs[k], k) for k in reversed(range(L))]   # (width, X channel, producing layer)
        if base_grad:
            slices.append((base, self.gen_base_off, -1))
        table, off = [len(slices)], 0
        self.gen_wc_off = []
        for gk, xk, k in slices:
            cin_s = 2 + xk
            segs = [(0, 2, self.offsets['gen_flow_model.predict_flow.weight'], self.gen_ctot, xk)]
            for j in range(L - 1, k, -1):               # later dense layers j > k (all of them for the base)
                segs.append((2 + outs[j], growth[j], self.offsets[names[j] + '.weight'],
                             self.gen_ctot - self.gen_in_off[j], xk - self.gen_in_off[j]))
            row = [off, gk, cin_s, len(segs)]
            for g in range(6):
                row += list(segs[g]) if g < len(segs) else [0, 0, 0, 0, 0]
            table += row
            self.gen_wc_off.append(off)
            off += gk * cin_s * 9
        self.gen_slices = slices
        self.gen_wc_table = table
        self.gen_wc = torch.zeros(off, **f32)

    def _alloc_classifier(self):
        dev, N = self.device, self.N
        H2, W2 = self.H // 2, self.W // 2
        f32 = dict(dtype=torch.float32, device=dev)
        # stem (planar)
        self.stem_Y = torch.zeros(N, 64, H2, W2, **f32)
        # stem conv on the tensor cores (im2col built in shared memory) when the shape allows
        self.stem_tc = (self.gemm_engine == 'tc' and self.W % 8 == 0 and 8 <= self.W <= 256
                        and self.H % 2 == 0)
        self.stem_wb = torch.zeros(128 * 128, dtype=torch.bfloat16, device=dev)
        self.stem_ws = torch.empty(ops.stem_wgrad_workspace_floats(), **f32) if self.stem_tc else None
        self.stem_dZ = torch.zeros(N, 64, H2, W2, **f32)
        # data gradient of the stem conv (GAN G-step: the classifier input is not detached) as a 16-tap
        # tensor-core GEMM over the H/2 x W/2 grid: input pixel (2u+a, 2v+b) reads dZ at (u+di, v+dj),
        # di, dj in {-1..2}, through kernel row r = a + 3 - 2*di (csrc/disc_pm.cu, "stem data gradient")
        self.stem_dgrad_tc = self.stem_tc and self.gan
        if self.stem_dgrad_tc:
            import numpy as np
            offs = [(di, dj) for di in (-1, 0, 1, 2) for dj in (-1, 0, 1, 2)]
            gmap = -np.ones((16, 32, 64), np.int32)
            co = np.arange(64)
            for t, (di, dj) in enumerate(offs):
                for a in range(2):
                    for b in range(2):
                        r, q = a + 3 - 2 * di, b + 3 - 2 * dj
                        if 0 <= r <= 6 and 0 <= q <= 6:
                            for c in range(2):
                                gmap[t, (a * 2 + b) * 2 + c, :] = ((co * 2 + c) * 7 + r) * 7 + q
            wp2 = W2 + 2
            self.stem_dg = {
                'gmap': torch.from_numpy(gmap.reshape(-1)).to(dev),
                'shift': [di * wp2 + dj for di, dj in offs],
                'hi': torch.zeros(N * (H2 + 2) * wp2, 64, dtype=torch.bfloat16, device=dev),
                'lo': torch.zeros(N * (H2 + 2) * wp2, 64, dtype=torch.bfloat16, device=dev),
                'W_hi': torch.zeros(16, 32, 64, dtype=torch.bfloat16, device=dev),
                'W_lo': torch.zeros(16, 32, 64, dtype=torch.bfloat16, device=dev),
                'out': torch.zeros(N * (H2 + 2) * wp2, 32, **f32)}
        self.stem = {k: torch.zeros(64, **f32) for k in ('scale', 'shift', 'mean', 'invstd')}
        self.stem['sums'] = torch.zeros(2, 64, dtype=torch.float64, device=dev)
        self.stem['sums2'] = torch.zeros(2, 64, dtype=torch.float64, device=dev)
        g = _Geo(N, self.H // 4, self.W // 4)
        self.geo0 = g
        self.A0_hi = torch.zeros(g.P, 64, dtype=torch.bfloat16, device=dev)
        self.A0_lo = torch.zeros(g.P, 64, dtype=torch.bfloat16, device=dev)
        self.pool_idx = torch.zeros(N * (self.H // 4) * (self.W // 4) * 64, dtype=torch.uint8, device=dev)
        # residual stages
        self.blocks = []
        cin, geo = 64, g
        max_pc, max_w = 0, 0
        for li, (width, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), start=1):
            for b in range(2):
                q = 'base_model.layer%d.%d' % (li, b)
                s = stride if b == 0 else 1
                geo_out = _Geo(N, geo.H // s, geo.W // s)
                blk = {'name': q, 'stride': s, 'geo_in': geo, 'geo': geo_out, 'cin': cin, 'width': width}
                blk['c1'] = _ConvBN(self, q + '.conv1', q + '.bn1', cin, width, 3, s, geo_out)
                blk['c2'] = _ConvBN(self, q + '.conv2', q + '.bn2', width, width, 3, 1, geo_out)
                bf = dict(dtype=torch.bfloat16, device=dev)
                if s != 1 or cin != width:
                    blk['ds'] = _ConvBN(self, q + '.downsample.0', q + '.downsample.1', cin, width, 1, s,
                                        geo_out)
                    # conv1 and downsample share one transposed-weight stack (dgrad B operand)
                    wt_hi = torch.zeros(10, cin, width, **bf)
                    wt_lo = torch.zeros(10, cin, width, **bf)
                    blk['c1'].Wt_hi, blk['c1'].Wt_lo = wt_hi, wt_lo
                    blk['ds'].Wt_hi, blk['ds'].Wt_lo = wt_hi[9:], wt_lo[9:]
                    blk['xp_hi'] = torch.zeros(4, geo_out.P, cin, **bf)     # phase-split block input
                    blk['xp_lo'] = torch.zeros(4, geo_out.P, cin, **bf)
                    blk['dxp'] = torch.zeros(4, geo_out.P, cin, **f32)
                else:
                    blk['c1'].Wt_hi = torch.zeros(9, cin, width, **bf)
                    blk['c1'].Wt_lo = torch.zeros(9, cin, width, **bf)
                blk['c2'].Wt_hi = torch.zeros(9, width, width, **bf)
                blk['c2'].Wt_lo = torch.zeros(9, width, width, **bf)
                self.blocks.append(blk)
                max_pc = max(max_pc, geo_out.P * width, geo.P * cin)
                max_w = max(max_w, 9 * width * max(cin, width))
                cin, geo = width, geo_out
        self.geo_last = geo
        # shared scratch
        self.G_hi = torch.zeros(2 * max_pc, dtype=torch.bfloat16, device=dev)   # dY planes (x2: conv1|ds stack)
        self.G_lo = None if self.grad_bf16 else torch.zeros(2 * max_pc, dtype=torch.bfloat16, device=dev)
        self.gbuf = [torch.zeros(max_pc, **f32) for _ in range(5)]             # T1, T2a, Ra, T2b, Rb
        # split-K workspace of the weight-gradient GEMM (deterministic reduction instead of atomics)
        self.wgrad_ws = None
        if self.gemm_engine == 'tc':
            need = 0
            for blk in self.blocks:
                for key in ('c1', 'c2', 'ds'):
                    if key in blk:
                        u = blk[key]
                        need = max(need, ops.wgrad_workspace_floats(u.geo.P, u.cout, u.cin, u.taps))
            self.wgrad_ws = torch.empty(need, **f32)
        self.pooled = torch.zeros(N, 512, **f32)
        self.d_pooled = torch.zeros(N, 512, **f32)
        self.logits = torch.zeros(N, self.num_class, **f32)
        self.d_logits = torch.zeros(N, self.num_class, **f32)

    def _alloc_discriminator(self):
        dev, H, W = self.device, self.H, self.W
        M = 2 * self.N
        f32 = dict(dtype=torch.float32, device=dev)
        self.d_layers = []
        h, w = H, W
        for name, ci, co, stride, bn in disc_blocks(self.arch_d):
            ho, wo = h // stride, w // stride
            L = {'name': 'discriminator.discriminator_block_%s' % name, 'cin': ci, 'cout': co,
                 'stride': stride, 'bn': bn, 'H': h, 'W': w, 'Ho': ho, 'Wo': wo}
            L['A'] = torch.zeros(M, co, ho, wo, **f32)         # lrelu(conv)*mask
            L['Z'] = torch.zeros(M, co, ho, wo, **f32) if bn else L['A']   # after BN (next input)
            L['mask'] = torch.ones(M, co, **f32)
            if bn:
                for k in ('scale', 'shift', 'mean', 'invstd'):
                    L[k] = torch.zeros(co, **f32)
                L['sums'] = torch.zeros(2, co, dtype=torch.float64, device=dev)
                L['sums2'] = torch.zeros(2, co, dtype=torch.float64, device=dev)
            self.d_layers.append(L)
            h, w = ho, wo
        big = max(L['A'].numel() for L in self.d_layers)
        self.d_in = torch.zeros(M, 2, H, W, **f32)
        self.d_g = [torch.zeros(max(big, self.d_in.numel()), **f32) for _ in range(2)]
        # stride-2 layers whose output width is a multiple of 4 run their backward on the
        # space-to-depth input through the stride-1 kernels (csrc/dense_conv.cu, "via space-to-depth")
        s2 = [L for L in self.d_layers if L['stride'] == 2 and L['Wo'] % 4 == 0 and L['H'] % 2 == 0]
        for L in s2:
            L['s2d'] = True
        if s2:
            n_in = max(M * L['cin'] * L['H'] * L['W'] for L in s2)
            n_w = max(L['cout'] * 4 * L['cin'] * 9 for L in s2)
            self.d_s2d = torch.zeros(n_in, **f32)          # S = space-to-depth(layer input)
            self.d_ds = torch.zeros(n_in, **f32)           # dS
            self.d_w3 = torch.zeros(n_w, **f32)            # W3 / flipped W3 / dW3
            self.d_w3t = torch.zeros(n_w, **f32)
        self.validity = torch.zeros(M, 2, **f32)
        self.d_validity = torch.zeros(M, 2, **f32)
        self.d_feat = torch.zeros(M, self.specs['discriminator.adv_layer.weight'][1], **f32)

    # ------------------------------------------------------------------ generator
    def _gen_forward(self, mv: torch.Tensor, res: torch.Tensor, n: int):
        """EstimatorDenseNet* / ...EarlyFusion*.forward (+ input_mv when gen_flow_or_delta == 1), at 1/f
        resolution between an average pool and a tiling when gen_flow_ds_factor = f
        (code/dmcnet/model.py:172-250, :330-348)."""
        H, W = self.gH, self.gW
        HW = H * W
        X = self.X.view(-1)
        ns = self.gen_ctot * HW
        bo = self.gen_base_off
        if self.gen_ds:
            f, FHW = self.gen_ds, self.H * self.W
            ops.avgpool_planar(mv, n * 2, self.H, self.W, f, self.in_small.view(-1)[:n * 2 * HW])
            ops.avgpool_planar(res, n * 3, self.H, self.W, f, self.in_small.view(-1)[self.N * 2 * HW:])
            mv = self.in_small.view(-1)[:self.N * 2 * HW]
            res = self.in_small.view(-1)[self.N * 2 * HW:]
        self._gen_in = (mv, res)
        q = 'gen_flow_model.'
        if self.gen_fusion is None:
            ops.copy_planar(mv, 2 * HW, X[bo * HW:], ns, 2 * HW, n)
            ops.copy_planar(res, 3 * HW, X[(bo + 2) * HW:], ns, 3 * HW, n)
        elif self.gen_fusion == 'stack':            # x = cat(lrelu(conv_mv(mv)), lrelu(conv_r(res)))
            ops.conv_fwd(mv, 2 * HW, 2, H, W, self.p(q + 'conv_0_mv.0.weight'), self.p(q + 'conv_0_mv.0.bias'), 8, 3, 1,
                         X[bo * HW:], ns, n, slope=0.1)
            ops.conv_fwd(res, 3 * HW, 3, H, W, self.p(q + 'conv_0_r.0.weight'), self.p(q + 'conv_0_r.0.bias'), 8, 3, 1,
                         X[(bo + 8) * HW:], ns, n, slope=0.1)
        else:                                       # x = lrelu(conv_mv(mv)) + lrelu(conv_r(res)); both kept for the backward
            a, b = self.ef[0].view(-1), self.ef[1].view(-1)
            ops.conv_fwd(mv, 2 * HW, 2, H, W, self.p(q + 'conv_0_mv.0.weight'), self.p(q + 'conv_0_mv.0.bias'), 8, 3, 1,
                         a, 8 * HW, n, slope=0.1)
            ops.conv_fwd(res, 3 * HW, 3, H, W, self.p(q + 'conv_0_r.0.weight'), self.p(q + 'conv_0_r.0.bias'), 8, 3, 1,
                         b, 8 * HW, n, slope=0.1)
            ops.conv_fwd(res, 3 * HW, 3, H, W, self.p(q + 'conv_0_r.0.weight'), self.p(q + 'conv_0_r.0.bias'), 8, 3, 1,
                         X[bo * HW:], ns, n, slope=0.1, add=a, add_ns=8 * HW)
        for k, g in enumerate(self.gen_dense):
            cin = self.gen_ctot - self.gen_in_off[k]
            ops.conv_fwd(X[self.gen_in_off[k] * HW:], ns, cin, H, W, self.p(self.gen_names[k] + '.weight'),
                         self.p(self.gen_names[k] + '.bias'), g, 3, 1, X[self.gen_out_off[k] * HW:], ns, n, slope=0.1)
        out = self.gen_small if self.gen_ds else self.gen_flow
        ops.conv_fwd(X, ns, self.gen_ctot, H, W, self.p(q + 'predict_flow.weight'), self.p(q + 'predict_flow.bias'), 2,
                     3, 1, out.view(-1), 2 * HW, n, slope=1.0,
                     add=(mv if self.gen_flow_or_delta == 1 else None), add_ns=2 * HW)
        if self.gen_ds:
            ops.tile_repeat(self.gen_small.view(-1), n * 2, H, W, self.gen_ds, self.gen_flow.view(-1))

    def _gen_backward(self, n: int):
        """Gradients of all generator parameters from d loss / d gen_flow (self.dD[:, 0:2])."""
        H, W = self.gH, self.gW
        HW = H * W
        X, dD = self.X.view(-1), self.gD.view(-1)
        ns, dns = self.gen_ctot * HW, self.gD.shape[1] * HW
        q = 'gen_flow_model.'
        if self.gen_ds:                              # backward of the tiling: sum of the f*f tiles
            ops.tile_sum(self.dD.view(-1), 2 * self.H * self.W, 2, H, W, self.gen_ds, n, dD, dns)
        ops.dense_dgrad_weights(self.params, self.gen_wc_table, self.gen_wc)
        ops.conv_wgrad(X, ns, self.gen_ctot, H, W, dD, dns, 2, 3, 1, self.g(q + 'predict_flow.weight'),
                       self.g(q + 'predict_flow.bias'), n)
        mv, res = self._gen_in
        for idx, (g, oo, k) in enumerate(self.gen_slices):
            cin_s = 2 + oo                                   # d out + every later slice
            wc = self.gen_wc[self.gen_wc_off[idx]:self.gen_wc_off[idx] + g * cin_s * 9]
            masked = k >= 0 or self.gen_fusion == 'stack'    # slice = LeakyReLU(0.1) of ONE conv's output
            # slice gradient in ONE launch (no read-modify-write), LeakyReLU' fused where it applies
            ops.conv3x3_dgrad_fused(dD, dns, cin_s, H, W, wc, g, dD[(2 + oo) * HW:], dns, n, accumulate=False,
                                    act_src=(X[oo * HW:] if masked else None), act_ns=ns,
                                    act_c1=(g if masked else 0), act_slope=0.1)
            if k >= 0:
                io = self.gen_in_off[k]
                ops.conv_wgrad(X[io * HW:], ns, self.gen_ctot - io, H, W, dD[(2 + oo) * HW:], dns, g, 3, 1,
                               self.g(self.gen_names[k] + '.weight'), self.g(self.gen_names[k] + '.bias'), n)
            elif self.gen_fusion == 'stack':
                ops.conv_wgrad(mv, 2 * HW, 2, H, W, dD[(2 + oo) * HW:], dns, 8, 3, 1,
                               self.g(q + 'conv_0_mv.0.weight'), self.g(q + 'conv_0_mv.0.bias'), n)
                ops.conv_wgrad(res, 3 * HW, 3, H, W, dD[(2 + oo + 8) * HW:], dns, 8, 3, 1,
                               self.g(q + 'conv_0_r.0.weight'), self.g(q + 'conv_0_r.0.bias'), n)
            else:                                            # sum: the same dx through each branch's LeakyReLU
                dpre = self.ef[2].view(-1)
                for branch, src, cin_b, key in ((self.ef[0], mv, 2, 'conv_0_mv.0'), (self.ef[1], res, 3, 'conv_0_r.0')):
                    ops.act_bwd_planar(dD[(2 + oo) * HW:], dns, branch.view(-1), 8 * HW, None, 0.1, 8, HW, n, dpre,
                                       8 * HW)
                    ops.conv_wgrad(src, cin_b * HW, cin_b, H, W, dpre, 8 * HW, 8, 3, 1, self.g(q + key + '.weight'),
                                   self.g(q + key + '.bias'), n)

    def _dgrad_s1(self, dY, dy_ns, cout, wkey, cin, ci_count, dX, dx_ns, H, W, n, accumulate, act=None):
        """3x3 stride-1 data gradient = forward convolution of dY with the flipped, transposed weight;
        act = (tensor, stride, channels, slope) fuses LeakyReLU' on the first `channels` outputs."""
        wT = self.wflip[:ci_count * cout * 9]
        ops.weight_flip(self.p(wkey), cout, cin, ci_count, wT)
        if act is None:          # plain path (also covers widths that are not multiples of 4)
            ops.conv_fwd(dY, dy_ns, cout, H, W, wT, None, ci_count, 3, 1, dX, dx_ns, n, slope=1.0,
                         accumulate=accumulate)
            return
        a_src, a_ns, a_c1, a_slope = act
        ops.conv3x3_dgrad_fused(dY, dy_ns, cout, H, W, wT, ci_count, dX, dx_ns, n, accumulate=accumulate,
                                act_src=a_src, act_ns=a_ns, act_c1=a_c1, act_slope=a_slope)

    # ------------------------------------------------------------------ classifier
    def _prep_weights(self):
        """fp32 OIHW parameters -> bf16 hi/lo GEMM operands (fprop and transposed for dgrad),
        every classifier conv in one launch."""
        if getattr(self, '_prep_chunks', None) is None:
            rows = []
            for blk in self.blocks:
                for key in ('c1', 'c2', 'ds'):
                    if key not in blk:
                        continue
                    u = blk[key]
                    assert u.cout % 32 == 0 and u.cin % 32 == 0 and u.taps <= 9
                    for tile in range((u.cout // 32) * (u.cin // 32)):       # 32 x 32 (co, ci) tiles
                        rows.append([self.offsets[u.name_conv + '.weight'], u.W_hi.data_ptr(),
                                     u.W_lo.data_ptr(), u.Wt_hi.data_ptr(), u.Wt_lo.data_ptr(),
                                     u.cout | (u.cin << 32), u.taps | (tile << 32)])
            self._prep_chunks = torch.tensor(rows, dtype=torch.int64).to(self.device)
        ops.weight_prep_multi(self.params, self._prep_chunks, self._prep_chunks.shape[0])

    def _unit_bn(self, u: _ConvBN, train: bool):
        """Batch statistics (train) or running statistics (eval) -> scale/shift of unit u."""
        wk = u.name_bn
        if train:        # batch statistics were accumulated by the conv's GEMM epilogue
            ops.bn_finalize(u.sums, u.geo.count, self.p(wk + '.weight'), self.p(wk + '.bias'),
                            self.buffers[wk + '.running_mean'], self.buffers[wk + '.running_var'],
                            self.buffers[wk + '.num_batches_tracked'], BN_MOMENTUM, 1e-5, u.cout,
                            u.scale, u.shift, u.mean, u.invstd)
        else:
            ops.bn_eval_coeffs(self.p(wk + '.weight'), self.p(wk + '.bias'),
                               self.buffers[wk + '.running_mean'], self.buffers[wk + '.running_var'],
                               1e-5, u.cout, u.scale, u.shift)

    def _conv_fwd(self, u: _ConvBN, a_hi, a_lo, a_phases: int, train: bool = False):
        geo = u.geo
        if u.ks == 1:
            shift, phase, bsel = [0], [0], [0]
        elif u.stride == 1:
            shift, phase, bsel = _taps_s1(geo.Wp)
        else:
            shift, phase, bsel = _taps_s2(geo.Wp)
        ops.tap_gemm(a_hi, a_lo, u.W_hi, u.W_lo, u.Y, a_phases=a_phases, a_rows=geo.P, K=u.cin,
                     b_slices=u.taps, N=u.cout, M=geo.P, ldD=u.cout, Hp=geo.Hp, Wp=geo.Wp,
                     shift=shift, phase=phase, bsel=bsel, engine=self.gemm_engine,
                     stats=(u.sums if train else None))

    def _cls_forward(self, x_planar: torch.Tensor, n: int, train: bool):
        """ResNet-18 forward on a planar [n,2,H,W] input -> self.logits[:n]."""
        H, W = self.H, self.W
        if n != self.N:
            raise RuntimeError('engine was built for %d frames, got %d' % (self.N, n))
        folded = not train and self.fold_bn and self.gemm_engine == 'tc'
        if not folded:
            self._prep_weights()
        if train:
            ops.memset_zero(self._sums_pool[:self._sums_used])
        # stem: 7x7/2 conv (planar) -> BN -> ReLU -> maxpool -> pixel-major hi/lo
        H2, W2 = H // 2, W // 2
        if self.stem_tc:
            ops.stem_conv_tc_fwd(x_planar, 2 * H * W, H, W, self.p('base_model.conv1.weight'), self.stem_wb,
                                 self.stem_Y.view(-1), 64 * H2 * W2, n)
        else:
            ops.conv_fwd(x_planar, 2 * H * W, 2, H, W, self.p('base_model.conv1.weight'), None, 64, 7, 2,
                         self.stem_Y.view(-1), 64 * H2 * W2, n)
        st = self.stem
        if train:
            ops.bn_stats_planar(self.stem_Y, 64 * H2 * W2, 64, H2 * W2, n, st['sums'])
            ops.bn_finalize(st['sums'], float(n * H2 * W2), self.p('base_model.bn1.weight'),
                            self.p('base_model.bn1.bias'), self.buffers['base_model.bn1.running_mean'],
                            self.buffers['base_model.bn1.running_var'],
                            self.buffers['base_model.bn1.num_batches_tracked'], BN_MOMENTUM, 1e-5, 64,
                            st['scale'], st['shift'], st['mean'], st['invstd'])
        else:
            ops.bn_eval_coeffs(self.p('base_model.bn1.weight'), self.p('base_model.bn1.bias'),
                               self.buffers['base_model.bn1.running_mean'],
                               self.buffers['base_model.bn1.running_var'], 1e-5, 64, st['scale'],
                               st['shift'])
        ops.stem_pool_fwd(self.stem_Y, st['scale'], st['shift'], n, 64, H2, W2, self.A0_hi, self.A0_lo,
                          self.pool_idx)
        x_hi, x_lo = self.A0_hi, self.A0_lo
        if folded:
            x_hi, x_lo = self._cls_blocks_folded(x_hi, x_lo, n)
        for blk in ([] if folded else self.blocks):
            c1, c2, geo = blk['c1'], blk['c2'], blk['geo']
            if 'ds' in blk:
                gi = blk['geo_in']
                ops.phase_split(x_hi, x_lo, n, gi.H, gi.W, blk['cin'], blk['xp_hi'], blk['xp_lo'])
                self._conv_fwd(c1, blk['xp_hi'], blk['xp_lo'], 4, train)
            else:
                self._conv_fwd(c1, x_hi, x_lo, 1, train)
            self._unit_bn(c1, train)
            ops.bn_apply(c1.Y, c1.scale, c1.shift, geo.P, c1.cout, geo.Hp, geo.Wp, True, c1.act_hi,
                         c1.act_lo)
            self._conv_fwd(c2, c1.act_hi, c1.act_lo, 1, train)
            self._unit_bn(c2, train)
            if 'ds' in blk:
                ds = blk['ds']
                self._conv_fwd(ds, blk['xp_hi'], blk['xp_lo'], 4, train)
                self._unit_bn(ds, train)
                ops.bn_apply(c2.Y, c2.scale, c2.shift, geo.P, c2.cout, geo.Hp, geo.Wp, True, c2.act_hi,
                             c2.act_lo, resY=ds.Y, res_scale=ds.scale, res_shift=ds.shift)
            else:
                ops.bn_apply(c2.Y, c2.scale, c2.shift, geo.P, c2.cout, geo.Hp, geo.Wp, True, c2.act_hi,
                             c2.act_lo, res_hi=x_hi, res_lo=x_lo)
            x_hi, x_lo = c2.act_hi, c2.act_lo
        gl = self.geo_last
        ops.avgpool(x_hi, x_lo, n, gl.Hp, gl.Wp, 512, self.pooled)
        ops.linear_fwd(self.pooled, self.p('base_model.fc.weight'), self.p('base_model.fc.bias'), n,
                       512, self.num_class, self.logits)

    def _fold_unit(self, u: _ConvBN):
        """Eval-mode BatchNorm of unit u folded into its GEMM operand: W' = W * scale, bias = shift."""
        wk = u.name_bn
        ops.bn_eval_coeffs(self.p(wk + '.weight'), self.p(wk + '.bias'), self.buffers[wk + '.running_mean'],
                           self.buffers[wk + '.running_var'], 1e-5, u.cout, u.scale, u.shift)
        ops.weight_fold_prep(self.p(u.name_conv + '.weight'), u.scale, u.shift, u.cout, u.cin, u.taps, u.cout, u.cin,
                             u.W_hi, u.W_lo, u.fbias)

    def _conv_fold(self, u: _ConvBN, a_hi, a_lo, a_phases: int, slope: float, **out):
        geo = u.geo
        if u.ks == 1:
            shift, phase, bsel = [0], [0], [0]
        elif u.stride == 1:
            shift, phase, bsel = _taps_s1(geo.Wp)
        else:
            shift, phase, bsel = _taps_s2(geo.Wp)
        ops.tap_gemm_fold(a_hi, a_lo, u.W_hi, u.W_lo, a_phases=a_phases, a_rows=geo.P, K=u.cin, b_slices=u.taps,
                          N=u.cout, M=geo.P, Hp=geo.Hp, Wp=geo.Wp, shift=shift, phase=phase, bsel=bsel,
                          bias=u.fbias, slope=slope, **out)

    def _cls_blocks_folded(self, x_hi, x_lo, n: int):
        """The eight BasicBlocks in eval mode with BatchNorm folded: two (three) GEMMs per block, each
        writing the hi/lo activation the next one reads; relu(bn2(conv2) + identity) in conv2's epilogue."""
        for blk in self.blocks:
            c1, c2 = blk['c1'], blk['c2']
            self._fold_unit(c1)
            self._fold_unit(c2)
            if 'ds' in blk:
                ds, gi = blk['ds'], blk['geo_in']
                self._fold_unit(ds)
                ops.phase_split(x_hi, x_lo, n, gi.H, gi.W, blk['cin'], blk['xp_hi'], blk['xp_lo'])
                self._conv_fold(c1, blk['xp_hi'], blk['xp_lo'], 4, 0.0, out_hi=c1.act_hi, out_lo=c1.act_lo)
                self._conv_fold(ds, blk['xp_hi'], blk['xp_lo'], 4, 1.0, D=ds.Y)
                self._conv_fold(c2, c1.act_hi, c1.act_lo, 1, 0.0, res_f32=ds.Y, out_hi=c2.act_hi, out_lo=c2.act_lo)
            else:
                self._conv_fold(c1, x_hi, x_lo, 1, 0.0, out_hi=c1.act_hi, out_lo=c1.act_lo)
                self._conv_fold(c2, c1.act_hi, c1.act_lo, 1, 0.0, res_hi=x_hi, res_lo=x_lo, out_hi=c2.act_hi,
                                out_lo=c2.act_lo)
            x_hi, x_lo = c2.act_hi, c2.act_lo
        return x_hi, x_lo

    def _unit_bn_bwd(self, u: _ConvBN, g_a, g_b, act_hi, G_hi, G_lo, dz_out):
        geo, wk = u.geo, u.name_bn
        ops.bn_bwd_reduce(g_a, g_b, act_hi, u.Y, u.mean, u.invstd, geo.P, u.cout, geo.Hp, geo.Wp,
                          u.sums2)
        ops.bn_bwd_apply(g_a, g_b, act_hi, u.Y, u.mean, u.invstd, self.p(wk + '.weight'), u.sums2,
                         geo.count, geo.P, u.cout, geo.Hp, geo.Wp, G_hi, G_lo, dz_out,
                         self.g(wk + '.weight'), self.g(wk + '.bias'))

    def _bn_bwd_apply_only(self, u: _ConvBN, dz, G_hi, G_lo):
        """BN backward when dz (masked gradient) and u.sums2 were produced by a fused GEMM epilogue."""
        geo, wk = u.geo, u.name_bn
        ops.bn_bwd_apply(dz, None, None, u.Y, u.mean, u.invstd, self.p(wk + '.weight'), u.sums2, geo.count,
                         geo.P, u.cout, geo.Hp, geo.Wp, G_hi, G_lo, None, self.g(wk + '.weight'),
                         self.g(wk + '.bias'))

    def _conv_wgrad(self, u: _ConvBN, G_hi, G_lo, x_hi, x_lo, x_phases: int):
        geo = u.geo
        if u.ks == 1:
            shift, phase, bsel = [0], [0], [0]
        elif u.stride == 1:
            shift, phase, bsel = _taps_s1(geo.Wp)
        else:
            shift, phase, bsel = _taps_s2(geo.Wp)
        # the split-K epilogue accumulates straight into the (zeroed) OIHW gradient bucket
        ops.wgrad_gemm(G_hi, G_lo, x_hi, x_lo, self.g(u.name_conv + '.weight'), P=geo.P, Cout=u.cout,
                       x_phases=x_phases, Cin=u.cin, shift=shift, phase=phase, bsel=bsel,
                       engine=self.gemm_engine, oihw_taps=u.taps, workspace=self.wgrad_ws)

    def _cls_backward(self, x_planar: torch.Tensor, n: int, need_wgrad: bool, need_input_grad: bool,
                      d_input: Optional[torch.Tensor] = None):
        """Backward of _cls_forward from self.d_logits.  need_wgrad=False skips every
        weight gradient (GAN generator step: those grads are zeroed unused,
        code/dmcnet_GAN/train.py:367-371)."""
        C = self.num_class
        gl = self.geo_last
        ops.linear_bwd(self.d_logits, self.pooled, self.p('base_model.fc.weight'), n, 512, C,
                       self.d_pooled, self.g('base_model.fc.weight') if need_wgrad else None,
                       self.g('base_model.fc.bias') if need_wgrad else None)
        T1, T2a, Ra, T2b, Rb = self.gbuf
        fuse = self.gemm_engine == 'tc'         # fused BN-backward epilogue exists in the tcgen05 kernel
        ops.memset_zero(self._sums2_pool[:self._sums2_used])
        g_a = T2a[:gl.P * 512]
        ops.avgpool_bwd(self.d_pooled, n, gl.Hp, gl.Wp, 512, g_a)
        g_b = None
        g_is_dz = False    # True: g_a already holds dz = relu'(x_out) * (sum of gradients) of THIS block's
                           # output and c2.sums2 was accumulated by the GEMM epilogue that produced it
        flip = False       # which (T2, R) pair the CURRENT block writes
        for bi in reversed(range(len(self.blocks))):
            blk = self.blocks[bi]
            c1, c2, geo, width, cin = blk['c1'], blk['c2'], blk['geo'], blk['width'], blk['cin']
            T2, R = (T2b, Rb) if not flip else (T2a, Ra)
            flip = not flip
            x_hi, x_lo = (self.blocks[bi - 1]['c2'].act_hi, self.blocks[bi - 1]['c2'].act_lo) \
                if bi > 0 else (self.A0_hi, self.A0_lo)
            pc = geo.P * width
            G_hi, G_lo = self.G_hi[:pc], (None if self.grad_bf16 else self.G_lo[:pc])
            has_ds = 'ds' in blk
            # ---- bn2 backward (dz also feeds the identity / downsample branch)
            if g_is_dz:
                dz = g_a                                       # masked + summed already, sums2 ready
                self._bn_bwd_apply_only(c2, dz, G_hi, G_lo)
            else:
                dz = None if has_ds else R[:pc]
                self._unit_bn_bwd(c2, g_a, g_b, c2.act_hi, G_hi, G_lo, dz)
            if need_wgrad:
                self._conv_wgrad(c2, G_hi, G_lo, c1.act_hi, c1.act_lo, 1)
            # ---- conv2 dgrad -> dz1 (bn1's ReLU mask and both reductions fused into the epilogue)
            shift, phase, bsel = _taps_s1(geo.Wp)
            bw1 = (c1.Y, c1.act_hi, None, c1.mean, c1.invstd) if fuse else None
            ops.tap_gemm(G_hi, G_lo, c2.Wt_hi, c2.Wt_lo, T1[:pc], a_phases=1, a_rows=geo.P, K=width,
                         b_slices=9, N=width, M=geo.P, ldD=width, Hp=geo.Hp, Wp=geo.Wp,
                         shift=[-s for s in shift], phase=phase, bsel=bsel, engine=self.gemm_engine,
                         stats=(c1.sums2 if fuse else None), bw=bw1)
            if has_ds:
                ds = blk['ds']
                Gp_hi = self.G_hi[:2 * pc].view(2, pc)
                Gp_lo2 = None if self.grad_bf16 else self.G_lo[:2 * pc].view(2, pc)
                Gp_lo = (None, None) if self.grad_bf16 else Gp_lo2
                # downsample BN backward (same dz as bn2) -> slot 1 ; bn1 backward -> slot 0
                if g_is_dz:
                    self._unit_bn_bwd(ds, g_a, None, None, Gp_hi[1], Gp_lo[1], None)
                else:
                    self._unit_bn_bwd(ds, g_a, g_b, c2.act_hi, Gp_hi[1], Gp_lo[1], None)
                if fuse:
                    self._bn_bwd_apply_only(c1, T1[:pc], Gp_hi[0], Gp_lo[0])
                else:
                    self._unit_bn_bwd(c1, T1[:pc], None, c1.act_hi, Gp_hi[0], Gp_lo[0], None)
                if need_wgrad:
                    self._conv_wgrad(c1, Gp_hi[0], Gp_lo[0], blk['xp_hi'], blk['xp_lo'], 4)
                    self._conv_wgrad(ds, Gp_hi[1], Gp_lo[1], blk['xp_hi'], blk['xp_lo'], 4)
                # dgrad into the four input phases; the 1x1 downsample joins phase (0,0)
                fs, fp, fb = _taps_s2(geo.Wp)
                for ph in range(4):
                    sh = [-fs[t] for t in range(9) if fp[t] == ph]
                    bs = [fb[t] for t in range(9) if fp[t] == ph]
                    aph = [0] * len(sh)
                    if ph == 0:
                        sh, bs, aph = sh + [0], bs + [9], aph + [1]
                    ops.tap_gemm(Gp_hi, Gp_lo2, c1.Wt_hi, c1.Wt_lo, blk['dxp'][ph], a_phases=2,
                                 a_rows=geo.P, K=width, b_slices=10, N=cin, M=geo.P, ldD=cin,
                                 Hp=geo.Hp, Wp=geo.Wp, shift=sh, phase=aph, bsel=bs,
                                 engine=self.gemm_engine)
                gi = blk['geo_in']
                g_a = T2[:gi.P * cin]
                if fuse:
                    # one pass: interleave the four phases, apply the ReLU mask of the previous block's
                    # output and reduce for its bn2 -> g_a is that block's dz
                    pb = self.blocks[bi - 1]['c2']
                    ops.phase_unsplit_reduce(blk['dxp'], n, gi.H, gi.W, cin, pb.act_hi, pb.Y, pb.mean, pb.invstd,
                                             g_a, pb.sums2)
                    g_b, g_is_dz = None, True
                else:
                    ops.phase_unsplit(blk['dxp'], n, gi.H, gi.W, cin, g_a)
                    g_b, g_is_dz = None, False
            else:
                if fuse:
                    self._bn_bwd_apply_only(c1, T1[:pc], G_hi, G_lo)
                else:
                    self._unit_bn_bwd(c1, T1[:pc], None, c1.act_hi, G_hi, G_lo, None)
                if need_wgrad:
                    self._conv_wgrad(c1, G_hi, G_lo, x_hi, x_lo, 1)
                res = dz if dz is not None else None       # gradient of the identity branch
                g_a = T2[:pc]
                if fuse and bi > 0:
                    # conv1 dgrad: the epilogue adds the identity-branch gradient, applies the ReLU mask of
                    # the PREVIOUS block's output and reduces for its bn2 -> g_a is that block's dz
                    pb = self.blocks[bi - 1]['c2']
                    ops.tap_gemm(G_hi, G_lo, c1.Wt_hi, c1.Wt_lo, g_a, a_phases=1, a_rows=geo.P, K=width,
                                 b_slices=9, N=cin, M=geo.P, ldD=cin, Hp=geo.Hp, Wp=geo.Wp,
                                 shift=[-s for s in shift], phase=phase, bsel=bsel, engine=self.gemm_engine,
                                 stats=pb.sums2, bw=(pb.Y, pb.act_hi, res, pb.mean, pb.invstd))
                    g_b, g_is_dz = None, True
                else:
                    ops.tap_gemm(G_hi, G_lo, c1.Wt_hi, c1.Wt_lo, g_a, a_phases=1, a_rows=geo.P, K=width,
                                 b_slices=9, N=cin, M=geo.P, ldD=cin, Hp=geo.Hp, Wp=geo.Wp,
                                 shift=[-s for s in shift], phase=phase, bsel=bsel, engine=self.gemm_engine)
                    g_b, g_is_dz = res, False
        # stem backward
        H, W = self.H, self.W
        H2, W2 = H // 2, W // 2
        st = self.stem
        ops.stem_pool_bwd(g_a, g_b, self.pool_idx, self.stem_Y, st['scale'], st['shift'], n, 64, H2, W2,
                          self.stem_dZ)
        ns = 64 * H2 * W2
        ops.bn_bwd_reduce_planar(self.stem_dZ, ns, self.stem_Y, ns, st['mean'], st['invstd'], 64,
                                 H2 * W2, n, st['sums2'])
        ops.bn_bwd_apply_planar(self.stem_dZ, ns, self.stem_Y, ns, st['mean'], st['invstd'],
                                self.p('base_model.bn1.weight'), st['sums2'], float(n * H2 * W2), 64,
                                H2 * W2, n, self.stem_dZ, ns, self.g('base_model.bn1.weight'),
                                self.g('base_model.bn1.bias'))
        if need_wgrad:
            if self.stem_tc:
                ops.stem_conv_tc_wgrad(x_planar, 2 * H * W, H, W, self.stem_dZ.view(-1), ns,
                                       self.g('base_model.conv1.weight'), self.stem_ws, n)
            else:
                ops.conv_wgrad(x_planar, 2 * H * W, 2, H, W, self.stem_dZ, ns, 64, 7, 2,
                               self.g('base_model.conv1.weight'), None, n)
        if need_input_grad and self.stem_dgrad_tc:
            sd = self.stem_dg
            rows = n * (H2 + 2) * (W2 + 2)
            ops.planar_to_pm_ring2(self.stem_dZ.view(-1), ns, 64, H2, W2, n, sd['hi'], sd['lo'])
            ops.weight_gather_prep(self.p('base_model.conv1.weight'), sd['gmap'], 16, 32, 64, sd['W_hi'],
                                   sd['W_lo'])
            ops.tap_gemm(sd['hi'], sd['lo'], sd['W_hi'], sd['W_lo'], sd['out'], a_phases=1, a_rows=rows, K=64,
                         b_slices=16, N=32, M=rows, ldD=32, Hp=H2 + 2, Wp=W2 + 2, shift=sd['shift'],
                         phase=[0] * 16, bsel=list(range(16)), engine='tc')
            ops.s2d2_ring2_to_planar(sd['out'], 32, H, W, n, self.dD.view(-1), self.dD.shape[1] * H * W,
                                     accumulate=True)
        elif need_input_grad:
            ops.conv_dgrad(self.stem_dZ, ns, 64, self.p('base_model.conv1.weight'), 2, 2, 7, 2,
                           self.dD.view(-1), self.dD.shape[1] * H * W, H, W, n, accumulate=True)

    # ------------------------------------------------------------------ discriminator
    def _disc_forward(self, x: torch.Tensor, m: int, train: bool, use_masks: bool):
        """Discriminator*.forward on planar x [m,2,H,W] -> self.validity[:m]."""
        inp, in_c, h, w = x, 2, self.H, self.W
        for L in self.d_layers:
            p = L['name']
            mk = L['mask'] if (train and use_masks) else None
            if L.get('s2d', False):
                # stride-2 layer as a 2x2-tap stride-1 conv on the space-to-depth input
                s_ns, c4 = in_c * h * w, 4 * in_c
                W3 = self.d_w3[:L['cout'] * c4 * 9]
                ops.s2d_planar(inp.view(-1), s_ns, in_c, h, w, self.d_s2d, s_ns, m)
                ops.s2_weight_map(self.p(p + '.0.weight'), W3, L['cout'], in_c, True)
                ops.conv3x3_taps2(self.d_s2d, s_ns, c4, L['Ho'], L['Wo'], W3, self.p(p + '.0.bias'),
                                  L['cout'], 0, L['A'].view(-1), L['cout'] * L['Ho'] * L['Wo'], m,
                                  slope=0.2, mask=mk)
            else:
                ops.conv_fwd(inp.view(-1), in_c * h * w, in_c, h, w, self.p(p + '.0.weight'),
                             self.p(p + '.0.bias'), L['cout'], 3, L['stride'], L['A'].view(-1),
                             L['cout'] * L['Ho'] * L['Wo'], m, slope=0.2, mask=mk)
            hw = L['Ho'] * L['Wo']
            if L['bn']:
                ns = L['cout'] * hw
                if train:
                    ops.bn_stats_planar(L['A'], ns, L['cout'], hw, m, L['sums'])
                    ops.bn_finalize(L['sums'], float(m * hw), self.p(p + '.3.weight'),
                                    self.p(p + '.3.bias'), self.buffers[p + '.3.running_mean'],
                                    self.buffers[p + '.3.running_var'],
                                    self.buffers[p + '.3.num_batches_tracked'], BN_MOMENTUM, 0.8,
                                    L['cout'], L['scale'], L['shift'], L['mean'], L['invstd'])
                else:
                    ops.bn_eval_coeffs(self.p(p + '.3.weight'), self.p(p + '.3.bias'),
                                       self.buffers[p + '.3.running_mean'],
                                       self.buffers[p + '.3.running_var'], 0.8, L['cout'], L['scale'],
                                       L['shift'])
                ops.bn_apply_planar(L['A'], ns, L['scale'], L['shift'], L['cout'], hw, m, False,
                                    L['Z'], ns)
            inp, in_c, h, w = L['Z'], L['cout'], L['Ho'], L['Wo']
        K = self.specs['discriminator.adv_layer.weight'][1]
        ops.linear_fwd(inp, self.p('discriminator.adv_layer.weight'),
                       self.p('discriminator.adv_layer.bias'), m, K, 2, self.validity)

    def _disc_backward(self, x: torch.Tensor, m: int, need_wgrad: bool, use_masks: bool,
                       d_input: Optional[torch.Tensor], d_input_rows: int):
        """Backward from self.d_validity[:m]; optionally accumulates d(loss)/d(x[:rows]) into d_input."""
        K = self.specs['discriminator.adv_layer.weight'][1]
        last = self.d_layers[-1]
        ops.linear_bwd(self.d_validity, last['Z'], self.p('discriminator.adv_layer.weight'), m, K, 2,
                       self.d_feat, self.g('discriminator.adv_layer.weight') if need_wgrad else None,
                       self.g('discriminator.adv_layer.bias') if need_wgrad else None)
        g = self.d_feat.view(-1)
        for li in reversed(range(len(self.d_layers))):
            L = self.d_layers[li]
            p, co = L['name'], L['cout']
            hw = L['Ho'] * L['Wo']
            ns = co * hw
            buf = self.d_g[li % 2]
            if L['bn']:
                ops.bn_bwd_reduce_planar(g, ns, L['A'], ns, L['mean'], L['invstd'], co, hw, m, L['sums2'])
                ops.bn_bwd_apply_planar(g, ns, L['A'], ns, L['mean'], L['invstd'], self.p(p + '.3.weight'),
                                        L['sums2'], float(m * hw), co, hw, m, buf, ns,
                                        self.g(p + '.3.weight'), self.g(p + '.3.bias'))
                g = buf
            # LeakyReLU(0.2) + Dropout2d backward -> gradient of the conv pre-activation
            ops.act_bwd_planar(g, ns, L['A'], ns, L['mask'] if use_masks else None, 0.2, co, hw, m,
                               buf, ns)
            g = buf
            if li > 0:
                prev = self.d_layers[li - 1]
                inp, ci, h, w = prev['Z'], prev['cout'], prev['Ho'], prev['Wo']
            else:
                inp, ci, h, w = x, 2, self.H, self.W
            s2d = L.get('s2d', False)
            c4, s_ns = 4 * ci, ci * h * w                     # S: [m][4ci][Ho][Wo], same size as the input
            if need_wgrad:
                if s2d:
                    S, dW3 = self.d_s2d, self.d_w3[:co * c4 * 9]
                    ops.s2d_planar(inp.view(-1), s_ns, ci, h, w, S, s_ns, m)
                    ops.memset_zero(dW3)
                    ops.conv_wgrad(S, s_ns, c4, L['Ho'], L['Wo'], g, ns, co, 3, 1, dW3,
                                   self.g(p + '.0.bias'), m)
                    ops.s2_weight_map(dW3, self.g(p + '.0.weight'), co, ci, False)
                else:
                    ops.conv_wgrad(inp.view(-1), ci * h * w, ci, h, w, g, ns, co, 3, L['stride'],
                                   self.g(p + '.0.weight'), self.g(p + '.0.bias'), m)
            want_dx = li > 0 or d_input is not None
            rows = m if li > 0 else d_input_rows
            if want_dx and s2d:
                # dS = conv3x3/1(dPre, flip(W3)); dX = depth-to-space(dS)
                W3, W3t = self.d_w3[:co * c4 * 9], self.d_w3t[:co * c4 * 9]
                ops.s2_weight_map(self.p(p + '.0.weight'), W3, co, ci, True)
                ops.weight_flip(W3, co, c4, c4, W3t)
                ops.conv3x3_taps2(g, ns, co, L['Ho'], L['Wo'], W3t, None, c4, 1, self.d_ds, s_ns, rows)
                if li > 0:
                    nxt = self.d_g[(li - 1) % 2]
                    ops.d2s_planar(self.d_ds, s_ns, ci, h, w, nxt, ci * h * w, rows)
                    g = nxt
                else:
                    ops.d2s_planar(self.d_ds, s_ns, ci, h, w, self.dD.view(-1), self.dD.shape[1] * h * w,
                                   rows, accumulate=True)
            elif li > 0:
                nxt = self.d_g[(li - 1) % 2]
                if L['stride'] == 1:
                    self._dgrad_s1(g, ns, co, p + '.0.weight', ci, ci, nxt, ci * h * w, h, w, m, False)
                else:
                    ops.conv_dgrad(g, ns, co, self.p(p + '.0.weight'), ci, ci, 3, L['stride'], nxt,
                                   ci * h * w, h, w, m, accumulate=False)
                g = nxt
            elif d_input is not None:
                ops.conv_dgrad(g, ns, co, self.p(p + '.0.weight'), ci, ci, 3, L['stride'],
                               self.dD.view(-1), self.dD.shape[1] * h * w, h, w, d_input_rows,
                               accumulate=True)



    # ------------------------------------------------------------------ ContextNetwork generator
    def _context_layout(self):
        """[(parameter prefix, cin, cout, dilation)]: the trunk (code/dmcnet/model.py:45-71; the fifth
        dilation is 1 instead of 16 when gen_flow_ds_factor != 0, :58-66 / :85-93) and, with --att 1, the
        two heads of ContextNetworkAtt (:94-98) instead of the trunk's last layer."""
        trunk = list(CONTEXT_LAYERS)
        if self.gen_ds:
            trunk[4] = (64, 1)
        if self.att:
            trunk = trunk[:6]
        out, cin = [], GEN_IN
        for i, (co, d) in enumerate(trunk):
            out.append(('gen_flow_model.conv_context.%d' % i, cin, co, d))
            cin = co
        if self.att:
            out.append(('gen_flow_model.predict_flow', 32, 2, 1))
            out.append(('gen_flow_model.predict_att.0', 32, 2, 1))
        return out

    def _alloc_context(self):
        """ContextNetwork[Att] on the tensor-core path: pixel-major bf16 hi/lo maps with a shared zero ring
        as wide as the largest dilation (a dilated tap is a flat row shift d*(dr*Wp + ds)), channels
        zero-padded to 64 / 128 GEMM columns, weights through the gather tables of
        disc_plan.layer_plan('P1'), BatchNorm statistics / backward reductions in the GEMM epilogues."""
        from . import disc_plan as DP
        dev, N, H, W = self.device, self.N, self.H, self.W
        f32 = dict(dtype=torch.float32, device=dev)
        bf = dict(dtype=torch.bfloat16, device=dev)
        i32 = lambda a: torch.from_numpy(a.astype('int32')).contiguous().to(dev)
        f = self.gen_ds or 1
        self.gH, self.gW = H // f, W // f                       # resolution the estimator runs at
        gH, gW = self.gH, self.gW
        layout = self._context_layout()
        R = self.ctx_ring = max(d for _, _, _, d in layout)
        self.gen_ctot = GEN_IN
        self.ctx_geo = (gH + R, gW + R)
        Hp, Wp = self.ctx_geo
        P = N * Hp * Wp
        if P >= (1 << 31) // 128:
            raise ValueError('ContextNetwork plan: %d frames of %dx%d exceed the 32-bit tile index range' % (N, H, W))
        self.ctx_P = P
        self.dD = torch.zeros(N, 2, H, W, **f32)                  # gradient w.r.t. gen_flow (planar, frame size)
        self.d_gen_flow = self.dD
        self.gen_flow = torch.zeros(N, 2, H, W, **f32)
        self.ctx_in = torch.zeros(N, GEN_IN, gH, gW, **f32)        # cat(mv, residual) (pooled), planar staging
        self.ctx_in_hi, self.ctx_in_lo = torch.zeros(P, 64, **bf), torch.zeros(P, 64, **bf)
        if self.gen_ds:
            self.gen_small = torch.zeros(N, 2, gH, gW, **f32)
            self.gD = torch.zeros(N, 2, gH, gW, **f32)             # d loss / d (estimator output)
        else:
            self.gD = self.dD
        if self.att:
            self.att_flow = torch.zeros(N, 2, gH, gW, **f32)       # attention map (estimator resolution, :341-357)
            self.d_att = torch.zeros(N, 2, gH, gW, **f32)
        ncol = sum(DP.pad64(co) for _, _, co, _ in layout)
        self._csums = torch.zeros(2 * ncol, dtype=torch.float64, device=dev)
        self._csums2 = torch.zeros(2 * ncol, dtype=torch.float64, device=dev)
        self.ctx_layers = []
        col, wg_total, max_np, max_ws = 0, 0, 64, 0
        for name, cin, co, d in layout:
            lp = DP.layer_plan('P1', cin, co)
            T, Np, Kp = lp['gmap'].shape
            L = {'name': name, 'cin': cin, 'cout': co, 'dil': d, 'Np': Np, 'Kp': Kp,
                 'gmap': i32(lp['gmap'].reshape(-1)), 'inv': i32(lp['inv'].reshape(-1)), 'cmap': i32(lp['cmap']),
                 'R': lp['inv'].shape[1],
                 'shift': [(r - 1) * d * Wp + (s2 - 1) * d for r in range(3) for s2 in range(3)],
                 'W_hi': torch.zeros(9, Np, Kp, **bf), 'W_lo': torch.zeros(9, Np, Kp, **bf),
                 'Wt_hi': torch.zeros(9, Kp, Np, **bf), 'Wt_lo': torch.zeros(9, Kp, Np, **bf),
                 'Y': torch.zeros(P, Np, **f32),
                 'act_hi': torch.zeros(P, Np, **bf), 'act_lo': torch.zeros(P, Np, **bf),
                 'sums': self._csums[2 * col:2 * col + 2 * Np].view(2, Np),
                 'sums2': self._csums2[2 * col:2 * col + 2 * Np].view(2, Np),
                 'coef': torch.zeros(3, Np, **f32), 'wg_off': wg_total, 'slope': 0.1}
            for k in ('scale', 'shift_', 'mean', 'invstd'):
                L[k] = torch.zeros(Np, **f32)
            col += Np
            wg_total += 9 * Np * Kp
            max_np = max(max_np, Np, Kp)
            max_ws = max(max_ws, ops.wgrad_workspace_floats(P, Np, Kp, 9))
            self.ctx_layers.append(L)
        if self.att:
            self.ctx_layers[-1]['slope'] = 0.0                     # ReLU(LeakyReLU(x)) = ReLU(x): the attention head
        self.ctx_trunk = self.ctx_layers[:-2] if self.att else self.ctx_layers
        self._cwg = torch.zeros(wg_total, **f32)                   # GEMM-space weight gradients
        for L in self.ctx_layers:
            L['dWg'] = self._cwg[L['wg_off']:L['wg_off'] + 9 * L['Np'] * L['Kp']]
        self.ctx_dz = [torch.zeros(P * max_np, **f32) for _ in range(3 if self.att else 2)]
        self.ctx_G_hi, self.ctx_G_lo = torch.zeros(P * max_np, **bf), torch.zeros(P * max_np, **bf)
        self.ctx_ws = torch.empty(max_ws, **f32)

    def _ctx_layer_fwd(self, L, a_hi, a_lo, rows: int, count: float, train: bool):
        """dilated Conv3x3(bias=False) -> BatchNorm2d(eps 1e-5) -> LeakyReLU (code/dmcnet/model.py:31-42)."""
        p, Np, Kp, R = L['name'], L['Np'], L['Kp'], self.ctx_ring
        Hp, Wp = self.ctx_geo
        if not train and self.fold_bn:
            # eval: BatchNorm folded into the operand, activation in the epilogue, hi/lo written directly
            ops.pm_bn_finalize(None, L['cmap'], Np, L['cout'], count, self.p(p + '.1.weight'), self.p(p + '.1.bias'),
                               self.buffers[p + '.1.running_mean'], self.buffers[p + '.1.running_var'], None,
                               BN_MOMENTUM, 1e-5, L['scale'], L['shift_'], L['mean'], L['invstd'])
            ops.weight_fold_prep(self.p(p + '.0.weight'), L['scale'], L['shift_'], L['cout'], L['cin'], 9, Np, Kp,
                                 L['W_hi'], L['W_lo'], L['coef'].view(-1)[:Np])
            ops.tap_gemm_fold(a_hi, a_lo, L['W_hi'], L['W_lo'], a_phases=1, a_rows=rows, K=Kp, b_slices=9, N=Np,
                              M=rows, Hp=Hp, Wp=Wp, ring=R, shift=L['shift'], phase=[0] * 9, bsel=list(range(9)),
                              bias=L['coef'].view(-1)[:Np], slope=L['slope'], out_hi=L['act_hi'], out_lo=L['act_lo'])
            return
        ops.weight_gather_prep(self.p(p + '.0.weight'), L['gmap'], 9, Np, Kp, L['W_hi'], L['W_lo'], L['Wt_hi'],
                               L['Wt_lo'])
        ops.tap_gemm_ring(a_hi, a_lo, L['W_hi'], L['W_lo'], L['Y'], a_rows=rows, K=Kp, b_slices=9, N=Np,
                          M=rows, ldD=Np, Hp=Hp, Wp=Wp, ring=R, shift=L['shift'],
                          stats=(L['sums'] if train else None))
        ops.pm_bn_finalize(L['sums'] if train else None, L['cmap'], Np, L['cout'], count,
                           self.p(p + '.1.weight'), self.p(p + '.1.bias'), self.buffers[p + '.1.running_mean'],
                           self.buffers[p + '.1.running_var'],
                           self.buffers[p + '.1.num_batches_tracked'] if train else None, BN_MOMENTUM, 1e-5,
                           L['scale'], L['shift_'], L['mean'], L['invstd'])
        ops.bn_apply_lrelu(L['Y'], L['scale'], L['shift_'], rows, Np, Hp, Wp, R, L['slope'], L['act_hi'], L['act_lo'])

    def _ctx_forward(self, mv: torch.Tensor, res: torch.Tensor, n: int, train: bool):
        """ContextNetwork[Att].forward (+ input_mv when gen_flow_or_delta == 1), at 1/f resolution between an
        average pool and a tiling when gen_flow_ds_factor = f (code/dmcnet/model.py:330-357)."""
        H, W, R = self.gH, self.gW, self.ctx_ring
        Hp, Wp = self.ctx_geo
        HW = H * W
        rows, count = n * Hp * Wp, float(n * HW)
        cin = self.ctx_in.view(-1)
        if self.gen_ds:
            ops.avgpool_planar(mv, n * 2, self.H, self.W, self.gen_ds, self.ctx_dz[0][:n * 2 * HW])
            ops.avgpool_planar(res, n * 3, self.H, self.W, self.gen_ds, self.ctx_dz[1][:n * 3 * HW])
            mv, res = self.ctx_dz[0][:n * 2 * HW], self.ctx_dz[1][:n * 3 * HW]
        ops.copy_planar(mv, 2 * HW, cin, GEN_IN * HW, 2 * HW, n)
        ops.copy_planar(res, 3 * HW, cin[2 * HW:], GEN_IN * HW, 3 * HW, n)
        ops.planar_to_pm_ring(cin, GEN_IN * HW, GEN_IN, 64, H, W, R, n, self.ctx_in_hi, self.ctx_in_lo)
        if train:
            ops.memset_zero(self._csums)
        a_hi, a_lo = self.ctx_in_hi, self.ctx_in_lo
        for L in self.ctx_trunk:
            self._ctx_layer_fwd(L, a_hi, a_lo, rows, count, train)
            a_hi, a_lo = L['act_hi'], L['act_lo']
        if self.att:
            head_f, head_a = self.ctx_layers[-2], self.ctx_layers[-1]
            self._ctx_layer_fwd(head_f, a_hi, a_lo, rows, count, train)
            self._ctx_layer_fwd(head_a, a_hi, a_lo, rows, count, train)
            ops.pm_ring_to_planar(head_a['act_hi'], head_a['act_lo'], head_a['Np'], 2, H, W, R, n, None, 0,
                                  self.att_flow.view(-1), 2 * HW)
            out = head_f
        else:
            out = self.ctx_trunk[-1]
        # "+ input_mv" (model.py:344-345) uses the POOLED mv when the estimator runs at reduced resolution
        add = cin if self.gen_flow_or_delta == 1 else None          # channels 0:2 of the staged input = mv
        dst = self.gen_small if self.gen_ds else self.gen_flow
        ops.pm_ring_to_planar(out['act_hi'], out['act_lo'], out['Np'], 2, H, W, R, n, add, GEN_IN * HW,
                              dst.view(-1), 2 * HW)
        if self.gen_ds:
            ops.tile_repeat(self.gen_small.view(-1), n * 2, H, W, self.gen_ds, self.gen_flow.view(-1))

    def _ctx_layer_bwd(self, L, g, x_hi, x_lo, rows: int, count: float):
        """BatchNorm backward of one block from dz (already through its activation) and its two
        reductions, then the weight gradient; returns the hi/lo gradient of the conv output."""
        p, Np, Kp, R = L['name'], L['Np'], L['Kp'], self.ctx_ring
        Hp, Wp = self.ctx_geo
        G_hi, G_lo = self.ctx_G_hi[:rows * Np], self.ctx_G_lo[:rows * Np]
        ops.pm_bn_bwd_fold(L['sums2'], L['cmap'], Np, L['cout'], count, self.p(p + '.1.weight'), L['invstd'],
                           L['coef'], self.g(p + '.1.weight'), self.g(p + '.1.bias'))
        ops.pm_bn_bwd_apply(g, L['Y'], L['mean'], L['invstd'], L['coef'], rows, Np, Hp, Wp, R, G_hi, G_lo)
        ops.wgrad_gemm(G_hi, G_lo, x_hi, x_lo, L['dWg'], P=rows, Cout=Np, x_phases=1, Cin=Kp, shift=L['shift'],
                       phase=[0] * 9, bsel=list(range(9)), oihw_taps=0, workspace=self.ctx_ws)
        ops.weight_grad_gather(L['dWg'], L['inv'], L['cout'] * L['cin'] * 9, L['R'], self.g(p + '.0.weight'))
        return G_hi, G_lo

    def _ctx_head_dz(self, L, d_planar, g, n: int, rows: int):
        """planar loss gradient -> pixel-major dz of block L (through its activation) + its BN reductions."""
        H, W, R = self.gH, self.gW, self.ctx_ring
        Hp, Wp = self.ctx_geo
        ops.planar_to_pm_ring(d_planar.view(-1), 2 * H * W, 2, L['Np'], H, W, R, n, None, None, out_f32=g,
                              act_hi=L['act_hi'], slope=L['slope'])
        ops.bn_bwd_reduce(g, None, None, L['Y'], L['mean'], L['invstd'], rows, L['Np'], Hp, Wp, L['sums2'])

    def _ctx_backward(self, n: int):
        """Gradients of every ContextNetwork[Att] parameter from self.dD (= d loss / d gen_flow) and, with
        --att 1, self.d_att (= d loss / d att_flow)."""
        H, W, R = self.gH, self.gW, self.ctx_ring
        Hp, Wp = self.ctx_geo
        rows, count = n * Hp * Wp, float(n * H * W)
        ops.memset_zero(self._csums2)
        ops.memset_zero(self._cwg)
        if self.gen_ds:                              # backward of the tiling: sum of the f*f tiles
            ops.tile_sum(self.dD.view(-1), 2 * self.H * self.W, 2, H, W, self.gen_ds, n, self.gD.view(-1), 2 * H * W)
        trunk = self.ctx_trunk
        g = self.ctx_dz[0]
        if self.att:
            head_f, head_a, top = self.ctx_layers[-2], self.ctx_layers[-1], trunk[-1]
            # flow head: its data gradient (raw) goes to a side buffer ...
            self._ctx_head_dz(head_f, self.gD, g, n, rows)
            G_hi, G_lo = self._ctx_layer_bwd(head_f, g, top['act_hi'], top['act_lo'], rows, count)
            side = self.ctx_dz[2]
            ops.tap_gemm_ring(G_hi, G_lo, head_f['Wt_hi'], head_f['Wt_lo'], side, a_rows=rows, K=head_f['Np'],
                              b_slices=9, N=head_f['Kp'], M=rows, ldD=head_f['Kp'], Hp=Hp, Wp=Wp, ring=R,
                              shift=[-s2 for s2 in head_f['shift']])
            # ... and joins the attention head's in the fused epilogue (gb), which also applies the trunk's
            # LeakyReLU' and reduces for its BatchNorm
            self._ctx_head_dz(head_a, self.d_att, g, n, rows)
            G_hi, G_lo = self._ctx_layer_bwd(head_a, g, top['act_hi'], top['act_lo'], rows, count)
            nxt = self.ctx_dz[1]
            ops.tap_gemm_ring(G_hi, G_lo, head_a['Wt_hi'], head_a['Wt_lo'], nxt, a_rows=rows, K=head_a['Np'],
                              b_slices=9, N=head_a['Kp'], M=rows, ldD=head_a['Kp'], Hp=Hp, Wp=Wp, ring=R,
                              shift=[-s2 for s2 in head_a['shift']], stats=top['sums2'],
                              bw=(top['Y'], top['act_hi'], side, top['mean'], top['invstd']), bw_slope=top['slope'])
            g, cur = nxt, 1
        else:
            self._ctx_head_dz(trunk[-1], self.gD, g, n, rows)
            cur = 0
        for li in reversed(range(len(trunk))):
            L = trunk[li]
            if li > 0:
                prev = trunk[li - 1]
                x_hi, x_lo = prev['act_hi'], prev['act_lo']
            else:
                x_hi, x_lo = self.ctx_in_hi, self.ctx_in_lo
            G_hi, G_lo = self._ctx_layer_bwd(L, g, x_hi, x_lo, rows, count)
            if li == 0:
                break
            nxt = self.ctx_dz[1 - cur]
            ops.tap_gemm_ring(G_hi, G_lo, L['Wt_hi'], L['Wt_lo'], nxt, a_rows=rows, K=L['Np'], b_slices=9,
                              N=L['Kp'], M=rows, ldD=L['Kp'], Hp=Hp, Wp=Wp, ring=R, shift=[-s2 for s2 in L['shift']],
                              stats=prev['sums2'], bw=(prev['Y'], prev['act_hi'], None, prev['mean'], prev['invstd']),
                              bw_slope=prev['slope'])
            g, cur = nxt, 1 - cur

    # ------------------------------------------------------------------ discriminator, tensor-core plan
    def _alloc_discriminator_tc(self):
        """Buffers and tables of the tensor-core discriminator plan (disc_plan.py): per block the index
        tables, bf16 hi/lo operands, the saved post-dropout activation A (fp32) and the block output Z
        (hi/lo, what the next GEMM reads); statistics, bias-gradient and GEMM-space weight-gradient
        pools that one memset clears."""
        from . import disc_plan as DP
        dev, H, W = self.device, self.H, self.W
        M = 2 * self.N
        f32 = dict(dtype=torch.float32, device=dev)
        bf = dict(dtype=torch.bfloat16, device=dev)
        i32 = lambda a: torch.from_numpy(a.astype('int32')).contiguous().to(dev)
        plans = DP.plan(disc_blocks(self.arch_d), H, W)
        ncol = sum(lp['Np'] for lp in plans)
        self._dsums = torch.zeros(2 * ncol, dtype=torch.float64, device=dev)
        self._dsums2 = torch.zeros(2 * ncol, dtype=torch.float64, device=dev)
        self._dbias = torch.zeros(ncol, dtype=torch.float64, device=dev)
        self._dwg = torch.zeros(sum(lp['gmap'].size for lp in plans), **f32)
        g0 = _Geo(M, H // 4, W // 4)
        self.d_s4_hi = torch.zeros(g0.P, 64, **bf)            # 4x4 space-to-depth input, hi / lo
        self.d_s4_lo = torch.zeros(g0.P, 64, **bf)
        self.d_layers = []
        col = wg = 0
        max_pn = g0.P * 64
        max_xp = max_ws = 0
        for lp in plans:
            Hg, Wg = lp['grid']
            geo = _Geo(M, Hg, Wg)
            T, Np, Kp = lp['gmap'].shape
            L = {'name': 'discriminator.discriminator_block_%s' % lp['name'], 'kind': lp['kind'], 'bn': lp['bn'],
                 'cin': lp['cin'], 'cout': lp['cout'], 'T': T, 'Np': Np, 'Kp': Kp, 'geo': geo,
                 'Ho': lp['out_hw'][0], 'Wo': lp['out_hw'][1], 'form_out': lp['form_out'],
                 'gmap': i32(lp['gmap'].reshape(-1)), 'inv': i32(lp['inv'].reshape(-1)),
                 'cmap': i32(lp['cmap']), 'binv': i32(lp['binv'].reshape(-1)), 'R': lp['inv'].shape[1]}
            if lp['kind'] == 'P2':
                L['taps'], L['phases'] = _taps_s2(geo.Wp), 4
                max_xp = max(max_xp, 4 * geo.P * Kp)
                # phase-split copy of the block input: kept for the weight gradient
                L['xp_hi'], L['xp_lo'] = torch.zeros(4 * geo.P * Kp, **bf), torch.zeros(4 * geo.P * Kp, **bf)
            else:
                sh = [di * geo.Wp + dj for di, dj in lp['offsets']]
                L['taps'], L['phases'] = (sh, [0] * T, list(range(T))), 1
            L['W_hi'], L['W_lo'] = torch.zeros(T, Np, Kp, **bf), torch.zeros(T, Np, Kp, **bf)
            L['Wt_hi'], L['Wt_lo'] = torch.zeros(T, Kp, Np, **bf), torch.zeros(T, Kp, Np, **bf)
            L['bias_exp'] = torch.zeros(Np, **f32)
            L['mask'] = torch.ones(M, Np, **f32)
            L['mask_idx'] = torch.from_numpy(lp['cmap'].clip(min=0).astype('int64')).to(dev)
            L['A'] = torch.zeros(geo.P, Np, **f32)
            L['Z_hi'], L['Z_lo'] = torch.zeros(geo.P, Np, **bf), torch.zeros(geo.P, Np, **bf)
            L['sums'] = self._dsums[2 * col:2 * col + 2 * Np].view(2, Np)
            L['sums2'] = self._dsums2[2 * col:2 * col + 2 * Np].view(2, Np)
            L['dbias_exp'] = self._dbias[col:col + Np]
            L['dWg'] = self._dwg[wg:wg + T * Np * Kp]
            for k in ('scale', 'shift', 'mean', 'invstd'):
                L[k] = torch.zeros(Np, **f32)
            L['coef'] = torch.zeros(3, Np, **f32)
            col += Np
            wg += T * Np * Kp
            max_pn = max(max_pn, geo.P * Np, geo.P * Kp)
            max_ws = max(max_ws, ops.wgrad_workspace_floats(geo.P, Np, Kp, T))
            self.d_layers.append(L)
        self.d_dz = [torch.zeros(max_pn, **f32) for _ in range(2)]      # gradient ping-pong (w.r.t. Z)
        self.d_G_hi, self.d_G_lo = torch.zeros(max_pn, **bf), torch.zeros(max_pn, **bf)
        if max_xp:
            self.d_dxp = torch.zeros(max_xp, **f32)
        self.d_wgrad_ws = torch.empty(max_ws, **f32)
        self.validity = torch.zeros(M, 2, **f32)
        self.d_validity = torch.zeros(M, 2, **f32)

    def _disc_input_tc(self, li: int, m: int):
        """(hi, lo, phases, rows) of the GEMM A operand of block li for m frames."""
        L = self.d_layers[li]
        rows = m * L['geo'].Hp * L['geo'].Wp
        if li == 0:
            return self.d_s4_hi, self.d_s4_lo, 1, rows
        if L['phases'] == 4:
            n = 4 * rows * L['Kp']
            return L['xp_hi'][:n].view(4, rows, L['Kp']), L['xp_lo'][:n].view(4, rows, L['Kp']), 4, rows
        prev = self.d_layers[li - 1]
        return prev['Z_hi'], prev['Z_lo'], 1, rows

    def _disc_forward_tc(self, x: torch.Tensor, m: int, train: bool, use_masks: bool):
        """Discriminator*.forward on planar x [m,2,H,W] -> self.validity[:m]; every conv is a tap GEMM
        whose epilogue applies bias + LeakyReLU(0.2) + Dropout2d and accumulates the statistics of the
        BatchNorm2d(eps 0.8) that follows (code/dmcnet_GAN/model.py:254-279)."""
        H, W = self.H, self.W
        if x is not None:          # planar input; None = the caller already filled the space-to-depth operand
            ops.planar_to_s2d4(x.view(-1), 2 * H * W, H, W, m, self.d_s4_hi, self.d_s4_lo)
        if train:
            ops.memset_zero(self._dsums)
        for li, L in enumerate(self.d_layers):
            p, geo, Np, Kp, T = L['name'], L['geo'], L['Np'], L['Kp'], L['T']
            ops.weight_gather_prep(self.p(p + '.0.weight'), L['gmap'], T, Np, Kp, L['W_hi'], L['W_lo'],
                                   L['Wt_hi'], L['Wt_lo'], self.p(p + '.0.bias'), L['cmap'], L['bias_exp'])
            if L['phases'] == 4:
                prev = self.d_layers[li - 1]
                a_hi, a_lo, ph, rows = self._disc_input_tc(li, m)
                ops.phase_split(prev['Z_hi'], prev['Z_lo'], m, prev['geo'].H, prev['geo'].W, Kp, a_hi, a_lo)
            else:
                a_hi, a_lo, ph, rows = self._disc_input_tc(li, m)
            shift, phase, bsel = L['taps']
            ops.tap_gemm_act(a_hi, a_lo, L['W_hi'], L['W_lo'], L['A'], a_phases=ph, a_rows=rows, K=Kp,
                             b_slices=T, N=Np, M=rows, ldD=Np, Hp=geo.Hp, Wp=geo.Wp, shift=shift, phase=phase,
                             bsel=bsel, bias=L['bias_exp'], mask=(L['mask'] if (train and use_masks) else None),
                             slope=0.2, stats=(L['sums'] if (train and L['bn']) else None))
            if L['bn']:
                ops.pm_bn_finalize(L['sums'] if train else None, L['cmap'], Np, L['cout'],
                                   float(m * L['Ho'] * L['Wo']), self.p(p + '.3.weight'), self.p(p + '.3.bias'),
                                   self.buffers[p + '.3.running_mean'], self.buffers[p + '.3.running_var'],
                                   self.buffers[p + '.3.num_batches_tracked'] if train else None, BN_MOMENTUM,
                                   0.8, L['scale'], L['shift'], L['mean'], L['invstd'])
                ops.bn_apply(L['A'], L['scale'], L['shift'], rows, Np, geo.Hp, geo.Wp, False, L['Z_hi'], L['Z_lo'])
            else:
                ops.split_planes(L['A'], rows, Np, geo.Hp, geo.Wp, L['Z_hi'], L['Z_lo'])
        last = self.d_layers[-1]
        ops.pm_linear_fwd(last['Z_hi'], last['Z_lo'], self.p('discriminator.adv_layer.weight'),
                          self.p('discriminator.adv_layer.bias'), m, last['Np'], last['geo'].H, last['geo'].W,
                          self.validity)

    def _disc_backward_tc(self, m: int, need_wgrad: bool, use_masks: bool, to_input: bool, input_rows: int,
                          defer_input: bool = False):
        """Backward from self.d_validity[:m].  Per block: BatchNorm-backward coefficients from the two
        reductions (accumulated by the epilogue of the data-gradient GEMM above it), ONE pass producing
        the hi/lo gradient of the pre-activation (BN backward + Dropout2d + LeakyReLU'), then the
        weight-gradient and data-gradient GEMMs.  to_input: accumulate d(loss)/d(x[:input_rows]) into the
        generator's gradient buffer (GAN G-step)."""
        H, W = self.H, self.W
        last = self.d_layers[-1]
        ops.memset_zero(self._dsums2)
        ops.memset_zero(self._dbias)
        if need_wgrad:
            ops.memset_zero(self._dwg)
        g = self.d_dz[0]
        ops.pm_linear_bwd(self.d_validity, last['Z_hi'], last['Z_lo'], self.p('discriminator.adv_layer.weight'),
                          m, last['Np'], last['geo'].H, last['geo'].W, g,
                          self.g('discriminator.adv_layer.weight') if need_wgrad else None,
                          self.g('discriminator.adv_layer.bias') if need_wgrad else None)
        if last['bn']:
            rows = m * last['geo'].Hp * last['geo'].Wp
            ops.bn_bwd_reduce(g, None, None, last['A'], last['mean'], last['invstd'], rows, last['Np'],
                              last['geo'].Hp, last['geo'].Wp, last['sums2'])
        cur = 0
        for li in reversed(range(len(self.d_layers))):
            L = self.d_layers[li]
            p, geo, Np, Kp, T = L['name'], L['geo'], L['Np'], L['Kp'], L['T']
            rows = m * geo.Hp * geo.Wp
            G_hi, G_lo = self.d_G_hi[:rows * Np], self.d_G_lo[:rows * Np]
            if L['bn']:
                ops.pm_bn_bwd_fold(L['sums2'], L['cmap'], Np, L['cout'], float(m * L['Ho'] * L['Wo']),
                                   self.p(p + '.3.weight'), L['invstd'], L['coef'],
                                   self.g(p + '.3.weight') if need_wgrad else None,
                                   self.g(p + '.3.bias') if need_wgrad else None)
            ops.pm_act_bwd(g, L['A'], L['mean'] if L['bn'] else None, L['invstd'] if L['bn'] else None,
                           L['coef'] if L['bn'] else None, L['mask'] if use_masks else None, 0.2, rows, Np,
                           geo.Hp, geo.Wp, G_hi, G_lo, L['dbias_exp'] if need_wgrad else None)
            a_hi, a_lo, ph, _ = self._disc_input_tc(li, m)
            shift, phase, bsel = L['taps']
            if need_wgrad:
                ops.wgrad_gemm(G_hi, G_lo, a_hi, a_lo, L['dWg'], P=rows, Cout=Np, x_phases=ph, Cin=Kp, shift=shift,
                               phase=phase, bsel=bsel, oihw_taps=0, workspace=self.d_wgrad_ws)
                ops.weight_grad_gather(L['dWg'], L['inv'], L['cout'] * L['cin'] * 9, L['R'], self.g(p + '.0.weight'),
                                       L['dbias_exp'], L['binv'], L['cout'], self.g(p + '.0.bias'))
            if li == 0 and not to_input:
                break
            prev = self.d_layers[li - 1] if li > 0 else None
            nxt = self.d_dz[1 - cur]
            fuse = prev is not None and prev['bn']
            if L['phases'] == 1:
                bw = (prev['A'], None, None, prev['mean'], prev['invstd']) if fuse else None
                ops.tap_gemm(G_hi, G_lo, L['Wt_hi'], L['Wt_lo'], nxt, a_phases=1, a_rows=rows, K=Np, b_slices=T,
                             N=Kp, M=rows, ldD=Kp, Hp=geo.Hp, Wp=geo.Wp, shift=[-s for s in shift],
                             phase=phase, bsel=bsel, stats=(prev['sums2'] if fuse else None), bw=bw)
            else:
                dxp = self.d_dxp[:4 * rows * Kp].view(4, rows, Kp)
                for q in range(4):
                    sh = [-shift[t] for t in range(T) if phase[t] == q]
                    bs = [bsel[t] for t in range(T) if phase[t] == q]
                    ops.tap_gemm(G_hi, G_lo, L['Wt_hi'], L['Wt_lo'], dxp[q], a_phases=1, a_rows=rows, K=Np,
                                 b_slices=T, N=Kp, M=rows, ldD=Kp, Hp=geo.Hp, Wp=geo.Wp, shift=sh,
                                 phase=[0] * len(sh), bsel=bs)
                pg = prev['geo']
                ops.phase_unsplit(dxp, m, pg.H, pg.W, Kp, nxt)
                if fuse:
                    ops.bn_bwd_reduce(nxt, None, None, prev['A'], prev['mean'], prev['invstd'],
                                      m * pg.Hp * pg.Wp, Kp, pg.Hp, pg.Wp, prev['sums2'])
            g, cur = nxt, 1 - cur
        self._disc_dx = g
        if to_input and not defer_input:
            self.disc_input_accumulate(input_rows)

    def disc_input_accumulate(self, input_rows: int):
        """d(loss_adv)/d(gen_flow) (space-to-depth form left by _disc_backward_tc) += into the generator's
        gradient buffer.  Separate from the backward so that a caller running the classifier and the
        discriminator on two streams can order the two accumulations after the join."""
        ops.s2d4_to_planar(self._disc_dx, self.H, self.W, input_rows, self.dD.view(-1),
                           self.dD.shape[1] * self.H * self.W, accumulate=True)

    # ------------------------------------------------------------------ public passes
    def forward(self, input_mv: torch.Tensor, input_residual: torch.Tensor,
                input_flow: Optional[torch.Tensor] = None, *, train: bool = True,
                masks: Optional[Sequence[torch.Tensor]] = None, use_dropout: bool = True):
        """Model.forward.  Returns views of engine-owned outputs:
        (logits [n,C], gen_flow [n,2,H,W]) or, for GAN, (logits, validity [m,2], gen_flow).
        The three parts below can also be called one by one (FusedTrainStep runs the classifier and the
        discriminator, which only share gen_flow, on two streams)."""
        n = self.forward_generator(input_mv, input_residual, train=train)
        self.forward_classifier(n, train=train)
        extra = (self.att_flow[:n],) if self.att else ()        # (base_out, [validity,] gen_flow, att_flow), model.py:354-357
        if not self.gan:
            return (self.logits[:n], self.gen_flow[:n]) + extra
        m = self.forward_discriminator(n, input_flow, train=train, masks=masks, use_dropout=use_dropout)
        return (self.logits[:n], self.validity[:m], self.gen_flow[:n]) + extra

    def forward_generator(self, input_mv: torch.Tensor, input_residual: torch.Tensor, *, train: bool = True) -> int:
        """cat(mv, residual) -> estimator (+ input_mv) -> self.gen_flow (model.py:330-348); returns the frame count."""
        H, W = self.H, self.W
        mv = input_mv.reshape(-1, 2, H, W)
        res = input_residual.reshape(-1, 3, H, W)
        n = mv.shape[0]
        if self.gen_arch == 'context':
            self._ctx_forward(mv, res, n, train)
        else:
            self._gen_forward(mv, res, n)
        return n

    def forward_classifier(self, n: int, *, train: bool = True) -> None:
        """ResNet-18 over self.gen_flow -> self.logits (model.py:349-353)."""
        self._cls_forward(self.gen_flow, n, train)

    def forward_discriminator(self, n: int, input_flow: Optional[torch.Tensor] = None, *, train: bool = True,
                              masks: Optional[Sequence[torch.Tensor]] = None, use_dropout: bool = True) -> int:
        """Discriminator over cat(gen_flow, input_flow) (D-step) or gen_flow alone (G-step) ->
        self.validity[:m]; returns m (GAN/model.py:546-551)."""
        H, W = self.H, self.W
        HW2 = 2 * H * W
        m = n if input_flow is None else 2 * n
        flow = None if input_flow is None else input_flow.reshape(-1, 2, H, W)
        self._use_masks = bool(train and use_dropout)
        if self._use_masks and not (isinstance(masks, str) and masks == 'preloaded'):
            self.set_masks(masks if masks is not None else self.draw_dropout_masks(m), m)
        if self.disc_engine == 'tc':
            # "first fake then real" (GAN/model.py:546-548): the two halves are converted straight into
            # the 4x4 space-to-depth operand, no planar concatenation
            rows = n * ops.padded(H // 4) * ops.padded(W // 4)
            ops.planar_to_s2d4(self.gen_flow.view(-1), HW2, H, W, n, self.d_s4_hi, self.d_s4_lo)
            if flow is not None:
                ops.planar_to_s2d4(flow, HW2, H, W, n, self.d_s4_hi[rows:], self.d_s4_lo[rows:])
            self._disc_forward_tc(None, m, train, self._use_masks)
        else:
            ops.copy_planar(self.gen_flow, HW2, self.d_in.view(-1), HW2, HW2, n)
            if flow is not None:
                ops.copy_planar(flow, HW2, self.d_in.view(-1)[n * HW2:], HW2, HW2, n)
            self._disc_forward(self.d_in, m, train, self._use_masks)
        self._m = m
        return m

    def set_masks(self, masks: Sequence[torch.Tensor], m: int):
        """Stage Dropout2d masks ([m, C] per block) into the static device buffers."""
        for L, mk in zip(self.d_layers, masks):
            mk = mk.reshape(m, -1).to(torch.float32)
            if self.disc_engine == 'tc':
                # one value per (frame, true channel) -> one per GEMM column (space-to-depth columns
                # repeat the channel's value, padding columns are irrelevant: their activation is 0)
                L['mask'][:m].copy_(mk.to(self.device, non_blocking=True)[:, L['mask_idx']], non_blocking=True)
            else:
                L['mask'][:m].copy_(mk, non_blocking=True)

    def draw_dropout_masks(self, m: int, generator: Optional[torch.Generator] = None):
        """Dropout2d(0.25) feature masks drawn with the same ATen calls, shapes and
        order as F.dropout2d inside the reference blocks (GAN/model.py:254-279)."""
        out = []
        for L in self.d_layers:
            noise = torch.empty(m, L['cout'], 1, 1).bernoulli_(0.75, generator=generator).div_(0.75)
            out.append(noise.view(m, L['cout']))
        return out

    def zero_grads(self):
        ops.memset_zero(self.grads)

    def backward(self, n: int, *, cls: bool = True, cls_wgrad: bool = True, gen_grad: bool = True,
                 cls_to_gen: bool = False, disc: bool = False, disc_wgrad: bool = False,
                 disc_to_gen: bool = False, disc_defer_input: bool = False):
        """Backward pass.  Inputs: self.d_logits (classifier), self.d_gen_flow (direct
        gradient on the generated map, e.g. MSE; must be initialised -- zeros if none)
        and self.d_validity (discriminator).  Flags select which parameter gradients are
        produced (dead-work elimination)."""
        if cls:
            self._cls_backward(self.gen_flow, n, cls_wgrad, cls_to_gen, self.d_gen_flow)
        if disc and self.disc_engine == 'tc':
            self._disc_backward_tc(self._m, disc_wgrad, self._use_masks, disc_to_gen, n,
                                   defer_input=disc_defer_input)
        elif disc:
            self._disc_backward(self.d_in, self._m, disc_wgrad, self._use_masks,
                                self.d_gen_flow if disc_to_gen else None, n)
        if gen_grad and self.gen_arch == 'context':
            self._ctx_backward(n)
        elif gen_grad:
            self._gen_backward(n)
