"""I3D classifier behind the DMC generator: execution plan on the tcgen05 tap GEMMs.

Mirrors ``I3D`` of the reference (code/dmcnet_I3D/network/i3d.py:435-533): the per-frame estimator
(``EstimatorDenseNet*``, the kernels of ``DmcEngine``) maps ``input[:, :5]`` to a 2-channel map per frame,
the Inception-3D trunk classifies the ``[B, 2, T, H, W]`` stack.  ``state_dict`` keys and shapes are the
reference's (358 entries for ``I3D(51, 'flow+mp4', arch_estimator='DenseNetTiny')``).

Data layout: every map of the trunk is pixel-major ``[clips][T+1][H+1][W+1][C]`` with
the shared zero ring on the low side of all three dimensions; activations are bf16 hi/lo pairs, conv outputs
and gradients fp32.  Channel counts are padded per BRANCH to multiples of 64, so the concatenation of an
inception block (i3d.py:425-432) is one map whose column ranges are the branches: each branch GEMM writes
its slice (row pitch = the map's width), the BatchNorm3d + ReLU of all four branches is one pass, and the next
block reads the whole map as its K dimension (weights of padding columns are zero).

  * 3x3x3 "SAME" conv  = 27-tap GEMM, row shifts dt*Hp*Wp + dh*Wp + dw (csrc/gemm_tc.cu, dmc_tc_tap_gemm_ex)
  * 1x1x1 conv          = one-tap GEMM; the two 1x1x1 convs in front of the 3x3x3 branches are ONE GEMM
  * 7x7x7 / 2 stem      = 7-tap GEMM over the temporal kernel index on per-frame 7x7x2 patches (K = 98 -> 128),
                          even / odd input frames as two phases (csrc/i3d.cu)
  * MaxPool3dTFPadding  = csrc/i3d.cu (argmax kept as a window code for the backward)
  * BatchNorm3d (train) = statistics in the GEMM epilogue, dmc_bn_finalize over a whole column group,
                          dmc_bn_apply; backward reductions in the epilogue of the data-gradient GEMM
  * head                = AvgPool3d((2,7,7)) + temporal mean as one weighted mean, two small linears, dropout
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .engine import DmcEngine, GEN_GROWTH

BN_MOMENTUM, BN_EPS = 0.1, 1e-5
STEM_KP = 128          # 7 x 7 x 2 spatial patch (98) padded; the temporal kernel index is the tap

# (name, in_channels, [b0, b1a, b1b, b2a, b2b, b3]) -- i3d.py:471-489
MIXED = [('mixed_3b', 192, [64, 96, 128, 16, 32, 32]), ('mixed_3c', 256, [128, 128, 192, 32, 96, 64]),
         ('mixed_4b', 480, [192, 96, 208, 16, 48, 64]), ('mixed_4c', 512, [160, 112, 224, 24, 64, 64]),
         ('mixed_4d', 512, [128, 128, 256, 24, 64, 64]), ('mixed_4e', 512, [112, 144, 288, 32, 64, 64]),
         ('mixed_4f', 528, [256, 160, 320, 32, 128, 128]), ('mixed_5b', 832, [256, 160, 320, 32, 128, 128]),
         ('mixed_5c', 832, [384, 192, 384, 48, 128, 128])]
GEN_TABLE = {'DenseNetTiny': GEN_GROWTH, 'DenseNetSmall': (32, 32, 24, 16, 8), 'DenseNet': (128, 128, 96, 64, 32)}


def pad64(c: int) -> int:
    return (c + 63) // 64 * 64


def i3d_param_specs(num_class: int) -> "OrderedDict[str, Tuple[int, ...]]":
    """Parameters of the trunk in the reference's module order (i3d.py:435-533)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def unit(name, cin, cout, k):
        s[name + '.conv3d.weight'] = (cout, cin, k, k, k)
        s[name + '.batch3d.weight'] = (cout,)
        s[name + '.batch3d.bias'] = (cout,)

    unit('conv3d_1a_7x7', 2, 64, 7)
    unit('conv3d_2b_1x1', 64, 64, 1)
    unit('conv3d_2c_3x3', 64, 192, 3)
    for name, cin, oc in MIXED:
        unit(name + '.branch_0', cin, oc[0], 1)
        unit(name + '.branch_1.0', cin, oc[1], 1)
        unit(name + '.branch_1.1', oc[1], oc[2], 3)
        unit(name + '.branch_2.0', cin, oc[3], 1)
        unit(name + '.branch_2.1', oc[3], oc[4], 3)
        unit(name + '.branch_3.1', cin, oc[5], 1)
    s['conv3d_0c_1x1.conv3d.weight'] = (400, 1024, 1, 1, 1)
    s['conv3d_0c_1x1.conv3d.bias'] = (400,)
    s['classifier.weight'] = (num_class, 400)
    s['classifier.bias'] = (num_class,)
    return s


class _Geo3:
    """[clips][T+1(+t_hi)][H+1][W+1] map: zero ring in front of every dimension; t_hi = 1 (the stem's map)
    keeps one more zero frame behind each clip (its temporal taps reach two frames ahead)."""

    def __init__(self, clips: int, T: int, H: int, W: int, t_hi: int = 0):
        self.clips, self.T, self.H, self.W, self.t_hi = clips, T, H, W, t_hi
        self.Tp, self.Hp, self.Wp = T + 1 + t_hi, H + 1, W + 1
        self.P = clips * self.Tp * self.Hp * self.Wp
        self.count = float(clips * T * H * W)
        self.hp = ops.pack_hp(self.Hp, self.Tp, t_hi)
        self.thw = (T, H, W)

    def taps(self, k: int):
        """(row shifts, weight slices) of a k x k x k "SAME" conv, slice = (kt*k + kh)*k + kw."""
        if k == 1:
            return [0], [0]
        sh, bs = [], []
        for kt in range(3):
            for kh in range(3):
                for kw in range(3):
                    sh.append((kt - 1) * self.Hp * self.Wp + (kh - 1) * self.Wp + (kw - 1))
                    bs.append((kt * 3 + kh) * 3 + kw)
        return sh, bs


class _Group:
    """Columns of one map that share a BatchNorm pass: the units whose outputs are concatenated in it."""

    def __init__(self, width: int):
        self.width = width
        self.units: List['_Unit'] = []


class _Unit:
    """Unit3Dpy (conv -> BatchNorm3d -> ReLU, i3d.py:318-373) as a GEMM: columns [col0, col0 + cop) of a group."""

    def __init__(self, name: str, cin: int, cout: int, k: int, in_cols: Sequence[int], Kp: int):
        self.name, self.cin, self.cout, self.k = name, cin, cout, k
        self.cop = pad64(cout)
        self.in_cols, self.Kp = list(in_cols), Kp
        self.T = k ** 3 if k <= 3 else 7          # taps of its GEMM (the 7x7x7 stem: temporal index only)
        self.group: Optional[_Group] = None
        self.col0 = 0


class I3DEngine(DmcEngine):
    """Execution plan of ``I3D(num_classes, modality, dropout_prob, arch_estimator)`` for `clips` clips of
    `clip_len` frames.  ``arch_estimator=None`` is the plain two-channel I3D (modality 'flow' / 'mv')."""

    def __init__(self, num_class: int, clips: int, clip_len: int = 16, *, arch_estimator: Optional[str] = 'DenseNetTiny',
                 arch_d: Optional[str] = None, height: int = 224, width: int = 224,
                 device: Optional[torch.device] = None, share_from: Optional['I3DEngine'] = None):
        if arch_estimator is not None and arch_estimator not in GEN_TABLE:
            # i3d.py:460-465 builds no estimator for any other string and forward() then feeds the 5-channel
            # input to a 2-channel convolution
            raise ValueError('I3D: arch_estimator must be None, DenseNet, DenseNetSmall or DenseNetTiny')
        if clip_len % 8 or clip_len < 16:
            raise ValueError('I3D: clip_len must be a multiple of 8, at least 16 (AvgPool3d((2,7,7)) needs T/8 >= 2)')
        if height != 224 or width != 224:
            raise ValueError('I3D: AvgPool3d((2,7,7)) + squeeze (i3d.py:484,352-354) fix the frame size at 224 x 224')
        if arch_d is not None and arch_estimator is None:
            raise ValueError('I3D: a discriminator (arch_d) judges generated maps; it needs an estimator')
        self.clips, self.clip_len = clips, clip_len
        self.has_gen = arch_estimator is not None
        self.i3d_arch_estimator = arch_estimator
        self._plan_trunk(clips, clip_len, height, width)
        # arch_d: the frame discriminator of the adversarial stages (i3d.py:466-476, node='D'): the 2-D plans of
        # DmcEngine over the clips*T generated frames [+ as many real ones]
        super().__init__(num_class, clip_len, clips * clip_len, gan=arch_d is not None, arch_d=arch_d,
                         gen_flow_or_delta=0, height=height, width=width, device=device, gemm_engine='tc',
                         share_from=share_from, gen_growth=GEN_TABLE.get(arch_estimator or 'DenseNetTiny'))

    # ------------------------------------------------------------------ plan
    def _plan_trunk(self, clips: int, T: int, H: int, W: int):
        g1 = _Geo3(clips, T // 2, H // 2, W // 2, t_hi=1)
        t2 = ops.maxpool3d_out_shape(g1.thw, (1, 3, 3), (1, 2, 2))
        g2 = _Geo3(clips, *t2)
        t3 = ops.maxpool3d_out_shape(g2.thw, (1, 3, 3), (1, 2, 2))
        g3 = _Geo3(clips, *t3)
        t4 = ops.maxpool3d_out_shape(g3.thw, (3, 3, 3), (2, 2, 2))
        g4 = _Geo3(clips, *t4)
        t5 = ops.maxpool3d_out_shape(g4.thw, (2, 2, 2), (2, 2, 2))
        g5 = _Geo3(clips, *t5)
        self.geos = [g1, g2, g3, g4, g5]
        self.units: "OrderedDict[str, _Unit]" = OrderedDict()
        self.groups: List[_Group] = []

        def group(units: Sequence[_Unit]) -> _Group:
            gr = _Group(sum(u.cop for u in units))
            col = 0
            for u in units:
                u.group, u.col0 = gr, col
                col += u.cop
                gr.units.append(u)
                self.units[u.name] = u
            self.groups.append(gr)
            return gr

        ident = lambda c: list(range(c))
        self.stem = _Unit('conv3d_1a_7x7', 2, 64, 7, [], STEM_KP)
        group([self.stem])
        self.u2b = _Unit('conv3d_2b_1x1', 64, 64, 1, ident(64), 64)
        group([self.u2b])
        self.u2c = _Unit('conv3d_2c_3x3', 64, 192, 3, ident(64), 64)
        group([self.u2c])
        cols, Kp = ident(192), 192                    # channel -> column of the map the next block reads
        self.mixed = []
        for name, cin, oc in MIXED:
            assert cin == len(cols)
            b0 = _Unit(name + '.branch_0', cin, oc[0], 1, cols, Kp)
            b1a = _Unit(name + '.branch_1.0', cin, oc[1], 1, cols, Kp)
            b2a = _Unit(name + '.branch_2.0', cin, oc[3], 1, cols, Kp)
            b1b = _Unit(name + '.branch_1.1', oc[1], oc[2], 3, ident(oc[1]), b1a.cop)
            b2b = _Unit(name + '.branch_2.1', oc[3], oc[4], 3, ident(oc[3]), b2a.cop)
            b3 = _Unit(name + '.branch_3.1', cin, oc[5], 1, cols, Kp)
            mid = group([b1a, b2a])
            cat = group([b0, b1b, b2b, b3])
            self.mixed.append({'name': name, 'b0': b0, 'b1a': b1a, 'b2a': b2a, 'b1b': b1b, 'b2b': b2b, 'b3': b3,
                               'mid': mid, 'cat': cat, 'Kin': Kp,
                               'stage': 2 if name.startswith('mixed_3') else (3 if name.startswith('mixed_4') else 4)})
            cols = []
            for u in (b0, b1b, b2b, b3):
                cols += [u.col0 + c for c in range(u.cout)]
            Kp = cat.width
        assert Kp == 1024 and cols == ident(1024), 'the head reads mixed_5c unpadded'

    # ------------------------------------------------------------------ parameters
    def _param_specs(self):
        specs: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
        if self.has_gen:
            for k, v in super()._param_specs().items():          # reference order: generator, discriminator, trunk
                if k.startswith('gen_flow_model') or k.startswith('discriminator'):
                    specs[k] = v
        specs.update(i3d_param_specs(self.num_class))
        return specs

    def _build_param_table(self):
        """Flat buckets.  BatchNorm vectors are stored per GROUP at the padded column positions (gamma of all
        the group's units adjacent, then beta), so one kernel call covers a whole concatenated map; the
        state_dict views are the first `cout` entries of each unit's span.  Padding entries stay 0."""
        dev = self.device
        self.specs = self._param_specs()
        self.offsets: Dict[str, int] = {}
        self.group_range: Dict[str, Tuple[int, int]] = {}
        off = 0
        al = lambda n: (n + 63) // 64 * 64
        for tag in ('gen_flow_model', 'discriminator'):         # each with its own optimizer
            start = off
            for k, shp in self.specs.items():
                if k.startswith(tag):
                    self.offsets[k] = off
                    off += al(self._numel(shp))
            self.group_range[tag] = (start, off)
        start = off
        for k, shp in self.specs.items():
            if not k.startswith(('gen_flow_model', 'discriminator')) and 'batch3d' not in k:
                self.offsets[k] = off
                off += al(self._numel(shp))
        for gr in self.groups:
            gr.gamma_off, gr.beta_off = off, off + gr.width
            for u in gr.units:
                self.offsets[u.name + '.batch3d.weight'] = gr.gamma_off + u.col0
                self.offsets[u.name + '.batch3d.bias'] = gr.beta_off + u.col0
            off += 2 * gr.width
        self.group_range['i3d'] = (start, off)
        self.total = off
        o = getattr(self, '_share_from', None)
        if o is not None:
            if list(o.specs.items()) != list(self.specs.items()):
                raise ValueError('share_from: the two engines describe different models')
            self.params, self.grads, self.exp_avg, self.exp_avg_sq = o.params, o.grads, o.exp_avg, o.exp_avg_sq
            self.buffers, self._rm, self._rv, self._nbt = o.buffers, o._rm, o._rv, o._nbt
            for gr, og in zip(self.groups, o.groups):
                gr.rm_off = og.rm_off
            return
        z = lambda: torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.params, self.grads, self.exp_avg, self.exp_avg_sq = z(), z(), z(), z()
        ncol = sum(gr.width for gr in self.groups)
        self._rm = torch.zeros(ncol, dtype=torch.float32, device=dev)
        self._rv = torch.ones(ncol, dtype=torch.float32, device=dev)
        self._nbt = torch.zeros(len(self.units), dtype=torch.int64, device=dev)
        self.buffers: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        col = 0
        for gr in self.groups:
            gr.rm_off = col
            col += gr.width
        for k, shp in self.specs.items():                       # BatchNorm2d of the discriminator blocks
            if k.startswith('discriminator') and k.endswith('.weight') and len(shp) == 1:
                base = k[:-len('.weight')]
                self.buffers[base + '.running_mean'] = torch.zeros(shp, dtype=torch.float32, device=dev)
                self.buffers[base + '.running_var'] = torch.ones(shp, dtype=torch.float32, device=dev)
                self.buffers[base + '.num_batches_tracked'] = torch.zeros((), dtype=torch.int64, device=dev)
        for i, u in enumerate(self.units.values()):
            a = u.group.rm_off + u.col0
            self.buffers[u.name + '.batch3d.running_mean'] = self._rm[a:a + u.cout]
            self.buffers[u.name + '.batch3d.running_var'] = self._rv[a:a + u.cout]
            self.buffers[u.name + '.batch3d.num_batches_tracked'] = self._nbt[i]

    @staticmethod
    def _numel(shp) -> int:
        n = 1
        for d in shp:
            n *= d
        return n

    # ------------------------------------------------------------------ allocation
    def _alloc_generator(self):
        if self.has_gen:
            super()._alloc_generator()
        else:
            H, W = self.H, self.W
            self.gen_flow = torch.zeros(self.N, 2, H, W, dtype=torch.float32, device=self.device)
            self.dD = torch.zeros(self.N, 2, H, W, dtype=torch.float32, device=self.device)
            self.d_gen_flow = self.dD

    def _weight_tables(self, u: _Unit, n_total: int, n0: int, gmap: torch.Tensor, inv: torch.Tensor):
        """Fill u's part of a GEMM weight operand [T][n_total][Kp] (columns n0.. of the N dimension):
        gmap[i] = absolute offset of the parameter element (or -1), inv[e] = i for every element e."""
        T, Kp = u.T, u.Kp
        base = self.offsets[u.name + '.conv3d.weight']
        if u.k == 7:                                           # stem: slice kt, column (kh*7+kw)*2 + ci
            co = torch.arange(64).view(64, 1, 1, 1)
            ci = torch.arange(2).view(1, 2, 1, 1)
            kt = torch.arange(7).view(1, 1, 7, 1)
            sp = torch.arange(49).view(1, 1, 1, 49)
            e = ((co * 2 + ci) * 7 + kt) * 49 + sp
            i = (kt * n_total + n0 + co) * Kp + sp * 2 + ci
        else:
            co = torch.arange(u.cout).view(-1, 1, 1)
            ci = torch.arange(u.cin).view(1, -1, 1)
            t = torch.arange(T).view(1, 1, -1)
            e = (co * u.cin + ci) * T + t                      # OIDHW: taps (kt, kh, kw) innermost
            col = torch.tensor(u.in_cols, dtype=torch.int64).view(1, -1, 1)
            i = (t * n_total + n0 + co) * Kp + col
        e, i = torch.broadcast_tensors(e, i)
        e, i = e.reshape(-1), i.reshape(-1)
        gmap[i] = (base + e).to(torch.int32)
        inv[e] = i.to(torch.int32)

    def _make_gemm(self, units: Sequence[_Unit], geo: _Geo3) -> dict:
        """Operands of one GEMM whose N dimension stacks `units` (same input map, same kernel size)."""
        dev = self.device
        u0 = units[0]
        n_total = sum(u.cop for u in units)
        T, Kp = u0.T, u0.Kp
        gmap = torch.full((T * n_total * Kp,), -1, dtype=torch.int32)
        invs, n0 = [], 0
        for u in units:
            inv = torch.full((u.cout * u.cin * (343 if u.k == 7 else u.T),), -1, dtype=torch.int32)
            self._weight_tables(u, n_total, n0, gmap, inv)
            invs.append((u, inv.to(dev)))
            n0 += u.cop
        bf = dict(dtype=torch.bfloat16, device=dev)
        G = {'units': list(units), 'N': n_total, 'Kp': Kp, 'T': T, 'gmap': gmap.to(dev), 'inv': invs,
             'W_hi': torch.zeros(T * n_total * Kp, **bf), 'W_lo': torch.zeros(T * n_total * Kp, **bf),
             'Wt_hi': torch.zeros(T * n_total * Kp, **bf), 'Wt_lo': torch.zeros(T * n_total * Kp, **bf)}
        if u0.k == 7:
            # output frame `to` reads input frame 2*to + kt - 2: even kt -> phase 0 (even frames) shifted by
            # kt/2 - 1 frames, odd kt -> phase 1 shifted by (kt - 3)/2
            fr = geo.Hp * geo.Wp
            G['shift'] = [((kt // 2 - 1) if kt % 2 == 0 else (kt - 3) // 2) * fr for kt in range(7)]
            G['phase'] = [kt % 2 for kt in range(7)]
            G['bsel'] = list(range(7))
        else:
            G['shift'], G['bsel'] = geo.taps(u0.k)
        self._i3d_n_dwg = getattr(self, '_i3d_n_dwg', 0)
        G['dwg_off'] = self._i3d_n_dwg
        self._i3d_n_dwg += T * n_total * Kp
        self._gemms.append(G)
        return G

    def _alloc_classifier(self):
        dev, clips = self.device, self.clips
        f32 = dict(dtype=torch.float32, device=dev)
        bf = dict(dtype=torch.bfloat16, device=dev)
        g1, g2, g3, g4, g5 = self.geos
        self._gemms: List[dict] = []
        # BatchNorm work arrays, one span per group
        ncol = sum(gr.width for gr in self.groups)
        self._scale, self._shift = torch.zeros(ncol, **f32), torch.zeros(ncol, **f32)
        self._mean, self._invstd = torch.zeros(ncol, **f32), torch.zeros(ncol, **f32)
        self._sums = torch.zeros(2 * ncol, dtype=torch.float64, device=dev)
        self._sums2 = torch.zeros(2 * ncol, dtype=torch.float64, device=dev)
        for gr in self.groups:
            a, w = gr.rm_off, gr.width
            gr.scale, gr.shift = self._scale[a:a + w], self._shift[a:a + w]
            gr.mean, gr.invstd = self._mean[a:a + w], self._invstd[a:a + w]
            gr.sums, gr.sums2 = self._sums[2 * a:2 * a + 2 * w], self._sums2[2 * a:2 * a + 2 * w]
            gr.rm, gr.rv = self._rm[a:a + w], self._rv[a:a + w]
            gr.gamma = self.params[gr.gamma_off:gr.gamma_off + w]
            gr.beta = self.params[gr.beta_off:gr.beta_off + w]
            gr.dgamma = self.grads[gr.gamma_off:gr.gamma_off + w]
            gr.dbeta = self.grads[gr.beta_off:gr.beta_off + w]

        def act(geo, width):                 # fp32 conv output + bf16 hi/lo activation of one map
            return {'Y': torch.zeros(geo.P * width, **f32), 'hi': torch.zeros(geo.P * width, **bf),
                    'lo': torch.zeros(geo.P * width, **bf), 'geo': geo, 'width': width}

        # stem: patch operand (two frame-parity phases), conv output, activation; pool 2a
        self.s_A_hi = torch.zeros(2 * g1.P * STEM_KP, **bf)
        self.s_A_lo = torch.zeros(2 * g1.P * STEM_KP, **bf)
        self.m_stem = act(g1, 64)
        self.g_stem = self._make_gemm([self.stem], g1)
        self.p2a = {'hi': torch.zeros(g2.P * 64, **bf), 'lo': torch.zeros(g2.P * 64, **bf),
                    'idx': torch.zeros(g2.P * 64, dtype=torch.uint8, device=dev), 'k': (1, 3, 3), 's': (1, 2, 2),
                    'gin': g1, 'gout': g2, 'C': 64}
        self.m_2b, self.g_2b = act(g2, 64), self._make_gemm([self.u2b], g2)
        self.m_2c, self.g_2c = act(g2, 192), self._make_gemm([self.u2c], g2)

        def pool(gin, gout, C, k, s):
            return {'hi': torch.zeros(gout.P * C, **bf), 'lo': torch.zeros(gout.P * C, **bf),
                    'idx': torch.zeros(gout.P * C, dtype=torch.uint8, device=dev), 'k': k, 's': s, 'gin': gin,
                    'gout': gout, 'C': C}
        self.p3a = pool(g2, g3, 192, (1, 3, 3), (1, 2, 2))
        self.p4a = pool(g3, g4, self.mixed[1]['cat'].width, (3, 3, 3), (2, 2, 2))
        self.p5a = pool(g4, g5, self.mixed[6]['cat'].width, (2, 2, 2), (2, 2, 2))
        for M in self.mixed:
            geo = self.geos[M['stage']]
            M['geo'] = geo
            M['m_mid'] = act(geo, M['mid'].width)
            M['m_cat'] = act(geo, M['cat'].width)
            M['g_mid'] = self._make_gemm([M['b1a'], M['b2a']], geo)
            M['g_b0'] = self._make_gemm([M['b0']], geo)
            M['g_b1'] = self._make_gemm([M['b1b']], geo)
            M['g_b2'] = self._make_gemm([M['b2b']], geo)
            M['g_b3'] = self._make_gemm([M['b3']], geo)
            M['pool'] = pool(geo, geo, M['Kin'], (3, 3, 3), (1, 1, 1))
        # backward scratch: three fp32 gradient maps, hi/lo gradient operands, GEMM-space weight gradients
        biggest = max([g1.P * 64, g2.P * 192] + [M['geo'].P * max(M['cat'].width, M['Kin']) for M in self.mixed])
        self.gbuf = [torch.zeros(biggest, **f32) for _ in range(3)]
        self.G_hi, self.G_lo = torch.zeros(biggest, **bf), torch.zeros(biggest, **bf)
        big_mid = max(M['geo'].P * M['mid'].width for M in self.mixed)
        self.dz_mid = torch.zeros(big_mid, **f32)
        self.Gm_hi, self.Gm_lo = torch.zeros(big_mid, **bf), torch.zeros(big_mid, **bf)
        self._i3d_dwg = torch.zeros(self._i3d_n_dwg, **f32)
        ws = 0
        self._gemm_geo = {id(self.g_stem): g1, id(self.g_2b): g2, id(self.g_2c): g2}
        for M in self.mixed:
            for key in ('g_mid', 'g_b0', 'g_b1', 'g_b2', 'g_b3'):
                self._gemm_geo[id(M[key])] = M['geo']
        for G in self._gemms:
            geo = self._gemm_geo[id(G)]
            ws = max(ws, ops.wgrad_workspace_floats(geo.P, G['N'], G['Kp'], G['T']))
        self.wgrad_ws = torch.zeros(ws, **f32)
        self.stem_dA = None                                   # [2][P1][128] fp32, allocated on first use (data gradient)
        # head
        nc = self.num_class
        self.pooled = torch.zeros(clips, 1024, **f32)
        self.feat = torch.zeros(clips, 400, **f32)
        self.featd = torch.zeros(clips, 400, **f32)
        self.drop_mask = torch.ones(clips, 400, **f32)
        self.logits = torch.zeros(clips, nc, **f32)
        self.d_logits = torch.zeros(clips, nc, **f32)
        self.d_h400d = torch.zeros(clips, 400, **f32)
        self.d_h400 = torch.zeros(clips, 400, **f32)
        self.d_pooled = torch.zeros(clips, 1024, **f32)
        self.dropout_p = 0.0

    # ------------------------------------------------------------------ pieces
    def _prep_weights(self):
        for G in self._gemms:
            ops.weight_gather_prep(self.params, G['gmap'], G['T'], G['N'], G['Kp'], G['W_hi'], G['W_lo'],
                                   G['Wt_hi'], G['Wt_lo'])

    def _gemm_fwd(self, G, a_hi, a_lo, lda, geo, D, ldD, gr: _Group, col0: int, train: bool):
        """conv of G's units: A (row pitch lda) x W -> columns of D; BatchNorm statistics into gr.sums."""
        stats = gr.sums[col0:] if train else None
        ops.tap_gemm_ex(a_hi, a_lo, G['W_hi'], G['W_lo'], D, lda=lda, a_rows=geo.P, K=G['Kp'], b_slices=G['T'],
                        N=G['N'], M=geo.P, ldD=ldD, Hp=geo.hp, Wp=geo.Wp, shift=G['shift'], bsel=G['bsel'],
                        stats=stats, stats_ld=gr.width)

    def _bn_fwd(self, gr: _Group, m: dict, train: bool):
        geo, w = m['geo'], gr.width
        if train:
            ops.bn_finalize(gr.sums, geo.count, gr.gamma, gr.beta, gr.rm, gr.rv, None, BN_MOMENTUM, BN_EPS, w,
                            gr.scale, gr.shift, gr.mean, gr.invstd)
        else:
            ops.bn_eval_coeffs(gr.gamma, gr.beta, gr.rm, gr.rv, BN_EPS, w, gr.scale, gr.shift)
        ops.bn_apply(m['Y'], gr.scale, gr.shift, geo.P, w, geo.hp, geo.Wp, True, m['hi'], m['lo'])

    def _pool_fwd(self, p: dict, x_hi, x_lo):
        ops.maxpool3d_fwd(x_hi, x_lo, self.clips, p['C'], p['gin'].thw, p['k'], p['s'], p['hi'], p['lo'], p['idx'],
                          in_t_hi=p['gin'].t_hi)

    def _pool_bwd(self, p: dict, gout, add, dX):
        ops.maxpool3d_bwd(gout, p['idx'], self.clips, p['C'], p['gin'].thw, p['k'], p['s'], add, dX,
                          in_t_hi=p['gin'].t_hi)

    def _mixed_fwd(self, M: dict, x_hi, x_lo, train: bool):
        geo, mid, cat = M['geo'], M['mid'], M['cat']
        mm, mc, Kin = M['m_mid'], M['m_cat'], M['Kin']
        self._gemm_fwd(M['g_mid'], x_hi, x_lo, Kin, geo, mm['Y'], mid.width, mid, 0, train)
        self._gemm_fwd(M['g_b0'], x_hi, x_lo, Kin, geo, mc['Y'][M['b0'].col0:], cat.width, cat, M['b0'].col0, train)
        p = M['pool']
        self._pool_fwd(p, x_hi, x_lo)
        self._gemm_fwd(M['g_b3'], p['hi'], p['lo'], Kin, geo, mc['Y'][M['b3'].col0:], cat.width, cat, M['b3'].col0,
                       train)
        self._bn_fwd(mid, mm, train)
        for key, ua, ub in (('g_b1', M['b1a'], M['b1b']), ('g_b2', M['b2a'], M['b2b'])):
            self._gemm_fwd(M[key], mm['hi'][ua.col0:], mm['lo'][ua.col0:], mid.width, geo, mc['Y'][ub.col0:],
                           cat.width, cat, ub.col0, train)
        self._bn_fwd(cat, mc, train)
        return mc['hi'], mc['lo']

    # ------------------------------------------------------------------ forward
    def _cls_forward(self, x_planar: torch.Tensor, n: int, train: bool):
        """I3D.forward from conv3d_1a_7x7 on (i3d.py:507-529) over the planar map [clips*T][2][H][W]."""
        if n != self.N:
            raise RuntimeError('engine was built for %d frames, got %d' % (self.N, n))
        g1, g2, g3, g4, g5 = self.geos
        H, W, T, clips = self.H, self.W, self.clip_len, self.clips
        self._prep_weights()
        if train:
            ops.memset_zero(self._sums)
            ops.add_i64(self._nbt, 1)
        ops.i3d_stem_patches(x_planar, 2 * H * W, clips, T, H, W, self.s_A_hi, self.s_A_lo)
        gs, G = self.stem.group, self.g_stem
        ops.tap_gemm(self.s_A_hi, self.s_A_lo, G['W_hi'], G['W_lo'], self.m_stem['Y'], a_phases=2, a_rows=g1.P,
                     K=STEM_KP, b_slices=7, N=64, M=g1.P, ldD=64, Hp=g1.hp, Wp=g1.Wp, shift=G['shift'],
                     phase=G['phase'], bsel=G['bsel'], stats=(gs.sums if train else None))
        self._bn_fwd(gs, self.m_stem, train)
        self._pool_fwd(self.p2a, self.m_stem['hi'], self.m_stem['lo'])
        self._gemm_fwd(self.g_2b, self.p2a['hi'], self.p2a['lo'], 64, g2, self.m_2b['Y'], 64, self.u2b.group, 0, train)
        self._bn_fwd(self.u2b.group, self.m_2b, train)
        self._gemm_fwd(self.g_2c, self.m_2b['hi'], self.m_2b['lo'], 64, g2, self.m_2c['Y'], 192, self.u2c.group, 0,
                       train)
        self._bn_fwd(self.u2c.group, self.m_2c, train)
        self._pool_fwd(self.p3a, self.m_2c['hi'], self.m_2c['lo'])
        x_hi, x_lo = self.p3a['hi'], self.p3a['lo']
        for i, M in enumerate(self.mixed):
            x_hi, x_lo = self._mixed_fwd(M, x_hi, x_lo, train)
            if i == 1:
                self._pool_fwd(self.p4a, x_hi, x_lo)
                x_hi, x_lo = self.p4a['hi'], self.p4a['lo']
            elif i == 6:
                self._pool_fwd(self.p5a, x_hi, x_lo)
                x_hi, x_lo = self.p5a['hi'], self.p5a['lo']
        # head: AvgPool3d((2,7,7)) -> conv3d_0c_1x1 (bias, no BN, no activation) -> mean over T' -> dropout -> classifier
        ops.i3d_head_pool_fwd(x_hi, x_lo, clips, g5.T, g5.H, g5.W, 1024, self.pooled)
        ops.linear_fwd(self.pooled, self.p('conv3d_0c_1x1.conv3d.weight'), self.p('conv3d_0c_1x1.conv3d.bias'),
                       clips, 1024, 400, self.feat)
        use_drop = train and self.dropout_p > 0.0
        if use_drop:
            ops.mul(self.feat, self.drop_mask, self.featd)
        self._head_in = self.featd if use_drop else self.feat
        ops.linear_fwd(self._head_in, self.p('classifier.weight'), self.p('classifier.bias'), clips, 400,
                       self.num_class, self.logits)
        self._used_drop = use_drop

    def draw_dropout_mask(self, p: float, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """The mask torch.nn.Dropout(p) applies to the [clips, 400] features (i3d.py:485,526), drawn with the
        ATen call F.dropout uses in train mode: bernoulli(1 - p) / (1 - p)."""
        if p >= 1.0:
            return torch.zeros(self.clips, 400)
        return torch.empty(self.clips, 400).bernoulli_(1.0 - p, generator=generator).div_(1.0 - p)

    def set_dropout(self, p: float, mask: Optional[torch.Tensor] = None):
        self.dropout_p = float(p)
        if mask is not None:
            self.drop_mask.copy_(mask.to(self.device, torch.float32), non_blocking=True)

    # ------------------------------------------------------------------ backward
    def _bn_bwd(self, gr: _Group, m: dict, g_a, g_b, g_is_dz: bool, G_hi, G_lo):
        """BatchNorm3d + ReLU backward of a whole map: gradient of the activation(s) -> hi/lo gradient of the
        conv outputs; d gamma / d beta of every unit of the group."""
        geo, w = m['geo'], gr.width
        if g_is_dz:            # dz and gr.sums2 were produced by the epilogue of the data-gradient GEMM
            ops.bn_bwd_apply(g_a, None, None, m['Y'], gr.mean, gr.invstd, gr.gamma, gr.sums2, geo.count, geo.P, w,
                             geo.hp, geo.Wp, G_hi, G_lo, None, gr.dgamma, gr.dbeta)
        else:
            ops.bn_bwd_reduce(g_a, g_b, m['hi'], m['Y'], gr.mean, gr.invstd, geo.P, w, geo.hp, geo.Wp, gr.sums2)
            ops.bn_bwd_apply(g_a, g_b, m['hi'], m['Y'], gr.mean, gr.invstd, gr.gamma, gr.sums2, geo.count, geo.P, w,
                             geo.hp, geo.Wp, G_hi, G_lo, None, gr.dgamma, gr.dbeta)

    def _wgrad(self, G: dict, g_hi, g_lo, ldg, x_hi, x_lo, ldx, geo):
        dW = self._i3d_dwg[G['dwg_off']:G['dwg_off'] + G['T'] * G['N'] * G['Kp']]
        sh, bs = G['shift'], G['bsel']
        ops.wgrad_gemm_ex(g_hi, g_lo, x_hi, x_lo, dW, ldg=ldg, ldx=ldx, P=geo.P, Cout=G['N'], Cin=G['Kp'],
                          shift=sh, bsel=bs, workspace=self.wgrad_ws)
        for u, inv in G['inv']:
            ops.weight_grad_gather(dW, inv, inv.numel(), 1, self.g(u.name + '.conv3d.weight'))

    def _dgrad(self, G: dict, g_hi, g_lo, ldg, geo, D, ldD, *, gb=None, fuse=None):
        """Data gradient of G's conv: gradient operand (row pitch ldg, K = G.N) x W^T -> D (N = G.Kp columns,
        row pitch ldD).  fuse = (group, map, col0, width): ReLU mask + BatchNorm-backward reductions of the
        producer of the conv's input in the epilogue (D = dz); gb: a second gradient source added first."""
        bw, stats, sld, N = None, None, 0, G['Kp']
        if fuse is not None:
            gr, m, c0 = fuse
            bw = (m['Y'][c0:], m['hi'][c0:], gb, gr.mean[c0:], gr.invstd[c0:])
            stats, sld, gb = gr.sums2[c0:], gr.width, None
        ops.tap_gemm_ex(g_hi, g_lo, G['Wt_hi'], G['Wt_lo'], D, lda=ldg, a_rows=geo.P, K=G['N'], b_slices=G['T'],
                        N=N, M=geo.P, ldD=ldD, Hp=geo.hp, Wp=geo.Wp, shift=[-s for s in G['shift']], bsel=G['bsel'],
                        stats=stats, stats_ld=sld, bw=bw, gb=gb)

    def _mixed_bwd(self, M: dict, x_hi, x_lo, cur: int, g_b, g_is_dz: bool, need_wgrad: bool, prev):
        """Backward of one inception block.  self.gbuf[cur] holds the gradient of its output map; returns the
        index of the buffer holding the gradient of its INPUT map (dz of the producer when prev is given)."""
        geo, mid, cat = M['geo'], M['mid'], M['cat']
        mm, mc, Kin = M['m_mid'], M['m_cat'], M['Kin']
        P, Wc, Wm = geo.P, cat.width, mid.width
        G_hi, G_lo = self.G_hi[:P * Wc], self.G_lo[:P * Wc]
        self._bn_bwd(cat, mc, self.gbuf[cur][:P * Wc], g_b, g_is_dz, G_hi, G_lo)
        o1, o2 = [i for i in range(3) if i != cur]
        p = M['pool']
        if need_wgrad:
            self._wgrad(M['g_b0'], G_hi[M['b0'].col0:], G_lo[M['b0'].col0:], Wc, x_hi, x_lo, Kin, geo)
            self._wgrad(M['g_b3'], G_hi[M['b3'].col0:], G_lo[M['b3'].col0:], Wc, p['hi'], p['lo'], Kin, geo)
        dzm = self.dz_mid[:P * Wm]
        for key, ua, ub in (('g_b1', M['b1a'], M['b1b']), ('g_b2', M['b2a'], M['b2b'])):
            if need_wgrad:
                self._wgrad(M[key], G_hi[ub.col0:], G_lo[ub.col0:], Wc, mm['hi'][ua.col0:], mm['lo'][ua.col0:], Wm, geo)
            self._dgrad(M[key], G_hi[ub.col0:], G_lo[ub.col0:], Wc, geo, dzm[ua.col0:], Wm, fuse=(mid, mm, ua.col0))
        Gm_hi, Gm_lo = self.Gm_hi[:P * Wm], self.Gm_lo[:P * Wm]
        self._bn_bwd(mid, mm, dzm, None, True, Gm_hi, Gm_lo)
        if need_wgrad:
            self._wgrad(M['g_mid'], Gm_hi, Gm_lo, Wm, x_hi, x_lo, Kin, geo)
        Ta = self.gbuf[o1][:P * Kin]
        self._dgrad(M['g_b3'], G_hi[M['b3'].col0:], G_lo[M['b3'].col0:], Wc, geo, Ta, Kin)
        Tb = self.gbuf[o2][:P * Kin]
        self._pool_bwd(p, Ta, None, Tb)
        Tc = self.gbuf[o1][:P * Kin]
        self._dgrad(M['g_b0'], G_hi[M['b0'].col0:], G_lo[M['b0'].col0:], Wc, geo, Tc, Kin, gb=Tb)
        out = self.gbuf[o2][:P * Kin]
        self._dgrad(M['g_mid'], Gm_hi, Gm_lo, Wm, geo, out, Kin, gb=Tc, fuse=prev)
        return o2

    def _cls_backward(self, x_planar: torch.Tensor, n: int, need_wgrad: bool, need_input_grad: bool,
                      d_input: Optional[torch.Tensor] = None):
        """Backward from self.d_logits; need_input_grad accumulates d loss / d gen_flow into self.dD[:, 0:2]."""
        g1, g2, g3, g4, g5 = self.geos
        clips, nc = self.clips, self.num_class
        gw = (lambda k: self.g(k)) if need_wgrad else (lambda k: None)
        ops.linear_bwd(self.d_logits, self._head_in, self.p('classifier.weight'), clips, 400, nc, self.d_h400d,
                       gw('classifier.weight'), gw('classifier.bias'))
        d_feat = self.d_h400d
        if self._used_drop:
            ops.mul(self.d_h400d, self.drop_mask, self.d_h400)
            d_feat = self.d_h400
        ops.linear_bwd(d_feat, self.pooled, self.p('conv3d_0c_1x1.conv3d.weight'), clips, 1024, 400, self.d_pooled,
                       gw('conv3d_0c_1x1.conv3d.weight'), gw('conv3d_0c_1x1.conv3d.bias'))
        ops.memset_zero(self._sums2)
        if need_wgrad:
            ops.memset_zero(self._i3d_dwg)
        cur = 0
        ops.i3d_head_pool_bwd(self.d_pooled, clips, g5.T, g5.H, g5.W, 1024, self.gbuf[cur][:g5.P * 1024])
        g_is_dz = False
        for i in reversed(range(len(self.mixed))):
            M = self.mixed[i]
            geo = M['geo']
            first_of_stage = i in (0, 2, 7)
            if i == 0:
                x_hi, x_lo = self.p3a['hi'], self.p3a['lo']
            elif i == 2:
                x_hi, x_lo = self.p4a['hi'], self.p4a['lo']
            elif i == 7:
                x_hi, x_lo = self.p5a['hi'], self.p5a['lo']
            else:
                x_hi, x_lo = self.mixed[i - 1]['m_cat']['hi'], self.mixed[i - 1]['m_cat']['lo']
            prev = None if first_of_stage else (self.mixed[i - 1]['cat'], self.mixed[i - 1]['m_cat'], 0)
            cur = self._mixed_bwd(M, x_hi, x_lo, cur, None, g_is_dz, need_wgrad, prev)
            g_is_dz = prev is not None
            if first_of_stage:
                # the block's input is a max-pooled map: route the gradient back through the pool
                p = {0: self.p3a, 2: self.p4a, 7: self.p5a}[i]
                nxt = (cur + 1) % 3
                self._pool_bwd(p, self.gbuf[cur][:geo.P * M['Kin']], None, self.gbuf[nxt][:p['gin'].P * p['C']])
                cur = nxt
        # conv3d_2c_3x3 (its output was pooled: plain gradient), conv3d_2b_1x1 (dz from 2c's epilogue)
        P2 = g2.P
        G_hi, G_lo = self.G_hi[:P2 * 192], self.G_lo[:P2 * 192]
        self._bn_bwd(self.u2c.group, self.m_2c, self.gbuf[cur][:P2 * 192], None, False, G_hi, G_lo)
        if need_wgrad:
            self._wgrad(self.g_2c, G_hi, G_lo, 192, self.m_2b['hi'], self.m_2b['lo'], 64, g2)
        o1 = (cur + 1) % 3
        self._dgrad(self.g_2c, G_hi, G_lo, 192, g2, self.gbuf[o1][:P2 * 64], 64, fuse=(self.u2b.group, self.m_2b, 0))
        G_hi, G_lo = self.G_hi[:P2 * 64], self.G_lo[:P2 * 64]
        self._bn_bwd(self.u2b.group, self.m_2b, self.gbuf[o1][:P2 * 64], None, True, G_hi, G_lo)
        if need_wgrad:
            self._wgrad(self.g_2b, G_hi, G_lo, 64, self.p2a['hi'], self.p2a['lo'], 64, g2)
        o2 = (o1 + 1) % 3
        self._dgrad(self.g_2b, G_hi, G_lo, 64, g2, self.gbuf[o2][:P2 * 64], 64)
        o3 = (o2 + 1) % 3
        self._pool_bwd(self.p2a, self.gbuf[o2][:P2 * 64], None, self.gbuf[o3][:g1.P * 64])
        # stem
        P1 = g1.P
        G_hi, G_lo = self.G_hi[:P1 * 64], self.G_lo[:P1 * 64]
        self._bn_bwd(self.stem.group, self.m_stem, self.gbuf[o3][:P1 * 64], None, False, G_hi, G_lo)
        G = self.g_stem
        if need_wgrad:
            dW = self._i3d_dwg[G['dwg_off']:G['dwg_off'] + 7 * 64 * STEM_KP]
            ops.wgrad_gemm(G_hi, G_lo, self.s_A_hi, self.s_A_lo, dW, P=P1, Cout=64, x_phases=2, Cin=STEM_KP,
                           shift=G['shift'], phase=G['phase'], bsel=G['bsel'], oihw_taps=0, workspace=self.wgrad_ws)
            for u, inv in G['inv']:
                ops.weight_grad_gather(dW, inv, inv.numel(), 1, self.g(u.name + '.conv3d.weight'))
        if need_input_grad:
            if self.stem_dA is None:
                self.stem_dA = torch.zeros(2 * P1 * STEM_KP, dtype=torch.float32, device=self.device)
            for ph in range(2):                  # gradient of each phase's patches: the taps that read it
                kts = [kt for kt in range(7) if G['phase'][kt] == ph]
                ops.tap_gemm_ex(G_hi, G_lo, G['Wt_hi'], G['Wt_lo'], self.stem_dA[ph * P1 * STEM_KP:], lda=64, a_rows=P1,
                                K=64, b_slices=7, N=STEM_KP, M=P1, ldD=STEM_KP, Hp=g1.hp, Wp=g1.Wp,
                                shift=[-G['shift'][kt] for kt in kts], bsel=kts)
            H, W = self.H, self.W
            ops.i3d_stem_patches_bwd(self.stem_dA, clips, self.clip_len, H, W, self.dD.view(-1),
                                     self.dD.shape[1] * H * W, True)

    # ------------------------------------------------------------------ public passes
    def forward_data(self, data: torch.Tensor, *, train: bool = True):
        """I3D.forward(inp, node='flow+logit') on the sample tensor [B, 5 or 7, T, H, W]: returns
        (logits [B, num_class], gen_flow [B*T, 2, H, W]); with 7 channels the flow target
        input[:, 5:7] is unpacked into self.in_flow (train/model.py:139-158, :176)."""
        B, Cd, T = int(data.shape[0]), int(data.shape[1]), int(data.shape[2])
        if B != self.clips or T != self.clip_len or Cd not in (5, 7):
            raise RuntimeError('I3DEngine built for [%d, 5|7, %d, H, W], got %s' % (self.clips, self.clip_len,
                                                                                      tuple(data.shape)))
        self._ensure_inputs()
        ops.i3d_unpack(data, B, Cd, T, self.H * self.W, self.in_mv, self.in_res, self.in_flow)
        return self.forward_frames(train=train)

    def _ensure_inputs(self):
        if not hasattr(self, 'in_mv'):
            f32 = dict(dtype=torch.float32, device=self.device)
            self.in_mv = torch.zeros(self.N, 2, self.H, self.W, **f32)
            self.in_res = torch.zeros(self.N, 3, self.H, self.W, **f32)
            self.in_flow = torch.zeros(self.N, 2, self.H, self.W, **f32)

    def forward_frames(self, *, train: bool = True):
        if self.has_gen:
            self.forward_generator(self.in_mv, self.in_res, train=train)
            self._cls_forward(self.gen_flow, self.N, train)
        else:                       # modality 'flow' / 'mv': the two-channel stack itself
            self._cls_forward(self.in_mv, self.N, train)
        return self.logits, self.gen_flow
