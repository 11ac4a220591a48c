"""One iteration of the I3D trainer (code/dmcnet_I3D/train/model.py:286-446, train_model.py:20-272) on the
I3DEngine: the non-adversarial branch of ``model.fit`` for modality 'flow+mp4' -- forward
``node='flow+logit'``, CE + MSE, one backward, the classifier's optimizer (``optimizer``: base layers with
``lr_mult``, new layers) and the generator's (``optimizer_mse``), gradients accumulated over ``iter_size``
batches and divided by it before the step (:423-432), the two-stage learning-rate rule of
``adjust_learning_rate`` (:268-283) with fresh optimizers at ``epoch_thre`` (:351-356).

Optimizers over the flat buckets: Adam (``dmc_adam_step``; eps 1e-8, the generator's stage-two Adam 1e-3,
train_model.py:122-176) or SGD with Nesterov momentum 0.9 (``dmc_sgd_nesterov_step``), weight decay 1e-4 on
every tensor (train_model.py:112-113).

With ``adv > 0`` (``--adv``, ``--arch-d``: ``optimizer_3``, :357-446) batches alternate in runs of ``iter_size``:
a D stage (loss = CE + adv * adversarial CE of the frame discriminator on [generated | real] frames, steps the
classifier's optimizer and the discriminator's Adam(eps 1e-3)) and a G stage (loss = [0 during epoch 0] * CE
+ MSE + adv * the same adversarial CE, steps the generator's optimizer only).  As in the reference every
backward adds into every parameter's gradient and a stage zeroes only what it steps, so the generator's
gradient of a D stage is part of the next G step and vice versa.
Data parallel: clips sharded over ranks, BatchNorm3d statistics per rank (nn.DataParallel, train_model.py:119),
loss gradients pre-scaled by the global batch, one sum all-reduce of the flat gradient bucket per optimizer step.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch

from . import ops
from .i3d_engine import I3DEngine


@dataclass
class I3DHParams:
    """Defaults of train_hmdb51.py:20-110 (``--optimizer sgd``, ``--fine_tune 1``, ``--drop-out 0.5``)."""
    optim: str = 'sgd'
    lr_base: float = 0.005
    lr_base2: float = 0.002
    weight_decay: float = 1e-4
    iter_size: int = 1
    epoch_thre: int = 1
    fine_tune: bool = True
    detach: bool = False
    dropout: float = 0.5
    momentum: float = 0.9
    betas: tuple = (0.9, 0.999)
    adv: float = 0.0              # --adv: weight of the adversarial loss; > 0 needs an engine built with arch_d
    lr_d: Optional[float] = None  # --lr-d (train_model.py:249-258); None = lr_base


def lr_mult_rule(lr_mult: float, epoch: int, epoch_thre: int) -> float:
    """model.adjust_learning_rate (train/model.py:268-283): groups with lr_mult 0.2 / 0.5 (the convolutional
    part of I3D) get 0 during stage one; 0.5 becomes 1.0 afterwards."""
    if lr_mult in (0.2, 0.5):
        if epoch_thre > 0 and epoch + 1 <= epoch_thre:
            return 0.0
        if lr_mult == 0.5:
            return 1.0
    return lr_mult


def param_group_of(key: str) -> str:
    """train_model.py:62-86 with modality 'flow+mp4'."""
    if key.startswith('gen_flow_model'):
        return 'gf'
    if key.startswith('discriminator'):
        return 'd'
    if key.startswith('conv3d_0c_1x1') or key.startswith('classifier'):
        return 'new'
    return 'base'


class I3DTrainStep:
    def __init__(self, engine: I3DEngine, hp: I3DHParams, *, world_size: int = 1, process_group=None,
                 use_graph: bool = False):
        if hp.optim not in ('sgd', 'adam'):
            raise ValueError("optimizer must be 'sgd' or 'adam' (train_hmdb51.py:80-82)")
        self.eng, self.hp, self.world, self.pg = engine, hp, world_size, process_group
        # use_graph: forward + heads + backward (+ accumulation) of a batch is one CUDA graph per stage, the
        # optimizer step another (the gradient all-reduce of N > 1 sits between them): ~400 launches per step,
        # which bound the step once a rank holds only a few clips (strong scaling of BASELINE config 5)
        self.use_graph = use_graph
        self._graphs: Dict[object, object] = {}
        self.in_data = None
        dev = engine.device
        self.B = engine.clips
        self.target = torch.zeros(self.B, dtype=torch.int64, device=dev)
        self.consensus = torch.zeros(self.B, engine.num_class, dtype=torch.float32, device=dev)
        self.ce_stats = torch.zeros(4, dtype=torch.float32, device=dev)
        self.mse_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        self.adv = hp.adv > 0
        if self.adv and not engine.gan:
            raise ValueError('adv > 0 needs I3DEngine(..., arch_d=...) (train_hmdb51.py --arch-d)')
        self.acc = torch.zeros_like(engine.grads) if (hp.iter_size > 1 or self.adv) else None
        if self.adv:
            n = engine.N
            self.adv_t = torch.cat((torch.zeros(n, dtype=torch.int64), torch.ones(n, dtype=torch.int64))).to(dev)
            self.adv_stats = torch.zeros(4, dtype=torch.float32, device=dev)
        self.lr_mul = 0.2 if hp.fine_tune else 0.5                      # train_model.py:100-105
        self.keys: List[str] = list(engine.specs.keys())
        self.group: List[str] = [param_group_of(k) for k in self.keys]
        chunks = {'cls': [], 'gf': [], 'd': []}
        for ti, (k, g) in enumerate(zip(self.keys, self.group)):
            off, n = engine.offsets[k], engine.numel(k)
            for c0 in range(0, n, 1024):
                chunks[g if g in ('gf', 'd') else 'cls'].append((off + c0, min(1024, n - c0), ti, 0))
        self.chunks = {}
        for g, rows in chunks.items():
            t = torch.tensor(rows, dtype=torch.int32).reshape(-1, 4).contiguous().to(dev)
            self.chunks[g] = (t, len(rows))
        self.hyper = torch.zeros(len(self.keys), 2, dtype=torch.float32, device=dev)
        self.steps = torch.zeros(3, dtype=torch.int32, device=dev)        # Adam step counts: classifier, generator, D
        self.i = 0
        self.i_batch = 0
        self.epoch = -1
        self.stage2 = False
        self.set_epoch(0)

    # ------------------------------------------------------------------ schedule
    def set_epoch(self, epoch: int, lr: Optional[float] = None, lr2: Optional[float] = None):
        """Learning rates for `epoch` (lr / lr2: the MultiFactorScheduler values of the two stages, default
        their base rates).  Entering stage two replaces both optimizers by fresh ones (train/model.py:351-356):
        moments and step counts restart."""
        hp = self.hp
        stage2 = epoch >= hp.epoch_thre
        if stage2 and not self.stage2:                  # optimizer_3 is NOT replaced: the discriminator's state stays
            for tag in ('gen_flow_model', 'i3d'):
                lo, hi = self.eng.group_range[tag]
                if hi > lo:
                    ops.memset_zero(self.eng.exp_avg[lo:hi])
                    ops.memset_zero(self.eng.exp_avg_sq[lo:hi])
            ops.memset_zero(self.steps[0:2])
        self.stage2, self.epoch, self.i_batch = stage2, epoch, 0
        if not stage2:
            lr_g = hp.lr_base if lr is None else lr
            lr_c = 0.0 if hp.detach else lr_g                               # :405-411
        else:
            lr_g = lr_c = hp.lr_base2 if lr2 is None else lr2
        rows = []
        for g in self.group:
            if g == 'gf':
                rows.append((lr_g, hp.weight_decay))
            elif g == 'd':
                rows.append((hp.lr_base if hp.lr_d is None else hp.lr_d, hp.weight_decay))
            else:
                mult = lr_mult_rule(self.lr_mul if g == 'base' else 1.0, epoch, hp.epoch_thre)
                rows.append((lr_c * mult, hp.weight_decay))
        self.hyper.copy_(torch.tensor(rows, dtype=torch.float32))

    # ------------------------------------------------------------------ step
    def _optimizers(self, which, src: torch.Tensor, scale: float):
        eng, hp = self.eng, self.hp
        for g in which:
            t, n = self.chunks[g]
            gi = ('cls', 'gf', 'd').index(g)
            if hp.optim == 'adam' or g == 'd':                  # optimizer_3 is Adam whatever --optimizer says
                eps = 1e-3 if (g == 'd' or (g == 'gf' and self.stage2)) else 1e-8
                ops.adam_step(eng.params, src, eng.exp_avg, eng.exp_avg_sq, t, n, self.hyper.view(-1),
                              self.steps[gi:gi + 1], hp.betas[0], hp.betas[1], eps, scale)
            else:
                ops.sgd_nesterov_step(eng.params, src, eng.exp_avg, t, n, self.hyper.view(-1), hp.momentum, scale)

    def _ranges(self, which):
        gr = self.eng.group_range
        out = []
        for g in which:
            lo, hi = gr[{'cls': 'i3d', 'gf': 'gen_flow_model', 'd': 'discriminator'}[g]]
            if hi > lo:
                out.append((lo, hi))
        return out

    def _fwd_bwd(self, data: torch.Tensor, d_stage: bool, w_ce: float):
        """zero_grad -> forward -> losses -> backward -> accumulate; static buffers only (graph-capturable)."""
        eng, hp, B = self.eng, self.hp, self.B
        n, H, W = eng.N, eng.H, eng.W
        numel = n * 2 * H * W
        eng.zero_grads()
        eng.forward_data(data, train=True)
        ops.ce_head(eng.logits, B, 1, eng.num_class, self.target, w_ce / (B * self.world), self.consensus,
                    eng.d_logits, self.ce_stats)
        # the D stage's loss has no MSE term (:362-369); the kernel still evaluates it for the metrics
        ops.mse_head(eng.gen_flow, eng.in_flow, numel, 0.0 if d_stage else 2.0 / (numel * self.world), eng.dD,
                     self.mse_sum, frame_elems=2 * H * W, dgen_ns=eng.dD.shape[1] * H * W)
        if self.adv:
            # validity of [generated | real] frames against [0 | 1], in BOTH stages (static_model.forward :148-160)
            eng.forward_discriminator(n, eng.in_flow, train=True, masks='preloaded')
            ops.ce_head(eng.validity, 2 * n, 1, 2, self.adv_t, hp.adv / (2 * n * self.world), None, eng.d_validity,
                        self.adv_stats)
        # I3D.forward is never called with detach=True by fit (train/model.py:139-147): the classifier's
        # gradient reaches the generator through the stem.  A zero-weighted CE (G stage of epoch 0) is dead work
        eng.backward(n, cls=w_ce != 0.0, cls_wgrad=True, gen_grad=True, cls_to_gen=True, disc=self.adv,
                     disc_wgrad=self.adv, disc_to_gen=self.adv)
        if self.acc is not None:
            ops.axpy(self.acc, eng.grads, 1.0)

    def _apply(self, which, src: torch.Tensor):
        self._optimizers(which, src, 1.0 / self.hp.iter_size)
        if self.acc is not None:
            for lo, hi in self._ranges(which):
                ops.memset_zero(self.acc[lo:hi])

    def _graphed(self, key, fn):
        """Run fn eagerly (use_graph off, or the first call of a key: warm-up), else capture it once and replay."""
        if not self.use_graph:
            return fn()
        st = self._graphs.get(key)
        if st is None:
            fn()
            self._graphs[key] = 'warm'
            return
        if st == 'warm':
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side, capture_error_mode='thread_local'):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            self._graphs[key] = st = g
        st.replay()

    def step(self, data: torch.Tensor, target: torch.Tensor, dropout_mask: Optional[torch.Tensor] = None,
             metrics: bool = True, disc_masks=None) -> Dict[str, float]:
        """data [B, 7, T, H, W] (device or pinned host), target [B].  dropout_mask [B, 400] replaces the draw of
        nn.Dropout(hp.dropout) (None: drawn here with the same ATen call); disc_masks: the Dropout2d masks of the
        discriminator blocks (adv > 0; None: drawn)."""
        eng, hp, B = self.eng, self.hp, self.B
        dev = eng.device
        if not data.is_cuda and not self.use_graph:
            data = data.to(dev, non_blocking=True)
        self.target.copy_(target, non_blocking=True)
        if hp.dropout > 0.0:
            eng.set_dropout(hp.dropout, dropout_mask if dropout_mask is not None else eng.draw_dropout_mask(hp.dropout))
        else:
            eng.set_dropout(0.0)
        d_stage = self.adv and self.i_batch % (2 * hp.iter_size) < hp.iter_size
        self.i_batch += 1
        w_ce = 1.0 if (d_stage or not self.adv or self.epoch >= 1) else 0.0          # train/model.py:397-402
        n = eng.N
        H, W = eng.H, eng.W
        numel = n * 2 * H * W
        if self.adv:
            eng.set_masks(disc_masks if disc_masks is not None else eng.draw_dropout_masks(2 * n), 2 * n)
        if self.use_graph:
            if self.in_data is None or self.in_data.shape != data.shape:
                self.in_data = torch.empty(data.shape, dtype=torch.float32, device=dev)
            self.in_data.copy_(data, non_blocking=True)
            data = self.in_data
        self._graphed(('fb', bool(d_stage), w_ce), lambda: self._fwd_bwd(data, d_stage, w_ce))
        self.i += 1
        stepped = False
        if self.i % hp.iter_size == 0:
            src = self.acc if self.acc is not None else eng.grads
            which = ('cls', 'gf') if not self.adv else (('cls', 'd') if d_stage else ('gf',))
            if self.world > 1:
                import torch.distributed as dist
                rs = self._ranges(which)
                lo, hi = min(r[0] for r in rs), max(r[1] for r in rs)
                dist.all_reduce(src[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
            self._graphed(('apply', which, self.stage2), lambda: self._apply(which, src))
            self.i = 0
            stepped = True
        if not metrics:
            return {'stepped': stepped}
        ce = self.ce_stats.cpu().tolist()
        out = {'loss_ce': ce[0] / B, 'top1': ce[1] * 100.0 / B, 'top5': ce[2] * 100.0 / B,
               'loss_mse': float(self.mse_sum.cpu()[0]) / numel, 'stepped': stepped,
               'stage': 'D' if d_stage else 'G'}
        if self.adv:
            out['loss_adv'] = self.adv_stats.cpu().tolist()[0] / (2 * n)
        return out

    # ------------------------------------------------------------------ pipelined host input
    def step_pipelined(self, data: torch.Tensor, target: torch.Tensor, dropout_mask: Optional[torch.Tensor] = None,
                       disc_masks=None) -> Dict[str, float]:
        """``step`` for pinned HOST batches: the host->device copy of THIS batch runs on a copy stream while the
        previous batch computes, and the call returns the metrics of the PREVIOUS batch (``{}`` first;
        ``flush()`` returns the last).  Same arithmetic as ``step``."""
        dev = self.eng.device
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage = [torch.empty(data.shape, dtype=torch.float32, device=dev) for _ in range(2)]
            self._stage_t = [torch.empty(target.shape, dtype=torch.int64, device=dev) for _ in range(2)]
            self._ev_copied = [torch.cuda.Event() for _ in range(2)]
            self._ev_free = [torch.cuda.Event() for _ in range(2)]
            self._host_stats = torch.zeros(8, dtype=torch.float64).pin_memory()
            self._k, self._pending = 0, None
        s = self._k & 1
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_stream):
            if self._k >= 2:
                self._copy_stream.wait_event(self._ev_free[s])          # the step that read this slot is done
            self._stage[s].copy_(data, non_blocking=True)
            self._stage_t[s].copy_(target, non_blocking=True)
            self._ev_copied[s].record(self._copy_stream)
        prev = self._collect()
        main.wait_event(self._ev_copied[s])
        out = self.step(self._stage[s], self._stage_t[s], dropout_mask=dropout_mask, metrics=False,
                        disc_masks=disc_masks)
        self._ev_free[s].record(main)
        self._pending = out
        self._k += 1
        return prev

    def _collect(self) -> Dict[str, float]:
        if getattr(self, '_pending', None) is None:
            return {}
        B, numel = self.B, self.eng.N * 2 * self.eng.H * self.eng.W
        ce = self.ce_stats.cpu().tolist()                               # synchronises with the previous step only
        out = dict(self._pending)
        out.update({'loss_ce': ce[0] / B, 'top1': ce[1] * 100.0 / B, 'top5': ce[2] * 100.0 / B,
                    'loss_mse': float(self.mse_sum.cpu()[0]) / numel})
        self._pending = None
        return out

    def flush(self) -> Dict[str, float]:
        return self._collect()
