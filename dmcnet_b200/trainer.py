"""Fused DMC-Net train step (the product's fast path).

``FusedTrainStep.step(input_flow, input_mv, input_residual, target)`` performs
what one iteration of the reference loops does --
code/dmcnet/train.py:221-266 (``gan=False``) and
code/dmcnet_GAN/train.py:237-372 (``gan=True``: D-step on even iterations,
G-step on odd ones) -- entirely with the library's kernels: forward, segment
consensus + CE / MSE / adversarial-CE heads, backward, one gradient all-reduce
(multi-GPU) and the per-tensor Adam of train.py:121-142 / :398-408.

Data parallelism: one process per GPU, each with its own
shard of clips; loss gradients are pre-scaled by the GLOBAL batch so that a
plain sum all-reduce over the flat gradient bucket reproduces the reference's
gradient; BatchNorm statistics stay per rank (= nn.DataParallel semantics).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from .engine import DmcEngine


@dataclass
class HParams:
    """Defaults = exp_my/hmdb51_gen_flow/split1/run.sh:12-35 and exp_my/hmdb51_gan/split1/run.sh:12-39."""
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


GROUPS = ('base_model', 'gen_flow_model', 'discriminator')
FLOW_LOSS_KINDS = {'MSELoss': 0, 'SmoothL1Loss': 1, 'L1': 2}


def flow_loss_kind(name: str) -> int:
    """criterion_mse selection of code/dmcnet/train.py:166-172.  The reference leaves the
    criterion undefined for any other string (NameError at the first iteration); here the
    same mistake is reported when the step is built."""
    if name not in FLOW_LOSS_KINDS:
        raise NameError("name 'criterion_mse' is not defined (--loss-mse %r; expected one of %s)"
                        % (name, ', '.join(FLOW_LOSS_KINDS)))
    return FLOW_LOSS_KINDS[name]


def shard_range(rank: int, world: int, global_batch: int) -> Tuple[int, int]:
    """Clips [lo, hi) of the global batch owned by `rank` (shard on clips, never on
    frames: consensus groups the S frames of one clip)."""
    if global_batch % world:
        raise ValueError('global batch %d is not divisible by world size %d' % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def loss_grad_scales(hp: 'HParams', batch: int, world: int, frames: int, height: int, width: int):
    """Per-rank gradient pre-scales that make a SUM all-reduce reproduce the
    reference's global-mean losses: CE / B_global, 2*MSE / numel_global ('flow' is the same
    normaliser without the factor 2 of d(d^2): the kernel of dmc_flow_loss_head applies the
    criterion's own slope)."""
    numel = float(frames * 2 * height * width * world)
    return {'cls': hp.lr_cls / (batch * world), 'mse': 2.0 * hp.lr_mse / numel,
            'flow': hp.lr_mse / numel}


def allreduce_groups(flat: torch.Tensor, group_range: Dict[str, Tuple[int, int]],
                     groups: Sequence[str], world: int, pg=None) -> Tuple[int, int]:
    """ONE sum all-reduce over the contiguous slice of the flat gradient bucket that
    covers the optimizers about to step.  Returns the [lo, hi) element range."""
    lo = min(group_range[g][0] for g in groups)
    hi = max(group_range[g][1] for g in groups)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=pg)
    return lo, hi


class FusedTrainStep:
    def __init__(self, engine: DmcEngine, hp: HParams, batch: int, *, world_size: int = 1,
                 process_group=None, use_graph: bool = False, pipelined: bool = False,
                 graph_allreduce: bool = False, overlap: Optional[bool] = None):
        self.eng, self.hp, self.B = engine, hp, batch
        self.S = hp.num_segments
        if batch * self.S != engine.N:
            raise ValueError('engine was built for %d frames, need batch*segments = %d'
                             % (engine.N, batch * self.S))
        self.world = world_size
        self.pg = process_group
        self.use_graph = use_graph
        # graph_allreduce: capture the NCCL all-reduce INSIDE the step's CUDA graph (one graph = forward +
        # backward + all-reduce + Adam, no host round trip between them) instead of issuing it eagerly
        # between two graphs
        self.graph_allreduce = bool(graph_allreduce and world_size > 1)
        dev = engine.device
        # two-stream step (see _fwd_bwd).  Default: on for dmcnet_GAN (measured 22.5 -> 21.5 ms per D/G pair at
        # B=64), off for dmcnet (17.4 vs 17.5 ms: the generator's backward and the tap GEMMs each fill every
        # SM -- 200 KB of shared memory / 46 K registers per GEMM CTA leave no room for co-residency -- so the
        # two branches only interleave); DMC_OVERLAP=0/1 or overlap=False/True overrides
        if overlap is None:
            env = os.environ.get('DMC_OVERLAP')
            overlap = bool(engine.gan) if env is None else env != '0'
        self.overlap = bool(overlap) and torch.device(dev).type == 'cuda'
        self._side_stream = torch.cuda.Stream(device=dev) if self.overlap else None
        H, W, n = engine.H, engine.W, engine.N
        f32 = dict(dtype=torch.float32, device=dev)
        # static inputs (graph-replayable)
        self.in_flow = torch.zeros(n, 2, H, W, **f32)
        self.in_mv = torch.zeros(n, 2, H, W, **f32)
        self.in_res = torch.zeros(n, 3, H, W, **f32)
        self.target = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.consensus = torch.zeros(batch, engine.num_class, **f32)
        self.ce_stats = torch.zeros(4, **f32)
        self.adv_stats = torch.zeros(4, **f32)
        self.mse_sum = torch.zeros(1, dtype=torch.float64, device=dev)
        if engine.gan:
            self.adv_t_d = torch.cat((torch.zeros(n, dtype=torch.int64), torch.ones(n, dtype=torch.int64))).to(dev)
            self.adv_t_g = torch.ones(n, dtype=torch.int64, device=dev)
        self.iteration = 0
        self.flow_kind = flow_loss_kind(hp.loss_mse)
        self._build_adam_tables()
        self._graphs: Dict[str, object] = {}
        self._tails: Dict[int, 'FusedTrainStep'] = {}
        self.set_epoch(0, epoch_thre=0)
        self._u8_stages: Dict[int, object] = {}
        self.launches_per_step = 0
        # pipelined input staging: H2D of batch k+1 overlaps the compute of batch k
        self.pipelined = pipelined
        if pipelined:
            self._copy_stream = torch.cuda.Stream(device=dev)
            mk = lambda t: [torch.zeros_like(t) for _ in range(2)]
            self._stg = {'flow': mk(self.in_flow), 'mv': mk(self.in_mv), 'res': mk(self.in_res),
                         'target': mk(self.target)}
            self._ev_h2d = [torch.cuda.Event() for _ in range(2)]
            self._ev_free = [torch.cuda.Event() for _ in range(2)]
            self._ev_done = [torch.cuda.Event() for _ in range(2)]
            self._host_stats = [torch.zeros(16, dtype=torch.float64).pin_memory() for _ in range(2)]
            self._dev_stats = torch.zeros(16, dtype=torch.float64, device=dev)
            self._pending = None          # (slot, mode) of the step whose metrics are still in flight
            self._k = 0

    # ------------------------------------------------------------------ optimizer tables
    def _build_adam_tables(self):
        eng, dev = self.eng, self.eng.device
        self.tensor_keys: List[str] = list(eng.specs.keys())
        self.tensor_group: List[int] = []
        chunks = {g: [] for g in GROUPS}
        for ti, k in enumerate(self.tensor_keys):
            tag = next(g for g in GROUPS if k.startswith(g))
            self.tensor_group.append(GROUPS.index(tag))
            off, n = eng.offsets[k], eng.numel(k)
            for c0 in range(0, n, 1024):
                chunks[tag].append((off + c0, min(1024, n - c0), ti, 0))
        self.chunks = {}
        for g in GROUPS:
            t = torch.tensor(chunks[g], dtype=torch.int32).reshape(-1, 4).contiguous().to(dev)
            self.chunks[g] = (t, len(chunks[g]))
        self.hyper = torch.zeros(len(self.tensor_keys), 2, dtype=torch.float32, device=dev)
        self.steps = torch.zeros(len(GROUPS), dtype=torch.int32, device=dev)

    def set_epoch(self, epoch: int, epoch_thre: int = 0):
        """adjust_learning_rate (code/dmcnet/train.py:398-408).  dmcnet freezes the
        classifier while epoch < epoch_thre (train.py:177,183); the GAN script
        ignores epoch_thre (GAN/train.py:190)."""
        hp = self.hp
        self.freeze = (not self.eng.gan) and epoch < epoch_thre
        decay = hp.lr_decay ** sum(epoch >= s for s in hp.lr_steps)
        mults = (hp.lr_cls_mult, hp.lr_mse_mult, hp.lr_d_mult)
        rows = []
        for k, gi in zip(self.tensor_keys, self.tensor_group):
            lr, wd = hp.lr * decay, hp.weight_decay
            if gi == 0 and self.freeze:
                lr, wd = 0.0, 0.0
            rows.append((lr * mults[gi], wd * (0.0 if 'bias' in k else 1.0)))
        self.hyper.copy_(torch.tensor(rows, dtype=torch.float32))
        # the captured graphs read `hyper` from device memory and 'full' / 'freeze' are separate graph
        # keys, so nothing is re-captured at an epoch boundary.
        # code/dmcnet_GAN/train.py:261,331 alternate D / G on the loader index of the CURRENT epoch
        # (`i % 2`): every epoch starts with a D-step, whatever the length of the previous one
        self.iteration = 0

    # ------------------------------------------------------------------ checkpoint / resume
    def _hyper_rows(self) -> Dict[str, Tuple[float, float]]:
        rows = self.hyper.cpu().tolist()
        return {k: (rows[i][0], rows[i][1]) for i, k in enumerate(self.tensor_keys)}

    def checkpoint(self, epoch: int, arch: str = 'resnet18', best_prec1: float = 0.0) -> dict:
        """The dict the reference passes to ``save_checkpoint`` (code/dmcnet/train.py:190-201; GAN
        code/dmcnet_GAN/train.py:203-215): ``module.``-prefixed state_dict and one
        ``torch.optim.Adam.state_dict()`` per optimizer (one param group per tensor), built from
        the flat buckets -- loadable by the reference's ``--resume`` and by ``resume`` below."""
        from . import checkpoint as C
        hp, eng = self.hp, self.eng
        steps = self.steps.cpu().tolist()
        rows = self._hyper_rows()
        mults = (hp.lr_cls_mult, hp.lr_mse_mult, hp.lr_d_mult)
        out = {'epoch': int(epoch), 'arch': arch,
               'state_dict': C.add_module_prefix({k: v.cpu() for k, v in eng.state_dict().items()}),
               'best_prec1': best_prec1}
        for gi, tag in enumerate(GROUPS):
            if tag == 'discriminator' and not eng.gan:
                continue
            out[C.OPTIMIZER_KEYS[tag]] = C.adam_state_to_torch(eng, tag, int(steps[gi]), rows, mults[gi],
                                                               hp.betas, hp.eps)
        return out

    def resume(self, ckpt: dict) -> Tuple[int, float]:
        """``--resume`` (code/dmcnet/train.py:145-163): model, optimizer moments and step counts.
        Returns (start_epoch, best_prec1); call ``set_epoch`` with the epoch about to run."""
        from . import checkpoint as C
        eng = self.eng
        state = ckpt['state_dict']
        if all(k.startswith('module.') for k in state):                   # saved from nn.DataParallel
            state = C.strip_first_component(state)
        eng.load_state(state)                                             # strict, as train.py:151
        steps = self.steps.cpu().tolist()
        for gi, tag in enumerate(GROUPS):
            key = C.OPTIMIZER_KEYS[tag]
            if key in ckpt and (tag != 'discriminator' or eng.gan):
                steps[gi] = C.adam_state_from_torch(eng, tag, ckpt[key])
        self.steps.copy_(torch.tensor(steps, dtype=torch.int32))
        return int(ckpt['epoch']), ckpt['best_prec1']

    def warm_start(self, state_dict: Dict[str, torch.Tensor]) -> Tuple[List[str], List[str]]:
        """``--weights`` (code/dmcnet/train.py:64-68): drop the first key component, then
        ``load_state_dict(strict=False)``.  Returns (missing_keys, unexpected_keys)."""
        from . import checkpoint as C
        cur = self.eng.state_dict()
        merged, missing, unexpected = C.merge_non_strict(cur, C.strip_first_component(state_dict))
        self.eng.load_state(merged)
        return missing, unexpected

    # ------------------------------------------------------------------ step pieces
    def _adam(self, groups: Sequence[str]):
        eng, hp = self.eng, self.hp
        for g in groups:
            t, n = self.chunks[g]
            gi = GROUPS.index(g)
            ops.adam_step(eng.params, eng.grads, eng.exp_avg, eng.exp_avg_sq, t, n, self.hyper.view(-1),
                          self.steps[gi:gi + 1], hp.betas[0], hp.betas[1], hp.eps, 1.0)

    def _allreduce(self, groups: Sequence[str]):
        if self.world <= 1:
            return
        allreduce_groups(self.eng.grads, self.eng.group_range, groups, self.world, self.pg)

    def _mode(self) -> str:
        if not self.eng.gan:
            return 'freeze' if self.freeze else 'full'
        return 'D' if self.iteration % 2 == 0 else 'G'

    def _fwd_bwd(self, mode: str):
        """forward + heads + backward for one mode (graph-capturable: static buffers only).

        With ``overlap`` (default) the parts of the step that only meet at gen_flow run on two streams
        that fork after the generator's forward and join before Adam -- inside the captured graph they are
        two branches.  dmcnet: the flow loss + generator backward (fp32 FMA kernels) next to the classifier's
        forward + backward (tensor-core and HBM-bound kernels): gen_flow is detached (model.py:352), so
        neither reads what the other writes.  dmcnet_GAN: the discriminator's forward + backward next to
        the classifier's; in the G-step both feed d(gen_flow), so the discriminator's last accumulation is
        issued after the join, in the serial order (bit-identical results either way)."""
        eng, hp, B, S = self.eng, self.hp, self.B, self.S
        n = B * S
        sc = loss_grad_scales(hp, B, self.world, n, eng.H, eng.W)
        g_cls = sc['cls']
        eng.zero_grads()
        if not (self.overlap and (not eng.gan or eng.disc_engine == 'tc')):
            return self._fwd_bwd_serial(mode, n, sc)
        main, side = torch.cuda.current_stream(), self._side_stream
        eng.forward_generator(self.in_mv, self.in_res, train=True)
        side.wait_stream(main)

        def classifier_head():
            eng.forward_classifier(n, train=True)
            ops.ce_head(eng.logits, B, S, eng.num_class, self.target, g_cls, self.consensus,
                        eng.d_logits, self.ce_stats)

        if not eng.gan:
            with torch.cuda.stream(side):
                self._flow_head(sc)
                eng.backward(n, cls=False, gen_grad=True)
            classifier_head()
            eng.backward(n, cls=(mode == 'full'), cls_wgrad=True, gen_grad=False, cls_to_gen=False)
            main.wait_stream(side)
            return
        d_step = mode == 'D'
        m = 2 * n if d_step else n
        with torch.cuda.stream(side):
            eng.forward_discriminator(n, self.in_flow if d_step else None, train=True, masks='preloaded')
            ops.ce_head(eng.validity, m, 1, 2, self.adv_t_d if d_step else self.adv_t_g,
                        (hp.lr_adv_d if d_step else hp.lr_adv_g) / (m * self.world), None,
                        eng.d_validity, self.adv_stats)
            eng.backward(n, cls=False, gen_grad=False, disc=True, disc_wgrad=d_step, disc_to_gen=not d_step,
                         disc_defer_input=True)
        classifier_head()
        if not d_step:
            self._flow_head(sc)
        eng.backward(n, cls=True, cls_wgrad=d_step, gen_grad=False, cls_to_gen=not d_step)
        main.wait_stream(side)
        if not d_step:
            eng.disc_input_accumulate(n)
            eng.backward(n, cls=False, gen_grad=True)

    def _fwd_bwd_serial(self, mode: str, n: int, sc):
        """The same step on one stream."""
        eng, hp, B, S = self.eng, self.hp, self.B, self.S
        g_cls = sc['cls']
        if not eng.gan:
            eng.forward(self.in_mv, self.in_res, train=True)
            ops.ce_head(eng.logits, B, S, eng.num_class, self.target, g_cls, self.consensus,
                        eng.d_logits, self.ce_stats)
            self._flow_head(sc)
            eng.backward(n, cls=(mode == 'full'), cls_wgrad=True, gen_grad=True, cls_to_gen=False)
        elif mode == 'D':
            eng.forward(self.in_mv, self.in_res, self.in_flow, train=True, masks='preloaded')
            ops.ce_head(eng.logits, B, S, eng.num_class, self.target, g_cls, self.consensus,
                        eng.d_logits, self.ce_stats)
            ops.ce_head(eng.validity, 2 * n, 1, 2, self.adv_t_d, hp.lr_adv_d / (2 * n * self.world), None,
                        eng.d_validity, self.adv_stats)
            # generator gradients are dead work in the D-step (GAN/train.py:297-302)
            eng.backward(n, cls=True, cls_wgrad=True, gen_grad=False, cls_to_gen=False, disc=True,
                         disc_wgrad=True, disc_to_gen=False)
        else:
            eng.forward(self.in_mv, self.in_res, None, train=True, masks='preloaded')
            ops.ce_head(eng.logits, B, S, eng.num_class, self.target, g_cls, self.consensus,
                        eng.d_logits, self.ce_stats)
            ops.ce_head(eng.validity, n, 1, 2, self.adv_t_g, hp.lr_adv_g / (n * self.world), None,
                        eng.d_validity, self.adv_stats)
            self._flow_head(sc)
            # classifier / discriminator weight gradients are dead work in the G-step (:367-371)
            eng.backward(n, cls=True, cls_wgrad=False, gen_grad=True, cls_to_gen=True, disc=True,
                         disc_wgrad=False, disc_to_gen=True)

    def _flow_head(self, sc, grads: bool = True):
        """criterion_mse(gen_flow, input_flow) and its gradient into the generator's
        gradient buffer (code/dmcnet/train.py:245; GAN/train.py:350); with --att 1 both sides are weighted
        by the attention map (train.py:246-247) and the map receives a gradient too."""
        eng = self.eng
        numel, frame = eng.N * 2 * eng.H * eng.W, 2 * eng.H * eng.W
        dgen_ns = eng.dD.shape[1] * eng.H * eng.W if grads else frame
        dD = eng.dD if grads else None
        if getattr(eng, 'att', 0):
            if eng.gen_ds:
                # the reference multiplies a 1/f-resolution attention map with the tiled flow: a shape error
                raise RuntimeError('The size of tensor a (%d) must match the size of tensor b (%d): --att 1 with '
                                   '--gen_flow_ds_factor != 0 cannot evaluate its flow loss (reference behaviour, '
                                   'code/dmcnet/train.py:246)' % (eng.gW, eng.W))
            ops.att_flow_loss_head(self.flow_kind, eng.gen_flow, self.in_flow, eng.att_flow, numel,
                                   sc['flow'] if grads else 0.0, dD, eng.d_att if grads else None, self.mse_sum,
                                   frame_elems=frame, dgen_ns=dgen_ns)
        elif self.flow_kind == 0:
            ops.mse_head(eng.gen_flow, self.in_flow, numel, sc['mse'] if grads else 0.0, dD, self.mse_sum,
                         frame_elems=frame, dgen_ns=dgen_ns)
        else:
            ops.flow_loss_head(self.flow_kind, eng.gen_flow, self.in_flow, numel, sc['flow'] if grads else 0.0, dD,
                               self.mse_sum, frame_elems=frame, dgen_ns=dgen_ns)

    def _step_groups(self, mode: str) -> List[str]:
        return {'full': ['base_model', 'gen_flow_model'], 'freeze': ['gen_flow_model'],
                'D': ['base_model', 'discriminator'], 'G': ['gen_flow_model']}[mode]

    def _run(self, mode: str, apply: bool):
        groups = self._step_groups(mode)
        if not self.use_graph:
            c0 = ops.launch_count()
            self._fwd_bwd(mode)
            if apply:
                self._allreduce(groups)
                self._adam(groups)
            self.launches_per_step = ops.launch_count() - c0
            return
        key = mode
        if key not in self._graphs:
            # warm-up eagerly once (sets kernel attributes, fills caches), restoring BN / step state
            # is not needed: the eager run IS this step; later calls replay the captured graphs.
            c0 = ops.launch_count()
            self._fwd_bwd(mode)
            if apply:
                self._allreduce(groups)
                self._adam(groups)
            self.launches_per_step = ops.launch_count() - c0
            torch.cuda.synchronize()
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            self._graphs[key] = ('pending', ga, gb)
            return
        entry = self._graphs[key]
        if entry[0] == 'pending':
            _, ga, gb = entry
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                if self.graph_allreduce and apply:
                    with torch.cuda.graph(ga, stream=s, capture_error_mode='thread_local'):
                        self._fwd_bwd(mode)
                        self._allreduce(groups)
                        self._adam(groups)
                    gb = None
                else:
                    with torch.cuda.graph(ga, stream=s, capture_error_mode='thread_local'):
                        self._fwd_bwd(mode)
                    with torch.cuda.graph(gb, stream=s, capture_error_mode='thread_local'):
                        self._adam(groups)
            torch.cuda.current_stream().wait_stream(s)
            self._graphs[key] = ('ready', ga, gb)
            entry = self._graphs[key]          # capture does not execute: fall through and run the step now
        _, ga, gb = entry
        ga.replay()
        if apply and gb is not None:
            self._allreduce(groups)
            gb.replay()

    # ------------------------------------------------------------------ public API
    def load_inputs(self, input_flow, input_mv, input_residual, target):
        """Stage one batch (host or device tensors, CoviarDataSet layout
        [B,S,c,H,W], code/dmcnet/dataset.py:278) into the static device buffers."""
        H, W = self.eng.H, self.eng.W
        self.in_flow.copy_(input_flow.reshape(-1, 2, H, W), non_blocking=True)
        self.in_mv.copy_(input_mv.reshape(-1, 2, H, W), non_blocking=True)
        self.in_res.copy_(input_residual.reshape(-1, 3, H, W), non_blocking=True)
        self.target.copy_(target, non_blocking=True)

    def load_inputs_u8(self, frames_u8, target, flow_ds_factor: int = 0, flip=None):
        """Stage one batch given as the uint8 stack [B,S,H,W,7] the augmentation produces
        (flow | mv | residual channels, code/dmcnet/dataset.py:210): split, block-mean flow
        target (--flow_ds_factor) and normalisation run on the device (input_stage.py), so the
        host->device copy is 7 B/pixel instead of 28."""
        from .input_stage import U8InputStage
        st = self._u8_stages.get(flow_ds_factor)
        if st is None:
            st = U8InputStage(self.eng.N, self.eng.H, self.eng.W, flow_ds_factor=flow_ds_factor,
                              device=self.eng.device)
            self._u8_stages[flow_ds_factor] = st
        st(frames_u8, self.in_flow, self.in_mv, self.in_res, flip=flip)
        self.target.copy_(target, non_blocking=True)

    def step_u8(self, frames_u8, target, masks: Optional[Sequence[torch.Tensor]] = None,
                flow_ds_factor: int = 0, apply: bool = True, metrics: bool = True, flip=None) -> Dict[str, float]:
        """``step`` fed with the uint8 sample stack (see ``load_inputs_u8``); ``flip``: one boolean
        per clip, the random horizontal flip applied on the device (``U8InputStage``)."""
        self.load_inputs_u8(frames_u8, target, flow_ds_factor, flip)
        return self._step_staged(masks, apply, metrics)

    def step(self, input_flow, input_mv, input_residual, target,
             masks: Optional[Sequence[torch.Tensor]] = None, apply: bool = True,
             metrics: bool = True) -> Dict[str, float]:
        b = int(target.shape[0])
        if b != self.B:
            return self._tail(b).step_from(self, input_flow, input_mv, input_residual, target, masks, apply,
                                           metrics)
        self.load_inputs(input_flow, input_mv, input_residual, target)
        return self._step_staged(masks, apply, metrics)

    # ------------------------------------------------------------------ short final batch
    def _tail(self, b: int) -> 'FusedTrainStep':
        """The reference DataLoaders have no drop_last (code/dmcnet/train.py:72-114), so every epoch
        ends with a smaller batch (HMDB-51 split 1: 3570 clips at batch 45 leave 15).  Such a batch
        runs on a second execution plan sized for it that SHARES this step's parameter, gradient and
        Adam buckets, BatchNorm buffers, hyper table and step counters; losses are means over the
        actual batch, as in the reference.  Plans are cached per size; eager launches (no graph)."""
        if b < 1 or b > self.B:
            raise ValueError('batch of %d clips: this step was built for at most %d' % (b, self.B))
        if self.world > 1:
            raise ValueError('short final batches are not sharded: use drop_last=True with world_size > 1 '
                             '(every rank must see the same number of clips)')
        t = self._tails.get(b)
        if t is None:
            eng = self.eng.sibling(b * self.S)
            t = FusedTrainStep(eng, self.hp, b, use_graph=False)
            t.hyper, t.steps = self.hyper, self.steps
            self._tails[b] = t
        return t

    def step_from(self, parent: 'FusedTrainStep', input_flow, input_mv, input_residual, target, masks, apply,
                  metrics) -> Dict[str, float]:
        """One iteration on behalf of `parent` (see ``_tail``): same mode, same counters."""
        self.freeze, self.iteration = parent.freeze, parent.iteration
        self.load_inputs(input_flow, input_mv, input_residual, target)
        out = self._step_staged(masks, apply, metrics)
        parent.iteration = self.iteration
        return out

    def _step_staged(self, masks, apply: bool, metrics: bool) -> Dict[str, float]:
        """One iteration on the batch already in the static input buffers."""
        eng = self.eng
        mode = self._mode()
        if eng.gan:
            m = 2 * eng.N if mode == 'D' else eng.N
            eng.set_masks(masks if masks is not None else eng.draw_dropout_masks(m), m)
        self._run(mode, apply)
        self.iteration += 1
        return self.read_metrics(mode) if metrics else {}

    def validate_batch(self, input_flow, input_mv, input_residual, target) -> Dict[str, float]:
        """One iteration of ``validate`` (code/dmcnet/train.py:296-347; GAN
        code/dmcnet_GAN/train.py:414-459): eval-mode forward (BatchNorm running statistics, no
        dropout), consensus + CE + top-k, the flow criterion and -- GAN -- the adversarial CE of the
        generated maps against "valid".  No gradients, no parameter or statistic is modified."""
        if int(target.shape[0]) != self.B:                 # last validation batch (no drop_last)
            return self._tail(int(target.shape[0])).validate_batch(input_flow, input_mv, input_residual, target)
        eng, hp, B, n = self.eng, self.hp, self.B, self.eng.N
        self.load_inputs(input_flow, input_mv, input_residual, target)
        eng.forward(self.in_mv, self.in_res, None, train=False)
        ops.ce_head(eng.logits, B, self.S, eng.num_class, self.target, 0.0, self.consensus, None,
                    self.ce_stats)
        self._flow_head(None, grads=False)
        if eng.gan:
            ops.ce_head(eng.validity, n, 1, 2, self.adv_t_g, 0.0, None, None, self.adv_stats)
        return self.read_metrics('G' if eng.gan else 'full')

    # ------------------------------------------------------------------ pipelined API
    def step_pipelined(self, input_flow, input_mv, input_residual, target,
                       masks: Optional[Sequence[torch.Tensor]] = None) -> Dict[str, float]:
        """Same work as ``step`` but asynchronous: the host->device copy of THIS batch runs on a
        copy stream and the call returns the metrics of the PREVIOUS step (``{}`` on the first
        call; ``flush()`` returns the last), so the next batch's copy overlaps this step's
        compute.  Host tensors should be pinned."""
        if not self.pipelined:
            raise RuntimeError('construct FusedTrainStep(..., pipelined=True)')
        eng, H, W = self.eng, self.eng.H, self.eng.W
        s = self._k & 1
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._ev_free[s])
            self._stg['flow'][s].copy_(input_flow.reshape(-1, 2, H, W), non_blocking=True)
            self._stg['mv'][s].copy_(input_mv.reshape(-1, 2, H, W), non_blocking=True)
            self._stg['res'][s].copy_(input_residual.reshape(-1, 3, H, W), non_blocking=True)
            self._stg['target'][s].copy_(target, non_blocking=True)
            self._ev_h2d[s].record(self._copy_stream)
        main.wait_event(self._ev_h2d[s])
        self.in_flow.copy_(self._stg['flow'][s], non_blocking=True)
        self.in_mv.copy_(self._stg['mv'][s], non_blocking=True)
        self.in_res.copy_(self._stg['res'][s], non_blocking=True)
        self.target.copy_(self._stg['target'][s], non_blocking=True)
        self._ev_free[s].record(main)
        mode = self._mode()
        if eng.gan:
            m = 2 * eng.N if mode == 'D' else eng.N
            eng.set_masks(masks if masks is not None else eng.draw_dropout_masks(m), m)
        self._run(mode, True)
        self.iteration += 1
        # stats -> pinned host memory, asynchronously
        self._dev_stats[0:4].copy_(self.ce_stats)
        self._dev_stats[4:8].copy_(self.adv_stats)
        self._dev_stats[8:9].copy_(self.mse_sum)
        self._host_stats[s].copy_(self._dev_stats, non_blocking=True)
        self._ev_done[s].record(main)
        prev, self._pending = self._pending, (s, mode)
        self._k += 1
        return self._collect(prev) if prev is not None else {}

    def step_pipelined_u8(self, frames_u8, target, masks: Optional[Sequence[torch.Tensor]] = None,
                          flow_ds_factor: int = 0) -> Dict[str, float]:
        """``step_pipelined`` fed with the uint8 sample stack [B,S,H,W,7] (``load_inputs_u8``): the
        copy stream moves 7 B/pixel, the split / block-mean / normalisation kernels run on the main
        stream in front of the step.  (Kept separate from ``step_pipelined`` so the measured fp32
        path is untouched; not yet timed on a GPU.)"""
        if not self.pipelined:
            raise RuntimeError('construct FusedTrainStep(..., pipelined=True)')
        from .input_stage import check_stack, normalisation_divisors
        eng, H, W, n = self.eng, self.eng.H, self.eng.W, self.eng.N
        check_stack(frames_u8, n, H, W)
        if '_u8' not in self._stg:
            self._stg['_u8'] = [torch.empty(n, H, W, 7, dtype=torch.uint8, device=eng.device) for _ in range(2)]
            self._u8_divs = normalisation_divisors()
        s = self._k & 1
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._ev_free[s])
            self._stg['_u8'][s].copy_(frames_u8.reshape(n, H, W, 7), non_blocking=True)
            self._stg['target'][s].copy_(target, non_blocking=True)
            self._ev_h2d[s].record(self._copy_stream)
        main.wait_event(self._ev_h2d[s])
        div_motion, div_res = self._u8_divs
        stack = self._stg['_u8'][s]
        if flow_ds_factor == 0:
            ops.unpack_normalize_u8(stack, n, H, W, div_motion, div_res, self.in_flow, self.in_mv, self.in_res)
        else:
            ops.unpack_normalize_u8(stack, n, H, W, div_motion, div_res, None, self.in_mv, self.in_res)
            ops.flow_block_mean_u8(stack, n, H, W, flow_ds_factor, div_motion, self.in_flow)
        self.target.copy_(self._stg['target'][s], non_blocking=True)
        self._ev_free[s].record(main)
        mode = self._mode()
        if eng.gan:
            m = 2 * eng.N if mode == 'D' else eng.N
            eng.set_masks(masks if masks is not None else eng.draw_dropout_masks(m), m)
        self._run(mode, True)
        self.iteration += 1
        self._dev_stats[0:4].copy_(self.ce_stats)
        self._dev_stats[4:8].copy_(self.adv_stats)
        self._dev_stats[8:9].copy_(self.mse_sum)
        self._host_stats[s].copy_(self._dev_stats, non_blocking=True)
        self._ev_done[s].record(main)
        prev, self._pending = self._pending, (s, mode)
        self._k += 1
        return self._collect(prev) if prev is not None else {}

    def flush(self) -> Dict[str, float]:
        """Metrics of the last pipelined step (blocks until it finished)."""
        prev, self._pending = self._pending, None
        return self._collect(prev) if prev is not None else {}

    def _collect(self, pending) -> Dict[str, float]:
        s, mode = pending
        self._ev_done[s].synchronize()
        h = self._host_stats[s]
        return self._metrics_from(mode, [float(h[i]) for i in range(0, 4)], float(h[8]),
                                  [float(h[i]) for i in range(4, 8)])

    def read_metrics(self, mode: str) -> Dict[str, float]:
        """One small device->host read (the reference does five .data[0] syncs, train.py:251-255)."""
        ce = self.ce_stats.cpu().tolist()
        mse = float(self.mse_sum.cpu()[0]) if mode in ('full', 'freeze', 'G') else 0.0
        adv = self.adv_stats.cpu().tolist() if mode in ('D', 'G') else [0.0] * 4
        return self._metrics_from(mode, ce, mse, adv)

    def _metrics_from(self, mode: str, ce, mse_sum: float, adv) -> Dict[str, float]:
        eng, hp, B = self.eng, self.hp, self.B
        n = B * self.S
        out = {'loss_cls': float(ce[0]) / B, 'prec1': float(ce[1]) * 100.0 / B,
               'prec5': float(ce[2]) * 100.0 / B}
        loss = out['loss_cls'] * hp.lr_cls
        if mode in ('full', 'freeze', 'G'):
            out['loss_mse'] = mse_sum / float(n * 2 * eng.H * eng.W)
            loss += out['loss_mse'] * hp.lr_mse
        if mode in ('D', 'G'):
            m = 2 * n if mode == 'D' else n
            out['loss_adv'] = float(adv[0]) / m
            out['acc_adv'] = float(adv[1]) * 100.0 / m
            loss += out['loss_adv'] * (hp.lr_adv_d if mode == 'D' else hp.lr_adv_g)
        out['loss'] = loss
        return out
