"""CPU, world_size 2, gloo: the data-parallel host logic of FusedTrainStep.

Each rank runs the ORACLE on its shard of clips with the trainer's gradient
pre-scales (loss / global batch), packs the gradients into a flat bucket laid
out like the engine's (classifier | generator | discriminator groups), and the
trainer's single sum all-reduce must reproduce the gradient the reference
computes on the gathered batch with per-replica BatchNorm (nn.DataParallel
semantics, code/dmcnet/train.py:117)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from dmcnet_b200 import trainer as T
from oracle import dmc_oracle as O

HW = 32          # small frames keep the CPU test fast


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_grads(sd, flow, mv, res, target, lo, hi, scales):
    """Oracle forward/backward on clips [lo,hi) with pre-scaled loss gradients."""
    st = {k: (v.clone().requires_grad_(True) if not O.is_buffer(k) else v.clone()) for k, v in sd.items()}
    out, gen = O.model_forward(st, mv[lo:hi], res[lo:hi], train=True)
    out = out.view(-1, 3, out.shape[-1]).mean(1)
    fl = flow[lo:hi].reshape(gen.shape)
    # sum-form losses times the trainer's scales == d/dtheta of the global-mean losses
    loss = F.cross_entropy(out, target[lo:hi], reduction='sum') * scales['cls'] \
        + ((gen - fl) ** 2).sum() * (scales['mse'] / 2.0)
    loss.backward()
    return {k: v.grad for k, v in st.items() if not O.is_buffer(k)}


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    B = 4
    sd = O.build_state(11, None, seed=1)
    flow, mv, res, target = O.make_inputs(B, 3, 11, seed=0, hw=HW)
    lo, hi = T.shard_range(rank, world, B)
    hp = T.HParams()
    scales = T.loss_grad_scales(hp, hi - lo, world, (hi - lo) * 3, HW, HW)
    grads = _shard_grads(sd, flow, mv, res, target, lo, hi, scales)
    # flat bucket with the engine's group layout
    keys = list(grads.keys())
    offs, off, group_range = {}, 0, {}
    for tag in T.GROUPS:
        start = off
        for k in keys:
            if k.startswith(tag):
                offs[k] = off
                off += (grads[k].numel() + 63) // 64 * 64
        group_range[tag] = (start, off)
    flat = torch.zeros(off)
    for k in keys:
        flat[offs[k]:offs[k] + grads[k].numel()] = grads[k].reshape(-1)
    before_gen = flat[group_range['gen_flow_model'][0]:group_range['gen_flow_model'][1]].clone()
    # classifier-only step group first: the generator slice must stay untouched
    lo_e, hi_e = T.allreduce_groups(flat, group_range, ['base_model'], world)
    assert (lo_e, hi_e) == group_range['base_model']
    assert torch.equal(flat[group_range['gen_flow_model'][0]:group_range['gen_flow_model'][1]], before_gen)
    T.allreduce_groups(flat, group_range, ['gen_flow_model'], world)
    if rank == 0:
        # reference gradient of the gathered batch with per-replica BN = sum of the shard gradients
        tot = None
        for r in range(world):
            a, b = T.shard_range(r, world, B)
            g = _shard_grads(sd, flow, mv, res, target, a, b,
                             T.loss_grad_scales(hp, b - a, world, (b - a) * 3, HW, HW))
            tot = g if tot is None else {k: tot[k] + g[k] for k in g}
        worst = 0.0
        for k in keys:
            got = flat[offs[k]:offs[k] + tot[k].numel()].view_as(tot[k])
            worst = max(worst, float((got - tot[k]).abs().max() / (tot[k].abs().max() + 1e-12)))
        ret['worst'] = worst
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range():
    assert T.shard_range(0, 8, 512) == (0, 64) and T.shard_range(7, 8, 512) == (448, 512)
    with pytest.raises(ValueError):
        T.shard_range(0, 3, 64)


def test_loss_scales_reproduce_global_means():
    hp = T.HParams()
    s = T.loss_grad_scales(hp, 64, 8, 192, 224, 224)
    assert s['cls'] == pytest.approx(hp.lr_cls / 512)
    assert s['mse'] == pytest.approx(2 * hp.lr_mse / (512 * 3 * 2 * 224 * 224))


def test_two_rank_gradient_allreduce_matches_reference_semantics():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert ret['worst'] < 1e-5, ret['worst']


# ---------------------------------------------------------------- the real FusedTrainStep, two ranks
def _trainer_worker(rank, world, port, ret):
    """Each rank: FusedTrainStep(world_size=2) on a simulated engine (tests/sim_engine.py) fed with
    its shard of clips; the step's own all-reduce + Adam must leave every rank with the parameters
    that ONE Adam step on the summed shard gradients gives (the reference: nn.DataParallel gathers
    the outputs, takes global-mean losses, reduces the replica gradients, steps once)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _pytest.monkeypatch import MonkeyPatch
    from sim_engine import SimEngine, patch_ops
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    mp_ = MonkeyPatch()
    patch_ops(mp_)
    try:
        B, C = 4, 11
        sd = O.build_state(C, None, seed=1)
        flow, mv, res, target = O.make_inputs(B, 3, C, seed=0, hw=HW)
        lo, hi = T.shard_range(rank, world, B)
        eng = SimEngine(C, 3, (hi - lo) * 3, height=HW, width=HW)
        eng.load_state(sd)
        hp = T.HParams()
        tr = T.FusedTrainStep(eng, hp, hi - lo, world_size=world)
        m = tr.step(flow[lo:hi], mv[lo:hi], res[lo:hi], target[lo:hi])
        mine = {k: eng.param_view(k).clone() for k in eng.specs}
        # every rank holds the same parameters after the step
        flat = torch.cat([v.reshape(-1) for v in mine.values()])
        other = flat.clone()
        dist.broadcast(other, src=0)
        same = bool(torch.equal(flat, other))
        if rank == 0:
            tot = None
            for r in range(world):
                a, b = T.shard_range(r, world, B)
                g = _shard_grads(sd, flow, mv, res, target, a, b,
                                 T.loss_grad_scales(hp, b - a, world, (b - a) * 3, HW, HW))
                tot = g if tot is None else {k: tot[k] + g[k] for k in g}
            ref = O.OracleTrainer(sd, O.HParams(), gan=False)
            ref._zero()
            for k, v in ref.st.items():
                if not O.is_buffer(k):
                    v.grad = tot[k].clone()
            ref.opt_cls.step()
            ref.opt_gf.step()
            worst = max(float((mine[k] - ref.st[k].detach()).abs().max() / (ref.st[k].detach().abs().max() + 1e-12))
                        for k in mine)
            ret['worst'], ret['loss_cls'] = worst, m['loss_cls']
        ret['same_%d' % rank] = same
        dist.barrier()
    finally:
        mp_.undo()
        dist.destroy_process_group()


def test_two_rank_fused_step_matches_one_adam_step_on_the_summed_gradients():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_trainer_worker, args=(world, port, ret), nprocs=world, join=True)
        assert ret['same_0'] and ret['same_1']
        assert ret['worst'] < 2e-5, ret['worst']


# ---------------------------------------------------------------- dmcnet_GAN, two ranks (BASELINE config 4)
def _gan_worker(rank, world, port, ret):
    """D-step then G-step of dmcnet_GAN with one clip per rank: per-rank BatchNorm and Dropout2d
    masks, adversarial / CE / MSE gradients pre-scaled by the GLOBAL batch, one all-reduce over the
    stepped groups -- against the oracle run shard by shard (its shard-mean gradients / world, summed)
    followed by one step of the torch optimizers the reference would step."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _pytest.monkeypatch import MonkeyPatch
    from sim_engine import SimEngine, patch_ops
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(4)
    mp_ = MonkeyPatch()
    patch_ops(mp_)
    try:
        B, C, arch_d = 2, 51, 'Discriminator'
        sd = O.build_state(C, arch_d, seed=1)
        flow, mv, res, target = O.make_inputs(B, 3, C, seed=0)

        def masks_for(r, it):
            g = torch.Generator().manual_seed(1000 + 10 * r + it)
            return O.draw_dropout_masks(arch_d, 3 * (2 if it == 0 else 1), generator=g)

        lo, hi = T.shard_range(rank, world, B)
        eng = SimEngine(C, 3, 3, gan=True, arch_d=arch_d)
        eng.load_state(sd)
        tr = T.FusedTrainStep(eng, T.HParams(), 1, world_size=world)
        state = sd
        worst = 0.0
        for it in range(2):
            tr.step(flow[lo:hi], mv[lo:hi], res[lo:hi], target[lo:hi], masks=masks_for(rank, it))
            if rank == 0:
                total = None
                for r in range(world):
                    a, b = T.shard_range(r, world, B)
                    shard = O.OracleTrainer(state, O.HParams(), gan=True, arch_d=arch_d)
                    shard.iteration = it
                    shard.step(flow[a:b], mv[a:b], res[a:b], target[a:b], masks=masks_for(r, it), apply=False)
                    g = {k: v / world for k, v in shard.grads().items()}
                    total = g if total is None else {k: total[k] + g[k] for k in g}
                ref = O.OracleTrainer(state, O.HParams(), gan=True, arch_d=arch_d)
                if it == 1:                       # Adam moments / step counts of the D-step carry over
                    ref.opt_cls.load_state_dict(prev['cls']); ref.opt_d.load_state_dict(prev['d'])
                ref._zero()
                for k, v in ref.st.items():
                    if not O.is_buffer(k):
                        v.grad = total[k].clone()
                for opt in ((ref.opt_cls, ref.opt_d) if it == 0 else (ref.opt_gf,)):
                    opt.step()
                prev = {'cls': ref.opt_cls.state_dict(), 'd': ref.opt_d.state_dict()}
                for k in eng.specs:
                    # 1e-8 absolute: BatchNorm biases start at 0, so after one step the parameter IS the
                    # Adam update of a ~1e-6 gradient whose fp32 cancellation noise is ~1e-4 relative
                    err = float((eng.param_view(k) - ref.st[k].detach()).abs().max())
                    worst = max(worst, max(0.0, err - 1e-8) / (float(ref.st[k].detach().abs().max()) + 1e-12))
                # the next iteration starts from rank 0's state on the reference side (parameters are
                # identical on every rank; BatchNorm buffers are per rank and do not enter the gradients)
                state = eng.state_dict()
        if rank == 0:
            ret['worst'] = worst
            ret['steps'] = tr.steps.tolist()
        dist.barrier()
    finally:
        mp_.undo()
        dist.destroy_process_group()


def test_two_rank_gan_d_and_g_steps_match_the_summed_shard_gradients():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gan_worker, args=(world, port, ret), nprocs=world, join=True)
        assert ret['steps'] == [1, 1, 1]
        assert ret['worst'] < 5e-5, ret['worst']
