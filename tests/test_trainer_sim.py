"""CPU: the host logic of FusedTrainStep end to end on a simulated engine (tests/sim_engine.py:
the engine's own parameter table and buckets, autograd on the oracle's model instead of kernels,
torch expressions of the head / Adam C-ABI calls).  Checks against the oracle trainer what no
kernel test can: mode selection, loss-gradient scales, dead-work flags, Adam chunk and
hyper-parameter tables, step counters, metrics, validate_batch, checkpoint / resume and the epoch
driver.  The same scenarios run on the real engine in the -m gpu tests."""
import pytest
import torch

from dmcnet_b200 import checkpoint as C
from dmcnet_b200 import loop as L
from dmcnet_b200.trainer import FusedTrainStep, HParams
from oracle import dmc_oracle as O
from sim_engine import SimEngine, patch_ops

HW = 32


def _pair(monkeypatch, batch=2, num_class=11, arch_d=None, hw=HW, hp_kw=None, seed=0):
    patch_ops(monkeypatch)
    gan = arch_d is not None
    sd = O.build_state(num_class, arch_d, seed=1)
    data = O.make_inputs(batch, 3, num_class, seed=seed, hw=hw)
    hp_kw = hp_kw or {}
    ref = O.OracleTrainer(sd, O.HParams(**hp_kw), gan=gan, arch_d=arch_d)
    eng = SimEngine(num_class, 3, batch * 3, gan=gan, arch_d=arch_d, height=hw, width=hw)
    eng.load_state(sd)
    tr = FusedTrainStep(eng, HParams(**hp_kw), batch)
    return sd, data, ref, eng, tr


def _close_state(eng, ref, tol=2e-5):
    a, b = eng.state_dict(), ref.state_dict()
    assert list(a) == list(b)
    for k in a:
        err = float((a[k].double() - b[k].double()).abs().max())
        assert err <= tol * (1.0 + float(b[k].double().abs().max())), (k, err)


def _close_metrics(mg, mo, keys=None, rel=2e-5):
    for k in (keys or mo):
        assert mg[k] == pytest.approx(mo[k], rel=rel, abs=1e-6), k


@pytest.mark.parametrize('loss_mse', ['MSELoss', 'SmoothL1Loss', 'L1'])
def test_dmcnet_steps_frozen_then_full(monkeypatch, loss_mse):
    sd, (flow, mv, res, target), ref, eng, tr = _pair(monkeypatch, hp_kw=dict(loss_mse=loss_mse, lr_steps=(2,)))
    for epoch in range(3):                       # epoch 0 frozen, 1 full, 2 full at the decayed rate
        ref.set_epoch(epoch, epoch_thre=1)
        tr.set_epoch(epoch, epoch_thre=1)
        mo = ref.step(flow, mv, res, target)
        mg = tr.step(flow, mv, res, target)
        _close_metrics(mg, mo)
        _close_state(eng, ref)
    assert tr.steps.tolist() == [2, 3, 0]
    got = tr.checkpoint(3)['optimizer_cls']
    want = ref.opt_cls.state_dict()
    assert [g['lr'] for g in got['param_groups']] == pytest.approx([g['lr'] for g in want['param_groups']])
    assert [g['weight_decay'] for g in got['param_groups']] == \
        pytest.approx([g['weight_decay'] for g in want['param_groups']])


def test_gan_d_step_then_g_step_with_dead_work_flags(monkeypatch):
    arch_d = 'Discriminator'
    sd, (flow, mv, res, target), ref, eng, tr = _pair(monkeypatch, batch=1, num_class=51, arch_d=arch_d, hw=224)
    for it in range(2):
        torch.manual_seed(100 + it)
        masks = O.draw_dropout_masks(arch_d, 3 * (2 if it % 2 == 0 else 1))
        if it == 1:
            # The generator gradient of a G-step passes through ResNet-18's ReLU / max-pool switches:
            # the 1e-7 rounding difference the D-step's Adam leaves in the classifier flips a few of
            # them and moves that gradient by ~2e-3.  The host logic is what is
            # under test, so the G-step starts from the oracle's exact parameters.
            eng.load_state(ref.state_dict())
        mo = ref.step(flow, mv, res, target, masks=masks)
        mg = tr.step(flow, mv, res, target, masks=masks)
        _close_metrics(mg, mo)
        _close_state(eng, ref)
        og = ref.grads()
        stepped = ('base_model', 'discriminator') if it == 0 else ('gen_flow_model',)
        for k in eng.specs:
            if k.startswith(stepped):
                err = float((eng.grad_view(k) - og[k]).abs().max())
                assert err <= 1e-4 * float(og[k].abs().max()) + 1e-9, (it, k, err)
            else:                                # dead work: never produced (GAN/train.py:297-302, :367-371)
                assert float(eng.grad_view(k).abs().max()) == 0.0, (it, k)
    assert tr.steps.tolist() == [1, 1, 1]


@pytest.mark.parametrize('arch_d,hw', [(None, HW), ('Discriminator', 224)])
def test_validate_batch_changes_nothing_and_matches_oracle(monkeypatch, arch_d, hw):
    sd, (flow, mv, res, target), ref, eng, tr = _pair(monkeypatch, batch=1, num_class=51, arch_d=arch_d, hw=hw)
    before = eng.state_dict()
    mg = tr.validate_batch(flow, mv, res, target)
    mo = O.validate_batch(sd, O.HParams(), flow, mv, res, target, gan=arch_d is not None, arch_d=arch_d)
    assert set(mg) == set(mo)
    _close_metrics(mg, mo)
    after = eng.state_dict()
    assert all(torch.equal(before[k], after[k]) for k in before) and tr.steps.tolist() == [0, 0, 0]


def test_checkpoint_resume_interop_with_torch_adam(monkeypatch, tmp_path):
    """The scenario of the GPU test test_checkpoint_resumes_on_the_oracle_and_back."""
    sd, (flow, mv, res, target), _, eng, tr = _pair(monkeypatch)
    tr.step(flow, mv, res, target)
    path = C.save_checkpoint(tr.checkpoint(epoch=1, best_prec1=3.0), True, str(tmp_path / 'hmdb51'), 'mv')
    ck, ck_o = C.load_checkpoint(path), C.load_checkpoint(path)
    ref = O.OracleTrainer(C.strip_first_component(ck_o['state_dict']), O.HParams(), gan=False)
    ref.opt_cls.load_state_dict(ck_o['optimizer_cls'])
    ref.opt_gf.load_state_dict(ck_o['optimizer_gf'])
    mo = ref.step(flow, mv, res, target)
    eng2 = SimEngine(11, 3, 6, height=HW, width=HW)
    eng2.load_state({k: torch.zeros_like(v) if v.is_floating_point() else v for k, v in sd.items()})
    tr2 = FusedTrainStep(eng2, HParams(), 2)
    assert tr2.resume(ck) == (1, 3.0) and tr2.steps.tolist() == [1, 1, 0]
    m2 = tr2.step(flow, mv, res, target)
    m1 = tr.step(flow, mv, res, target)
    _close_metrics(m2, m1)
    _close_metrics(m2, mo)
    _close_state(eng2, ref)
    _close_state(eng, ref)
    # oracle-written checkpoint -> engine
    ck_back = {'epoch': 2, 'arch': 'resnet18', 'state_dict': C.add_module_prefix(ref.state_dict()),
               'best_prec1': 0.0, 'optimizer_cls': ref.opt_cls.state_dict(), 'optimizer_gf': ref.opt_gf.state_dict()}
    eng3 = SimEngine(11, 3, 6, height=HW, width=HW)
    eng3.load_state(sd)
    tr3 = FusedTrainStep(eng3, HParams(), 2)
    tr3.resume(ck_back)
    assert tr3.steps.tolist() == [2, 2, 0]
    _close_metrics(tr3.step(flow, mv, res, target), ref.step(flow, mv, res, target))
    _close_state(eng3, ref)


def test_warm_start_of_the_gan_stage(monkeypatch):
    patch_ops(monkeypatch)
    stage1, gan_sd = O.build_state(51, None, seed=3), O.build_state(51, 'Discriminator', seed=1)
    eng = SimEngine(51, 3, 3, gan=True, arch_d='Discriminator')
    eng.load_state(gan_sd)
    tr = FusedTrainStep(eng, HParams(), 1)
    missing, unexpected = tr.warm_start(C.add_module_prefix(stage1))
    assert unexpected == [] and missing and all(k.startswith('discriminator') for k in missing)
    now = eng.state_dict()
    for k in now:
        assert torch.equal(now[k], (gan_sd[k] if k.startswith('discriminator') else stage1[k]).to(now[k].dtype)), k


def test_epoch_driver_runs_the_real_step(monkeypatch, tmp_path):
    """The scenario of the GPU test test_epoch_driver_on_the_engine."""
    patch_ops(monkeypatch)
    sd = O.build_state(11, None, seed=1)
    data = [O.make_inputs(1, 3, 11, seed=s, hw=HW) for s in (0, 1)]
    eng = SimEngine(11, 3, 3, height=HW, width=HW)
    eng.load_state(sd)
    tr = FusedTrainStep(eng, HParams(lr_steps=(1,)), 1)
    lines = []
    best = L.fit(tr, data, data[:1], epochs=2, eval_freq=1, epoch_thre=1, model_prefix=str(tmp_path / 'x'),
                 log=lines.append)
    assert tr.steps.tolist() == [2, 4, 0] and 0.0 <= best <= 100.0
    ck = C.load_checkpoint(C.checkpoint_names(str(tmp_path / 'x'), 'mv')[0])
    assert ck['epoch'] in (1, 2) and float(ck['optimizer_gf']['state'][0]['step']) == 2.0 * ck['epoch']
    if ck['epoch'] == 1:
        assert ck['optimizer_cls']['state'] == {} and ck['optimizer_cls']['param_groups'][0]['lr'] == 0.0
    assert float(tr.hyper[len(C.group_keys(eng.specs, 'base_model')), 0]) == pytest.approx(0.01 * 0.1)
    # the same two epochs on the oracle end in the same state
    ref = O.OracleTrainer(sd, O.HParams(lr_steps=(1,)), gan=False)
    for epoch in range(2):
        ref.set_epoch(epoch, epoch_thre=1)
        for flow, mv, res, target in data:
            ref.step(flow, mv, res, target)
    _close_state(eng, ref)


def test_short_last_batch_runs_on_a_sibling_plan_over_the_same_state(monkeypatch):
    """The reference loaders have no drop_last (code/dmcnet/train.py:72-114): the last batch of an
    epoch is smaller.  It runs on a second plan sharing parameters / Adam state / counters; losses are
    means over the actual batch."""
    sd, (flow, mv, res, target), ref, eng, tr = _pair(monkeypatch, batch=3)
    for b in (3, 1, 3, 2):
        mo = ref.step(flow[:b], mv[:b], res[:b], target[:b])
        mg = tr.step(flow[:b], mv[:b], res[:b], target[:b])
        _close_metrics(mg, mo, rel=5e-4)          # four consecutive Adam steps: rounding differences compound
        _close_state(eng, ref, tol=2e-4)
    assert tr.steps.tolist() == [4, 4, 0]
    assert sorted(tr._tails) == [1, 2]
    assert tr._tails[1].eng.params.data_ptr() == eng.params.data_ptr()
    mg = tr.validate_batch(flow[:2], mv[:2], res[:2], target[:2])
    mo = O.validate_batch(ref.state_dict(), O.HParams(), flow[:2], mv[:2], res[:2], target[:2])
    _close_metrics(mg, mo, rel=5e-4)
    with pytest.raises(ValueError):
        tr.step(torch.cat((flow, flow)), torch.cat((mv, mv)), torch.cat((res, res)), torch.cat((target, target)))


def test_gan_alternation_restarts_with_a_d_step_every_epoch(monkeypatch):
    """code/dmcnet_GAN/train.py:261,331 alternate on the per-epoch loader index: with an odd number of
    batches per epoch the next epoch still opens with a D-step."""
    arch_d = 'Discriminator'
    sd, (flow, mv, res, target), ref, eng, tr = _pair(monkeypatch, batch=1, num_class=51, arch_d=arch_d, hw=224)
    modes = []
    for epoch in range(2):
        tr.set_epoch(epoch)
        for i in range(3):
            modes.append(tr._mode())
            tr.iteration += 1                    # what _step_staged does after running the mode
    assert modes == ['D', 'G', 'D', 'D', 'G', 'D']


def test_state_dicts_without_num_batches_tracked_load(monkeypatch):
    """Checkpoints written by torch 0.3.1 (the reference's, README.md:28) lack those keys."""
    sd, data, ref, eng, tr = _pair(monkeypatch)
    eng.buffers['base_model.bn1.num_batches_tracked'].fill_(7)
    old = {k: v for k, v in sd.items() if not k.endswith('num_batches_tracked')}
    eng.load_state(old)
    assert int(eng.buffers['base_model.bn1.num_batches_tracked']) == 7
    missing, unexpected = tr.warm_start({'module.' + k: v for k, v in old.items()})
    assert missing == [] and unexpected == []
