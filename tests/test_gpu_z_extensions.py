"""GPU: the path extensions of DESIGN.md section 10 -- the --loss-mse criteria, video-level
scoring (test.py protocol), the uint8 input stage and checkpoint / resume interop -- against the
oracle, torch fp32 ops and fixtures produced by the reference.  Same bars as
test_gpu_kernels.py / test_gpu_parity.py (bit-exact for the uint8 input stage).  The file sorts
last so the parity tests of the train step proper run first."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import dmc_oracle as O          # noqa: E402  (checker only)

if torch.cuda.is_available():
    from dmcnet_b200 import ops
    from dmcnet_b200.engine import DmcEngine
    from dmcnet_b200.trainer import FusedTrainStep, HParams


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rel2(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


# ------------------------------------------------------------------ --loss-mse criteria
@pytest.mark.parametrize('kind,crit', [(0, F.mse_loss), (1, F.smooth_l1_loss), (2, F.l1_loss)])
def test_flow_loss_head_kernel_vs_torch(kind, crit):
    g = torch.Generator().manual_seed(11 + kind)
    a = (torch.randn(6, 2, 32, 32, generator=g) * 1.5).requires_grad_(True)
    b = torch.randn(6, 2, 32, 32, generator=g)
    with torch.no_grad():
        a[0, 0, 0, :8] = b[0, 0, 0, :8]                       # exact zeros: sign(0) = 0 in ATen
        a[0, 0, 1, :4] = b[0, 0, 1, :4] + 1.0                 # the SmoothL1 knee
    loss = crit(a, b)
    (loss * 10).backward()
    dev = dict(device='cuda')
    # dgen as a 2-channel slice of a wider [N][5][H][W] buffer (the generator's gradient buffer)
    wide = torch.full((6, 5, 32, 32), 7.0, **dev)
    s = torch.full((1,), 123.0, dtype=torch.float64, **dev)
    ops.flow_loss_head(kind, a.detach().cuda(), b.cuda(), a.numel(), 10.0 / a.numel(), wide, s,
                       frame_elems=2 * 32 * 32, dgen_ns=5 * 32 * 32)
    assert abs(float(s.cpu()[0]) / a.numel() - float(loss.detach())) < 1e-6
    assert rel(wide[:, :2], a.grad) < 1e-6
    assert torch.equal(wide[:, 2:].cpu(), torch.full((6, 3, 32, 32), 7.0))
    assert torch.equal(wide[0, 0, 0, :8].cpu() == 0, a.grad[0, 0, 0, :8] == 0)


@pytest.mark.parametrize('loss_mse', ['SmoothL1Loss', 'L1'])
def test_train_step_with_alternative_flow_criterion(loss_mse):
    batch, num_class = 2, 51
    sd = O.build_state(num_class, None, seed=1)
    flow, mv, res, target = O.make_inputs(batch, 3, num_class, seed=0)
    ref = O.OracleTrainer(sd, O.HParams(loss_mse=loss_mse), gan=False)
    eng = DmcEngine(num_class, 3, batch * 3)
    eng.load_state(sd)
    tr = FusedTrainStep(eng, HParams(loss_mse=loss_mse), batch)
    mo = ref.step(flow, mv, res, target)
    mg = tr.step(flow.cuda(), mv.cuda(), res.cuda(), target.cuda())
    for k in ('loss', 'loss_cls', 'loss_mse'):
        assert mg[k] == pytest.approx(mo[k], rel=1e-3, abs=1e-6), k
    # dmcnet: generator gradients come from the flow loss only (plain fp32 kernels).  The L1
    # slope sign(gen - flow) is discontinuous: one element within ~1e-7 of its target may flip
    # between two correct fp32 forwards and moves a weight gradient by ~0.5 %, hence 2e-2 there.
    tol_g, tol_p = (2e-2, 1e-3) if loss_mse == 'L1' else (1e-4, 1e-4)
    og = ref.grads()
    for k in eng.specs:
        if k.startswith('gen_flow_model'):
            assert rel2(eng.grad_view(k), og[k]) < tol_g, k
    new = eng.state_dict()
    for k, v in ref.state_dict().items():
        if k.startswith('gen_flow_model'):
            assert rel(new[k], v) < tol_p, k


# ------------------------------------------------------------------ video-level scoring (test.py protocol)
@pytest.mark.parametrize('max_frames', [None, 3])
def test_video_scorer_vs_oracle_protocol(max_frames, monkeypatch):
    """2 videos x (3 segments x 2 'crops') frames, eval-mode forward, mean of the logits over the
    6 frames of a video (code/dmcnet/test.py:139-151); whole video per launch and 2 launches."""
    from dmcnet_b200 import inference as I
    from oracle import video_protocol as V
    monkeypatch.setattr(I, 'check_crops', lambda c: c)       # 2 crops keep the CPU oracle quick
    num_class, segs, crops = 51, 3, 2
    sd = O.build_state(num_class, 'Discriminator', seed=1)   # a GAN checkpoint: D keys are ignored
    # non-trivial running statistics, as after training
    g = torch.Generator().manual_seed(5)
    for k in sd:
        if k.endswith('running_mean'):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif k.endswith('running_var'):
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
    scorer = I.VideoScorer(sd, num_class, segs, crops, max_frames_per_launch=max_frames)
    assert scorer.launches == (1 if max_frames is None else 2)
    out, ref = [], []
    for v in range(2):
        flow, mv, res, target = O.make_inputs(2, 3, num_class, seed=10 + v)   # [2,3,c,H,W] = 6 frames
        mv, res = mv.reshape(1, 6, 2, 224, 224), res.reshape(1, 6, 3, 224, 224)
        label = int(target[0])
        s = scorer.forward_video(mv, res, label)
        r = V.forward_video(sd, mv, res, segs, crops)
        assert s.shape == r.shape == (1, num_class)
        np.testing.assert_allclose(s, r, rtol=1e-3, atol=1e-3 * np.abs(r).max())
        assert int(s.argmax()) == int(r.argmax())
        st = scorer.last_stats()
        assert st['loss'] == pytest.approx(float(F.cross_entropy(torch.from_numpy(r), target[:1])), rel=1e-3)
        assert st['top1'] == float(int(r.argmax()) == label)
        out.append((s, label)); ref.append((r, label))
    assert I.video_accuracy(out) == V.accuracy(ref)
    # the engine's running statistics are read, never updated, in eval mode
    new = scorer.eng.state_dict()
    for k in new:
        if 'running' in k or 'num_batches' in k:
            assert torch.equal(new[k].cpu(), sd[k].to(new[k].dtype)), k


# ------------------------------------------------------------------ uint8 input stage
def test_input_stage_kernels_bit_exact_vs_reference_golden(golden_dir):
    """tests/golden/input_pipe.npz: outputs of the reference's own dataset.py (oracle/pin_input_pipe.py)."""
    import os
    from dmcnet_b200.input_stage import U8InputStage
    z = np.load(os.path.join(golden_dir, 'input_pipe.npz'))
    for name in sorted({k.split('.')[0] for k in z.files}):
        if bool(z[name + '.flip']):
            continue                              # test_input_stage_flip_bit_exact_vs_reference_golden
        frames, factor = z[name + '.frames'], int(z[name + '.factor'])
        S, H, W, _ = frames.shape
        stage = U8InputStage(S, H, W, flow_ds_factor=factor)
        flow, mv, res = stage(torch.from_numpy(frames))
        torch.cuda.synchronize()
        assert np.array_equal(flow.cpu().numpy(), z[name + '.flow']), name
        assert np.array_equal(mv.cpu().numpy(), z[name + '.mv']), name
        assert np.array_equal(res.cpu().numpy(), z[name + '.res']), name


@pytest.mark.parametrize('factor', [0, 16, 3])
def test_input_stage_full_size_vs_oracle(factor):
    from dmcnet_b200.input_stage import U8InputStage
    from oracle import input_pipe as P
    frames = P.synthetic_frames(6, 224, 224, seed=factor)
    o_flow, o_mv, o_res = P.sample_from_frames(list(frames), factor)
    stage = U8InputStage(6, 224, 224, flow_ds_factor=factor)
    host = torch.from_numpy(frames).reshape(2, 3, 224, 224, 7).pin_memory()      # [B,S,H,W,7]
    flow, mv, res = stage(host)
    torch.cuda.synchronize()
    assert torch.equal(flow.cpu(), o_flow) and torch.equal(mv.cpu(), o_mv) and torch.equal(res.cpu(), o_res)
    assert stage.h2d_bytes == 6 * 224 * 224 * 7


def test_step_from_uint8_stack_equals_step_from_float_tensors():
    from oracle import input_pipe as P
    batch, num_class = 2, 51
    frames = P.synthetic_frames(batch * 3, 224, 224, seed=4)
    flow, mv, res = P.sample_from_frames(list(frames), 16)
    target = torch.tensor([3, 40])
    sd = O.build_state(num_class, None, seed=1)
    out = []
    for use_u8 in (False, True):
        eng = DmcEngine(num_class, 3, batch * 3)
        eng.load_state(sd)
        tr = FusedTrainStep(eng, HParams(), batch)
        if use_u8:
            m = tr.step_u8(torch.from_numpy(frames).reshape(batch, 3, 224, 224, 7), target.cuda(),
                           flow_ds_factor=16)
        else:
            m = tr.step(flow.cuda(), mv.cuda(), res.cuda(), target.cuda())
        out.append((m, eng.gen_flow.cpu().clone(), tr.consensus.cpu().clone()))
    (ma, ga, ca), (mb, gb, cb) = out
    assert torch.equal(ga, gb)                 # identical inputs -> identical generator output
    assert rel(ca, cb) < 1e-5                  # BatchNorm channel sums use fp32 atomics: last-bit noise
    for k in ('loss_cls', 'loss_mse', 'prec1', 'prec5'):
        assert ma[k] == pytest.approx(mb[k], rel=1e-6), k
    ref = O.OracleTrainer(sd, O.HParams(), gan=False)
    mo = ref.step(flow, mv, res, target)
    for k in ('loss', 'loss_cls', 'loss_mse'):
        assert mb[k] == pytest.approx(mo[k], rel=1e-3, abs=1e-6), k


def test_video_scorer_gan_flavour_returns_validity(monkeypatch):
    """code/dmcnet_GAN/test.py: the eval-mode discriminator (Dropout2d off, BatchNorm(eps 0.8) on
    running statistics) scores the generated map of every frame."""
    from dmcnet_b200 import inference as I
    from oracle import video_protocol as V
    monkeypatch.setattr(I, 'check_crops', lambda c: c)
    num_class, segs, crops, arch_d = 51, 3, 2, 'Discriminator3'
    sd = O.build_state(num_class, arch_d, seed=1)
    g = torch.Generator().manual_seed(6)
    for k in sd:
        if k.endswith('running_mean'):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif k.endswith('running_var'):
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
    scorer = I.VideoScorer(sd, num_class, segs, crops, arch_d=arch_d)
    flow, mv, res, target = O.make_inputs(2, 3, num_class, seed=12)
    mv, res = mv.reshape(1, 6, 2, 224, 224), res.reshape(1, 6, 3, 224, 224)
    out = scorer.evaluate([(None, mv, res, target[:1])])
    (s, lab, val), = out
    r, rv = V.forward_video_gan(sd, mv, res, segs, crops, arch_d)
    assert lab == int(target[0]) and val.shape == rv.shape == (6, 2)
    np.testing.assert_allclose(s, r, rtol=1e-3, atol=1e-3 * np.abs(r).max())
    np.testing.assert_allclose(val, rv, rtol=1e-3, atol=1e-3 * np.abs(rv).max())
    # 'Accuracy adv G' is an argmax over the flattened [frames, 2] logits (GAN/test.py:125): the six
    # frames score within 1e-6 of each other here, so the index is compared on one array only
    assert I.adversarial_accuracy([(r, lab, rv)]) == 100.0 * float(np.argmax(rv))
    assert I.adversarial_accuracy(out) == 100.0 * float(np.argmax(val))


# ------------------------------------------------------------------ checkpoint / resume
def test_gan_stage_warm_starts_from_a_stage1_checkpoint():
    """--weights (code/dmcnet_GAN/train.py:64-68): strict=False after stripping the prefix; the
    discriminator keeps its initialisation."""
    from dmcnet_b200 import checkpoint as C
    stage1 = O.build_state(51, None, seed=3)
    gan_sd = O.build_state(51, 'Discriminator', seed=1)
    eng = DmcEngine(51, 3, 3, gan=True, arch_d='Discriminator')
    eng.load_state(gan_sd)
    tr = FusedTrainStep(eng, HParams(), 1)
    missing, unexpected = tr.warm_start(C.add_module_prefix(stage1))
    assert unexpected == [] and missing and all(k.startswith('discriminator') for k in missing)
    now = eng.state_dict()
    for k in now:
        want = gan_sd[k] if k.startswith('discriminator') else stage1[k]
        assert torch.equal(now[k].cpu(), want.to(now[k].dtype)), k


# ------------------------------------------------------------------ validate() and the epoch driver
@pytest.mark.parametrize('arch_d,loss_mse', [(None, 'MSELoss'), (None, 'SmoothL1Loss'), ('Discriminator', 'MSELoss')])
def test_validate_batch_vs_oracle(arch_d, loss_mse):
    """validate() of code/dmcnet/train.py:296-347 / code/dmcnet_GAN/train.py:403-459: eval-mode batch,
    nothing is updated."""
    gan = arch_d is not None
    batch, num_class = 2, 51
    sd = O.build_state(num_class, arch_d, seed=1)
    g = torch.Generator().manual_seed(8)
    for k in sd:
        if k.endswith('running_mean'):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif k.endswith('running_var'):
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
    flow, mv, res, target = O.make_inputs(batch, 3, num_class, seed=3)
    mo = O.validate_batch(sd, O.HParams(loss_mse=loss_mse), flow, mv, res, target, gan=gan, arch_d=arch_d)
    eng = DmcEngine(num_class, 3, batch * 3, gan=gan, arch_d=arch_d)
    eng.load_state(sd)
    tr = FusedTrainStep(eng, HParams(loss_mse=loss_mse), batch)
    before = eng.state_dict()
    mg = tr.validate_batch(flow.cuda(), mv.cuda(), res.cuda(), target.cuda())
    assert set(mg) == set(mo)
    for k in mo:
        if k in ('prec1', 'prec5', 'acc_adv'):
            assert mg[k] == pytest.approx(mo[k], abs=1e-9), k
        else:
            assert mg[k] == pytest.approx(mo[k], rel=1e-3, abs=1e-6), k
    after = eng.state_dict()
    for k in before:
        assert torch.equal(before[k], after[k]), k
    assert tr.steps.cpu().tolist() == [0, 0, 0]


def test_epoch_driver_on_the_engine(tmp_path):
    """fit(): two epochs of two batches, validation, checkpoint in the reference format, resume."""
    from dmcnet_b200 import loop as L
    from dmcnet_b200 import checkpoint as C
    batch, num_class = 1, 51
    sd = O.build_state(num_class, None, seed=1)
    data = [O.make_inputs(batch, 3, num_class, seed=s) for s in (0, 1)]
    eng = DmcEngine(num_class, 3, batch * 3)
    eng.load_state(sd)
    tr = FusedTrainStep(eng, HParams(lr_steps=(1,)), batch)
    lines = []
    best = L.fit(tr, data, data[:1], epochs=2, eval_freq=1, epoch_thre=1, model_prefix=str(tmp_path / 'x'),
                 log=lines.append)
    assert tr.steps.cpu().tolist() == [2, 4, 0]          # epoch 0 frozen: only the generator stepped
    assert 0.0 <= best <= 100.0 and any(l.startswith('Testing Results') for l in lines)
    ck = C.load_checkpoint(C.checkpoint_names(str(tmp_path / 'x'), 'mv')[0])
    if ck['epoch'] == 1:       # saved after epoch 0 (0 % SAVE_FREQ == 0); epoch 1 did not improve Prec@1
        assert float(ck['optimizer_gf']['state'][0]['step']) == 2.0
        assert ck['optimizer_cls']['state'] == {} and ck['optimizer_cls']['param_groups'][0]['lr'] == 0.0
    else:                      # epoch 1 was the best so far and overwrote the file
        assert ck['epoch'] == 2 and ck['best_prec1'] == best
        assert float(ck['optimizer_gf']['state'][0]['step']) == 4.0
        assert float(ck['optimizer_cls']['state'][0]['step']) == 2.0
    # epoch 1 ran at the decayed rate (lr_steps=(1,)): lr * lr_decay * lr_mse_mult
    assert float(tr.hyper[len(C.group_keys(eng.specs, 'base_model')), 0]) == pytest.approx(0.01 * 0.1)


# ------------------------------------------------------------------ checkpoint / resume interop (GPU <-> torch.optim.Adam)
def test_checkpoint_resumes_on_the_oracle_and_back(tmp_path):
    """One step here -> checkpoint in the reference's format (code/dmcnet/train.py:190-201) ->
    torch optimizers wired as train.py:121-142 load it and take step 2; the engine resumed from the
    same file takes step 2 as well.  Then the other direction: an oracle-written checkpoint."""
    from dmcnet_b200 import checkpoint as C
    batch, num_class = 2, 51
    sd = O.build_state(num_class, None, seed=1)
    flow, mv, res, target = O.make_inputs(batch, 3, num_class, seed=0)
    dev = lambda *ts: [t.cuda() for t in ts]

    def engine_trainer(state=None):
        eng = DmcEngine(num_class, 3, batch * 3)
        eng.load_state(state if state is not None else sd)
        return eng, FusedTrainStep(eng, HParams(), batch)

    eng, tr = engine_trainer()
    tr.step(*dev(flow, mv, res, target))
    ck = tr.checkpoint(epoch=1, best_prec1=3.0)
    path = C.save_checkpoint(ck, False, str(tmp_path / 'hmdb51'), 'mv')
    ck = C.load_checkpoint(path)
    assert set(ck) == {'epoch', 'arch', 'state_dict', 'best_prec1', 'optimizer_cls', 'optimizer_gf'}
    assert all(k.startswith('module.') for k in ck['state_dict'])
    assert len(ck['optimizer_cls']['param_groups']) == 62 and len(ck['optimizer_gf']['param_groups']) == 12
    assert float(ck['optimizer_gf']['state'][0]['step']) == 1.0

    # (1) the oracle (torch.optim.Adam) resumes from our file.  torch's load_state_dict keeps the
    # tensors it is given (no copy) and Adam then updates them in place, so the oracle gets its own
    # read of the file
    ck_for_oracle = C.load_checkpoint(path)
    ref = O.OracleTrainer(C.strip_first_component(ck_for_oracle['state_dict']), O.HParams(), gan=False)
    ref.opt_cls.load_state_dict(ck_for_oracle['optimizer_cls'])
    ref.opt_gf.load_state_dict(ck_for_oracle['optimizer_gf'])
    mo = ref.step(flow, mv, res, target)
    assert float(ck['optimizer_gf']['state'][0]['step']) == 1.0          # our copy is untouched
    # (2) a fresh engine resumes from the same file
    eng2, tr2 = engine_trainer(state={k: torch.zeros_like(v) if v.is_floating_point() else v for k, v in sd.items()})
    assert tr2.resume(ck) == (1, 3.0)
    assert tr2.steps.cpu().tolist() == [1, 1, 0]
    m2 = tr2.step(*dev(flow, mv, res, target))
    # (3) the original engine simply continues
    m1 = tr.step(*dev(flow, mv, res, target))
    for k in ('loss', 'loss_cls', 'loss_mse'):
        assert m2[k] == pytest.approx(m1[k], rel=1e-5), k
        assert m2[k] == pytest.approx(mo[k], rel=1e-3, abs=1e-6), k
    after_o = ref.state_dict()
    a1, a2 = eng.state_dict(), eng2.state_dict()
    for k in a1:
        if k.startswith('gen_flow_model'):     # same state, same batch: only fp32-atomic ordering differs
            assert rel(a2[k], a1[k]) < 1e-4, k
            assert rel(a2[k], after_o[k]) < 1e-3, k
    st = ref.opt_gf.state_dict()['state'][0]
    ck2 = tr2.checkpoint(epoch=2)
    assert float(ck2['optimizer_gf']['state'][0]['step']) == float(st['step']) == 2.0
    assert rel(ck2['optimizer_gf']['state'][0]['exp_avg'], st['exp_avg']) < 1e-3

    # (4) oracle-written checkpoint (the reference's own layout) -> engine
    ck_o = {'epoch': 2, 'arch': 'resnet18', 'state_dict': C.add_module_prefix(after_o), 'best_prec1': 0.0,
            'optimizer_cls': ref.opt_cls.state_dict(), 'optimizer_gf': ref.opt_gf.state_dict()}
    eng3, tr3 = engine_trainer()
    tr3.resume(ck_o)
    assert tr3.steps.cpu().tolist()[:2] == [2, 2]
    m3 = tr3.step(*dev(flow, mv, res, target))
    mo3 = ref.step(flow, mv, res, target)
    for k in ('loss', 'loss_cls', 'loss_mse'):
        assert m3[k] == pytest.approx(mo3[k], rel=1e-3, abs=1e-6), k
    o3, e3 = ref.state_dict(), eng3.state_dict()
    for k in e3:
        if k.startswith('gen_flow_model'):
            assert rel(e3[k], o3[k]) < 1e-3, k



# ------------------------------------------------------------------ wider dense estimators (opt-in)
# The engine's generator code and kernels take the growth table as a parameter, so
# EstimatorDenseNetSmall / EstimatorDenseNet run on the DenseNetTiny kernels.  Gradient bar: 1e-4
# (measured 1e-5 .. 6e-5), 3e-4 for the 128-channel EstimatorDenseNet whose 3501-term fp32 sums
# differ from torch's summation order by 1.03e-4 on the B200 (profiles/r02_ext_tests_gpu.log).
import os as _os


@pytest.mark.parametrize('arch,batch', [('DenseNetSmall', 2), ('DenseNet', 1)])
def test_wider_dense_estimators_train_step_vs_oracle(arch, batch):
    num_class = 51
    sd = O.build_state(num_class, None, seed=1, arch_estimator=arch)
    flow, mv, res, target = O.make_inputs(batch, 3, num_class, seed=0)
    ref = O.OracleTrainer(sd, O.HParams(), gan=False)
    eng = DmcEngine(num_class, 3, batch * 3, gen_growth=O.DENSE_GROWTH[arch])
    eng.load_state(sd)
    assert list(eng.state_keys()) == list(sd.keys())
    tr = FusedTrainStep(eng, HParams(), batch)
    mo = ref.step(flow, mv, res, target)
    mg = tr.step(flow.cuda(), mv.cuda(), res.cuda(), target.cuda())
    for k in ('loss', 'loss_cls', 'loss_mse'):
        assert mg[k] == pytest.approx(mo[k], rel=1e-3, abs=1e-6), k
    assert rel(eng.gen_flow, ref.last_gen_flow) < 1e-3
    assert torch.equal(tr.consensus.argmax(1).cpu(), ref.last_output.argmax(1))
    og = ref.grads()
    for k in eng.specs:
        if k.startswith('gen_flow_model'):
            assert rel2(eng.grad_view(k), og[k]) < (3e-4 if arch == 'DenseNet' else 1e-4), k


def test_pipelined_uint8_step_matches_blocking_step():
    """step_pipelined_u8 (copy stream moves the uint8 stack, kernels normalise in front of the step)
    against step_u8 on the same batches."""
    from oracle import input_pipe as P
    batch, num_class = 2, 51
    sd = O.build_state(num_class, None, seed=1)
    stacks = [torch.from_numpy(P.synthetic_frames(batch * 3, 224, 224, seed=s)).reshape(batch, 3, 224, 224, 7)
              .pin_memory() for s in (1, 2, 3)]
    target = torch.tensor([5, 17]).pin_memory()
    runs = []
    for pipelined in (False, True):
        eng = DmcEngine(num_class, 3, batch * 3)
        eng.load_state(sd)
        tr = FusedTrainStep(eng, HParams(), batch, pipelined=pipelined)
        ms = []
        for st in stacks:
            if pipelined:
                m = tr.step_pipelined_u8(st, target, flow_ds_factor=16)
                if m:
                    ms.append(m)
            else:
                ms.append(tr.step_u8(st, target.cuda(), flow_ds_factor=16))
        if pipelined:
            ms.append(tr.flush())
        runs.append(ms)
    assert len(runs[0]) == len(runs[1]) == 3
    for a, b in zip(*runs):
        for k in ('loss', 'loss_cls', 'loss_mse', 'prec1'):
            assert b[k] == pytest.approx(a[k], rel=1e-4, abs=1e-6), k


def test_input_stage_flip_bit_exact_vs_reference_golden(golden_dir):
    """Device-side GroupRandomHorizontalFlip (code/dmcnet/transforms.py:47-58) + sample arithmetic against
    the reference's outputs (fixture cases *_flip, values 256 included), and a mixed batch against the oracle."""
    from dmcnet_b200.input_stage import U8InputStage
    from oracle import input_pipe as P
    z = np.load(_os.path.join(golden_dir, 'input_pipe.npz'))
    n_cases = 0
    for name in sorted({k.split('.')[0] for k in z.files}):
        if not bool(z[name + '.flip']):
            continue
        frames, factor = z[name + '.frames'], int(z[name + '.factor'])
        S, H, W, _ = frames.shape
        stage = U8InputStage(S, H, W, flow_ds_factor=factor)
        flow, mv, res = stage(torch.from_numpy(frames), flip=[True])         # one decision for the clip
        torch.cuda.synchronize()
        assert np.array_equal(flow.cpu().numpy(), z[name + '.flow']), name
        assert np.array_equal(mv.cpu().numpy(), z[name + '.mv']), name
        assert np.array_equal(res.cpu().numpy(), z[name + '.res']), name
        n_cases += 1
    assert n_cases == 3
    # 2 clips x 3 segments at 224 x 224, only the second clip flipped, block-mean flow target
    frames = P.synthetic_frames(6, 224, 224, seed=9)
    a = P.sample_from_frames(list(frames[:3]), 16)
    b = P.sample_from_frames(P.flip_group(list(frames[3:])), 16)
    want = [torch.cat((x, y)) for x, y in zip(a, b)]
    got = U8InputStage(6, 224, 224, flow_ds_factor=16)(torch.from_numpy(frames), flip=[False, True])
    torch.cuda.synchronize()
    assert all(torch.equal(g.cpu(), w) for g, w in zip(got, want))


def test_crop_resize_flip_normalise_chain_bit_exact_vs_oracle():
    """Decoded uint8 frames -> GroupMultiScaleCrop -> GroupRandomHorizontalFlip -> CoviarDataSet sample
    arithmetic, all on the device (crop_resize_u8 + the flip entry points), against the oracle chain
    (pinned on the reference's transforms.py / dataset.py and the installed cv2)."""
    import random
    from dmcnet_b200 import input_stage as S
    from oracle import input_pipe as P
    rng = np.random.default_rng(2)
    clips, segs, Hs, Ws = 2, 3, 256, 340
    src = rng.integers(0, 256, (clips * segs, Hs, Ws, 7), dtype=np.uint8)
    src[0, :8, :8, 0] = 0                                   # 0 -> 256 under the flip
    random.seed(5)
    crops = [S.sample_multi_scale_crop(Hs, Ws) for _ in range(clips)]
    flips = [True, False]
    tabs = np.stack([S.crop_tables(r0, c0, rows, cols, 224, 224, Hs, Ws) for r0, c0, rows, cols in crops])
    stage = S.CropResizeStage(clips * segs, Hs, Ws)
    cropped = stage(torch.from_numpy(src), tabs)
    want_u8, want = [], []
    for ci, ((r0, c0, rows, cols), fl) in enumerate(zip(crops, flips)):
        group = P.multi_scale_crop(list(src[ci * segs:(ci + 1) * segs]), rows, cols, r0, c0)
        want_u8.append(np.stack(group))
        want.append(P.sample_from_frames(P.flip_group(group) if fl else group, 16))
    torch.cuda.synchronize()
    assert np.array_equal(cropped.cpu().numpy(), np.concatenate(want_u8))
    flow, mv, res = S.U8InputStage(clips * segs, 224, 224, flow_ds_factor=16)(cropped, flip=flips)
    torch.cuda.synchronize()
    for got, idx in ((flow, 0), (mv, 1), (res, 2)):
        assert torch.equal(got.cpu(), torch.cat([w[idx] for w in want]))
    # validation transform: GroupScale(256) + GroupCenterCrop(224) as one table
    tab = S.scaled_crop_tables(Hs, Ws, 256, 256, 16, 16, 224, 224)
    val = stage(torch.from_numpy(src), tab[None])
    torch.cuda.synchronize()
    want_val = np.concatenate([np.stack(P.scale_center_crop(list(src[i:i + 1]))) for i in range(clips * segs)])
    assert np.array_equal(val.cpu().numpy(), want_val)


def test_video_scorer_from_decoded_uint8_frames_ten_crops():
    """test.py with --test-crops 10 from the decoded stacks: GroupOverSample on the device, then the
    eval forward and the mean over 3 segments x 10 crops, against the oracle chain."""
    from dmcnet_b200 import inference as I
    from dmcnet_b200 import input_stage as S
    from oracle import input_pipe as P, video_protocol as V
    num_class, segs = 51, 3
    sd = O.build_state(num_class, None, seed=1)
    rng = np.random.default_rng(3)
    decoded = rng.integers(0, 256, (segs, 256, 340, 7), dtype=np.uint8)
    scorer = I.VideoScorer(sd, num_class, segs, 10)
    s = scorer.forward_video_u8(torch.from_numpy(decoded), label=4)
    # oracle: GroupScale(256) -> five windows -> [crop, flipped crop] -> sample arithmetic
    scaled = P.resize_group(list(decoded), (256, 256))
    frames = []
    for r0, c0 in S.oversample_offsets(256, 256, 224, 224):
        crops = [img[r0:r0 + 224, c0:c0 + 224] for img in scaled]
        frames += [P.sample_from_frames(crops, 0), P.sample_from_frames(P.flip_group(crops), 0)]
    mv = torch.cat([f[1] for f in frames]).unsqueeze(0)
    res = torch.cat([f[2] for f in frames]).unsqueeze(0)
    r = V.forward_video(sd, mv, res, segs, 10)
    np.testing.assert_allclose(s, r, rtol=1e-3, atol=1e-3 * np.abs(r).max())
    assert int(s.argmax()) == int(r.argmax())


# ------------------------------------------------------------------ the remaining generator choices
@pytest.mark.parametrize('arch,ds', [('DenseNetTinyEarlyFusionSum', 0), ('DenseNetTinyEarlyFusionStack', 0),
                                     ('DenseNetTiny', 4), ('DenseNetTinyEarlyFusionStack', 4)])
def test_early_fusion_and_downsampled_generators_train_step_vs_oracle(arch, ds):
    """EstimatorDenseNetTinyEarlyFusionSum / ...Stack (code/dmcnet/model.py:197-250) and
    --gen_flow_ds_factor (AvgPool2d in, f x f tiling out; model.py:326-327, :335-337, :347-348) on the
    dense-generator kernels: forward, losses, generator gradients (1e-4) and post-Adam state vs the oracle."""
    num_class, batch = 51, 2
    sd = O.build_state(num_class, None, seed=1, arch_estimator=arch, gen_flow_ds_factor=ds)
    flow, mv, res, target = O.make_inputs(batch, 3, num_class, seed=0)
    ref = O.OracleTrainer(sd, O.HParams(), gan=False, arch_estimator=arch, gen_flow_ds_factor=ds)
    eng = DmcEngine(num_class, 3, batch * 3, arch_estimator=arch, gen_flow_ds_factor=ds)
    eng.load_state(sd)
    assert list(eng.state_keys()) == list(sd.keys())
    tr = FusedTrainStep(eng, HParams(), batch)
    for it in range(2):
        mo = ref.step(flow, mv, res, target)
        mg = tr.step(flow.cuda(), mv.cuda(), res.cuda(), target.cuda())
        for k in ('loss', 'loss_cls', 'loss_mse'):
            assert mg[k] == pytest.approx(mo[k], rel=1e-3, abs=1e-6), (it, k)
        assert rel(eng.gen_flow, ref.last_gen_flow) < 1e-3
        assert torch.equal(tr.consensus.argmax(1).cpu(), ref.last_output.argmax(1))
        if it == 0:
            og = ref.grads()
            for k in eng.specs:
                if k.startswith('gen_flow_model'):
                    assert rel2(eng.grad_view(k), og[k]) < 1e-4, k
    osd, gsd = ref.state_dict(), eng.state_dict()
    for k in osd:
        if k.startswith('gen_flow_model'):
            assert rel(gsd[k], osd[k]) < 1e-3, k


def test_gan_step_with_downsampled_generator_routes_gradients_through_the_tiling():
    """GAN G-step with gen_flow_ds_factor = 4: classifier and discriminator gradients arrive at frame
    resolution and are folded back through the tiling."""
    arch_d, ds = 'Discriminator', 4
    sd = O.build_state(51, arch_d, seed=1, gen_flow_ds_factor=ds)
    flow, mv, res, target = O.make_inputs(1, 3, 51, seed=0)
    ref = O.OracleTrainer(sd, O.HParams(), gan=True, arch_d=arch_d, gen_flow_ds_factor=ds)
    eng = DmcEngine(51, 3, 3, gan=True, arch_d=arch_d, gen_flow_ds_factor=ds)
    eng.load_state(sd)
    tr = FusedTrainStep(eng, HParams(), 1)
    for it in range(2):
        torch.manual_seed(100 + it)
        masks = O.draw_dropout_masks(arch_d, 3 * (2 if it == 0 else 1))
        mo = ref.step(flow, mv, res, target, masks=masks, apply=False)
        mg = tr.step(flow.cuda(), mv.cuda(), res.cuda(), target.cuda(), masks=masks, apply=False)
        for k in ('loss', 'loss_cls', 'loss_adv'):
            assert mg[k] == pytest.approx(mo[k], rel=1e-3, abs=1e-6), (it, k)
    og = ref.grads()
    errs = [rel2(eng.grad_view(k), og[k]) for k in eng.specs if k.startswith('gen_flow_model')]
    assert float(np.median(errs)) < 5e-2 and max(errs) < 1.2e-1, errs


# ------------------------------------------------------------------ eval-mode BatchNorm folding
@pytest.mark.parametrize('arch', ['DenseNetTiny', 'ContextNetwork'])
def test_eval_forward_with_folded_batchnorm_matches_oracle_and_unfolded_path(arch):
    """validate() / test.py forwards fold BatchNorm into the GEMM operands (W' = W * scale, shift and the
    residual added in the epilogue, activation applied there, hi/lo written directly):
    same outputs as the oracle's eval-mode forward and as the unfolded kernels, argmax exact."""
    num_class, n = 51, 6
    sd = O.build_state(num_class, None, seed=1, arch_estimator=arch)
    g = torch.Generator().manual_seed(9)
    for k in sd:                                   # non-trivial running statistics and affine parameters
        if k.endswith('running_mean'):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.2
        elif k.endswith('running_var'):
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
        elif k.endswith(('bn1.weight', 'bn2.weight', 'downsample.1.weight')) or ('.1.weight' in k and 'context' in k):
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
        elif k.endswith(('bn1.bias', 'bn2.bias', 'downsample.1.bias')) or ('.1.bias' in k and 'context' in k):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.2
    flow, mv, res, target = O.make_inputs(2, 3, num_class, seed=0)
    st = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        ref_logits, ref_gen = O.model_forward(st, mv, res, train=False, arch_estimator=arch)
    outs = {}
    for fold in (True, False):
        eng = DmcEngine(num_class, 3, n, arch_estimator=arch)
        eng.fold_bn = fold
        eng.load_state(sd)
        before = eng.state_dict()
        logits, gen_flow = eng.forward(mv.cuda(), res.cuda(), train=False)
        outs[fold] = (logits.clone(), gen_flow.clone())
        assert rel(logits, ref_logits) < 1e-3 and rel(gen_flow, ref_gen) < 1e-3, fold
        assert torch.equal(logits.view(2, 3, num_class).mean(1).argmax(1).cpu(),
                           ref_logits.view(2, 3, num_class).mean(1).argmax(1))
        after = eng.state_dict()
        assert all(torch.equal(before[k], after[k]) for k in before)          # eval changes no state
    assert rel(outs[True][0], outs[False][0]) < 1e-4
    # ... and a train step after an eval forward still sees un-folded operands
    eng = DmcEngine(num_class, 3, n, arch_estimator=arch)
    eng.load_state(sd)
    eng.forward(mv.cuda(), res.cuda(), train=False)
    tr = FusedTrainStep(eng, HParams(), 2)
    ref = O.OracleTrainer(sd, O.HParams(), arch_estimator=arch)
    mo = ref.step(flow, mv, res, target, apply=False)
    mg = tr.step(flow.cuda(), mv.cuda(), res.cuda(), target.cuda(), apply=False)
    assert mg['loss'] == pytest.approx(mo['loss'], rel=1e-3)
