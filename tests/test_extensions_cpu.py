"""CPU: oracle and host logic of the path extensions (--loss-mse criteria, video-level
inference protocol, uint8 input pipeline).  No kernels are launched."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from dmcnet_b200 import trainer as T
from oracle import dmc_oracle as O


# ------------------------------------------------------------------ --loss-mse criteria
def test_flow_criterion_selection_matches_reference_strings():
    # code/dmcnet/train.py:166-172: 'MSELoss' | 'SmoothL1Loss' | 'L1'; anything else -> NameError
    assert isinstance(O.flow_criterion('MSELoss'), torch.nn.MSELoss)
    assert isinstance(O.flow_criterion('SmoothL1Loss'), torch.nn.SmoothL1Loss)
    assert isinstance(O.flow_criterion('L1'), torch.nn.L1Loss)
    with pytest.raises(NameError):
        O.flow_criterion('L1Loss')
    assert [T.flow_loss_kind(k) for k in ('MSELoss', 'SmoothL1Loss', 'L1')] == [0, 1, 2]
    with pytest.raises(NameError, match='criterion_mse'):
        T.flow_loss_kind('L1Loss')


@pytest.mark.parametrize('name,crit', [('MSELoss', F.mse_loss), ('SmoothL1Loss', F.smooth_l1_loss),
                                       ('L1', F.l1_loss)])
def test_flow_loss_kernel_formulas_against_torch(name, crit):
    """The value / slope pair dmc_flow_loss_head implements, written out in numpy, against
    the torch criterion and its autograd gradient (mean reduction, weight lr_mse)."""
    g = torch.Generator().manual_seed(3)
    a = (torch.randn(4, 2, 16, 16, generator=g) * 1.5).requires_grad_(True)
    b = torch.randn(4, 2, 16, 16, generator=g)
    with torch.no_grad():
        a[0, 0, 0, :4] = b[0, 0, 0, :4]
        a[0, 0, 1, :4] = b[0, 0, 1, :4] + 1.0
    loss = crit(a, b)
    (loss * 10.0).backward()
    d = (a.detach() - b).numpy().astype(np.float32)
    kind = T.flow_loss_kind(name)
    if kind == 0:
        val, slope = d * d, 2 * d
    elif kind == 1:
        val, slope = np.where(np.abs(d) < 1, 0.5 * d * d, np.abs(d) - 0.5), np.clip(d, -1, 1)
    else:
        val, slope = np.abs(d), np.sign(d)
    scales = T.loss_grad_scales(T.HParams(lr_mse=10.0), 1, 1, 4, 16, 16)
    assert abs(val.astype(np.float64).sum() / d.size - float(loss.detach())) < 1e-6
    np.testing.assert_allclose(scales['flow'] * slope, a.grad.numpy(), rtol=1e-6, atol=1e-9)
    assert scales['mse'] == pytest.approx(2 * scales['flow'])


# ------------------------------------------------------------------ video-level protocol (host logic)
def test_launch_plan_and_crop_check():
    from dmcnet_b200 import inference as I
    assert I.plan_launches(250, None) == (250, 1)
    assert I.plan_launches(250, 250) == (250, 1)
    assert I.plan_launches(250, 128) == (125, 2)
    assert I.plan_launches(250, 60) == (50, 5)
    assert I.plan_launches(25, 7) == (5, 5)
    assert I.plan_launches(7, 3) == (1, 7)
    with pytest.raises(ValueError):
        I.plan_launches(0, None)
    assert I.check_crops(1) == 1 and I.check_crops(10) == 10
    with pytest.raises(ValueError, match='Only 1 and 10 crops are supported, but got 5'):
        I.check_crops(5)


def test_strip_module_prefix_like_test_py():
    from dmcnet_b200 import inference as I
    sd = {'module.base_model.conv1.weight': torch.zeros(1), 'module.gen_flow_model.conv_0.0.bias': torch.ones(1)}
    out = I.strip_module_prefix(sd)
    assert list(out) == ['.'.join(k.split('.')[1:]) for k in sd]          # test.py:84


def test_score_file_round_trip_and_fusion_against_oracle(tmp_path):
    from dmcnet_b200 import inference as I
    from oracle import video_protocol as V
    rng = np.random.default_rng(0)
    names = ['v_%s' % s for s in ('run', 'Archery', 'jump', 'dive', 'clap')]
    mk = lambda: [(rng.standard_normal((1, 7)).astype(np.float32), int(rng.integers(0, 7))) for _ in names]
    out_a, out_b = mk(), mk()
    out_b = [(s, la) for (s, _), (_, la) in zip(out_b, out_a)]              # same labels in both streams
    assert I.video_accuracy(out_a) == pytest.approx(V.accuracy(out_a))
    pa, pb = str(tmp_path / 'a.npz'), str(tmp_path / 'b.npz')
    ra, rb = str(tmp_path / 'ra.npz'), str(tmp_path / 'rb.npz')
    I.save_scores(pa, out_a, names); I.save_scores(pb, out_b, names)
    V.save_scores(ra, out_a, names); V.save_scores(rb, out_b, names)
    za, zr = np.load(pa, allow_pickle=True), np.load(ra, allow_pickle=True)
    assert list(za['names']) == list(zr['names']) == sorted(names)
    assert list(za['labels']) == list(zr['labels'])
    for x, y in zip(za['scores'], zr['scores']):                            # combine.py's indexing
        assert np.array_equal(x[0][0], y[0][0]) and x[1] == y[1]
    s, l, n = I.load_scores(ra)                                            # reads the reference's layout
    assert s.shape == (5, 7) and list(n) == sorted(names)
    assert I.combine_scores([pa, pb], [2.0, 1.0]) == pytest.approx(V.combine([ra, rb], [2.0, 1.0]))
    with pytest.raises(ValueError):
        I.save_scores(pa, out_a, names[:-1])


def test_video_scorer_refuses_to_run_without_a_gpu():
    from dmcnet_b200 import inference as I
    if torch.cuda.is_available():
        pytest.skip('CPU-only check')
    with pytest.raises(RuntimeError, match='CUDA'):
        I.VideoScorer({}, 51, 3, 1)


# ------------------------------------------------------------------ uint8 input stage
def _golden_cases(golden_dir, flipped=False):
    import os
    z = np.load(os.path.join(golden_dir, 'input_pipe.npz'))
    names = sorted({k.split('.')[0] for k in z.files})
    assert len(names) == 7
    for n in names:
        if bool(z[n + '.flip']) == flipped:
            yield n, z[n + '.frames'], int(z[n + '.factor']), z[n + '.flow'], z[n + '.mv'], z[n + '.res']


def test_input_oracle_pinned_against_reference_dataset():
    from oracle import ref_loader as R
    if not R.reference_available():
        pytest.skip('/root/reference not present')
    from oracle.pin_input_pipe import pin
    assert pin(write=False, verbose=False) == 2 * 7 * 3


def test_input_oracle_reproduces_reference_golden(golden_dir):
    """tests/golden/input_pipe.npz holds outputs of the reference's own dataset.py."""
    from oracle import input_pipe as P
    for flipped in (False, True):
        n_cases = 0
        for name, frames, factor, flow, mv, res in _golden_cases(golden_dir, flipped):
            group = P.flip_group(list(frames)) if flipped else list(frames)
            o_flow, o_mv, o_res = P.sample_from_frames(group, factor)
            assert np.array_equal(o_flow.numpy(), flow), name
            assert np.array_equal(o_mv.numpy(), mv), name
            assert np.array_equal(o_res.numpy(), res), name
            n_cases += 1
        assert n_cases == (3 if flipped else 4)


def _kernel_model_unpack(frames, div_motion, div_res):
    """numpy model of unpack_normalize_u8_kernel: thread g reads words 7g..7g+6 of the byte
    stream, byte b of its 28 is pixel b // 7, channel b % 7."""
    S, H, W, _ = frames.shape
    words = np.frombuffer(frames.tobytes(), dtype='<u4')
    groups = S * H * W // 4
    w = words.reshape(groups, 7)
    out = np.zeros((7, groups, 4), np.float32)
    divs = np.array([div_motion] * 4 + list(div_res), np.float32)
    for b in range(28):
        byte = (w[:, b >> 2] >> np.uint32(8 * (b & 3))) & np.uint32(0xff)
        v = byte.astype(np.float32) / np.float32(255.0)
        out[b % 7, :, b // 7] = (v - np.float32(0.5)) / divs[b % 7]
    planes = out.reshape(7, S, H * W)                      # group g = frame n, float4 p4 -> contiguous
    split = lambda lo, hi: planes[lo:hi].transpose(1, 0, 2).reshape(S, hi - lo, H, W)
    return split(0, 2), split(2, 4), split(4, 7)


def _kernel_model_block_mean(frames, f, div_motion):
    """numpy model of flow_block_mean_u8_kernel (integer block sums, one double division)."""
    S, H, W, _ = frames.shape
    out = np.zeros((S, 2, H, W), np.float32)
    for by in range(-(-H // f)):
        for bx in range(-(-W // f)):
            blk = frames[:, by * f:min(H, by * f + f), bx * f:min(W, bx * f + f), 0:2].astype(np.int64)
            mean = (blk.sum(axis=(1, 2)).astype(np.float64) / np.float64(f * f)).astype(np.float32)   # [S,2]
            val = (mean / np.float32(255.0) - np.float32(0.5)) / np.float32(div_motion)
            out[:, :, by * f:by * f + f, bx * f:bx * f + f] = val[:, :, None, None]
    return out


def _kernel_model_unpack_flip(frames, div_motion, div_res):
    """numpy model of unpack_normalize_flip_u8_kernel with every frame flipped: output group cg of a
    row reads source group w4-1-cg, pixel order inside the group reversed, channels 0 and 2 -> 256 - v."""
    S, H, W, _ = frames.shape
    w4 = W // 4
    words = np.frombuffer(frames.tobytes(), dtype='<u4').reshape(S, H, w4, 7)
    src = words[:, :, ::-1, :]                                   # src = row*w4 + (w4-1-cg)
    out = np.zeros((7, S, H, w4, 4), np.float32)
    divs = np.array([div_motion] * 4 + list(div_res), np.float32)
    for b in range(28):
        byte = ((src[..., b >> 2] >> np.uint32(8 * (b & 3))) & np.uint32(0xff)).astype(np.float32)
        c, p = b % 7, b // 7
        v = np.float32(256.0) - byte if c in (0, 2) else byte
        out[c, ..., 3 - p] = (v / np.float32(255.0) - np.float32(0.5)) / divs[c]
    planes = out.reshape(7, S, H, W)
    split = lambda lo, hi: planes[lo:hi].transpose(1, 0, 2, 3)
    return split(0, 2), split(2, 4), split(4, 7)


def _kernel_model_block_mean_flip(frames, f, div_motion):
    """numpy model of flow_block_mean_flip_u8_kernel with every frame flipped."""
    S, H, W, _ = frames.shape
    nbx = W // f
    out = np.zeros((S, 2, H, W), np.float32)
    for by in range(-(-H // f)):
        rows = min(f, H - by * f)
        for bx in range(nbx):
            sb = nbx - 1 - bx
            blk = frames[:, by * f:by * f + rows, sb * f:sb * f + f, 0:2].astype(np.int64)
            s = blk.sum(axis=(1, 2))
            s[:, 0] = 256 * rows * f - s[:, 0]
            mean = (s.astype(np.float64) / np.float64(f * f)).astype(np.float32)
            val = (mean / np.float32(255.0) - np.float32(0.5)) / np.float32(div_motion)
            out[:, :, by * f:by * f + f, bx * f:bx * f + f] = val[:, :, None, None]
    return out


def test_input_kernels_flip_model_is_bit_exact(golden_dir):
    """The flipped entry points against outputs of the reference's GroupRandomHorizontalFlip +
    dataset.py (fixture cases *_flip; they contain v = 0 -> 256)."""
    from dmcnet_b200.input_stage import normalisation_divisors
    div_motion, div_res = normalisation_divisors()
    n_cases = 0
    for name, frames, factor, flow, mv, res in _golden_cases(golden_dir, flipped=True):
        assert (frames[..., 0] == 0).any()
        k_flow, k_mv, k_res = _kernel_model_unpack_flip(frames, div_motion, div_res)
        assert np.array_equal(k_mv, mv) and np.array_equal(k_res, res), name
        if factor == 0:
            assert np.array_equal(k_flow, flow), name
            assert flow.max() > (255 / 255.0 - 0.5) / div_motion          # the 256 is really there
        else:
            assert np.array_equal(_kernel_model_block_mean_flip(frames, factor, div_motion), flow), name
        n_cases += 1
    assert n_cases == 3


def test_input_kernels_index_arithmetic_model_is_bit_exact(golden_dir):
    from dmcnet_b200.input_stage import normalisation_divisors
    div_motion, div_res = normalisation_divisors()
    std = torch.from_numpy(np.array([0.229, 0.224, 0.225]).reshape((1, 3, 1, 1))).float()   # dataset.py:111
    assert div_motion == float(torch.mean(std)) and list(div_res) == [float(v) for v in std.reshape(-1)]
    for name, frames, factor, flow, mv, res in _golden_cases(golden_dir):
        k_flow, k_mv, k_res = _kernel_model_unpack(frames, div_motion, div_res)
        assert np.array_equal(k_mv, mv) and np.array_equal(k_res, res), name
        if factor == 0:
            assert np.array_equal(k_flow, flow), name
        else:
            assert np.array_equal(_kernel_model_block_mean(frames, factor, div_motion), flow), name


def test_input_stage_argument_checks():
    from dmcnet_b200 import input_stage as S
    with pytest.raises(TypeError):
        S.check_stack(torch.zeros(2, 8, 8, 7), 2, 8, 8)
    with pytest.raises(ValueError, match='trailing'):
        S.check_stack(torch.zeros(2, 7, 8, 8, dtype=torch.uint8), 2, 8, 8)
    with pytest.raises(ValueError, match='frames'):
        S.check_stack(torch.zeros(3, 8, 8, 7, dtype=torch.uint8), 2, 8, 8)
    S.check_stack(torch.zeros(1, 2, 8, 8, 7, dtype=torch.uint8), 2, 8, 8)
    with pytest.raises(NotImplementedError, match='interp'):
        S.U8InputStage(2, 8, 8, upsample_interp=True)


def test_score_reader_and_fusion_on_files_written_by_the_reference(golden_dir, tmp_path):
    """tests/golden/score_files.npz: every 30th video of the score files the reference ships
    (exp_my/hmdb51_*/split1, written by its own test.py), with the accuracy combine.py's formula
    gives (tests/golden/make_score_fixture.py)."""
    import os
    from dmcnet_b200 import inference as I
    z = np.load(os.path.join(golden_dir, 'score_files.npz'), allow_pickle=True)
    files = []
    for k in z['order']:
        p = str(tmp_path / ('%s.npz' % k))
        np.savez(p, scores=z[k + '.scores'], labels=z[k + '.labels'], names=z[k + '.names'])
        files.append(p)
    acc, n = I.combine_scores(files, list(z['weights']))
    assert n == 51 and acc == pytest.approx(float(z['accuracy_subset'][0]), abs=1e-12)
    # write -> read round trip of a reference file, both flavours (the GAN file has a validity column)
    for k in ('mv', 'dmc_gan'):
        cols = z[k + '.scores'].shape[1]
        assert cols == (3 if k == 'dmc_gan' else 2)
        names = list(z[k + '.names'])
        output = [tuple(row) for row in z[k + '.scores']]
        p = str(tmp_path / ('again_%s.npz' % k))
        I.save_scores(p, output[::-1], names[::-1])                      # any input order
        w = np.load(p, allow_pickle=True)
        assert list(w['names']) == sorted(names) and w['scores'].shape == (51, cols)
        for a, b in zip(w['scores'], z[k + '.scores'][np.argsort(names, kind='stable')]):
            assert np.array_equal(a[0], b[0]) and a[1] == b[1]
            if cols == 3:
                assert np.array_equal(a[2], b[2]) and a[2].shape == (25, 2)
        assert I.video_accuracy(output) == pytest.approx(
            100.0 * np.mean([int(np.argmax(o[0])) == int(o[1]) for o in output]))
    adv = I.adversarial_accuracy([tuple(row) for row in z['dmc_gan.scores']])
    assert adv == pytest.approx(100.0 * np.mean([np.argmax(r[2]) for r in z['dmc_gan.scores']]))


def test_full_reference_score_files_reproduce_the_published_fusion_accuracy():
    import os
    from dmcnet_b200 import inference as I
    root = '/root/reference/exp_my'
    if not os.path.isdir(root):
        pytest.skip('/root/reference not present')
    expect = {1: 0.6405, 2: 0.6131, 3: 0.6007}
    for split, want in expect.items():
        files = [root + '/hmdb51_coviar/iframe/split%d/iframe_score_model_best.npz' % split,
                 root + '/hmdb51_coviar/mv/split%d/mv_score_model_best.npz' % split,
                 root + '/hmdb51_coviar/residual/split%d/residual_score_model_best.npz' % split,
                 root + '/hmdb51_gan/split%d/mv_score_model_best.npz' % split]
        acc, n = I.combine_scores(files, [2.0, 1.0, 1.0, 1.0])
        assert n == 1530 and acc == pytest.approx(want, abs=5e-5)


# ------------------------------------------------------------------ checkpoint / resume format
class _Bucket:
    """CPU stand-in for the engine's flat parameter table (specs / offsets / moment buckets)."""

    def __init__(self, state):
        from collections import OrderedDict
        self.specs = OrderedDict((k, tuple(v.shape)) for k, v in state.items() if not O.is_buffer(k))
        self.offsets, off = {}, 0
        for tag in ('base_model', 'gen_flow_model', 'discriminator'):
            for k, shp in self.specs.items():
                if k.startswith(tag):
                    self.offsets[k] = off
                    off += (int(np.prod(shp)) + 63) // 64 * 64
        self.exp_avg, self.exp_avg_sq = torch.zeros(off), torch.zeros(off)


def test_optimizer_state_round_trips_through_the_torch_adam_format():
    """The reference checkpoints torch.optim.Adam.state_dict() of one-group-per-tensor optimizers
    (code/dmcnet_GAN/train.py:122-153, :203-215).  Oracle optimizers after real steps -> flat
    buckets -> torch format again: identical, and loadable by a fresh torch Adam."""
    from dmcnet_b200 import checkpoint as C
    arch_d = 'Discriminator'
    sd = O.build_state(51, arch_d, seed=1)
    flow, mv, res, target = O.make_inputs(1, 3, 51, seed=0)
    ref = O.OracleTrainer(sd, O.HParams(), gan=True, arch_d=arch_d)
    for it in range(2):                                              # D-step then G-step
        torch.manual_seed(100 + it)
        ref.step(flow, mv, res, target, masks=O.draw_dropout_masks(arch_d, 3 * (2 if it % 2 == 0 else 1)))
    bucket = _Bucket(sd)
    hp = O.HParams()
    for tag, opt, mult in (('base_model', ref.opt_cls, hp.lr_cls_mult), ('gen_flow_model', ref.opt_gf, hp.lr_mse_mult),
                           ('discriminator', ref.opt_d, hp.lr_d_mult)):
        want = opt.state_dict()
        step = C.adam_state_from_torch(bucket, tag, want)
        assert step == 1                                             # each optimizer stepped once
        keys = C.group_keys(bucket.specs, tag)
        assert len(keys) == len(want['param_groups']) == {'base_model': 62, 'gen_flow_model': 12,
                                                          'discriminator': 16}[tag]
        rows = {k: (g['lr'], g['weight_decay']) for k, g in zip(keys, want['param_groups'])}
        got = C.adam_state_to_torch(bucket, tag, step, rows, mult, hp.betas, hp.eps)
        assert sorted(got['state']) == sorted(want['state'])
        for i in want['state']:
            assert float(got['state'][i]['step']) == float(want['state'][i]['step'])
            assert torch.equal(got['state'][i]['exp_avg'], want['state'][i]['exp_avg'])
            assert torch.equal(got['state'][i]['exp_avg_sq'], want['state'][i]['exp_avg_sq'])
        for g, w in zip(got['param_groups'], want['param_groups']):
            for f in ('lr', 'betas', 'eps', 'weight_decay', 'amsgrad', 'lr_mult', 'decay_mult', 'params'):
                assert g[f] == w[f] or tuple(g[f]) == tuple(w[f]), (tag, f)
        # a fresh optimizer wired as the reference's accepts the produced dict and reproduces it
        params = [{'params': torch.zeros(bucket.specs[k], requires_grad=True), 'lr': hp.lr, 'lr_mult': mult,
                   'decay_mult': 0.0 if 'bias' in k else 1.0} for k in keys]
        fresh = torch.optim.Adam(params, weight_decay=hp.weight_decay, eps=hp.eps)
        fresh.load_state_dict(got)
        again = fresh.state_dict()
        assert torch.equal(again['state'][0]['exp_avg'], want['state'][0]['exp_avg'])
        assert [g['lr'] for g in again['param_groups']] == [g['lr'] for g in want['param_groups']]
    # a never-stepped optimizer has no state entries
    assert C.adam_state_to_torch(bucket, 'gen_flow_model', 0, rows_all(bucket), 1.0)['state'] == {}
    with pytest.raises(ValueError, match='parameters'):
        C.adam_state_from_torch(bucket, 'gen_flow_model', ref.opt_cls.state_dict())


def rows_all(bucket):
    return {k: (0.01, 1e-4) for k in bucket.specs}


def test_checkpoint_file_names_prefix_handling_and_warm_start(tmp_path):
    from dmcnet_b200 import checkpoint as C
    assert C.checkpoint_names('exp/run1/hmdb51', 'MV') == ('exp/run1/hmdb51_mv_checkpoint.pth.tar',
                                                          'exp/run1/hmdb51_mv_model_best.pth.tar')   # train.py:372-377
    sd = O.build_state(51, None, seed=1)
    wrapped = C.add_module_prefix(sd)
    assert all(k.startswith('module.') for k in wrapped)
    assert list(C.strip_first_component(wrapped)) == list(sd)
    state = {'epoch': 3, 'arch': 'resnet18', 'state_dict': wrapped, 'best_prec1': 12.5}
    path = C.save_checkpoint(state, True, str(tmp_path / 'm'), 'mv')
    best = C.checkpoint_names(str(tmp_path / 'm'), 'mv')[1]
    a, b = C.load_checkpoint(path), C.load_checkpoint(best)
    assert a['epoch'] == b['epoch'] == 3 and torch.equal(a['state_dict']['module.base_model.fc.bias'],
                                                         sd['base_model.fc.bias'])
    # --weights: stage-2 (GAN) model warm-started from a stage-1 checkpoint (no discriminator keys)
    gan = O.build_state(51, 'Discriminator3', seed=2)
    merged, missing, unexpected = C.merge_non_strict(gan, C.strip_first_component(wrapped))
    assert all(k.startswith('discriminator') for k in missing) and unexpected == []
    assert torch.equal(merged['base_model.conv1.weight'], sd['base_model.conv1.weight'])
    assert torch.equal(merged['discriminator.adv_layer.bias'], gan['discriminator.adv_layer.bias'])
    bad = dict(sd)
    bad['base_model.fc.weight'] = torch.zeros(101, 512)
    with pytest.raises(RuntimeError, match='size mismatch'):
        C.merge_non_strict(gan, bad)


# ------------------------------------------------------------------ epoch driver (host logic)
class _FakeStep:
    """Records what the epoch driver asks of the fused step."""

    def __init__(self, precs):
        self.calls, self.precs, self.epoch = [], list(precs), None

    def set_epoch(self, epoch, epoch_thre=0):
        self.calls.append(('set_epoch', epoch, epoch_thre))
        self.epoch = epoch

    def step(self, flow, mv, res, target):
        self.calls.append(('step', self.epoch))
        return {'loss': 2.0, 'loss_cls': 1.0, 'loss_mse': 0.1, 'prec1': 50.0, 'prec5': 100.0}

    def validate_batch(self, flow, mv, res, target):
        self.calls.append(('val', self.epoch))
        return {'loss': 1.0, 'loss_cls': 1.0, 'loss_mse': 0.0, 'prec1': self.precs[self.epoch], 'prec5': 100.0}

    def checkpoint(self, epoch, arch, best_prec1):
        self.calls.append(('ckpt', epoch, best_prec1))
        return {'epoch': epoch, 'arch': arch, 'state_dict': {}, 'best_prec1': best_prec1}


def test_epoch_driver_follows_the_reference_schedule(tmp_path):
    """main() of code/dmcnet/train.py:173-201: set lr per epoch, validate when epoch % eval_freq == 0
    or on the last epoch, save when best or epoch % SAVE_FREQ == 0, 'epoch' stored as epoch + 1."""
    import os
    from dmcnet_b200 import loop as L
    from dmcnet_b200 import checkpoint as C
    batch = (None, None, None, torch.zeros(4, dtype=torch.int64))
    step = _FakeStep(precs=[10.0, 0.0, 30.0, 0.0, 20.0, 25.0])
    lines = []
    best = L.fit(step, [batch] * 3, [batch] * 2, epochs=6, eval_freq=2, epoch_thre=1,
                 model_prefix=str(tmp_path / 'hmdb51'), representation='mv', log=lines.append)
    assert best == 30.0
    assert [c for c in step.calls if c[0] == 'set_epoch'] == [('set_epoch', e, 1) for e in range(6)]
    assert [c[1] for c in step.calls if c[0] == 'val'] == [0, 0, 2, 2, 4, 4, 5, 5]       # epochs 0, 2, 4 and the last
    # epoch 0: best (10 > 0) and 0 % 40 == 0; epoch 2: best (30); epochs 4, 5: neither
    assert [c for c in step.calls if c[0] == 'ckpt'] == [('ckpt', 1, 10.0), ('ckpt', 3, 30.0)]
    ck, best_file = C.checkpoint_names(str(tmp_path / 'hmdb51'), 'mv')
    assert C.load_checkpoint(ck)['epoch'] == 3 and C.load_checkpoint(best_file)['best_prec1'] == 30.0
    assert lines[0] == 'current epoch freeze?: True' and 'current epoch freeze?: False' in lines
    assert sum(l.startswith('Testing Results: Prec@1') for l in lines) == 4
    m = L.AverageMeter()
    m.update(2.0, 3); m.update(4.0, 1)
    assert (m.val, m.sum, m.count, m.avg) == (4.0, 10.0, 4, 2.5)                          # train.py:380-395
    # GAN: D-step and G-step metrics are averaged separately
    class _Gan(_FakeStep):
        def step(self, flow, mv, res, target):
            self.k = getattr(self, 'k', 0) + 1
            return {'loss': 1.0, 'loss_adv': 0.5} if self.k % 2 else {'loss': 3.0, 'loss_adv': 0.7, 'loss_mse': 0.2}
    avg = L.train_epoch(_Gan([]), [batch] * 4, 0, gan=True, log=lines.append)
    assert avg['D_loss'] == 1.0 and avg['G_loss'] == 3.0 and avg['G_loss_mse'] == pytest.approx(0.2) and 'D_loss_mse' not in avg


# ------------------------------------------------------------------ generator choices (constructors)
GEN_CASES = [('ContextNetwork', 0, 0), ('ContextNetwork', 1, 0), ('ContextNetwork', 0, 4), ('ContextNetwork', 1, 4),
             ('DenseNet', 0, 0), ('DenseNetSmall', 0, 0), ('DenseNetTiny', 0, 0),
             ('DenseNetTinyEarlyFusionSum', 0, 0), ('DenseNetTinyEarlyFusionStack', 0, 0)]


def _our_model(arch, att, ds, gan=False):
    import contextlib, io
    from dmcnet_b200 import model as M
    torch.manual_seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        if gan:
            return M.GANModel(51, 3, 'mv', base_model='resnet18', arch_estimator=arch, att=att,
                              gen_flow_ds_factor=ds, use_databn=0, gen_flow_or_delta=1, arch_d='Discriminator')
        return M.Model(51, 3, 'mv', base_model='resnet18', arch_estimator=arch, att=att,
                       gen_flow_ds_factor=ds, use_databn=0, gen_flow_or_delta=1)


@pytest.mark.parametrize('arch,att,ds', GEN_CASES)
def test_every_generator_choice_constructs_the_reference_state(arch, att, ds, golden_dir):
    """state_dict keys / shapes / init values of gen_flow_model for every --arch_estimator
    (code/dmcnet/model.py:310-325) against the reference constructor's (fixture:
    tests/golden/make_generator_states.py; live comparison of the WHOLE state_dict when
    /root/reference is present).  Only the dense estimators have kernels; the others construct
    (checkpoints load, optimizers can be wired) and refuse to run."""
    import os
    from oracle.digest import digest
    z = np.load(os.path.join(golden_dir, 'generator_states.npz'))
    tag = '%s.att%d.ds%d' % (arch, att, ds)
    m = _our_model(arch, att, ds)
    sd = m.state_dict()
    keys = [k for k in sd if k.startswith('gen_flow_model')]
    assert keys == list(z[tag + '.keys'])
    assert [','.join(map(str, sd[k].shape)) for k in keys] == list(z[tag + '.shapes'])
    for k, d in zip(keys, z[tag + '.digests']):
        assert np.array_equal(digest(sd[k].float()), d), k
    n_params = sum(p.numel() for n, p in m.named_parameters() if n.startswith('gen_flow_model'))
    assert n_params == int(z[tag + '.params'])
    assert hasattr(m, 'downsample') == (ds != 0)
    from oracle import ref_loader as R
    if R.reference_available():
        torch.manual_seed(1)
        ref = R.build_reference_model('dmcnet', 51, 3, 'mv', base_model='resnet18', arch_estimator=arch,
                                      att=att, gen_flow_ds_factor=ds, use_databn=0, gen_flow_or_delta=1)
        rsd = ref.state_dict()
        assert list(rsd) == list(sd)
        for k in rsd:
            assert torch.equal(rsd[k], sd[k]), k
    if arch != 'DenseNetTiny':
        with pytest.raises((NotImplementedError, RuntimeError)):
            m(torch.zeros(1, 3, 2, 224, 224), torch.zeros(1, 3, 3, 224, 224))


def test_gan_model_with_other_generators_matches_reference_and_unknown_names_stay_undefined():
    from oracle import ref_loader as R
    m = _our_model('DenseNetSmall', 0, 0, gan=True)
    if R.reference_available():
        torch.manual_seed(1)
        ref = R.build_reference_model('dmcnet_GAN', 51, 3, 'mv', base_model='resnet18', arch_estimator='DenseNetSmall',
                                      att=0, gen_flow_ds_factor=0, use_databn=0, gen_flow_or_delta=1,
                                      arch_d='Discriminator')
        rsd, sd = ref.state_dict(), m.state_dict()
        assert list(rsd) == list(sd) and all(torch.equal(rsd[k], sd[k]) for k in rsd)
    bad = _our_model('NoSuchEstimator', 0, 0)
    assert not hasattr(bad, 'gen_flow_model')                          # model.py:310-325: silently undefined
    with pytest.raises((AttributeError, NotImplementedError)):
        bad(torch.zeros(1, 3, 2, 224, 224), torch.zeros(1, 3, 3, 224, 224))


@pytest.mark.parametrize('arch', ['DenseNetSmall', 'DenseNet'])
def test_oracle_with_wider_dense_estimators_pinned_against_reference(arch):
    """EstimatorDenseNetSmall / EstimatorDenseNet (code/dmcnet/model.py:122-169) share the dense
    structure of the Tiny variant: the oracle's state and forward are bit-identical to the
    reference Model's for them too."""
    from oracle import ref_loader as R
    from dmcnet_b200 import model as M
    sd = O.build_state(51, None, seed=1, arch_estimator=arch)
    ours = M.build_state(51, None, seed=1, arch_estimator=arch)
    assert list(sd) == list(ours) and all(torch.equal(sd[k], ours[k]) for k in sd)
    if not R.reference_available():
        pytest.skip('/root/reference not present')
    torch.manual_seed(1)
    ref = R.build_reference_model('dmcnet', 51, 3, 'mv', base_model='resnet18', arch_estimator=arch,
                                  use_databn=0, gen_flow_or_delta=1)
    rsd = ref.state_dict()
    assert list(rsd) == list(sd) and all(torch.equal(rsd[k], sd[k]) for k in rsd)
    flow, mv, res, target = O.make_inputs(1, 3, 51, seed=0)
    ref.eval()
    with torch.no_grad():
        r_out, r_gen = ref(mv, res)
        o_out, o_gen = O.model_forward({k: v.clone() for k, v in sd.items()}, mv, res, train=False)
    assert torch.equal(r_out, o_out) and torch.equal(r_gen, o_gen)


def test_dropin_modules_export_every_public_name_of_the_reference_with_its_signature():
    """``from model import X`` for every class / helper of code/dmcnet/model.py and
    code/dmcnet_GAN/model.py; with /root/reference present each one is constructed with the
    reference's own arguments under the same seed and its state_dict compared bit for bit."""
    import importlib.util, inspect, os
    from oracle import ref_loader as R
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mods = {}
    for variant in ('dmcnet', 'dmcnet_GAN'):
        spec = importlib.util.spec_from_file_location('_dropin_%s' % variant,
                                                      os.path.join(root, 'dmcnet_b200', 'dropin', variant, 'model.py'))
        mods[variant] = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mods[variant])
    cases = [('ContextNetwork', (5, True, 0)), ('ContextNetwork', (5, True, 4)), ('ContextNetwork', (5, False, 0)),
             ('ContextNetworkAtt', (5, True, 0)), ('EstimatorDenseNet', (5,)), ('EstimatorDenseNetSmall', (5,)),
             ('EstimatorDenseNetTiny', (5,)), ('EstimatorDenseNetTinyEarlyFusionSum', (5,)),
             ('EstimatorDenseNetTinyEarlyFusionStack', (5,)), ('conv', (5, 8)), ('predict_flow', (33,)),
             ('conv_dilation', (True, 5, 32, 3, 1, 2)), ('conv_dilation', (False, 5, 32, 3, 1, 2))]
    gan_cases = cases + [('Discriminator%s' % n, (2,)) for n in ('', '2', '3', '4', '5')] + \
        [('discriminator_block', (2, 16, False)), ('discriminator_block', (16, 32, True)),
         ('discriminator_block2', (16, 16, True))]
    for variant, todo in (('dmcnet', cases), ('dmcnet_GAN', gan_cases)):
        ours = mods[variant]
        ref = R.load_reference_model_module(variant) if R.reference_available() else None
        assert inspect.isclass(ours.Model) and issubclass(ours.Flatten, torch.nn.Module)
        assert ours.Flatten()(torch.zeros(2, 3, 4)).shape == (2, 12)
        for name, args in todo:
            assert hasattr(ours, name), (variant, name)
            torch.manual_seed(7)
            a = getattr(ours, name)(*args)
            if ref is None:
                continue
            torch.manual_seed(7)
            b = getattr(ref, name)(*args)
            sa, sb = a.state_dict(), b.state_dict()
            assert list(sa) == list(sb), (variant, name)
            assert all(torch.equal(sa[k], sb[k]) for k in sa), (variant, name)
            assert type(a).__name__ == type(b).__name__
            pa = list(inspect.signature(getattr(ours, name)).parameters)
            pb = list(inspect.signature(getattr(ref, name)).parameters)
            assert pa == pb, (variant, name, pa, pb)
        if ref is not None:
            pa = inspect.signature(ours.Model.__init__).parameters
            pb = inspect.signature(ref.Model.__init__).parameters
            assert list(pa) == list(pb) and all(pa[k].default == pb[k].default for k in pa), variant


def test_get_augmentation_composes_the_reference_transforms():
    """Model.get_augmentation (code/dmcnet/model.py:369-378) with the reference's own transforms
    module: multi-scale crop to 224 then the random flip that negates the x components."""
    import os, random, sys
    from oracle import ref_loader as R
    if not R.reference_available():
        pytest.skip('/root/reference not present')
    d = os.path.join(R.REFERENCE_ROOT, 'code', 'dmcnet')
    saved = sys.modules.pop('transforms', None)
    sys.path.insert(0, d)
    try:
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()) as out:
            aug = _our_model('DenseNetTiny', 0, 0).get_augmentation()
        assert 'Augmentation scales: [1, 0.875, 0.75]' in out.getvalue()
        names = [type(t).__name__ for t in aug.transforms]
        assert names == ['GroupMultiScaleCrop', 'GroupRandomHorizontalFlip']
        assert aug.transforms[0].scales == [1, .875, .75] and aug.transforms[0].input_size == [224, 224]
        random.seed(3)
        rng = np.random.default_rng(0)
        group = [rng.integers(0, 256, (256, 340, 7), dtype=np.uint8) for _ in range(3)]
        res = aug(group)
        assert len(res) == 3 and all(r.shape == (224, 224, 7) for r in res)
    finally:
        sys.path.remove(d)
        sys.modules.pop('transforms', None)
        if saved is not None:
            sys.modules['transforms'] = saved


# ------------------------------------------------------------------ crop + resize (transforms.py)
def _ref_transforms():
    import importlib.util, os
    from oracle import ref_loader as R
    if not R.reference_available():
        pytest.skip('/root/reference not present')
    spec = importlib.util.spec_from_file_location(
        '_ref_transforms_for_tests', os.path.join(R.REFERENCE_ROOT, 'code', 'dmcnet', 'transforms.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _crop_resize_kernel_model(src, tab, out_h, out_w):
    """numpy model of crop_resize_u8_kernel: src [N,Hs,Ws,7] uint8, one table for all frames."""
    x0, x1, a0, a1 = (tab[i * out_w:(i + 1) * out_w] for i in range(4))
    ty = tab[4 * out_w:]
    y0, y1, b0, b1 = (ty[i * out_h:(i + 1) * out_h] for i in range(4))
    s = src.astype(np.int32)
    d0 = s[:, y0][:, :, x0] * a0[None, None, :, None] + s[:, y0][:, :, x1] * a1[None, None, :, None]
    d1 = s[:, y1][:, :, x0] * a0[None, None, :, None] + s[:, y1][:, :, x1] * a1[None, None, :, None]
    bb0, bb1 = b0[None, :, None, None], b1[None, :, None, None]
    return ((((bb0 * (d0 >> 4)) >> 16) + ((bb1 * (d1 >> 4)) >> 16) + 2) >> 2).astype(np.uint8)


def test_resize_restatement_is_bit_exact_against_installed_opencv():
    cv2 = pytest.importorskip('cv2')
    from oracle import input_pipe as P
    rng = np.random.default_rng(0)
    for (sh, sw) in [(256, 256), (192, 192), (192, 224), (256, 224), (224, 192), (340, 256), (255, 191), (193, 257),
                     (224, 224), (170, 300), (256, 340)]:
        for dsize in [(224, 224), (256, 256), (224, 256)]:
            src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
            assert np.array_equal(P.resize_linear_u8(src, dsize), cv2.resize(src, dsize, interpolation=cv2.INTER_LINEAR)), \
                ((sh, sw), dsize)
    src = rng.integers(0, 256, (192, 256, 3), dtype=np.uint8)
    assert np.array_equal(P.resize_linear_u8(src, (224, 224)), cv2.resize(src, (224, 224), interpolation=cv2.INTER_LINEAR))


def test_crop_resize_oracle_tables_and_kernel_model_against_reference_transforms():
    """GroupMultiScaleCrop (train), GroupScale + GroupCenterCrop (validation) and GroupOverSample's
    windows (10-crop test) of the reference's transforms.py, executed with the same ``random`` draws,
    against the oracle restatement and against a numpy model of the kernel driven by the product's
    host-built tables."""
    import random
    import torchvision
    T_ref = _ref_transforms()
    from oracle import input_pipe as P
    from dmcnet_b200 import input_stage as S
    rng = np.random.default_rng(1)
    group = [rng.integers(0, 256, (256, 340, 7), dtype=np.uint8) for _ in range(3)]
    src = np.stack(group)
    # train: multi-scale crop, several draws (all crop sizes 256 / 224 / 192 and distortions occur)
    seen = set()
    for seed in range(12):
        random.seed(seed)
        want = T_ref.GroupMultiScaleCrop(224, [1, .875, .75])(group)
        random.seed(seed)
        row0, col0, rows, cols = S.sample_multi_scale_crop(256, 340, (224, 224), (1, .875, .75))
        seen.add((rows, cols))
        got = P.multi_scale_crop(group, rows, cols, row0, col0)
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), seed
        tab = S.crop_tables(row0, col0, rows, cols, 224, 224, 256, 340)
        assert np.array_equal(_crop_resize_kernel_model(src, tab, 224, 224), np.stack(want)), seed
    assert len(seen) >= 4
    # validation: scale to 256 x 256 then centre crop (train.py:98-101)
    val = torchvision.transforms.Compose([T_ref.GroupScale(256), T_ref.GroupCenterCrop(224)])(group)
    assert all(np.array_equal(a, b) for a, b in zip(P.scale_center_crop(group, 256, 224), val))
    tab = S.scaled_crop_tables(256, 340, 256, 256, 16, 16, 224, 224)
    assert np.array_equal(_crop_resize_kernel_model(src, tab, 224, 224), np.stack(val))
    # 10-crop test: five windows of the scaled frame, each also flipped (odd entries; the flip is
    # the device-side flag of the input stage)
    over = T_ref.GroupOverSample(224, 256)(group)
    offsets = S.oversample_offsets(256, 256, 224, 224)
    assert len(over) == 5 * 3 * 2
    for wi, (r0, c0) in enumerate(offsets):
        tab = S.scaled_crop_tables(256, 340, 256, 256, r0, c0, 224, 224)
        got = _crop_resize_kernel_model(src, tab, 224, 224)
        for fi in range(3):
            plain, flipped = over[(wi * 3 + fi) * 2], over[(wi * 3 + fi) * 2 + 1]
            assert np.array_equal(got[fi], plain), (wi, fi)
            assert np.array_equal(P.flip_group([got[fi]])[0], flipped), (wi, fi)
    with pytest.raises(ValueError, match='leaves the frame'):
        S.crop_tables(100, 0, 192, 192, 224, 224, 256, 340)


def test_synthetic_training_example_plumbing():
    """examples/train_synthetic.py: loader, uint8 adapter and the epoch driver wired together (the
    fused step itself is replaced by a recorder; the real thing needs a GPU)."""
    import importlib.util, os
    from dmcnet_b200 import loop as L
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('_example_train', os.path.join(root, 'examples', 'train_synthetic.py'))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)

    class Rec:
        def __init__(self):
            self.calls = []
            self.in_flow = self.in_mv = self.in_res = self.target = None
        def set_epoch(self, epoch, epoch_thre=0):
            self.calls.append(('epoch', epoch))
        def step_u8(self, stack, target, flow_ds_factor=0, flip=None):
            assert stack.dtype == torch.uint8 and tuple(stack.shape) == (2, 3, 224, 224, 7)
            assert len(flip) == 2 and flow_ds_factor == 16 and target.shape == (2,)
            self.calls.append('step')
            return {'loss': 1.0, 'loss_cls': 1.0, 'loss_mse': 0.0, 'prec1': 0.0, 'prec5': 0.0}
        def load_inputs_u8(self, stack, target, ds):
            self.calls.append('load')
        def validate_batch(self, *a):
            self.calls.append('val')
            return {'loss': 1.0, 'loss_cls': 1.0, 'loss_mse': 0.0, 'prec1': 50.0, 'prec5': 50.0}
        def checkpoint(self, *a):
            raise AssertionError('no model_prefix was given')

    rec = Rec()
    train = ex._Pairs(ex.SyntheticLoader(3, 2, 3, 51, seed=0))
    val = ex._Pairs(ex.SyntheticLoader(1, 2, 3, 51, seed=1))
    best = L.fit(ex.U8Step(rec, 16), train, val, epochs=2, eval_freq=1, log=lambda *_: None)
    assert best == 50.0
    assert rec.calls == [('epoch', 0)] + ['step'] * 3 + ['load', 'val'] + [('epoch', 1)] + ['step'] * 3 + ['load', 'val']
