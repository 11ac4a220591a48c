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
