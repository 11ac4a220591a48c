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
    assert abs(val.astype(np.float64).sum() / d.size - float(loss)) < 1e-6
    np.testing.assert_allclose(scales['flow'] * slope, a.grad.numpy(), rtol=1e-6, atol=1e-9)
    assert scales['mse'] == pytest.approx(2 * scales['flow'])
