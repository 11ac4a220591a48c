"""CPU: pin the oracle against the reference's own model.py when /root/reference
is present (build container); skipped on the GPU box."""
import pytest

from oracle import ref_loader as R


@pytest.mark.skipif(not R.reference_available(), reason='/root/reference not present')
def test_oracle_pinned_against_reference_model():
    from oracle.pin_against_reference import pin
    assert pin('dmcnet', 51, None, batch=1, verbose=False) <= 1e-6
    assert pin('dmcnet_GAN', 51, 'Discriminator3', batch=1, verbose=False) <= 1e-6


@pytest.mark.skipif(not R.reference_available(), reason='/root/reference not present')
@pytest.mark.parametrize('kw', [dict(arch_estimator='DenseNetTinyEarlyFusionStack'),
                                dict(arch_estimator='DenseNetTiny', ds=4),
                                dict(arch_estimator='ContextNetwork', att=1)])
def test_oracle_pinned_for_other_generator_choices(kw):
    """--arch_estimator / --att / --gen_flow_ds_factor.  The full list
    (ContextNetwork gradients included) runs in ``python oracle/pin_against_reference.py``; the
    reference's att=1 backward raises on torch >= 1.x, so that case pins forwards only."""
    from oracle.pin_against_reference import pin
    assert pin('dmcnet', 51, None, batch=1, verbose=False, **kw) <= 1e-6
