"""CPU: pin the oracle against the reference's own model.py when /root/reference
is present (build container); skipped on the GPU box."""
import pytest

from oracle import ref_loader as R


@pytest.mark.skipif(not R.reference_available(), reason='/root/reference not present')
def test_oracle_pinned_against_reference_model():
    from oracle.pin_against_reference import pin
    assert pin('dmcnet', 51, None, batch=1, verbose=False) <= 1e-6
    assert pin('dmcnet_GAN', 51, 'Discriminator3', batch=1, verbose=False) <= 1e-6
