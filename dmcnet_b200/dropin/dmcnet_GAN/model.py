"""Drop-in replacement for code/dmcnet_GAN/model.py (``from model import Model``,
code/dmcnet_GAN/train.py:21).  See INTEGRATION.md."""
from dmcnet_b200.model import GANModel as Model, EstimatorDenseNetTiny, conv, predict_flow  # noqa: F401
from dmcnet_b200.model import discriminator_block  # noqa: F401
