"""Drop-in replacement for code/dmcnet_GAN/model.py (``from model import Model``,
code/dmcnet_GAN/train.py:21).  Every public name of the reference module is provided with
the same signature; see INTEGRATION.md."""
from dmcnet_b200.model import GANModel as Model  # noqa: F401
from dmcnet_b200.model import (ContextNetwork, ContextNetworkAtt, EstimatorDenseNet,  # noqa: F401
                               EstimatorDenseNetSmall, EstimatorDenseNetTiny,
                               EstimatorDenseNetTinyEarlyFusionSum, EstimatorDenseNetTinyEarlyFusionStack,
                               Flatten, conv, conv_dilation, predict_flow, discriminator_block,
                               discriminator_block2, Discriminator, Discriminator2, Discriminator3,
                               Discriminator4, Discriminator5)
