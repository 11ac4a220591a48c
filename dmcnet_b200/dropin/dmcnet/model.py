"""Drop-in replacement for code/dmcnet/model.py (``from model import Model``,
code/dmcnet/train.py:21, test.py:14).  Every public name of the reference module is
provided with the same signature; see INTEGRATION.md."""
from dmcnet_b200.model import (Model, ContextNetwork, ContextNetworkAtt, EstimatorDenseNet,  # noqa: F401
                               EstimatorDenseNetSmall, EstimatorDenseNetTiny,
                               EstimatorDenseNetTinyEarlyFusionSum, EstimatorDenseNetTinyEarlyFusionStack,
                               Flatten, conv, conv_dilation, predict_flow)
