"""Drop-in replacement for code/dmcnet/model.py: put this directory on sys.path
ahead of the reference's so that ``from model import Model`` (code/dmcnet/
train.py:21, test.py) resolves here.  See INTEGRATION.md."""
from dmcnet_b200.model import Model, EstimatorDenseNetTiny, conv, predict_flow  # noqa: F401
