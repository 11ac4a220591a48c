"""Drop-in replacement for code/dmcnet_I3D/network/i3d.py (``from network.i3d import I3D`` /
``from .i3d import I3D``, network/symbol_builder.py:10): the public classes with the reference's
signatures; see INTEGRATION.md."""
from dmcnet_b200.i3d_model import I3D, Mixed, MaxPool3dTFPadding, Unit3Dpy, get_padding_shape  # noqa: F401
from dmcnet_b200.model import (EstimatorDenseNet, EstimatorDenseNetSmall, EstimatorDenseNetTiny,  # noqa: F401
                               conv, predict_flow)
