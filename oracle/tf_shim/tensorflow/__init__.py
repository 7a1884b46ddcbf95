"""Stand-in for the `tensorflow` package, just wide enough for the reference's hot-path files
(model/{weight_norm,fc,graph_att_layer,graph_att_net,relation_encoder,fusion,classifier,
rel_graph_net}.py and train.py) to be imported and executed UNMODIFIED.  See ../README.md.

TEST INFRASTRUCTURE: used only by oracle/make_golden_ref.py in the build container.
"""
from ._core import *  # noqa: F401,F403
from ._core import (DType, GradientTape, Tensor, Variable, bool_ as bool, float32, float64, int32, int64, math, newaxis,  # noqa: F401
                    nn, random, set_floatx)
from . import keras  # noqa: F401,E402

__version__ = "0.0-regat-standin"
