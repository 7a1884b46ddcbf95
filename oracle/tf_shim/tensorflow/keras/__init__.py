"""tf.keras stand-in: layers, Model, backend.set_floatx, optimizers.experimental.Adamax.  TEST INFRASTRUCTURE."""
from . import initializers, layers, optimizers, preprocessing, utils  # noqa: F401
from .. import _core


class Model(layers.Layer):
    """keras.Model as the reference uses it: a Layer with compile() (rel_graph_net.py:9, train.py:51)."""
    optimizer = None

    def compile(self, loss=None, optimizer=None, **_):
        self.loss, self.optimizer = loss, optimizer


class backend:
    set_floatx = staticmethod(_core.set_floatx)

    @staticmethod
    def floatx():
        return "float64" if _core.floatx() is __import__("torch").float64 else "float32"


def Input(*a, **k):
    raise NotImplementedError("functional API is not part of the stand-in")
