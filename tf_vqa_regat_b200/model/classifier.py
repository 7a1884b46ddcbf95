"""SimpleClassifier -- mirrors model/classifier.py:11-26: WN Dense(hid) -> relu -> Dropout -> WN Dense(out)."""
from .weight_norm import Activation, Dense, Dropout, Layer, WeightNorm


class SimpleClassifier(Layer):
    def __init__(self, in_dim, hid_dim, out_dim, dropout):
        first = WeightNorm(Dense(hid_dim, input_shape=(in_dim,)))
        first.fused_relu = True
        self.layers = [first, Activation('relu'), Dropout(dropout), WeightNorm(Dense(out_dim))]

    def call(self, x):
        for layer in self.layers:
            if isinstance(layer, Activation):
                continue
            x = layer(x)
        return x
