"""FullyConnected -- mirrors model/fc.py:11-50: [Dropout(p) if p > 0] -> WeightNorm(Dense) -> [Activation] per layer.

Quirk kept on purpose: the activation is applied only when `activation` is EXACTLY the string 'relu'
(or 'tanh'); fusion.py passes a float there, which silently makes those FCs linear with no dropout."""
from .weight_norm import Activation, Dense, Dropout, Layer, WeightNorm


class FullyConnected(Layer):
    def __init__(self, dims, activation='relu', dropout=0, bias=True):
        self.layers = []
        for i in range(len(dims) - 2):                                   # fc.py:17-31
            self._block(dims[i + 1], activation, dropout, bias)
        self._block(dims[-1], activation, dropout, bias)                 # fc.py:33-43

    def _block(self, out_dim, activation, dropout, bias):
        if isinstance(dropout, (int, float)) and dropout > 0:
            self.layers.append(Dropout(dropout))
        wn = WeightNorm(Dense(out_dim, use_bias=bias, activation=None))
        self.layers.append(wn)
        if activation == 'relu':
            wn.fused_relu = True                                         # ReLU runs in the GEMM epilogue
            self.layers.append(Activation('relu'))
        elif activation == 'tanh':
            raise NotImplementedError("FullyConnected(activation='tanh') never occurs on the hot path")

    def call(self, x):
        for layer in self.layers:
            if isinstance(layer, Activation):
                continue                                                 # fused into the preceding WeightNorm(Dense)
            x = layer(x)
        return x

    @property
    def dense(self):
        """The last WeightNorm(Dense) of the stack (all hot-path FCs have exactly one)."""
        return [l for l in self.layers if isinstance(l, WeightNorm)][-1]
