"""tf.experimental.dlpack stand-in: torch's DLPack exporter / importer (capsule name "dltensor", like TensorFlow's)."""
import torch.utils.dlpack as _d


def to_dlpack(t):
    return _d.to_dlpack(t)


def from_dlpack(capsule):
    return _d.from_dlpack(capsule)
