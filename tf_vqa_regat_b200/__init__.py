"""B200-native implementation of ReGAT's implicit-relation graph attention + BUTD fusion hot path
(reference: jhss/TF_VQA_ReGAT).  CUDA kernels + C ABI in csrc/ -> libregat.so; this package is the host-side
mirror of the reference's layer API plus the fused engine.  Importing the package does not touch CUDA;
any op without a CUDA device / built library raises (there is no CPU fallback)."""
from .config import HotPathConfig, param_layout, num_trainable  # noqa: F401

__all__ = ["HotPathConfig", "param_layout", "num_trainable"]
