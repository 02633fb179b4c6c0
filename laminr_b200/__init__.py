"""laminr_b200 -- B200-native (sm_100a) implementation of the LAMINR invariance-manifold hot path.

Drop-in for the public API of sinzlab/laminr (reference laminr/__init__.py:1-3):
    from laminr_b200 import neuron_models, get_mei_dict, InvarianceManifold
The compute path is the C-ABI library of include/laminr_b200.h (hand-written CUDA for sm_100a); importing this package
fails if the library has not been built (no CPU / eager fallback).
"""
from . import _lib

_lib.load()  # fail loudly at import time if liblaminr_b200.so is missing

from . import neuron_models  # noqa: E402,F401
from .inv_temp_learn_match import InvarianceManifold  # noqa: E402,F401
from .utils import get_mei_dict  # noqa: E402,F401
from .ops import set_default_precision  # noqa: E402,F401
from . import torch_ops  # noqa: E402,F401  (registers torch.ops.laminr_b200.*)

__version__ = "0.1.0"
