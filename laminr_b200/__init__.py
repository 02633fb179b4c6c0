"""laminr_b200 -- B200-native (sm_100a) implementation of the LAMINR invariance-manifold hot path.

Drop-in for the public API of sinzlab/laminr (reference laminr/__init__.py:1-3):
    from laminr_b200 import neuron_models, get_mei_dict, InvarianceManifold
"""
__version__ = "0.1.0"
