"""Image-constraint projection (differentiable).

Mirror of the reference interface laminr/utils/img_transforms.py:5-65: `FixEnergyNormClip(norm, mean, pixel_min,
pixel_max)` and `StandardizeClip(mean, std, pixel_min, pixel_max, dim)` return an nn.Module mapping [B,C,H,W] ->
[B,C,H,W].  Norm/standardise + clip is one fused op (lmr_constrain_fwd/bwd) instead of ~8 elementwise launches.
"""
from torch import nn

from .. import ops, _lib


class _ConstrainClip(nn.Module):
    def __init__(self, mode, target, mean, pixel_min, pixel_max, eps=1e-12):
        super().__init__()
        self.mode, self.target, self.mean, self.pixel_min, self.pixel_max, self.eps = mode, target, mean, pixel_min, pixel_max, eps

    def forward(self, x):
        return ops.constrain(x, self.mode, self.target, self.mean, self.pixel_min, self.pixel_max, self.eps)

    def __repr__(self):
        kind = "FixEnergyNormClip(norm" if self.mode == _lib.MODE_NORM else "StandardizeClip(std"
        return f"{kind}={self.target}, mean={self.mean}, min={self.pixel_min}, max={self.pixel_max})"


def FixEnergyNormClip(norm, mean=0, pixel_min=None, pixel_max=None):
    return _ConstrainClip(_lib.MODE_NORM, norm, mean, pixel_min, pixel_max)


def StandardizeClip(mean, std, pixel_min=None, pixel_max=None, dim=[1, 2, 3]):
    if list(dim) != [1, 2, 3]:
        raise NotImplementedError("laminr_b200.StandardizeClip: statistics are per image (dim=[1,2,3]) on the fused path")
    if mean is None or std is None:
        raise NotImplementedError("laminr_b200.StandardizeClip: mean and std are required")
    return _ConstrainClip(_lib.MODE_STD, std, mean, pixel_min, pixel_max)
