"""Post-update operators of the MEI ascent (reference featurevis/ops.py:294-307,441-481): non-differentiable
renormalisation (eps 1e-9) and range clipping, each one fused launch pair of lmr_constrain_fwd."""
import torch

from .. import _lib
from .. import ops as _ops
from .utils import varargin


class ClipRange:
    def __init__(self, x_min, x_max):
        self.x_min, self.x_max = x_min, x_max

    @varargin
    def __call__(self, x):
        return torch.clamp(x, self.x_min, self.x_max)


class _Renorm:
    mode = None

    def _apply(self, x, target, mean, pmin=None, pmax=None):
        with torch.no_grad():
            z, _ = _ops.raw_constrain_fwd(self.mode, x.contiguous(), target, mean, pmin, pmax, 1e-9)
        return z


class ChangeNorm(_Renorm):
    """x -> (x-mean)/(||x-mean|| + 1e-9) * norm + mean, per image."""
    mode = _lib.MODE_NORM

    def __init__(self, norm, mean=0):
        self.norm, self.mean = norm, mean

    @varargin
    def __call__(self, x):
        return self._apply(x, self.norm, self.mean)


class ChangeStats(_Renorm):
    """x -> (x-mu) * std/(std_unbiased(x) + 1e-9) + mean, per image."""
    mode = _lib.MODE_STD

    def __init__(self, std, mean):
        self.std, self.mean = std, mean

    @varargin
    def __call__(self, x):
        return self._apply(x, self.std, self.mean)


class GaussianBlur:
    """Blur with a Gaussian window -- mirror of featurevis.ops.GaussianBlur (featurevis/ops.py:378-423; same arguments), evaluated by
    lmr_gaussian_blur.  The reference builds the window with `scipy.signal.gaussian`, which no longer exists in scipy >= 1.13 (it is
    `scipy.signal.windows.gaussian`); the window here is the same exp(-0.5 (k / sigma)^2)."""

    def __init__(self, sigma, decay_factor=None, truncate=4, pad_mode="reflect"):
        self.sigma = sigma if isinstance(sigma, tuple) else (sigma,) * 2
        self.decay_factor = decay_factor
        self.truncate = truncate
        self.pad_mode = pad_mode

    def __call__(self, x, iteration=None):
        from .. import ops as _ops

        sigma = self.sigma if self.decay_factor is None else tuple(s + self.decay_factor * (iteration - 1) for s in self.sigma)
        return _ops.raw_gaussian_blur(x, sigma[0], sigma[1], truncate=self.truncate, pad_mode=self.pad_mode)
