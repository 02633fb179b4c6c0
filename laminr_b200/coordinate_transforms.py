"""Per-target affine coordinate transform (match stage).

Mirror of the reference interface laminr/coordinate_transforms.py:226-354 for the path the pipeline uses
(`CoordinateTransform(only_affine=True)` -> `AffineTransform`): same constructor arguments, same parameter names
(`transforms.Affine.{angles,scalings,shears,translation}`) and initial values, so `state_dict()` round-trips with the
reference.  The arithmetic (M_d = R.S.Hs, coordinate map, clamp, gradients of the 7 scalars) runs inside the INR
kernels (lmr_inr_forward / lmr_inr_backward); these modules own the parameters, and `forward` is a plain tensor-op mirror
of the reference's for callers that map coordinates themselves.

Out of scope: the RealNVP warp (`only_affine=False`) and the stochastic transform.
"""
from collections import OrderedDict

import torch
from torch import nn


class AffineTransform(nn.Module):
    def __init__(self, n_neurons, in_dim=2, stochastic=False, init_noise_scale=0.1, allow_scale=True, allow_shear=True, uniform_scale=False):
        super().__init__()
        if in_dim != 2:
            raise NotImplementedError("laminr_b200: AffineTransform supports in_dim=2 (image coordinates) only")
        if stochastic:
            raise NotImplementedError("laminr_b200: stochastic coordinate transforms are not on the fused path (reference default is False)")
        self.n_neurons, self.in_dim, self.stochastic, self.uniform_scale = n_neurons, in_dim, stochastic, uniform_scale
        self.angles = nn.Parameter(torch.zeros(n_neurons))
        if allow_scale:
            self.scalings = nn.Parameter(torch.ones(n_neurons, 1 if uniform_scale else in_dim))
        else:
            self.register_buffer("scalings", torch.ones(n_neurons, in_dim))
        if allow_shear:
            self.shears = nn.Parameter(torch.zeros(n_neurons, in_dim))
        else:
            self.register_buffer("shears", torch.zeros(n_neurons, in_dim))
        self.translation = nn.Parameter(torch.zeros(n_neurons, 1, in_dim))

    @property
    def As(self):
        """[D,2,2] matrices M = R(theta).diag(s).[[1,sh0],[sh1,1]] (coordinate_transforms.py:276-296); host-side inspection only."""
        c, s = torch.cos(self.angles), torch.sin(self.angles)
        R = torch.stack([torch.stack([c, -s], -1), torch.stack([s, c], -1)], -2)
        sc = self.scalings.repeat(1, 2) if self.scalings.shape[1] == 1 else self.scalings
        S = torch.diag_embed(sc)
        one = torch.ones_like(self.shears[:, 0])
        Hs = torch.stack([torch.stack([one, self.shears[:, 0]], -1), torch.stack([self.shears[:, 1], one], -1)], -2)
        return R @ S @ Hs

    def forward(self, x, stochastic=None):
        """x [D, P, 2] -> M_d x + t_d (coordinate_transforms.py:299-309).  Host-side mirror for callers that map coordinates themselves (plots,
        inspection of a matched transform): plain differentiable tensor ops on whatever device the parameters live on.  The optimisation loops
        never call it -- learn / match evaluate the transform inside the fused INR kernels (lmr_inr_forward / lmr_inr_backward)."""
        if stochastic:
            raise NotImplementedError("laminr_b200: stochastic coordinate transforms are not supported (reference default is False)")
        return torch.matmul(x, self.As.transpose(-1, -2)) + self.translation


class CoordinateTransform(nn.Module):
    def __init__(self, n_neurons, only_affine=False, flow_kwgs=None, in_dim=2, stochastic=False, init_noise_scale=0.1, allow_scale=True,
                 allow_shear=True, uniform_scale=False):
        super().__init__()
        if not only_affine:
            raise NotImplementedError("laminr_b200: only_affine=False (RealNVP warp) is outside the fused path; match() uses only_affine=True")
        self.n_dim, self.only_affine, self.stochastic, self.init_noise_scale = n_neurons, only_affine, stochastic, init_noise_scale
        self.allow_scale, self.allow_shear, self.uniform_scale = allow_scale, allow_shear, uniform_scale
        self.transforms = nn.Sequential(OrderedDict([("Affine", AffineTransform(
            n_neurons, in_dim=in_dim, stochastic=stochastic, init_noise_scale=init_noise_scale, allow_scale=allow_scale,
            allow_shear=allow_shear, uniform_scale=uniform_scale))]))

    def forward(self, coords):
        return self.transforms(coords)
