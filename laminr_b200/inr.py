"""INR image generator: a CPPN-style MLP over (y, x) pixel coordinates and a periodic latent.

Mirror of the reference interface laminr/inr.py:8-413 (`INRTemplates`): same constructor signature, buffers
(`B`, `auxB`, `coords`, `template_input`, `coords_shift_to_center`, `template_center_in_coords_space`,
`coords_shift_from_center`) and parameter names (`func.layer{i}.weight/bias`), same consumption of the CPU RNG at
construction (inr.py:49-51), so a seed gives the same network as the reference.  `forward` is ONE fused kernel
(lmr_inr_forward) instead of sin/cos + cartesian-product materialisation + 5 GEMMs; autograd goes through
lmr_inr_backward.
"""
from collections import OrderedDict

import torch
from torch import nn

from . import ops
from .coordinate_transforms import CoordinateTransform


class INRTemplates(nn.Module):
    # Hidden width the kernels are built for.  Any list of widths up to it (uniform or not, the reference's `widths=` argument, inr.py:156-174) runs
    # on the same kernels ZERO-PADDED: the parameters are strided views ([w_out, w_in] corners) of 50-wide blocks of the flat vector; a padded unit
    # has zero weights and bias, so it outputs tanh(0) = 0, feeds nothing forward, receives a zero gradient (its outgoing weights are zero) and
    # passes a zero gradient to its incoming weights (its activation is zero) -- the padding stays exactly zero under SGD / Adam, and the
    # function and every gradient are those of the narrow network.
    KERNEL_WIDTH = 50

    def __init__(self, img_res, num_neurons=0, num_templates=1, out_channels=1, aux_dim=1, widths=[15] * 8, periodic_invariance=False,
                 positional_encoding_dim=None, positional_encoding_projection_scale=1.0, aux_positional_encoding_dim=None,
                 aux_positional_encoding_projection_scale=1.0, nonlinearity=nn.Tanh, final_nonlinearity=nn.Sigmoid, bias=False,
                 batchnorm=False, weights_scale=1.0, jitter_coords=False, jitter_coords_scale=1.0, precision=None):
        super().__init__()
        unsupported = []
        if num_templates != 1: unsupported.append("num_templates != 1")
        if aux_dim != 1: unsupported.append("aux_dim != 1")
        if not periodic_invariance: unsupported.append("periodic_invariance=False")
        if positional_encoding_dim is None or aux_positional_encoding_dim is None: unsupported.append("no positional encoding")
        if nonlinearity is not nn.Tanh or final_nonlinearity is not nn.Tanh: unsupported.append("non-tanh nonlinearity")
        if not bias: unsupported.append("bias=False")
        if batchnorm: unsupported.append("batchnorm=True")
        if jitter_coords: unsupported.append("jitter_coords=True")
        if max(widths) > self.KERNEL_WIDTH or min(widths) < 1:
            unsupported.append("hidden widths %s (the kernels serve widths up to %d: narrower layers run zero-padded)" % (list(widths), self.KERNEL_WIDTH))
        if not 2 <= len(widths) <= 8: unsupported.append("%d hidden layers (supported: 2..8)" % len(widths))
        if not 1 <= out_channels <= 4: unsupported.append("%d output channels (supported: 1..4)" % out_channels)
        if unsupported:
            raise NotImplementedError("laminr_b200.INRTemplates: configuration outside the fused kernels (%s); the pipeline "
                                      "(InvarianceManifold) uses widths=[50]*4, tanh, bias, periodic latent, positional encodings" % ", ".join(unsupported))
        self.img_res = tuple(img_res)
        self.out_channels, self.num_neurons, self.num_templates, self.aux_dim, self.template_dim = out_channels, num_neurons, num_templates, aux_dim, 1
        self.weights_scale, self.periodic_invariance, self.batchnorm = weights_scale, periodic_invariance, batchnorm
        self.positional_encoding_dim, self.positional_encoding_projection_scale = positional_encoding_dim, positional_encoding_projection_scale
        self.aux_positional_encoding_dim = aux_positional_encoding_dim
        self.aux_positional_encoding_projection_scale = aux_positional_encoding_projection_scale
        self.jitter_coords, self.jitter_coords_scale = jitter_coords, jitter_coords_scale
        self.precision = precision
        self.in_dim = 2 * positional_encoding_dim + self.template_dim + 2 * aux_positional_encoding_dim  # inr.py:142-154
        self.widths = list(widths)

        # (1) nn.Linear construction draws from the CPU generator even though the values are overwritten (inr.py:156-174)
        elements, in_widths = [], [self.in_dim] + self.widths[:-1]
        for i, w in enumerate(self.widths):
            elements.append((f"layer{i}", nn.Linear(in_widths[i], w, bias=True)))
            elements.append((f"nonlinearity{i}", nn.Tanh()))
        elements.append((f"layer{len(self.widths)}", nn.Linear(self.widths[-1], out_channels, bias=True)))
        elements.append((f"nonlinearity{len(self.widths)}", nn.Tanh()))
        self.func = nn.Sequential(OrderedDict(elements))
        # (2),(3) random Fourier projections (inr.py:176-192)
        self.register_buffer("B", torch.randn(2, positional_encoding_dim) * positional_encoding_projection_scale)
        self.register_buffer("auxB", torch.randn(2 * aux_dim, aux_positional_encoding_dim) * aux_positional_encoding_projection_scale)
        coords = torch.meshgrid([torch.linspace(-1, 1, r) for r in self.img_res], indexing="ij")  # (row=y, col=x), inr.py:194-198
        self.register_buffer("coords", torch.stack(coords, dim=-1))
        self.register_buffer("template_input", torch.zeros(1))
        # (4) weights ~ N(0, weights_scale), bias = 0, in module order (inr.py:207-211)
        for m in self.func:
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(0.0, weights_scale)
                m.bias.data.fill_(0)
        self._flat = None
        self._desc = ops.make_desc(self.KERNEL_WIDTH, len(self.widths), positional_encoding_dim, aux_positional_encoding_dim, out_channels)

    # ---- flat parameter vector shared with the kernels -------------------------------------------------------
    def _blocks(self):
        """(offset, padded shape) in the flat vector of every parameter of `func`, in module order: W0 [50, in_dim], b0 [50], W_l [50, 50],
        b_l [50], ..., W_out [C, 50], b_out [C] -- the layout include/laminr_b200.h gives lmr_inr_forward for hidden width 50."""
        K, out, off = self.KERNEL_WIDTH, [], 0
        k_in = self.in_dim
        for _ in self.widths:
            out.append((off, (K, k_in))); off += K * k_in
            out.append((off, (K,))); off += K
            k_in = K
        out.append((off, (self.out_channels, K))); off += self.out_channels * K
        out.append((off, (self.out_channels,))); off += self.out_channels
        return out, off

    def flat_params(self):
        """The parameters of `func` as ONE contiguous vector (the layout lmr_inr_forward takes); every nn.Parameter is a view of it (the
        [w_out, w_in] corner of its zero-padded block when a layer is narrower than the kernels' width)."""
        ps = list(self.func.parameters())
        blocks, total = self._blocks()
        flat = self._flat

        def view_of(fl, off, shape, p):
            blk = fl[off:off + int(torch.Size(shape).numel())].view(shape)
            return blk[tuple(slice(0, n) for n in p.shape)]

        ok = flat is not None and flat.device == ps[0].device and flat.numel() == total
        if ok:
            for p, (off, shape) in zip(ps, blocks):
                want = view_of(flat, off, shape, p)
                if p.data_ptr() != want.data_ptr() or p.stride() != want.stride():
                    ok = False
                    break
        if not ok:
            flat = torch.zeros(total, dtype=torch.float32, device=ps[0].device)
            for p, (off, shape) in zip(ps, blocks):
                v = view_of(flat, off, shape, p)
                v.copy_(p.detach())
                p.data = v
            self._flat = flat
        return flat

    # ---- coordinate transform bookkeeping (inr.py:213-296) ------------------------------------------------------
    def initialize_coordinate_transform(self, num_neurons, only_affine, stochastic, init_noise_scale, allow_scale, allow_shear, uniform_scale,
                                        clamp_boundaries):
        device = next(self.parameters()).device
        self.num_neurons = num_neurons
        self.coordinate_transform = CoordinateTransform(
            num_neurons, only_affine=only_affine, stochastic=stochastic, init_noise_scale=init_noise_scale, allow_scale=allow_scale,
            allow_shear=allow_shear, uniform_scale=uniform_scale).to(device)
        self.coordinate_transform_clamp_boundaries = clamp_boundaries

    def reset_coordinate_transform(self, num_neurons=None):
        ct = self.coordinate_transform
        self.initialize_coordinate_transform(num_neurons if num_neurons is not None else ct.n_dim, ct.only_affine, ct.stochastic,
                                             ct.init_noise_scale, ct.allow_scale, ct.allow_shear, ct.uniform_scale,
                                             self.coordinate_transform_clamp_boundaries)

    def get_coords_shift_to_center(self, template_rf_center, img_res=None):
        device = next(self.parameters()).device
        res = torch.tensor(img_res if img_res is not None else self.img_res).to(device)
        centre = res / 2
        return (template_rf_center.flip(dims=[-1]).to(device) - centre) / centre

    def get_coords_shift_from_center(self, template_rf_center, img_res=None):
        return -1 * self.get_coords_shift_to_center(template_rf_center, img_res=img_res)

    def get_template_center_in_coords_space(self, template_rf_center, img_res=None):
        device = next(self.parameters()).device
        res = torch.tensor(img_res if img_res is not None else self.img_res).to(device)
        return template_rf_center.flip(dims=[-1]).to(device) / res * 2 - 1

    def register_coords_shifts(self, template_rf_center, target_rf_centers):
        self.register_buffer("coords_shift_to_center", self.get_coords_shift_to_center(template_rf_center))
        self.register_buffer("template_center_in_coords_space", self.get_template_center_in_coords_space(template_rf_center))
        self.register_buffer("coords_shift_from_center", self.get_coords_shift_from_center(target_rf_centers))

    def get_coordinates(self, img_res, device):
        if img_res is None:
            return self.coords
        return torch.stack(torch.meshgrid([torch.linspace(-1, 1, s, device=device) for s in img_res], indexing="ij"), dim=-1)

    # ---- kernel argument assembly ---------------------------------------------------------------------------------
    def affine_args(self):
        aff = self.coordinate_transform.transforms.Affine
        return ops.AffineArgs(
            aff.angles.detach(), aff.scalings.detach(), aff.shears.detach(), aff.translation.detach(),
            **self._affine_const())

    def _affine_const(self):
        def f(name):
            t = getattr(self, name, None)
            return None if t is None else t.detach().to(torch.float32).contiguous()
        return dict(tc=f("template_center_in_coords_space"), shift_to=f("coords_shift_to_center"), shift_from=f("coords_shift_from_center"),
                    clamp=bool(getattr(self, "coordinate_transform_clamp_boundaries", True)))

    def forward(self, img_res=None, aux=None, return_template=False, centralize_template=False):
        """Same contract as the reference forward (inr.py:53-100): aux [N,1] latents in [0, 2pi) ->
        [N, C, H, W] (return_template) or [D, N, C, H, W] (one image set per target neuron)."""
        return_template = True if self.num_neurons == 0 else return_template
        device = next(self.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("laminr_b200.INRTemplates runs on CUDA (sm_100a) only; move the module to a B200 (no CPU fallback)")
        coords = self.get_coordinates(img_res, device)
        res = tuple(img_res) if img_res is not None else self.img_res
        coords = coords.reshape(-1, 2).contiguous()
        if aux is None:
            aux = torch.zeros(1, 1, device=device)
        grid = aux.to(device=device, dtype=torch.float32).reshape(-1)
        N = grid.numel()
        flat = self.flat_params()
        views = list(self.func.parameters())
        meta = dict(desc=self._desc, precision=self.precision, affine_const={})
        if return_template:
            shift = None
            if centralize_template and hasattr(self, "coords_shift_to_center"):
                shift = self.coords_shift_to_center.to(torch.float32).contiguous()
            out = ops.inr_images(meta, flat, self.B, self.auxB, grid, coords, shift, None, views)
            return out.reshape(N, self.out_channels, *res)
        aff = self.coordinate_transform.transforms.Affine
        meta["affine_const"] = self._affine_const()
        out = ops.inr_images(meta, flat, self.B, self.auxB, grid, coords, None, (aff.angles, aff.scalings, aff.shears, aff.translation), views)
        return out.reshape(self.num_neurons, N, self.out_channels, *res)
