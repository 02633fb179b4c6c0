"""Stacked-conv core + Gaussian readout response models (BASELINE config 4).

Host-side mirrors of the three neuralpredictors classes the LAMINR pipeline can be pointed at:
  * `Stacked2dCore`      -- neuralpredictors/layers/cores/conv2d.py:33-278 (plain `nn.Conv2d` + eval-mode BatchNorm + shifted ELU per
                            layer; the features of the layers in `stack` are concatenated)
  * `FullGaussian2d`     -- neuralpredictors/layers/readouts/gaussian.py:210-599 (one learned position per neuron, bilinear read of the
                            feature map there, dot with a per-neuron feature vector, + bias)
  * `FiringRateEncoder`  -- neuralpredictors/layers/encoders/firing_rate.py:4-61 (core -> readout -> elu(x + offset) + 1)
Same constructor arguments (the subset the pipeline uses), same parameter / buffer names and the same RNG draw order, so a seed
gives the reference's weights and a reference state_dict loads unchanged.  Variants that LAMINR never instantiates for this
configuration (depth-separable / attention convolutions, skip > 1, grid predictors, shared features) raise NotImplementedError.

On a CUDA device `FiringRateEncoder` evaluates responses with lmr_convcore_response (csrc/convcone.cu): one CTA per (image, neuron) pair
runs the receptive-field cone of the selected neuron through all layers in shared memory and returns the response and its image gradient
in one launch: `forward_neurons(x, idxs)`, what SingleNeuronModel / the fused steppers / the batched MEI ascent use.  `forward(x)` (all neurons)
runs the whole frame through the core once with lmr_ff_forward / lmr_ff_backward (csrc/fullframe.cu) and reads every neuron out of that.  There is NO stock-PyTorch / CPU evaluation path: the classes below only build and hold the
parameters (same names, shapes and RNG draw order as the reference, so a reference state_dict loads unchanged); CPU tensors raise, and
configurations the kernel does not cover (non-"same" padding, stride/dilation != 1, more than the last layer stacked, training mode) raise
NotImplementedError instead of silently falling back.
"""
import warnings
from collections import OrderedDict
from collections.abc import Iterable

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


class AdaptiveELU(nn.Module):
    """elu(x - xshift) + yshift  (neuralpredictors/layers/activations.py:31-51)."""

    def __init__(self, xshift, yshift):
        super().__init__()
        self.xshift, self.yshift = xshift, yshift

    def forward(self, x):
        return F.elu(x - self.xshift) + self.yshift


class _LaplaceBuffer(nn.Module):
    """Holds the 3x3 Laplace filter buffer the reference core registers for its (training-only) input regulariser, so that state_dict
    keys match (`_input_weights_regularizer.laplace.filter`, neuralpredictors/regularizers.py:12-17,136-159)."""

    def __init__(self):
        super().__init__()
        self.laplace = nn.Module()
        self.laplace.register_buffer("filter", torch.tensor([[0.0, -1.0, 0.0], [-1.0, 4.0, -1.0], [0.0, -1.0, 0.0]])[None, None])


class Stacked2dCore(nn.Module):
    def __init__(self, input_channels, hidden_channels, input_kern, hidden_kern, layers=3, gamma_hidden=0, gamma_input=0.0, skip=0, stride=1,
                 final_nonlinearity=True, elu_shift=(0, 0), bias=True, momentum=0.1, pad_input=True, hidden_padding=None, batch_norm=True,
                 batch_norm_scale=True, independent_bn_bias=True, hidden_dilation=1, laplace_padding=0, input_regularizer="LaplaceL2", stack=None,
                 use_avg_reg=True, depth_separable=False, attention_conv=False, linear=False):
        super().__init__()
        if depth_separable or attention_conv or skip > 1 or not independent_bn_bias or input_regularizer != "LaplaceL2":
            raise NotImplementedError("laminr_b200 Stacked2dCore: only plain convolutions with the default batch-norm arrangement are mirrored")
        self._input_weights_regularizer = _LaplaceBuffer()
        self.num_layers, self.input_channels, self.hidden_channels = layers, input_channels, hidden_channels
        self.gamma_input, self.gamma_hidden, self.skip, self.stride = gamma_input, gamma_hidden, skip, stride
        self.input_kern, self.hidden_dilation, self.final_nonlinearity = input_kern, hidden_dilation, final_nonlinearity
        self.elu_xshift, self.elu_yshift = elu_shift
        self.bias, self.momentum, self.pad_input, self.batch_norm, self.linear = bias, momentum, pad_input, batch_norm, linear
        self.stack = range(layers) if stack is None else ([*range(layers)[stack:]] if isinstance(stack, int) else stack)
        self.hidden_kern = list(hidden_kern) if isinstance(hidden_kern, Iterable) else [hidden_kern] * (layers - 1)
        self.hidden_padding = hidden_padding

        self.features = nn.Sequential()
        for l in range(layers):  # conv2d.py:202-234
            layer = OrderedDict()
            if l == 0:
                # first conv: zero-padded k//2 only if pad_input; no conv bias when batch norm carries it (:204-210)
                layer["conv"] = nn.Conv2d(input_channels, hidden_channels, input_kern, padding=input_kern // 2 if pad_input else 0,
                                          bias=bias and not batch_norm)
            else:
                if self.hidden_padding is None:  # NB: computed once, from the first hidden kernel (:220-222)
                    self.hidden_padding = ((self.hidden_kern[l - 1] - 1) * hidden_dilation + 1) // 2
                layer["conv"] = nn.Conv2d(hidden_channels, hidden_channels, self.hidden_kern[l - 1], stride=stride, padding=self.hidden_padding,
                                          dilation=hidden_dilation, bias=bias)
            if batch_norm:
                layer["norm"] = nn.BatchNorm2d(hidden_channels, momentum=momentum)
            if not linear and (l < layers - 1 or final_nonlinearity):
                layer["nonlin"] = AdaptiveELU(self.elu_xshift, self.elu_yshift)
            self.features.add_module(f"layer{l}", nn.Sequential(layer))
        for m in self.modules():  # Core.init_conv (cores/base.py:17-30), module order
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_normal_(m.weight.data)
                if m.bias is not None:
                    m.bias.data.fill_(0)

    def forward(self, x):
        raise NotImplementedError("laminr_b200: the core is evaluated only inside FiringRateEncoder on a CUDA device (csrc/convcone.cu); "
                                  "there is no stock-PyTorch / CPU path")

    @property
    def outchannels(self):
        return len(self.features) * self.hidden_channels


class FullGaussian2d(nn.Module):
    def __init__(self, in_shape, outdims, bias, init_mu_range=0.1, init_sigma=1, batch_sample=True, align_corners=True, gauss_type="full",
                 grid_mean_predictor=None, shared_features=None, shared_grid=None, source_grid=None, mean_activity=None, feature_reg_weight=1.0,
                 **kwargs):
        super().__init__()
        if grid_mean_predictor is not None or shared_features is not None or shared_grid is not None:
            raise NotImplementedError("laminr_b200 FullGaussian2d: grid predictors / shared features or grids are not mirrored")
        if init_mu_range > 1.0 or init_mu_range <= 0.0 or init_sigma <= 0.0:
            raise ValueError("either init_mu_range doesn't belong to [0.0, 1.0] or init_sigma_range is non-positive")
        if gauss_type not in ("full", "uncorrelated", "isotropic"):
            raise ValueError(f'gauss_type "{gauss_type}" not known')
        self.in_shape, self.outdims, self.batch_sample, self.align_corners = in_shape, outdims, batch_sample, align_corners
        self.gauss_type, self.init_mu_range, self.init_sigma, self.mean_activity = gauss_type, init_mu_range, init_sigma, mean_activity
        self.feature_reg_weight = feature_reg_weight
        c = in_shape[0]
        self.grid_shape = (1, outdims, 1, 2)
        sigma_shape = {"full": (1, outdims, 2, 2), "uncorrelated": (1, outdims, 1, 2), "isotropic": (1, outdims, 1, 1)}[gauss_type]
        self._mu = nn.Parameter(torch.empty(*self.grid_shape))
        self.sigma = nn.Parameter(torch.empty(*sigma_shape))
        self._features = nn.Parameter(torch.empty(1, c, 1, outdims))
        self.bias = nn.Parameter(torch.empty(outdims)) if bias else None
        # gaussian.py:452-469: mu ~ U(+-init_mu_range), sigma ~ U(+-init_sigma) ('full') or the constant, features = 1/c, bias = mean activity or 0
        self._mu.data.uniform_(-init_mu_range, init_mu_range)
        if gauss_type != "full":
            self.sigma.data.fill_(init_sigma)
        else:
            self.sigma.data.uniform_(-init_sigma, init_sigma)
        self._features.data.fill_(1 / c)
        if self.bias is not None:
            if mean_activity is None:
                warnings.warn("Readout is NOT initialized with mean activity but with 0!")
                self.bias.data.fill_(0)
            else:
                self.bias.data = mean_activity

    @property
    def mu(self):
        return self._mu

    @property
    def features(self):
        return self._features

    def sample_grid(self, batch_size, sample=None):
        """gaussian.py:396-429.  Eval mode (what LAMINR must run in): grid = clamp(mu); mu itself is clamped in place."""
        with torch.no_grad():
            self._mu.clamp_(min=-1, max=1)
        shape = (batch_size,) + self.grid_shape[1:]
        sample = self.training if sample is None else sample
        norm = self._mu.new_empty(*shape).normal_() if sample else self._mu.new_zeros(*shape)
        if self.gauss_type != "full":
            return torch.clamp(norm * self.sigma + self._mu, min=-1, max=1)
        return torch.clamp(torch.einsum("ancd,bnid->bnic", self.sigma, norm) + self._mu, min=-1, max=1)

    def forward(self, x, sample=None, shift=None, out_idx=None, **kwargs):
        raise NotImplementedError("laminr_b200: the readout is evaluated only inside FiringRateEncoder on a CUDA device (csrc/convcone.cu); "
                                  "there is no stock-PyTorch / CPU path")


class FiringRateEncoder(nn.Module):
    def __init__(self, core, readout, *, shifter=None, modulator=None, elu_offset=0.0):
        super().__init__()
        if shifter is not None or modulator is not None:
            raise NotImplementedError("laminr_b200 FiringRateEncoder: shifter / modulator branches are not mirrored")
        self.core, self.readout, self.shifter, self.modulator, self.offset = core, readout, None, None, elu_offset
        self._cone, self._cone_sig, self._ff, self._ff_sig = None, None, None, None
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module.refresh_cone())

    # ---- B200 path --------------------------------------------------------------------------------------------------
    def _state_signature(self):
        # identity + in-place version counter of every tensor the pack is derived from: load_state_dict / optimiser steps / .copy_() bump the
        # version, .to(device) / .float() replace the tensors
        return (self.training,) + tuple((id(t), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def cone(self):
        """The lmr_convcore_response weight pack of this (frozen, eval-mode) model: built on first use and rebuilt whenever the module's
        parameters / buffers were replaced or modified in place since (load_state_dict, .to(), .train()/.eval(), manual edits), or by
        `refresh_cone()`."""
        sig = self._state_signature()
        if self._cone is None or sig != self._cone_sig:
            self._cone = self._build_cone()
            self._cone_sig = self._state_signature()  # (building clamps the readout positions in place)
        return self._cone

    def refresh_cone(self):
        self._cone = self._ff = None

    def _apply(self, fn, *args, **kwargs):
        self._cone = self._ff = None
        return super()._apply(fn, *args, **kwargs)

    def train(self, mode=True):
        self._cone = self._ff = None
        return super().train(mode)

    def _build_cone(self):
        from .. import ops
        core, ro = self.core, self.readout
        if self.training:
            raise NotImplementedError("laminr_b200: the conv-core response kernel implements EVAL mode (running batch-norm statistics, readout at "
                                      "its mean position); call model.eval() first (SURVEY 8a row a-9)")
        L = len(core.features)
        if list(core.stack) != [L - 1]:
            raise NotImplementedError("laminr_b200: only stack=-1 (last layer read out) is covered by the conv-core kernel")
        layers = []
        for l, feat in enumerate(core.features):
            conv = feat.conv
            k = conv.kernel_size[0]
            if conv.kernel_size[0] != conv.kernel_size[1] or k % 2 == 0 or tuple(conv.padding) != (k // 2, k // 2) or tuple(conv.stride) != (1, 1) \
                    or tuple(conv.dilation) != (1, 1) or conv.groups != 1:
                raise NotImplementedError("laminr_b200: the conv-core kernel covers stride-1, dilation-1, odd square kernels with 'same' zero padding")
            bn = None
            if hasattr(feat, "norm"):
                n = feat.norm
                bn = (n.running_mean, n.running_var, n.weight, n.bias, n.eps)
            layers.append(dict(weight=conv.weight, conv_bias=conv.bias, bn=bn, nonlin=hasattr(feat, "nonlin")))
        with torch.no_grad():
            ro._mu.clamp_(min=-1, max=1)  # sample_grid does this in place on every call (gaussian.py:408)
        c, h, w = ro.in_shape
        return ops.ConvCone(layers, ro._mu, ro._features, ro.bias, h, w, xshift=core.elu_xshift, yshift=core.elu_yshift, offset=self.offset,
                            align_corners=ro.align_corners)

    def forward_neurons(self, x, idxs):
        """Responses [B, len(idxs)] of the listed neurons (same values as forward(x)[:, idxs]); differentiable wrt x."""
        from .. import ops
        cone = self.cone()
        if tuple(x.shape[1:]) != (cone.c_in, cone.desc.height, cone.desc.width):
            raise ValueError(f"expected images [B,{cone.c_in},{cone.desc.height},{cone.desc.width}], got {tuple(x.shape)}")
        nog = torch.as_tensor(list(idxs), dtype=torch.int32, device=x.device)
        return ops.convresp(x, cone, nog).t()  # raises on CPU tensors: no fallback

    def full_frame(self):
        """The lmr_ff_forward weight pack (csrc/fullframe.cu): the whole frame through the core once, every neuron read out in one launch.  Cached and
        invalidated like `cone()`."""
        from .. import ops
        sig = self._state_signature()
        if self._ff is None or sig != self._ff_sig:
            if self.training:
                raise NotImplementedError("laminr_b200: the full-frame kernels implement EVAL mode; call model.eval() first")
            core, ro = self.core, self.readout
            if list(core.stack) != [len(core.features) - 1]:
                raise NotImplementedError("laminr_b200: only stack=-1 (last layer read out) is covered by the full-frame kernels")
            specs = []
            for feat in core.features:
                conv = feat.conv
                if tuple(conv.stride) != (1, 1) or tuple(conv.dilation) != (1, 1) or conv.groups != 1 or conv.kernel_size[0] != conv.kernel_size[1] \
                        or conv.padding[0] != conv.padding[1]:
                    raise NotImplementedError("laminr_b200: the full-frame kernels cover stride-1, dilation-1 square convolutions for this core")
                n = getattr(feat, "norm", None)
                specs.append(dict(kind="dense", weight=conv.weight, pad=conv.padding[0], conv_bias=conv.bias,
                                  bn=None if n is None else (n.running_mean, n.running_var, n.weight, n.bias, n.eps),
                                  elu=(core.elu_xshift, core.elu_yshift) if hasattr(feat, "nonlin") else None))
            with torch.no_grad():
                ro._mu.clamp_(min=-1, max=1)
            c, h, w = ro.in_shape
            self._ff = ops.FullFrameNet(specs, (core.input_channels, h, w), ro._mu, ro._features, ro.bias, offset=self.offset, align_corners=ro.align_corners)
            self._ff_sig = self._state_signature()
        return self._ff

    def forward(self, inputs, *args, shift=None, detach_core=False, **kwargs):
        """firing_rate.py:36-52: elu(readout(core(x)) + offset) + 1 for all neurons -> [B, n]; one full-frame evaluation of the core."""
        from .. import ops
        if shift is not None or detach_core:
            raise NotImplementedError("laminr_b200: shift / detach_core are not covered by the conv-core kernel")
        return ops.fullframe_response(inputs, self.full_frame())


def export_arrays(model):
    """Flat numpy export of a FiringRateEncoder(Stacked2dCore, FullGaussian2d) (the layout oracle.convcore_oracle.model_from_arrays and
    oracle.torch_port.ConvResponsePort read; same keys as oracle/make_golden.py:export_convcore_arrays writes for the reference's modules)."""
    core, ro = model.core, model.readout
    npy = lambda t: t.detach().cpu().numpy().copy()
    d = dict(n_layers=np.int64(len(core.features)), xs=np.float64(core.elu_xshift), ys=np.float64(core.elu_yshift), offset=np.float64(model.offset),
             align_corners=np.bool_(ro.align_corners), mu=npy(ro._mu).reshape(-1, 2).clip(-1, 1), features=npy(ro._features).reshape(ro._features.shape[1], -1))
    if ro.bias is not None:
        d["bias"] = npy(ro.bias)
    for l, feat in enumerate(core.features):
        d[f"w{l}"] = npy(feat.conv.weight)
        if feat.conv.bias is not None:
            d[f"b{l}"] = npy(feat.conv.bias)
        if hasattr(feat, "norm"):
            d[f"bn{l}_mean"], d[f"bn{l}_var"] = npy(feat.norm.running_mean), npy(feat.norm.running_var)
            d[f"bn{l}_gamma"], d[f"bn{l}_beta"], d[f"bn{l}_eps"] = npy(feat.norm.weight), npy(feat.norm.bias), np.float64(feat.norm.eps)
        d[f"nonlin{l}"] = np.bool_(hasattr(feat, "nonlin"))
    return d


def conv_core_gaussian_model(input_shape=(1, 200, 200), n_neurons=128, hidden_channels=32, input_kern=9, hidden_kern=7, layers=3, seed=0):
    """BASELINE config 4: random-init stacked-conv core (1 -> 32 channels, kernels 9/7/7, last layer read out) + full-Gaussian
    readout with `n_neurons` outputs, as a firing-rate encoder in EVAL mode (batch-norm running statistics, readout at its mean positions)."""
    torch.manual_seed(seed)
    c, h, w = input_shape
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        core = Stacked2dCore(input_channels=c, hidden_channels=hidden_channels, input_kern=input_kern, hidden_kern=hidden_kern, layers=layers, stack=-1,
                             pad_input=True)
        readout = FullGaussian2d(in_shape=(hidden_channels, h, w), outdims=n_neurons, bias=True)
    return FiringRateEncoder(core, readout).eval()
