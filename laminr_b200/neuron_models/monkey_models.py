"""Macaque-V1 response models: squeeze-excitation core with depth-separable, dilated hidden convolutions and a
point-pooled or Gaussian readout, the per-session -> 458-neuron "keyless" wrappers and the three-model ensembles.

Host-side mirrors of
  * laminr/neuron_models/utils/monkey_model.py:95-108 (SQ_EX_Block), :114-317 (SE2dCore), :330-397 (the per-session readout dicts),
    :400-505 / :508-634 (se_core_point_readout / se_core_full_gauss_readout)
  * neuralpredictors/layers/conv.py:4-30 (DepthSeparableConv2d), layers/readouts/point_pooled.py:15-179 (PointPooled2d)
  * laminr/neuron_models/monkey_models.py:7-259 (session table and the two configurations), :262-374 (keyless wrappers, ensembles),
    :377-417 (monkey_v1)
with the reference's constructor arguments, parameter / buffer names and RNG draw order: a seed gives the reference's initial weights and the
published checkpoints (state_dicts of the per-session models) load unchanged.

The classes only build and hold parameters.  Responses are evaluated on a CUDA device by lmr_ff_forward / lmr_ff_backward (csrc/fullframe.cu):
the whole frame is run through the core -- a squeeze-excitation block pools over the entire feature map, so the single-neuron receptive-field
cone of csrc/convcone.cu does not apply -- and every neuron is read out in one launch.  Eval mode only (running batch-norm statistics); CPU
tensors raise, there is no stock-PyTorch path.

`monkey_v1()` needs the published checkpoints, which the reference downloads from the Hugging Face hub; that download is the one thing not
reproduced here.  Pass `model_paths=[...]` (three state_dict files) to build the same ensemble offline, or `pretrained=False` for the same
architecture with random initial weights.
"""
import random
import warnings
from collections import OrderedDict
from collections.abc import Iterable

import numpy as np
import torch
from torch import nn

from .conv_models import FullGaussian2d, _LaplaceBuffer

# session key -> number of recorded neurons (laminr/neuron_models/monkey_models.py:7-231); 458 in total, all sessions saw 93 x 93 grey images
_SESSIONS = (
    ("3631896544452", 32), ("3632669014376", 21), ("3632932714885", 11), ("3633364677437", 10), ("3634055946316", 21), ("3634142311627", 14),
    ("3634658447291", 5), ("3634744023164", 12), ("3635178040531", 6), ("3635949043110", 10), ("3636034866307", 20), ("3636552742293", 24),
    ("3637161140869", 22), ("3637248451650", 7), ("3637333931598", 9), ("3637760318484", 20), ("3637851724731", 14), ("3638367026975", 16),
    ("3638456653849", 2), ("3638885582960", 5), ("3638373332053", 10), ("3638541006102", 17), ("3638802601378", 7), ("3638973674012", 18),
    ("3639060843972", 12), ("3639406161189", 14), ("3640011636703", 2), ("3639664527524", 18), ("3639492658943", 17), ("3639749909659", 8),
    ("3640095265572", 25), ("3631807112901", 29),
)
data_info = {
    key: {"input_dimensions": torch.Size([128, 1, 93, 93]), "input_channels": 1, "output_dimension": n, "img_mean": 114.54466, "img_std": 64.07781}
    for key, n in _SESSIONS
}
N_NEURONS = sum(n for _, n in _SESSIONS)

_shared_config = dict(pad_input=False, stack=-1, depth_separable=True, input_kern=24, gamma_input=10, gamma_readout=0.5, hidden_dilation=2, hidden_kern=9,
                      se_reduction=16, hidden_channels=32, linear=False)
pointpool_model_config = dict(_shared_config, n_se_blocks=2)
gaussian_model_config = dict(_shared_config, n_se_blocks=0)

HF_REPO = "mobashiri/bashiri_baroni_iclr2025"


def _no_torch_path(what):
    raise NotImplementedError(f"laminr_b200: {what} is evaluated only inside a response model on a CUDA device (csrc/fullframe.cu); "
                              "there is no stock-PyTorch / CPU path")


def set_random_seed(seed, deterministic=True):
    """monkey_model.py:63-92: python, numpy and torch generators (+ the cuDNN determinism switches the reference flips)."""
    random.seed(seed)
    np.random.seed(seed)
    if deterministic:
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.deterministic = True
    torch.manual_seed(seed)


class GlobalAvgPool(nn.Module):
    def forward(self, x):
        _no_torch_path("GlobalAvgPool")


class SQ_EX_Block(nn.Module):
    """x * sigmoid(W2 relu(W1 mean_hw(x) + b1) + b2); the parameter-free stages keep their Sequential slots so the Linear layers are `se.1` / `se.3`."""

    def __init__(self, in_ch, reduction=16):
        super().__init__()
        self.se = nn.Sequential(GlobalAvgPool(), nn.Linear(in_ch, in_ch // reduction), nn.ReLU(inplace=True), nn.Linear(in_ch // reduction, in_ch), nn.Sigmoid())

    def forward(self, x):
        _no_torch_path("SQ_EX_Block")


class DepthSeparableConv2d(nn.Sequential):
    """1 x 1 -> per-channel k x k (dilated) -> 1 x 1."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, bias=True):
        super().__init__()
        self.add_module("in_depth_conv", nn.Conv2d(in_channels, out_channels, 1, bias=bias))
        self.add_module("spatial_conv", nn.Conv2d(out_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation, bias=bias,
                                                  groups=out_channels))
        self.add_module("out_depth_conv", nn.Conv2d(out_channels, out_channels, 1, bias=bias))

    def forward(self, x):
        _no_torch_path("DepthSeparableConv2d")


def _conv_spec(kind, conv, **extra):
    if tuple(conv.stride) != (1, 1) or conv.kernel_size[0] != conv.kernel_size[1] or conv.padding[0] != conv.padding[1] or conv.dilation[0] != conv.dilation[1]:
        raise NotImplementedError("laminr_b200: the full-frame kernels cover stride-1 square convolutions")
    return dict(kind=kind, weight=conv.weight, pad=conv.padding[0], dilation=conv.dilation[0], conv_bias=conv.bias, **extra)


class SE2dCore(nn.Module):
    def __init__(self, input_channels, hidden_channels, input_kern, hidden_kern, layers=3, gamma_input=0.0, gamma_hidden=0.0, skip=0,
                 final_nonlinearity=True, final_batch_norm=True, bias=False, momentum=0.1, pad_input=True, batch_norm=True, hidden_dilation=1,
                 laplace_padding=None, input_regularizer="LaplaceL2norm", stack=None, se_reduction=32, n_se_blocks=1, depth_separable=False,
                 attention_conv=False, linear=False, first_layer_stride=1):
        super().__init__()
        assert not bias or not batch_norm, "bias and batch_norm should not both be true"
        assert not depth_separable or not attention_conv, "depth_separable and attention_conv should not both be true"
        if attention_conv or skip > 1 or first_layer_stride != 1 or input_regularizer not in ("LaplaceL2norm", "LaplaceL2"):
            raise NotImplementedError("laminr_b200 SE2dCore: attention convolutions, skip connections and strided first layers are not mirrored")
        self._input_weights_regularizer = _LaplaceBuffer()
        self.layers, self.depth_separable, self.gamma_input, self.gamma_hidden = layers, depth_separable, gamma_input, gamma_hidden
        self.input_channels, self.hidden_channels, self.skip, self.n_se_blocks = input_channels, hidden_channels, skip, n_se_blocks
        self.stack = range(layers) if stack is None else ([*range(layers)[stack:]] if isinstance(stack, int) else stack)
        self.features = nn.Sequential()
        if not isinstance(hidden_kern, Iterable):
            hidden_kern = [hidden_kern] * (layers - 1)
        for l in range(layers):  # monkey_model.py:213-276
            layer = OrderedDict()
            if l == 0:
                layer["conv"] = nn.Conv2d(input_channels, hidden_channels, input_kern, padding=input_kern // 2 if pad_input else 0, bias=bias)
            else:
                pad = ((hidden_kern[l - 1] - 1) * hidden_dilation + 1) // 2
                if depth_separable:
                    layer["ds_conv"] = DepthSeparableConv2d(hidden_channels, hidden_channels, kernel_size=hidden_kern[l - 1], dilation=hidden_dilation,
                                                            padding=pad, bias=False, stride=1)
                else:
                    layer["conv"] = nn.Conv2d(hidden_channels, hidden_channels, hidden_kern[l - 1], padding=pad, bias=bias, dilation=hidden_dilation)
            last = l == layers - 1
            if batch_norm and (l == 0 or final_batch_norm or not last):
                layer["norm"] = nn.BatchNorm2d(hidden_channels, momentum=momentum)
            if not linear and ((layers > 1 or final_nonlinearity) if l == 0 else (final_nonlinearity or not last)):
                layer["nonlin"] = nn.ELU(inplace=True)
            if l > 0 and (layers - l) <= n_se_blocks:
                layer["seg_ex_block"] = SQ_EX_Block(in_ch=hidden_channels, reduction=se_reduction)
            self.features.add_module(f"layer{l}", nn.Sequential(layer))
        for m in self.modules():  # Core.init_conv via self.apply (cores/base.py:17-30): same order over the convolutions
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_normal_(m.weight.data)
                if m.bias is not None:
                    m.bias.data.fill_(0)

    def forward(self, x):
        _no_torch_path("SE2dCore")

    @property
    def outchannels(self):
        return len(self.features) * self.hidden_channels

    def out_shape(self, in_shape):
        """(channels, height, width) the readouts see for an input [.., c, h, w] (what the reference finds by pushing zeros through the core)."""
        h, w = in_shape[-2:]
        for feat in self.features:
            convs = [feat.conv] if hasattr(feat, "conv") else list(feat.ds_conv)
            for c in convs:
                h = h + 2 * c.padding[0] - c.dilation[0] * (c.kernel_size[0] - 1)
                w = w + 2 * c.padding[1] - c.dilation[1] * (c.kernel_size[1] - 1)
        return torch.Size([self.hidden_channels * len(self.stack), h, w])

    def layer_specs(self):
        """The layer list ops.FullFrameNet packs."""
        if list(self.stack) != [len(self.features) - 1]:
            raise NotImplementedError("laminr_b200: only stack=-1 (last layer read out) is covered by the full-frame kernels")
        specs = []
        for feat in self.features:
            post = {}
            if hasattr(feat, "norm"):
                n = feat.norm
                post["bn"] = (n.running_mean, n.running_var, n.weight, n.bias, n.eps)
            if hasattr(feat, "nonlin"):
                post["elu"] = (0.0, 0.0)
            if hasattr(feat, "ds_conv"):
                ds = feat.ds_conv
                if ds.in_depth_conv.bias is not None:
                    raise NotImplementedError("laminr_b200: depth-separable convolutions with biases are not packed")
                specs += [_conv_spec("pointwise", ds.in_depth_conv), _conv_spec("depthwise", ds.spatial_conv), _conv_spec("pointwise", ds.out_depth_conv, **post)]
            else:
                conv = feat.conv
                specs.append(_conv_spec("dense" if conv.groups == 1 else "depthwise", conv, **post))
            if hasattr(feat, "seg_ex_block"):
                se = feat.seg_ex_block.se
                specs.append(dict(kind="se", w1=se[1].weight, b1=se[1].bias, w2=se[3].weight, b2=se[3].bias))
        return specs


class PointPooled2d(nn.Module):
    def __init__(self, in_shape, outdims, pool_steps, bias, pool_kern, init_range, align_corners=True, mean_activity=None, feature_reg_weight=1.0,
                 gamma_readout=None, **kwargs):
        super().__init__()
        if init_range > 1.0 or init_range <= 0.0:
            raise ValueError("init_range is not within required limit!")
        self._pool_steps, self.in_shape, self.outdims = pool_steps, in_shape, outdims
        self.feature_reg_weight = feature_reg_weight if gamma_readout is None else gamma_readout
        self.mean_activity, self.pool_kern, self.init_range, self.align_corners = mean_activity, pool_kern, init_range, align_corners
        c = in_shape[0]
        self.grid = nn.Parameter(torch.empty(1, outdims, 1, 2))
        self.features = nn.Parameter(torch.empty(1, c * (pool_steps + 1), 1, outdims))
        self.bias = nn.Parameter(torch.empty(outdims)) if bias else None
        self.avg = nn.AvgPool2d((pool_kern, pool_kern), stride=pool_kern, count_include_pad=False)
        # point_pooled.py:96-105: grid ~ U(+-init_range), features = 1/c, bias = mean activity or 0
        self.grid.data.uniform_(-init_range, init_range)
        self.features.data.fill_(1 / c)
        if self.bias is not None:
            if mean_activity is None:
                warnings.warn("Readout is NOT initialized with mean activity but with 0!")
                self.bias.data.fill_(0)
            else:
                self.bias.data = mean_activity

    @property
    def pool_steps(self):
        return self._pool_steps

    def forward(self, x, shift=None, out_idx=None, **kwargs):
        _no_torch_path("PointPooled2d")


class _ReadoutDict(nn.ModuleDict):
    def forward(self, *args, **kwargs):
        _no_torch_path("a per-session readout")


class MultiplePointPooled2d(_ReadoutDict):
    def __init__(self, core, in_shape_dict, n_neurons_dict, pool_steps, pool_kern, bias, init_range, gamma_readout):
        super().__init__()
        for k in n_neurons_dict:
            self.add_module(k, PointPooled2d(core.out_shape(in_shape_dict[k]), n_neurons_dict[k], pool_steps=pool_steps, pool_kern=pool_kern, bias=bias,
                                             init_range=init_range))
        self.gamma_readout = gamma_readout


class MultipleFullGaussian2d(_ReadoutDict):
    def __init__(self, core, in_shape_dict, n_neurons_dict, init_mu_range, init_sigma_range, bias, gamma_readout):
        super().__init__()
        for k in n_neurons_dict:
            # the reference passes init_sigma_range as a stray keyword (monkey_model.py:377-385), so the readout's init_sigma stays at its default 1
            self.add_module(k, FullGaussian2d(in_shape=core.out_shape(in_shape_dict[k]), outdims=n_neurons_dict[k], init_mu_range=init_mu_range,
                                              init_sigma_range=init_sigma_range, bias=bias))
        self.gamma_readout = gamma_readout


def _readout_pack(readout):
    """(grid, features, bias, n_scales, pool_kern, align_corners) of a PointPooled2d or an eval-mode FullGaussian2d."""
    if isinstance(readout, PointPooled2d):
        with torch.no_grad():
            readout.grid.data = torch.clamp(readout.grid.data, -1, 1)  # point_pooled.py:133
        return readout.grid, readout.features, readout.bias, readout.pool_steps + 1, readout.pool_kern, readout.align_corners
    if isinstance(readout, FullGaussian2d):
        with torch.no_grad():
            readout._mu.clamp_(min=-1, max=1)  # gaussian.py:408
        return readout._mu, readout._features, readout.bias, 1, 1, readout.align_corners
    raise NotImplementedError(f"laminr_b200: no full-frame readout for {type(readout).__name__}")


class _FullFrameModel(nn.Module):
    """core + readout evaluated by the full-frame kernels; the weight pack is rebuilt when parameters / buffers change (see FiringRateEncoder.cone)."""

    def __init__(self):
        super().__init__()
        self._packs = {}
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module._packs.clear())

    def _signature(self):
        return (self.training,) + tuple((id(t), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def _apply(self, fn, *args, **kwargs):
        self._packs.clear()
        return super()._apply(fn, *args, **kwargs)

    def train(self, mode=True):
        self.__dict__.get("_packs", {}).clear()
        return super().train(mode)

    def _pack(self, key, readout, offset):
        from .. import ops
        hit = self._packs.get(key)
        if hit is not None and hit[0] == self._signature():
            return hit[1]
        if self.training:
            raise NotImplementedError("laminr_b200: the full-frame kernels implement EVAL mode (running batch-norm statistics, readout at its mean "
                                      "position); call model.eval() first")
        grid, features, bias, n_scales, pool_kern, align = _readout_pack(readout)
        in_shape = (self.core.input_channels,) + tuple(self.input_hw)
        net = ops.FullFrameNet(self.core.layer_specs(), in_shape, grid, features, bias, n_scales=n_scales, pool_kern=pool_kern, offset=offset,
                               align_corners=align)
        self._packs[key] = (self._signature(), net)
        return net


class Encoder(_FullFrameModel):
    """The per-session model the published checkpoints are state_dicts of (monkey_model.py:443-457): elu(readout[data_key](core(x)) + offset) + 1."""

    def __init__(self, core, readout, elu_offset, input_hw=(93, 93)):
        super().__init__()
        self.core, self.readout, self.offset, self.input_hw = core, readout, elu_offset, input_hw

    def forward(self, x, data_key=None, **kwargs):
        from .. import ops
        if data_key is None and len(self.readout) == 1:
            data_key = list(self.readout.keys())[0]
        return ops.fullframe_response(x, self._pack(data_key, self.readout[data_key], self.offset))


def _build(data_info, seed, readout_kind, hidden_channels=32, input_kern=13, hidden_kern=3, layers=3, gamma_input=15.5, skip=0, final_nonlinearity=True,
           momentum=0.9, pad_input=False, batch_norm=True, hidden_dilation=1, laplace_padding=None, input_regularizer="LaplaceL2norm", readout_bias=True,
           gamma_readout=4, elu_offset=0, stack=None, se_reduction=32, n_se_blocks=1, depth_separable=False, linear=False, **readout_args):
    n_neurons = {k: v["output_dimension"] for k, v in data_info.items()}
    in_shapes = {k: v["input_dimensions"] for k, v in data_info.items()}
    channels = [v["input_channels"] for v in data_info.values()][0]
    set_random_seed(seed)
    core = SE2dCore(input_channels=channels, hidden_channels=hidden_channels, input_kern=input_kern, hidden_kern=hidden_kern, layers=layers,
                    gamma_input=gamma_input, skip=skip, final_nonlinearity=final_nonlinearity, bias=False, momentum=momentum, pad_input=pad_input,
                    batch_norm=batch_norm, hidden_dilation=hidden_dilation, laplace_padding=laplace_padding, input_regularizer=input_regularizer, stack=stack,
                    se_reduction=se_reduction, n_se_blocks=n_se_blocks, depth_separable=depth_separable, linear=linear)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        if readout_kind == "point":
            readout = MultiplePointPooled2d(core, in_shape_dict=in_shapes, n_neurons_dict=n_neurons, bias=readout_bias, gamma_readout=gamma_readout,
                                            **readout_args)
        else:
            readout = MultipleFullGaussian2d(core, in_shape_dict=in_shapes, n_neurons_dict=n_neurons, bias=readout_bias, gamma_readout=gamma_readout,
                                             **readout_args)
    hw = tuple(next(iter(in_shapes.values()))[-2:])
    return Encoder(core, readout, elu_offset, input_hw=hw)


def se_core_point_readout(data_info, seed, pool_steps=2, pool_kern=3, init_range=0.2, **kwargs):
    return _build(data_info, seed, "point", pool_steps=pool_steps, pool_kern=pool_kern, init_range=init_range, **kwargs)


def se_core_full_gauss_readout(data_info, seed, init_mu_range=0.2, init_sigma=1.0, **kwargs):
    return _build(data_info, seed, "gauss", init_mu_range=init_mu_range, init_sigma_range=init_sigma, **kwargs)


def get_pointpool_monkey_model(seed=None, data_info=data_info, model_config=pointpool_model_config):
    seed = seed if seed is not None else np.random.randint(0, 100)
    return se_core_point_readout(seed=seed, data_info=data_info, **model_config)


def get_gaussian_monkey_model(seed=None, data_info=data_info, model_config=gaussian_model_config):
    seed = seed if seed is not None else np.random.randint(0, 100)
    return se_core_full_gauss_readout(seed=seed, data_info=data_info, **model_config)


class _Keyless(_FullFrameModel):
    """One readout over all sessions' neurons, in session order (monkey_models.py:262-294, :313-345); response = elu(readout(core(x))) + 1."""

    def __init__(self, model_initial, readout):
        super().__init__()
        model_initial.eval()
        start = 0
        with torch.no_grad():
            for r in model_initial.readout.values():
                end = start + r.outdims
                self._merge(readout, r, start, end)
                start = end
        self.core, self.readout, self.input_hw = model_initial.core, readout, model_initial.input_hw
        self.eval()

    def forward(self, x):
        from .. import ops
        return ops.fullframe_response(x, self._pack("all", self.readout, 0.0))


class keyless_pointpool_monkey_model(_Keyless):
    def __init__(self, model_initial):
        first = next(iter(model_initial.readout.values()))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)
            readout = PointPooled2d(first.in_shape, N_NEURONS, pool_steps=2, bias=True, pool_kern=first.pool_kern, init_range=first.init_range,
                                    align_corners=first.align_corners, mean_activity=first.mean_activity, feature_reg_weight=first.feature_reg_weight)
        super().__init__(model_initial, readout)

    @staticmethod
    def _merge(readout, r, start, end):
        readout.features.data[:, :, :, start:end] = r.features.data
        readout.grid.data[:, start:end, :, :] = r.grid.data
        readout.bias.data[start:end] = r.bias.data


class keyless_gaussian_monkey_model(_Keyless):
    def __init__(self, model_initial):
        first = next(iter(model_initial.readout.values()))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)
            readout = FullGaussian2d(in_shape=first.in_shape, outdims=N_NEURONS, bias=True, init_mu_range=first.init_mu_range, init_sigma=first.init_sigma,
                                     batch_sample=first.batch_sample, align_corners=first.align_corners, gauss_type=first.gauss_type,
                                     mean_activity=first.mean_activity, feature_reg_weight=first.feature_reg_weight)
        super().__init__(model_initial, readout)

    @staticmethod
    def _merge(readout, r, start, end):
        # Features and biases are taken over.  The positions are NOT: the reference assigns into `readout.grid.data[...]` (monkey_models.py:332), and
        # FullGaussian2d.grid is a property that returns a freshly computed tensor (gaussian.py:353-355), so that assignment never reaches `_mu` -- the
        # merged readout keeps the positions it drew at construction.  Reproduced as is (results identical to the reference's).
        readout._features.data[:, :, :, start:end] = r._features.data
        readout.bias.data[start:end] = r.bias.data


class _Ensemble(nn.Module):
    """Mean response of the member models (monkey_models.py:297-310, :348-361)."""

    def __init__(self, models_paths, get_model, keyless):
        super().__init__()
        members = []
        for path in models_paths:
            model = get_model()
            if path is not None:
                model.load_state_dict(torch.load(path))
            members.append(model)
        self.models = nn.ModuleList(keyless(m) for m in members)
        self.eval()

    def forward(self, x):
        return torch.stack([m(x) for m in self.models]).mean(dim=0)


class ensamble_keyless_pointpool_monkey_model(_Ensemble):
    def __init__(self, models_paths):
        super().__init__(models_paths, get_pointpool_monkey_model, keyless_pointpool_monkey_model)


class ensamble_keyless_gaussian_monkey_model(_Ensemble):
    def __init__(self, models_paths):
        super().__init__(models_paths, get_gaussian_monkey_model, keyless_gaussian_monkey_model)


def monkey_v1(model_type="pointpool", model_paths=None, pretrained=True):
    """monkey_models.py:377-417.  The reference fetches `monkey_<model_type>_{1,2,3}.pt` from the Hugging Face hub; here the three checkpoint files are
    passed as `model_paths`, or `pretrained=False` builds the ensemble with random initial weights."""
    if model_type not in ("pointpool", "gaussian"):
        raise ValueError("readout_type must be 'pointpool' or 'gaussian'")
    if model_paths is None:
        if pretrained:
            raise NotImplementedError(
                f"laminr_b200: the pretrained monkey-V1 ensembles are downloaded by the reference from the Hugging Face hub ({HF_REPO}, "
                f"monkey_{model_type}_1..3.pt) and there is no network here; pass model_paths=[three state_dict files] or pretrained=False")
        model_paths = [None, None, None]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        if model_type == "pointpool":
            return ensamble_keyless_pointpool_monkey_model(model_paths)
        return ensamble_keyless_gaussian_monkey_model(model_paths)
