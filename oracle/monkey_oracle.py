"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the macaque-V1 response models on plain torch functional ops, from a
state_dict of the per-session model (the published checkpoint format).  Pinned by tests/golden/monkey_*.npz, which oracle/make_golden_monkey.py
generates from the reference's own classes constructed offline.  Nothing here is imported by the product (laminr_b200/).

Restates, for eval mode:
  * SE2dCore.forward               laminr/neuron_models/utils/monkey_model.py:278-286 over the layers built at :213-276
      layer0: conv(k = input_kern, no padding) -> batch norm (running stats) -> ELU
      layer l: 1x1 conv -> depthwise conv(k = hidden_kern, dilation, padding ((k-1)*dilation+1)//2) -> 1x1 conv -> batch norm -> ELU
               [-> SQ_EX_Block :95-108: x * sigmoid(W2 relu(W1 mean_hw(x) + b1) + b2) on the last n_se_blocks layers]
  * PointPooled2d.forward          neuralpredictors/layers/readouts/point_pooled.py:121-169
  * FullGaussian2d.forward (eval)  neuralpredictors/layers/readouts/gaussian.py:532-581 with the grid of :396-429 at its mean
  * the keyless merge              laminr/neuron_models/monkey_models.py:262-294 / :313-345 and the ensemble mean :297-310 / :348-361
"""
import numpy as np
import torch
import torch.nn.functional as F

CONFIG = dict(hidden_dilation=2, hidden_kern=9, pool_kern=3, pool_steps=2, bn_eps=1e-5)


def _bn(x, sd, p, eps):
    m, v, g, b = (sd[p + k].to(x.dtype) for k in ("running_mean", "running_var", "weight", "bias"))
    return (x - m[None, :, None, None]) / torch.sqrt(v[None, :, None, None] + eps) * g[None, :, None, None] + b[None, :, None, None]


def core_forward(sd, x, cfg=CONFIG, prefix="core.features."):
    l = 0
    while True:
        p = f"{prefix}layer{l}."
        if p + "conv.weight" in sd:
            x = F.conv2d(x, sd[p + "conv.weight"].to(x.dtype))
        elif p + "ds_conv.in_depth_conv.weight" in sd:
            k, d = cfg["hidden_kern"], cfg["hidden_dilation"]
            x = F.conv2d(x, sd[p + "ds_conv.in_depth_conv.weight"].to(x.dtype))
            w = sd[p + "ds_conv.spatial_conv.weight"].to(x.dtype)
            x = F.conv2d(x, w, padding=((k - 1) * d + 1) // 2, dilation=d, groups=w.shape[0])
            x = F.conv2d(x, sd[p + "ds_conv.out_depth_conv.weight"].to(x.dtype))
        else:
            return x
        x = F.elu(_bn(x, sd, p + "norm.", cfg["bn_eps"]))
        if p + "seg_ex_block.se.1.weight" in sd:
            s = x.flatten(2).mean(-1)
            s = torch.relu(s @ sd[p + "seg_ex_block.se.1.weight"].to(x.dtype).t() + sd[p + "seg_ex_block.se.1.bias"].to(x.dtype))
            s = torch.sigmoid(s @ sd[p + "seg_ex_block.se.3.weight"].to(x.dtype).t() + sd[p + "seg_ex_block.se.3.bias"].to(x.dtype))
            x = x * s[:, :, None, None]
        l += 1


def merged_readout(sd, names=("grid", "features", "bias")):
    """The per-session readout parameters concatenated over sessions in state_dict order (the keyless merge)."""
    keys = []
    for k in sd:
        if k.startswith("readout.") and k.endswith("." + names[2]):
            keys.append(k.split(".")[1])
    cat = lambda name, dim: torch.cat([sd[f"readout.{k}.{name}"] for k in keys], dim=dim)
    return cat(names[0], 1), cat(names[1], 3), cat(names[2], 0)


def point_readout(feat, grid, features, bias, cfg=CONFIG):
    N, n = feat.shape[0], grid.shape[1]
    g = grid.to(feat.dtype).clamp(-1, 1).expand(N, n, 1, 2)
    pools = [F.grid_sample(feat, g, align_corners=True)]
    for _ in range(features.shape[1] // feat.shape[1] - 1):
        feat = F.avg_pool2d(feat, cfg["pool_kern"], stride=cfg["pool_kern"])
        pools.append(F.grid_sample(feat, g, align_corners=True))
    y = torch.cat(pools, dim=1).squeeze(-1)
    return (y * features.to(feat.dtype).view(1, -1, n)).sum(1) + bias.to(feat.dtype)


def keyless_pointpool(sd, x, cfg=CONFIG):
    grid, features, bias = merged_readout(sd)
    return F.elu(point_readout(core_forward(sd, x, cfg), grid, features, bias, cfg)) + 1


def keyless_gaussian(sd, x, mu, cfg=CONFIG):
    """`mu` [1, n, 1, 2]: the merged readout's own positions (the reference's merge leaves them at their construction-time draw)."""
    _, features, bias = merged_readout(sd, names=("_mu", "_features", "bias"))
    return F.elu(point_readout(core_forward(sd, x, cfg), mu, features, bias, cfg)) + 1


def session_pointpool(sd, x, key, cfg=CONFIG, offset=0.0):
    r = f"readout.{key}."
    return F.elu(point_readout(core_forward(sd, x, cfg), sd[r + "grid"], sd[r + "features"], sd[r + "bias"], cfg) + offset) + 1


def perturb_state(model, seed):
    """Moves a freshly initialised per-session model (the reference's or the mirror's: same state_dict keys) away from its trivial initial state --
    batch-norm statistics and affine terms, squeeze-excitation weights, readout positions / features / biases -- deterministically from `seed`."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda t, s: torch.randn(t.shape, generator=g) * s
    ru = lambda t, a, b: torch.rand(t.shape, generator=g) * (b - a) + a
    sd = model.state_dict()
    new = {}
    for k, t in sd.items():
        if k.endswith("running_mean"):
            new[k] = rn(t, 0.3)
        elif k.endswith("running_var"):
            new[k] = ru(t, 0.5, 1.5)
        elif ".norm.weight" in k:
            new[k] = ru(t, 0.7, 1.3)
        elif ".norm.bias" in k:
            new[k] = rn(t, 0.2)
        elif ".seg_ex_block." in k:
            new[k] = t * 3 + rn(t, 0.3)
        elif k.endswith(".grid") or k.endswith("._mu"):
            new[k] = ru(t, -0.97, 0.97)
        elif k.endswith(".features") or k.endswith("._features"):
            new[k] = rn(t, 0.25)
        elif k.startswith("readout.") and k.endswith(".bias"):
            new[k] = rn(t, 0.3)
    sd.update(new)
    model.load_state_dict(sd)
    return model


def sd_to_numpy(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def sd_from_numpy(d, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.asarray(d[k])) for k in d if k.startswith(prefix)}
