"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the stacked-conv core + Gaussian readout response model.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product path never does.

Restates, in float64 numpy with explicit analytic backward passes, what the reference executes for BASELINE config 4:
  * `Stacked2dCore.forward`      neuralpredictors/layers/cores/conv2d.py:246-253 (layers built at :202-234): per layer
        conv2d (zero "same" padding k//2, stride 1) -> eval-mode BatchNorm2d -> AdaptiveELU `elu(x - xs) + ys`
        (neuralpredictors/layers/activations.py:31-51); the features of the layers in `stack` are concatenated;
  * `FullGaussian2d.forward`     neuralpredictors/layers/readouts/gaussian.py:532-581 with the eval-mode grid of `sample_grid`
        (:396-429: grid = clamp(mu, -1, 1)): bilinear `grid_sample` (zero padding) of the feature map at one position per
        neuron, dot with the neuron's feature vector, + bias;
  * `FiringRateEncoder.forward`  neuralpredictors/layers/encoders/firing_rate.py:36-52: `elu(x + offset) + 1`.
Pinned against the reference's own modules by tests/golden/convcore_*.npz (oracle/make_golden.py:case_convcore).

The oracle is FULL-FRAME like the reference (every layer evaluated on the whole image, all neurons read out); the CUDA path
evaluates only the receptive-field cone of the selected neuron, which must give the same numbers.
"""
import numpy as np


def elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def conv2d_same(x, w, bias=None):
    """x [B,Ci,H,W], w [Co,Ci,k,k] (k odd) -> [B,Co,H,W]: cross-correlation with zero padding k//2 (torch.nn.Conv2d, stride 1)."""
    B, Ci, H, W = x.shape
    Co, _, k, _ = w.shape
    p = k // 2
    xp = np.zeros((B, Ci, H + 2 * p, W + 2 * p), np.float64)
    xp[:, :, p:p + H, p:p + W] = x
    out = np.zeros((B, Co, H, W), np.float64)
    for ky in range(k):
        for kx in range(k):
            out += np.einsum("oc,bchw->bohw", w[:, :, ky, kx].astype(np.float64), xp[:, :, ky:ky + H, kx:kx + W], optimize=True)
    if bias is not None:
        out += bias.astype(np.float64)[None, :, None, None]
    return out


def conv2d_same_input_grad(g_out, w):
    """Gradient of conv2d_same wrt x: the same correlation with the kernel flipped and in/out channels swapped."""
    return conv2d_same(g_out, np.ascontiguousarray(w[:, :, ::-1, ::-1].transpose(1, 0, 2, 3)))


def core_forward(x, layers, keep=False):
    """layers: list of dict(w, b|None, bn=(running_mean, running_var, gamma, beta, eps)|None, nonlin: bool, xs, ys).
    Returns the LAST layer's feature map [B,C,H,W] (stack=-1) and, if keep, the per-layer caches for the backward."""
    h = x.astype(np.float64)
    caches = []
    for L in layers:
        v = conv2d_same(h, L["w"], L.get("b"))
        scale = 1.0
        if L.get("bn") is not None:
            rm, rv, gamma, beta, eps = L["bn"]
            inv = 1.0 / np.sqrt(rv.astype(np.float64) + eps)
            scale = (inv * gamma.astype(np.float64))[None, :, None, None]
            v = (v - rm.astype(np.float64)[None, :, None, None]) * scale + beta.astype(np.float64)[None, :, None, None]
        if L.get("nonlin", True):
            pre = v - L.get("xs", 0.0)
            out = elu(pre) + L.get("ys", 0.0)
            dact = np.where(pre > 0, 1.0, np.exp(np.minimum(pre, 0)))
        else:
            out, dact = v, 1.0
        caches.append((scale, dact))
        h = out
    return (h, caches) if keep else h


def core_backward(g_feat, layers, caches):
    g = g_feat
    for L, (scale, dact) in zip(layers[::-1], caches[::-1]):
        g = conv2d_same_input_grad(g * dact * scale, L["w"])
    return g


def bilinear_taps(mu_xy, H, W, align_corners=True):
    """torch grid_sample (bilinear, padding_mode='zeros'): mu_xy [n,2] = (x -> width, y -> height) in [-1,1].
    Returns per neuron the 4 taps ((iy, ix, weight, inbounds), ...)."""
    mu = np.clip(np.asarray(mu_xy, np.float32), -1, 1)  # sample_grid clamps (gaussian.py:408-429); unnormalisation in float32 as ATen does
    if align_corners:
        ix = ((mu[:, 0] + np.float32(1)) / np.float32(2)) * np.float32(W - 1)
        iy = ((mu[:, 1] + np.float32(1)) / np.float32(2)) * np.float32(H - 1)
    else:
        ix = ((mu[:, 0] + np.float32(1)) * np.float32(W) - np.float32(1)) / np.float32(2)
        iy = ((mu[:, 1] + np.float32(1)) * np.float32(H) - np.float32(1)) / np.float32(2)
    x0, y0 = np.floor(ix), np.floor(iy)
    taps = []
    for dy in (0, 1):
        for dx in (0, 1):
            xx, yy = x0 + dx, y0 + dy
            wgt = (1 - np.abs(ix - xx)).astype(np.float64) * (1 - np.abs(iy - yy)).astype(np.float64)
            ok = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
            taps.append((yy.astype(np.int64), xx.astype(np.int64), wgt, ok))
    return taps


def readout_forward(feat_map, mu_xy, features, bias, align_corners=True):
    """feat_map [B,C,H,W]; mu_xy [n,2]; features [C,n]; bias [n]|None -> y [B,n]."""
    B, C, H, W = feat_map.shape
    n = mu_xy.shape[0]
    samp = np.zeros((B, C, n), np.float64)
    for yy, xx, wgt, ok in bilinear_taps(mu_xy, H, W, align_corners):
        yc, xc = np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)
        samp += feat_map[:, :, yc, xc] * (wgt * ok)[None, None, :]
    y = (samp * features.astype(np.float64)[None]).sum(1)
    return y + bias.astype(np.float64)[None] if bias is not None else y


def readout_backward(g_y, shape, mu_xy, features, align_corners=True):
    """g_y [B,n] -> gradient wrt feat_map [B,C,H,W]."""
    B, C, H, W = shape
    g = np.zeros(shape, np.float64)
    for yy, xx, wgt, ok in bilinear_taps(mu_xy, H, W, align_corners):
        for j in range(mu_xy.shape[0]):
            if ok[j]:
                g[:, :, yy[j], xx[j]] += g_y[:, j, None] * features.astype(np.float64)[None, :, j] * wgt[j]
    return g


def model_forward(x, model):
    """model: dict(layers, mu [n,2], features [C,n], bias [n]|None, offset, align_corners).  Returns act [B,n] = elu(y + offset) + 1."""
    fm = core_forward(x, model["layers"])
    y = readout_forward(fm, model["mu"], model["features"], model.get("bias"), model.get("align_corners", True))
    return elu(y + model.get("offset", 0.0)) + 1.0


def response_and_grad(x, model, neuron_of_image, g_act=None):
    """act[b] = model(x)[b, neuron_of_image[b]] (SingleNeuronModel, laminr/utils/general.py:5-12) and d sum_b(g_act[b] act[b]) / dx."""
    fm, caches = core_forward(x, model["layers"], keep=True)
    y = readout_forward(fm, model["mu"], model["features"], model.get("bias"), model.get("align_corners", True))
    pre = y + model.get("offset", 0.0)
    act_all = elu(pre) + 1.0
    B = x.shape[0]
    idx = np.asarray(neuron_of_image, np.int64)
    act = act_all[np.arange(B), idx]
    g_act = np.ones(B) if g_act is None else np.asarray(g_act, np.float64)
    g_y = np.zeros_like(y)
    dpre = np.where(pre > 0, 1.0, np.exp(np.minimum(pre, 0)))
    g_y[np.arange(B), idx] = g_act * dpre[np.arange(B), idx]
    g_fm = readout_backward(g_y, fm.shape, model["mu"], model["features"], model.get("align_corners", True))
    return act, core_backward(g_fm, model["layers"], caches), act_all


def model_from_arrays(d, prefix=""):
    """Rebuilds the `model` dict from the flat arrays of a golden fixture / a state_dict export (see laminr_b200 conv_models.export_arrays)."""
    layers = []
    L = int(d[prefix + "n_layers"])
    for l in range(L):
        bn = None
        if prefix + f"bn{l}_mean" in d:
            bn = (d[prefix + f"bn{l}_mean"], d[prefix + f"bn{l}_var"], d[prefix + f"bn{l}_gamma"], d[prefix + f"bn{l}_beta"], float(d[prefix + f"bn{l}_eps"]))
        layers.append(dict(w=d[prefix + f"w{l}"], b=d[prefix + f"b{l}"] if prefix + f"b{l}" in d else None, bn=bn, nonlin=bool(d[prefix + f"nonlin{l}"]),
                           xs=float(d[prefix + "xs"]), ys=float(d[prefix + "ys"])))
    return dict(layers=layers, mu=d[prefix + "mu"], features=d[prefix + "features"], bias=d[prefix + "bias"] if prefix + "bias" in d else None,
                offset=float(d[prefix + "offset"]), align_corners=bool(d[prefix + "align_corners"]))
