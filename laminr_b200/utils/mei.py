"""MEI generation.  Mirror of laminr/utils/mei.py:9-120 (`generate_mei`, `get_mei_dict`).

For simulated populations (`ArbitraryMultiNeuronModel`) all requested neurons are optimised at once by ONE launch of
lmr_mei_ascent (one CTA per neuron, image resident in shared memory for all iterations); the reference loops neurons
sequentially with ~10 launches and 2 host syncs per iteration.  Initial images are drawn from the CPU generator in the
reference's order (mei.py:33,93), so a seed gives the same starting points.  Any other nn.Module goes through the generic
`featurevis.gradient_ascent`."""
import torch

from .. import _lib, featurevis
from .. import ops as _ops
from ..featurevis import ops as fops
from ..featurevis.utils import Compose
from ..neuron_models.utils.simulated_model import ArbitraryMultiNeuronModel
from .general import SingleNeuronModel, infer_device
from .mask import create_mask, create_masks


def _check(std, norm):
    if std is None and norm is None:
        raise ValueError("Either std or norm should be provided.")
    if std is not None and norm is not None:
        raise ValueError("Only one of std or norm should be provided.")


def generate_mei(model, input_shape, pixel_min, pixel_max, std=None, norm=None, gb=None, zscore=0.3, step_size=10, num_iterations=1000,
                 print_iters=1001):
    """Generic single-neuron path (any differentiable `model`: [1,C,H,W] -> scalar)."""
    items = list(model.parameters()) + list(model.buffers())
    device = items[0].device
    _check(std, norm)
    mean = (pixel_min + pixel_max) / 2
    initial_image = torch.randn(1, *input_shape, dtype=torch.float32).to(device) + mean
    renorm = fops.ChangeStats(std=std, mean=mean) if std is not None else fops.ChangeNorm(norm=norm, mean=mean)
    post_update = Compose([renorm, fops.ClipRange(pixel_min, pixel_max)])
    initial_image = post_update(initial_image)
    if gb is not None:
        gb = fops.GaussianBlur(gb)
    mei, act, _ = featurevis.gradient_ascent(model, initial_image, step_size=step_size, num_iterations=num_iterations, post_update=post_update,
                                             gradient_f=gb, print_iters=print_iters)
    mei = mei[0]
    mask, mask_center = create_mask(mei.mean(dim=0), return_mask_centroid=True, zscore_thresh=zscore)
    return mei.cpu().data.numpy(), act[-1], mask.cpu().data.numpy(), mask_center


def batched_simulated_meis(model, neuron_idxs, initial_images, pixel_min, pixel_max, std=None, norm=None, step_size=10, num_iterations=1000):
    """All neurons in one kernel launch.  initial_images: [G,C,H,W] on the model's device, BEFORE the first post-update.
    Returns (meis [G,C,H,W], fevals [G, iters+1]) as device tensors."""
    _check(std, norm)
    mean = (pixel_min + pixel_max) / 2
    x = initial_images.to(torch.float32).contiguous().clone()
    nog = torch.as_tensor(list(neuron_idxs), dtype=torch.int32, device=x.device)
    mode, target = (_lib.MODE_STD, std) if std is not None else (_lib.MODE_NORM, norm)
    fevals = _ops.raw_mei_ascent(x, model.packed_filters, model.nfilt, nog, step_size, num_iterations, mode, target, mean, pixel_min, pixel_max,
                                 act_eps=model.eps, apply_post_first=True)
    return x, fevals


def batched_conv_meis(model, neuron_idxs, initial_images, pixel_min, pixel_max, std=None, norm=None, step_size=10, num_iterations=1000):
    """Gradient ascent for all listed neurons of a conv-core + Gaussian-readout encoder at once: per iteration ONE launch of
    lmr_convcore_response (response and image gradient of every neuron's receptive-field cone) and ONE lmr_ascent_update
    (x <- clip(renorm(x + step*g))), no host sync (featurevis/core.py:54-136 driven as in laminr/utils/mei.py:33-59).
    initial_images: [G,C,H,W] BEFORE the first post-update.  Returns (meis [G,C,H,W], fevals [G, iters+1]) as device tensors."""
    _check(std, norm)
    mean = (pixel_min + pixel_max) / 2
    cone = model.cone()
    x = initial_images.to(torch.float32).contiguous().clone()
    G, per_img = x.shape[0], x[0].numel()
    nog = torch.as_tensor(list(neuron_idxs), dtype=torch.int32, device=x.device)
    mode, target = (_lib.MODE_STD, std) if std is not None else (_lib.MODE_NORM, norm)
    ws = torch.empty(max(int(_lib.load().lmr_constrain_workspace_bytes(G, per_img)), 16), dtype=torch.uint8, device=x.device)
    _ops.raw_constrain_fwd(mode, x, target, mean, pixel_min, pixel_max, 1e-9, out=x, workspace=ws)  # post_update(initial_image), mei.py:42
    fev = torch.empty((num_iterations + 1, G), dtype=torch.float32, device=x.device)
    g = torch.zeros_like(x)  # every iteration rewrites the same receptive-field windows (accumulate = 2); the rest stays zero
    for it in range(num_iterations):
        _ops.raw_convcore_response(cone, x, per_img, 1, G, nog, act=fev[it], g_img=g, accumulate=2)
        _ops.raw_ascent_update(mode, x, g, step_size, target, mean, pixel_min, pixel_max, eps=1e-9, workspace=ws)
    _ops.raw_convcore_response(cone, x, per_img, 1, G, nog, act=fev[num_iterations])
    return x, fev.t().contiguous()


def get_mei_dict(response_predicting_model, input_shape, pixel_value_lower_bound, pixel_value_upper_bound, neuron_idxs=None, required_img_std=None,
                 required_img_norm=None, gb=None, zscore=0.3, step_size=10, num_iterations=1000, print_iters=1001, device=None):
    if device is None:
        device = infer_device(response_predicting_model)
    if neuron_idxs is None:
        initial_image = torch.randn(1, *input_shape, dtype=torch.float32).to(device)  # consumes the RNG like mei.py:93
        neuron_idxs = list(range(response_predicting_model(initial_image).shape[1]))

    meis_dict = {}
    batched = (batched_simulated_meis if isinstance(response_predicting_model, ArbitraryMultiNeuronModel) else
               batched_conv_meis if hasattr(response_predicting_model, "cone") else None)
    if batched is not None and gb is None:
        _check(required_img_std, required_img_norm)
        mean = (pixel_value_lower_bound + pixel_value_upper_bound) / 2
        # one CPU draw per neuron, in index order (mei.py:33 inside the loop of :98)
        x0 = torch.cat([torch.randn(1, *input_shape, dtype=torch.float32) for _ in neuron_idxs]).to(device) + mean
        meis, fevals = batched(response_predicting_model, neuron_idxs, x0, pixel_value_lower_bound, pixel_value_upper_bound,
                                              std=required_img_std, norm=required_img_norm, step_size=step_size, num_iterations=num_iterations)
        # all RF masks in one launch group on the device (the reference: one scipy/skimage host pass per neuron, mei.py:56)
        masks, centres = create_masks(meis.mean(dim=1), zscore_thresh=zscore)
        meis_host, acts_host, masks_host = meis.cpu(), fevals[:, -1].cpu().tolist(), masks.cpu()
        for idx, neuron_idx in enumerate(neuron_idxs):
            print(f"neuron_idx = {neuron_idx}: neuron number {idx}/{len(neuron_idxs)}")
            meis_dict[idx] = {"mei": meis_host[idx].numpy(), "activation": acts_host[idx], "mask": masks_host[idx].numpy(), "center_pos": centres[idx]}
        return meis_dict

    for idx, neuron_idx in enumerate(neuron_idxs):
        print(f"neuron_idx = {neuron_idx}: neuron number {idx}/{len(neuron_idxs)}")
        model = SingleNeuronModel(response_predicting_model, neuron_idx).to(device)
        mei, mei_act, mask, mask_center = generate_mei(model, input_shape, std=required_img_std, norm=required_img_norm,
                                                       pixel_min=pixel_value_lower_bound, pixel_max=pixel_value_upper_bound, gb=gb, zscore=zscore,
                                                       step_size=step_size, num_iterations=num_iterations, print_iters=print_iters)
        meis_dict[idx] = {"mei": mei, "activation": mei_act, "mask": mask, "center_pos": mask_center}
    return meis_dict
