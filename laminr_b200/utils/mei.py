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
from .mask import create_mask


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


def get_mei_dict(response_predicting_model, input_shape, pixel_value_lower_bound, pixel_value_upper_bound, neuron_idxs=None, required_img_std=None,
                 required_img_norm=None, gb=None, zscore=0.3, step_size=10, num_iterations=1000, print_iters=1001, device=None):
    if device is None:
        device = infer_device(response_predicting_model)
    if neuron_idxs is None:
        initial_image = torch.randn(1, *input_shape, dtype=torch.float32).to(device)  # consumes the RNG like mei.py:93
        neuron_idxs = list(range(response_predicting_model(initial_image).shape[1]))

    meis_dict = {}
    if isinstance(response_predicting_model, ArbitraryMultiNeuronModel) and gb is None:
        _check(required_img_std, required_img_norm)
        mean = (pixel_value_lower_bound + pixel_value_upper_bound) / 2
        # one CPU draw per neuron, in index order (mei.py:33 inside the loop of :98)
        x0 = torch.cat([torch.randn(1, *input_shape, dtype=torch.float32) for _ in neuron_idxs]).to(device) + mean
        meis, fevals = batched_simulated_meis(response_predicting_model, neuron_idxs, x0, pixel_value_lower_bound, pixel_value_upper_bound,
                                              std=required_img_std, norm=required_img_norm, step_size=step_size, num_iterations=num_iterations)
        meis_host, acts_host = meis.cpu(), fevals[:, -1].cpu().tolist()
        for idx, neuron_idx in enumerate(neuron_idxs):
            print(f"neuron_idx = {neuron_idx}: neuron number {idx}/{len(neuron_idxs)}")
            mask, mask_center = create_mask(meis_host[idx].mean(dim=0), return_mask_centroid=True, zscore_thresh=zscore)
            meis_dict[idx] = {"mei": meis_host[idx].numpy(), "activation": acts_host[idx], "mask": mask.numpy(), "center_pos": mask_center}
        return meis_dict

    for idx, neuron_idx in enumerate(neuron_idxs):
        print(f"neuron_idx = {neuron_idx}: neuron number {idx}/{len(neuron_idxs)}")
        model = SingleNeuronModel(response_predicting_model, neuron_idx).to(device)
        mei, mei_act, mask, mask_center = generate_mei(model, input_shape, std=required_img_std, norm=required_img_norm,
                                                       pixel_min=pixel_value_lower_bound, pixel_max=pixel_value_upper_bound, gb=gb, zscore=zscore,
                                                       step_size=step_size, num_iterations=num_iterations, print_iters=print_iters)
        meis_dict[idx] = {"mei": mei, "activation": mei_act, "mask": mask, "center_pos": mask_center}
    return meis_dict
