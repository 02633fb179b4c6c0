"""RF masks.  Mirror of laminr/utils/mask.py:7-64.

`apply_mask` is differentiable elementwise glue (the fused learn path uses SimCLROnGrid.masked_reg_term instead).
`create_mask` (z-score threshold -> binary closing -> largest 8-connected component -> convex hull -> Gaussian blur ->
centroid) runs on the device through `lmr_create_mask` (csrc/mask.cu), batched over MEIs by `create_masks`; the
reference does this on the host with scipy.ndimage + skimage.morphology once per neuron (mask.py:49-64).  There is no
host implementation here (the scipy restatement lives in oracle/mask_oracle.py as the checker)."""
import torch

from .. import ops as _ops


def rescale(x, in_min, in_max, out_min, out_max):
    in_mean, out_mean = (in_min + in_max) / 2, (out_min + out_max) / 2
    gain = (out_max - out_min) / (in_max - in_min)
    return (x - in_mean) * gain + out_mean


def apply_mask(mei, mask, pixel_min, pixel_max):
    mei = rescale(mei, pixel_min, pixel_max, -1, 1)
    mask = mask.repeat(len(mei), 1, 1, 1).reshape(mei.shape)
    return rescale(mei * mask, -1, 1, pixel_min, pixel_max)


def create_masks(meis, zscore_thresh=1.5, closing_iters=2, gaussian_sigma=1.5):
    """Batched create_mask: meis [G,H,W] (channel-averaged) on a CUDA device -> (masks [G,H,W] device tensor, centres list of
    {"y","x"} dicts as mask.py:61 computes them).  One launch group for all G images; raises like the reference (argmax of an
    empty sequence, mask.py:54) when an MEI has no pixel above the threshold."""
    masks, stats = _ops.raw_create_mask(meis, zscore_thresh, closing_iters, gaussian_sigma)
    st = stats.cpu().numpy().astype("int64") & 0xFFFFFFFF
    if (st[:, 3] != 0).any():
        raise ValueError(f"create_mask: no pixel above the z-score threshold for image(s) {[int(i) for i in (st[:, 3] != 0).nonzero()[0]]} "
                         "(attempt to get argmax of an empty sequence)")
    centres = [{"y": float(st[i, 1]) / float(st[i, 0]) + 0.5, "x": float(st[i, 2]) / float(st[i, 0]) + 0.5} for i in range(st.shape[0])]
    return masks, centres


def create_mask(mei, zscore_thresh=1.5, closing_iters=2, gaussian_sigma=1.5, return_mask_centroid=False):
    """Signature of laminr/utils/mask.py:25-31.  `mei`: tensor that squeezes to [H,W]; must live on (or is moved to) the current CUDA
    device -- the mask is built by the CUDA kernel; the returned mask is a host tensor as in the reference (`torch.Tensor(smoothed)`)."""
    if not torch.is_tensor(mei):
        mei = torch.as_tensor(mei)
    x = mei.detach().squeeze().to(torch.float32)
    if not x.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("laminr_b200.create_mask runs on CUDA (sm_100a) only; there is no CPU fallback.")
        x = x.cuda()
    masks, centres = create_masks(x[None], zscore_thresh, closing_iters, gaussian_sigma)
    mask = masks[0].cpu()
    return (mask, centres[0]) if return_mask_centroid else mask
