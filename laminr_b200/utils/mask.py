"""RF masks.  Mirror of laminr/utils/mask.py:7-64.

`apply_mask` is differentiable elementwise glue (the fused learn path uses SimCLROnGrid.masked_reg_term instead).
`create_mask` is one-off host post-processing of an MEI (z-score threshold -> binary closing -> largest 8-connected
component -> convex hull -> Gaussian blur); the reference uses skimage for `label` / `convex_hull_image`, restated
here with scipy (the hull test uses pixel-corner offsets and a half-plane test at pixel centres)."""
import numpy as np
import torch


def rescale(x, in_min, in_max, out_min, out_max):
    in_mean, out_mean = (in_min + in_max) / 2, (out_min + out_max) / 2
    gain = (out_max - out_min) / (in_max - in_min)
    return (x - in_mean) * gain + out_mean


def apply_mask(mei, mask, pixel_min, pixel_max):
    mei = rescale(mei, pixel_min, pixel_max, -1, 1)
    mask = mask.repeat(len(mei), 1, 1, 1).reshape(mei.shape)
    return rescale(mei * mask, -1, 1, pixel_min, pixel_max)


def convex_hull_image(img):
    from scipy.spatial import ConvexHull

    pts = np.argwhere(img).astype(float)
    offsets = np.array([[-0.5, 0.0], [0.5, 0.0], [0.0, -0.5], [0.0, 0.5]])
    cloud = (pts[:, None, :] + offsets[None]).reshape(-1, 2)
    eq = ConvexHull(cloud).equations
    centres = np.stack(np.mgrid[: img.shape[0], : img.shape[1]], -1).reshape(-1, 2)
    return np.all(centres @ eq[:, :2].T + eq[:, 2] <= 1e-10, axis=1).reshape(img.shape)


def create_mask(mei, zscore_thresh=1.5, closing_iters=2, gaussian_sigma=1.5, return_mask_centroid=False):
    from scipy import ndimage

    mei_np = mei.detach().cpu().squeeze().numpy() if torch.is_tensor(mei) else np.asarray(mei).squeeze()
    norm_mei = (mei_np - mei_np.mean()) / mei_np.std()
    thresholded = (np.abs(norm_mei) > zscore_thresh).squeeze()
    closed = ndimage.binary_closing(thresholded, iterations=closing_iters)
    labeled = ndimage.label(closed, structure=np.ones((3, 3)))[0]
    most_frequent = np.argmax(np.bincount(labeled.ravel())[1:]) + 1
    hull = convex_hull_image(labeled == most_frequent)
    mask = torch.Tensor(ndimage.gaussian_filter(hull.astype(np.float32), sigma=gaussian_sigma))
    if return_mask_centroid:
        py, px = (c.mean() + 0.5 for c in np.nonzero(hull))
        return mask, {"y": py, "x": px}
    return mask
