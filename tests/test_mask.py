"""create_mask (laminr/utils/mask.py:25-64).  CPU: the scipy restatement (oracle/mask_oracle.py) against masks produced by the unmodified
reference function (tests/golden/create_mask.npz, oracle/make_golden.py::case_create_mask) and against an independent exact-integer hull.
GPU: lmr_create_mask (csrc/mask.cu) through the C ABI against the golden masks / centroids and the oracle on fresh random images."""
import numpy as np
import pytest
import torch

from oracle import mask_oracle as MO
from tests._util import golden

CASES = ["a", "b", "c"]


def _meis(d, tag):
    n, H, W, seed = (int(v) for v in d[f"{tag}_shape_seed"])
    return MO.synthetic_meis(n, H, W, seed)


@pytest.mark.parametrize("tag", CASES)
@pytest.mark.parametrize("zt", [0.3, 1.5])
def test_oracle_matches_reference_golden(tag, zt):
    d = golden("create_mask")
    meis = _meis(d, tag)
    for i, mei in enumerate(meis):
        mask, c, _ = MO.create_mask(mei, zscore_thresh=zt)
        np.testing.assert_array_equal(mask, d[f"{tag}_mask_z{zt}"][i])  # same scipy calls on the same float32 input: bit-identical
        assert (c["y"], c["x"]) == tuple(d[f"{tag}_centre_z{zt}"][i])


def test_exact_integer_hull_equals_qhull_hull():
    d = golden("create_mask")
    for tag in ("a", "c"):
        for mei in _meis(d, tag)[:6]:
            for zt in (0.3, 1.5):
                _, _, h1 = MO.create_mask(mei, zscore_thresh=zt)
                _, _, h2 = MO.create_mask(mei, zscore_thresh=zt, hull_fn=MO.convex_hull_exact)
                np.testing.assert_array_equal(h1, h2)


def test_gaussian_weights_match_scipy():
    from scipy.ndimage import correlate1d, gaussian_filter1d

    from laminr_b200.ops import gaussian_blur_weights

    for sigma in (0.7, 1.5, 3.2):
        w, r = gaussian_blur_weights(sigma)
        full = np.concatenate([w, w[:-1][::-1]])
        x = np.random.RandomState(0).rand(50)
        np.testing.assert_array_equal(correlate1d(x, full, mode="reflect"), gaussian_filter1d(x, sigma))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", CASES)
@pytest.mark.parametrize("zt", [0.3, 1.5])
def test_gpu_create_mask_golden(tag, zt):
    from laminr_b200.utils.mask import create_mask, create_masks

    d = golden("create_mask")
    meis = _meis(d, tag)
    masks, centres = create_masks(torch.from_numpy(meis).cuda(), zscore_thresh=zt)
    got = masks.cpu().numpy()
    want = d[f"{tag}_mask_z{zt}"]
    # integer pipeline up to the hull is exact; the blur accumulates in float64 in scipy's order and rounds to float32 like scipy
    assert np.abs(got - want).max() <= 1e-7, np.abs(got - want).max()
    assert ((got > 0.5) == (want > 0.5)).all()
    for i, c in enumerate(centres):
        assert abs(c["y"] - d[f"{tag}_centre_z{zt}"][i][0]) < 1e-9 and abs(c["x"] - d[f"{tag}_centre_z{zt}"][i][1]) < 1e-9
    # the single-image signature of the reference (host tensor out)
    m0, c0 = create_mask(torch.from_numpy(meis[0])[None], zscore_thresh=zt, return_mask_centroid=True)
    assert not m0.is_cuda and np.abs(m0.numpy() - want[0]).max() <= 1e-7 and abs(c0["x"] - centres[0]["x"]) < 1e-12
    m1 = create_mask(torch.from_numpy(meis[1]).cuda(), zscore_thresh=zt)
    assert np.abs(m1.numpy() - want[1]).max() <= 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,n,sigma,iters", [(100, 100, 64, 1.5, 2), (200, 200, 8, 1.5, 2), (23, 61, 32, 0.8, 1), (64, 48, 16, 3.0, 3), (16, 16, 40, 1.5, 0)])
def test_gpu_create_mask_vs_oracle_random(H, W, n, sigma, iters):
    from laminr_b200 import ops

    meis = MO.synthetic_meis(n, H, W, seed=H * 1000 + W)
    for zt in (0.3, 1.0):
        masks, stats = ops.raw_create_mask(torch.from_numpy(meis).cuda(), zt, iters, sigma)
        got, st = masks.cpu().numpy(), stats.cpu().numpy().astype(np.int64)
        for i in range(n):
            want, c, hull = MO.create_mask(meis[i], zscore_thresh=zt, closing_iters=iters, gaussian_sigma=sigma) if iters > 0 else _no_closing(meis[i], zt, sigma)
            assert st[i, 3] == 0 and st[i, 0] == hull.sum(), (i, st[i], hull.sum())
            assert abs(st[i, 1] / st[i, 0] + 0.5 - c["y"]) < 1e-9 and abs(st[i, 2] / st[i, 0] + 0.5 - c["x"]) < 1e-9
            assert np.abs(got[i] - want).max() <= 1e-7, (i, np.abs(got[i] - want).max())


def _no_closing(mei, zt, sigma):
    """closing_iters=0 means 'no closing' in the kernel; scipy's binary_closing(iterations=0) would iterate to convergence instead."""
    from scipy import ndimage

    norm = (mei - mei.mean()) / mei.std()
    lab = MO.label8(np.abs(norm) > zt)
    hull = MO.convex_hull_image(lab == np.argmax(np.bincount(lab.ravel())[1:]) + 1)
    py, px = (c.mean() + 0.5 for c in np.nonzero(hull))
    return ndimage.gaussian_filter(hull.astype(np.float32), sigma=sigma), {"y": float(py), "x": float(px)}, hull


@pytest.mark.gpu
def test_gpu_create_mask_empty_raises():
    from laminr_b200.utils.mask import create_masks

    x = torch.zeros(2, 12, 12, device="cuda")
    x[0, 5, 5] = 1.0
    with pytest.raises(ValueError, match="empty sequence"):
        create_masks(x, zscore_thresh=0.3)  # image 1 is constant: std 0 -> NaN z-scores -> nothing above the threshold
