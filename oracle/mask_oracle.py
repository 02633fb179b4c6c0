"""TEST INFRASTRUCTURE ONLY -- CPU restatement of `create_mask` (laminr/utils/mask.py:25-64).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product path
(laminr_b200/utils/mask.py -> lmr_create_mask, csrc/mask.cu) never does.

Steps, each following the reference line it restates:
  mask.py:49-50  z-score of the (squeezed) MEI with the population std (numpy default ddof=0), float32 arithmetic
  mask.py:51     |z| > zscore_thresh
  mask.py:52     scipy.ndimage.binary_closing(iterations): default 3x3 cross structure, border value 0
  mask.py:53     skimage.morphology.label(connectivity=2)   == scipy.ndimage.label(structure=ones((3,3))) (8-connectivity, raster label order)
  mask.py:54-55  most frequent non-zero label (argmax of bincount: first on ties)
  mask.py:56     skimage.morphology.convex_hull_image (offset_coordinates=True): hull of the four edge mid-points of every object pixel,
                 pixel centres tested against the hull's half planes with tolerance 1e-10
  mask.py:57     scipy.ndimage.gaussian_filter(float32 hull, sigma)
  mask.py:61     centroid = mean index of the hull pixels + 0.5

PARITY PIN: thresholding, closing, bincount, Gaussian blur and centroid are pinned by running the unmodified reference function
(tests/golden/create_mask.npz, written by oracle/make_golden.py::case_create_mask).  scikit-image (unpinned `>=0.19.1` in the reference's
pyproject) is absent from this image, so `label` and `convex_hull_image` are the published-algorithm restatements above on BOTH sides of
that golden: for these two steps parity is unpinned.  `convex_hull_exact` below is a second, independent statement of the hull step
(integer monotone chain, closed half-plane test) used to cross-check the qhull-based one.
"""
import numpy as np
from scipy import ndimage


def label8(img):
    return ndimage.label(img, structure=np.ones((3, 3)))[0]


def convex_hull_image(img):
    from scipy.spatial import ConvexHull

    pts = np.argwhere(img).astype(float)
    offsets = np.array([[-0.5, 0.0], [0.5, 0.0], [0.0, -0.5], [0.0, 0.5]])
    cloud = (pts[:, None, :] + offsets[None]).reshape(-1, 2)
    eq = ConvexHull(cloud).equations
    centres = np.stack(np.mgrid[: img.shape[0], : img.shape[1]], -1).reshape(-1, 2)
    return np.all(centres @ eq[:, :2].T + eq[:, 2] <= 1e-10, axis=1).reshape(img.shape)


def convex_hull_exact(img):
    """Same set in exact integer arithmetic: coordinates doubled (+1), Andrew's monotone chain, closed hull (boundary counts as inside)."""
    pts = set()
    for r, c in np.argwhere(img):
        R, Cc = 2 * int(r) + 1, 2 * int(c) + 1
        pts.update({(R - 1, Cc), (R + 1, Cc), (R, Cc - 1), (R, Cc + 1)})
    pts = sorted(pts)

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    def chain(seq):
        h = []
        for p in seq:
            while len(h) >= 2 and cross(h[-2], h[-1], p) <= 0:
                h.pop()
            h.append(p)
        return h

    lower, upper = chain(pts), chain(reversed(pts))
    poly = lower[:-1] + upper[:-1]  # counter-clockwise in (row, col) orientation of `cross`
    out = np.zeros(img.shape, bool)
    for r in range(img.shape[0]):
        for c in range(img.shape[1]):
            q = (2 * r + 1, 2 * c + 1)
            out[r, c] = all(cross(poly[i], poly[(i + 1) % len(poly)], q) >= 0 for i in range(len(poly)))
    return out


def create_mask(mei, zscore_thresh=1.5, closing_iters=2, gaussian_sigma=1.5, hull_fn=convex_hull_image):
    """mei: array that squeezes to [H,W] float32.  Returns (mask float32 [H,W], {"y","x"} centroid, hull bool [H,W])."""
    mei_np = np.asarray(mei, np.float32).squeeze()
    norm = (mei_np - mei_np.mean()) / mei_np.std()
    thresholded = (np.abs(norm) > zscore_thresh).squeeze()
    closed = ndimage.binary_closing(thresholded, iterations=closing_iters)
    labeled = label8(closed)
    most_frequent = np.argmax(np.bincount(labeled.ravel())[1:]) + 1
    hull = hull_fn(labeled == most_frequent)
    mask = ndimage.gaussian_filter(hull.astype(np.float32), sigma=gaussian_sigma)
    py, px = (c.mean() + 0.5 for c in np.nonzero(hull))
    return mask, {"y": float(py), "x": float(px)}, hull


def synthetic_meis(n, H, W, seed=0):
    """MEI-like test images [n,H,W] float32: Gabor patches (some clipped by the frame), plus secondary blobs and pixel noise so that the
    closing, the largest-component choice and the hull all have work to do."""
    rng = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    out = np.zeros((n, H, W), np.float32)
    for i in range(n):
        img = np.zeros((H, W))
        for b in range(1 + i % 3):
            cy, cx = rng.uniform(-0.05, 1.05) * H, rng.uniform(-0.05, 1.05) * W
            if b == 0:
                cy, cx = rng.uniform(0.1, 0.9) * H, rng.uniform(0.1, 0.9) * W
            sy, sx = rng.uniform(0.05, 0.16) * H / (1 + b), rng.uniform(0.05, 0.16) * W / (1 + b)
            th, fr, ph = rng.uniform(0, np.pi), rng.uniform(0.15, 0.9), rng.uniform(0, 2 * np.pi)
            u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
            img += np.exp(-0.5 * (((xx - cx) / sx) ** 2 + ((yy - cy) / sy) ** 2)) * np.cos(fr * u + ph) / (1 + b)
        img += rng.randn(H, W) * rng.choice([0.0, 0.02, 0.08])
        out[i] = img.astype(np.float32)
    return out
