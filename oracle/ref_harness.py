"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference (sinzlab/laminr) for golden-vector generation.

Only usable in the build container, where the reference is mounted at /root/reference (read-only).  It does not
travel to the GPU box; nothing under tests/ -m gpu, bench.py or __graft_entry__.smoke() may call this module.

The reference declares three dependencies that are absent from this image (matplotlib, lipstick, scikit-image).
None of them touches the arithmetic of the hot path: matplotlib/lipstick are used by the GIF/plot helpers
(inv_temp_learn_match.py:6-7,424-450; simulated_model.py:4) and skimage.morphology by create_mask
(utils/mask.py:52,55).  They are replaced by import stubs; `label` is restated with scipy.ndimage and
`convex_hull_image` with scipy.spatial (pixel-corner offsets, half-plane test at pixel centres).  The hull
restatement cannot be checked against real skimage here, which is why learn/match parity takes `meis_dict`
(mask, centre) as a shared *input fixture* instead of regenerating masks on both sides.
"""
import os
import sys
import types

import numpy as np

REFERENCE_SRC = os.environ.get("LAMINR_REFERENCE_SRC", "/root/reference/src")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "laminr"))


def _convex_hull_image(img):
    from scipy.spatial import ConvexHull

    pts = np.argwhere(img).astype(float)
    offs = np.array([[-0.5, 0.0], [0.5, 0.0], [0.0, -0.5], [0.0, 0.5]])
    cloud = (pts[:, None, :] + offs[None]).reshape(-1, 2)
    eq = ConvexHull(cloud).equations
    centres = np.stack(np.mgrid[: img.shape[0], : img.shape[1]], -1).reshape(-1, 2)
    return np.all(centres @ eq[:, :2].T + eq[:, 2] <= 1e-10, axis=1).reshape(img.shape)


def install_stubs():
    from scipy import ndimage

    for name in ("matplotlib", "matplotlib.pyplot", "lipstick", "skimage", "skimage.morphology"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["lipstick"].GifMaker = object
    mor = sys.modules["skimage.morphology"]
    sys.modules["skimage"].morphology = mor
    mor.label = lambda img, connectivity=2: ndimage.label(img, structure=np.ones((3, 3)))[0]
    mor.convex_hull_image = _convex_hull_image


def import_reference():
    """Returns the reference's `laminr` package (plus featurevis / neuralpredictors on sys.path)."""
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    install_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import laminr  # noqa: F401  (the reference package, not laminr_b200)

    return laminr
