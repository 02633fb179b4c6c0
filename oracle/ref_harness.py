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

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Where the unmodified reference can live: the mounted sources (build container only) or the offline pip install of those sources under
# baseline/_ref (`pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`, done by
# __graft_entry__.build(); git-ignored, but it travels to the GPU box -- that is what `bench.py --impl reference` times there).
INSTALLED_REF = os.path.join(_REPO, "baseline", "_ref")
_CANDIDATES = [p for p in (os.environ.get("LAMINR_REFERENCE_SRC"), "/root/reference/src", INSTALLED_REF) if p]


def _find(prefer_installed=False):
    cands = ([INSTALLED_REF] + _CANDIDATES) if prefer_installed else _CANDIDATES
    for p in cands:
        if os.path.isdir(os.path.join(p, "laminr")):
            return p
    return None


REFERENCE_SRC = _find() or "/root/reference/src"


def reference_available(prefer_installed=False):
    return _find(prefer_installed) is not None


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


def import_reference(prefer_installed=False):
    """Returns the reference's `laminr` package (plus featurevis / neuralpredictors on sys.path).  prefer_installed: take the pip-installed copy
    under baseline/_ref first (bench.py's reference arm), else the mounted sources first (golden generation)."""
    src = _find(prefer_installed)
    if src is None:
        raise RuntimeError(f"reference not found (looked in {_CANDIDATES})")
    install_stubs()
    if src not in sys.path:
        sys.path.insert(0, src)
    import laminr  # noqa: F401  (the reference package, not laminr_b200)

    return laminr


def reference_location():
    """Path the imported reference package came from (for the bench line's `sample`)."""
    import laminr

    return os.path.dirname(os.path.dirname(os.path.abspath(laminr.__file__)))
