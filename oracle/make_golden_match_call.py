"""TEST INFRASTRUCTURE ONLY -- golden output of a whole `InvarianceManifold.match([1, 2], num_epochs=5)` call of the UNMODIFIED reference at 20x20
(coordinate-transform initialisation with its RNG draws, 60-angle sweep, five jittered epochs of Adam, final un-jittered snapshot), generated in the
build container (needs /root/reference; see oracle/ref_harness.py).

    python -m oracle.make_golden_match_call        ->  tests/golden/match_call_20.npz  (defaults)
                                                       tests/golden/match_call_variants_20.npz  (non-default transform flags, see VARIANTS)
"""
import contextlib
import io
import os

import numpy as np
import torch

from oracle.make_golden import OUT, npy, synthetic_meis, template_arrays
from oracle.ref_harness import import_reference


# non-default flags of `match` (inv_temp_learn_match.py:257-278): which of the 7 affine scalars are parameters, their shapes, the clamp, the sweep
VARIANTS = {
    "uniform_noshear": dict(uniform_scale_coordinate_transformation=True, allow_shear_coordinate_transformation=False),
    "noscale_noclamp": dict(allow_scale_coordinate_transformation=False, coordinate_transform_clamp_boundaries=False),
    "nosweep_3steps": dict(rotate_angle_and_scale=False, steps_per_epoch=3),
    "translation_sweep": dict(find_best_translation=True, num_epochs=2),
}


def run(laminr, res, targets, **kw):
    model = laminr.neuron_models.simulated("demo1", img_res=[res, res])
    g = torch.Generator().manual_seed(res)
    meis = synthetic_meis(model, res, g)
    torch.manual_seed(0)
    im = laminr.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    with torch.no_grad():
        for p in im.template.func.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.1)
    im.template_neuron_idx = 0
    d = template_arrays(im.template)
    d.update(masks=np.stack([meis[k]["mask"] for k in range(3)]), centers=np.array([[meis[k]["center_pos"]["x"], meis[k]["center_pos"]["y"]] for k in range(3)]),
             mei_acts=np.array([meis[k]["activation"] for k in range(3)]))
    torch.manual_seed(5)
    kw.setdefault("num_epochs", 5)
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        imgs, acts = im.match(targets, **kw)
    aff = im.template.coordinate_transform.transforms.Affine
    d.update(imgs=np.asarray(imgs, np.float32), acts=np.asarray(acts, np.float32), angles=npy(aff.angles), scalings=npy(aff.scalings), shears=npy(aff.shears),
             translation=npy(aff.translation)[:, 0], rng_after=torch.get_rng_state().numpy())
    return d


def main():
    laminr = import_reference()
    torch.set_num_threads(os.cpu_count())
    d = run(laminr, 20, [1, 2])
    np.savez_compressed(os.path.join(OUT, "match_call_20.npz"), **d)
    print({k: d[k].shape for k in ("imgs", "acts", "angles", "scalings", "shears", "translation")}, d["angles"], d["acts"].mean(axis=1))
    out = {k: d[k] for k in d if k[0] in "WbB" or k in ("auxB", "coords", "masks", "centers", "mei_acts")}  # the shared inputs, once
    for tag, kw in VARIANTS.items():
        v = run(laminr, 20, [1, 2], **kw)
        out.update({f"{tag}_{k}": v[k] for k in ("imgs", "acts", "angles", "scalings", "shears", "translation", "rng_after")})
        print(tag, v["angles"], v["scalings"].tolist(), v["shears"].tolist(), v["acts"].mean(axis=1))
    np.savez_compressed(os.path.join(OUT, "match_call_variants_20.npz"), **out)


if __name__ == "__main__":
    main()
