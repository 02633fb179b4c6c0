"""TEST INFRASTRUCTURE ONLY -- golden output of whole `learn` and `match` calls of the UNMODIFIED reference with the STD image constraint
(`InvarianceManifold(required_img_std=0.1)`: StandardizeClip, inv_temp_learn_match.py:52-73, img_transforms.py:38-51) and intermediate snapshots
(`save_results_every_n_epochs=2`: results stacked over epochs 0, 2 and the final state, :180-184, :344-348, :384-388), demo1 at 20x20, from seeds alone.

    python -m oracle.make_golden_std_calls        ->  tests/golden/std_calls_20.npz
"""
import contextlib
import io
import os

import numpy as np
import torch

from oracle.make_golden import OUT, npy, synthetic_meis, template_arrays
from oracle.ref_harness import import_reference


def main():
    laminr = import_reference()
    torch.set_num_threads(os.cpu_count())
    res = 20
    model = laminr.neuron_models.simulated("demo1", img_res=[res, res])
    meis = synthetic_meis(model, res, torch.Generator().manual_seed(res))
    torch.manual_seed(0)
    im = laminr.InvarianceManifold(model, meis, required_img_std=0.1, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    torch.manual_seed(21)
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        l_imgs, l_acts = im.learn(0, steps_per_epoch=5, num_max_epochs=4, save_results_every_n_epochs=2)
    d = {"learned_" + k: v for k, v in template_arrays(im.template).items() if k[0] in "Wb"}
    d.update(learn_imgs=np.asarray(l_imgs, np.float32), learn_acts=np.asarray(l_acts, np.float32), rng_after_learn=torch.get_rng_state().numpy())
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        m_imgs, m_acts = im.match([1, 2], num_epochs=4, save_results_every_n_epochs=2)
    aff = im.template.coordinate_transform.transforms.Affine
    d.update(match_imgs=np.asarray(m_imgs, np.float32), match_acts=np.asarray(m_acts, np.float32), rng_after_match=torch.get_rng_state().numpy(),
             angles=npy(aff.angles), scalings=npy(aff.scalings), shears=npy(aff.shears), translation=npy(aff.translation)[:, 0],
             masks=np.stack([meis[k]["mask"] for k in range(3)]), centers=np.array([[meis[k]["center_pos"]["x"], meis[k]["center_pos"]["y"]] for k in range(3)]),
             mei_acts=np.array([meis[k]["activation"] for k in range(3)]))
    np.savez_compressed(os.path.join(OUT, "std_calls_20.npz"), **d)
    print({k: d[k].shape for k in ("learn_imgs", "learn_acts", "match_imgs", "match_acts")}, d["learn_acts"].mean(axis=1), d["match_acts"].mean(axis=2), d["angles"])


if __name__ == "__main__":
    main()
