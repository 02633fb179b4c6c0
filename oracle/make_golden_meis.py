"""TEST INFRASTRUCTURE ONLY -- golden output of the UNMODIFIED reference's `get_mei_dict` on demo1 at 100x100 (BASELINE configs[0], the first call
of the README quick start), generated in the build container (needs /root/reference; see oracle/ref_harness.py).

    python -m oracle.make_golden_meis        ->  tests/golden/mei_dict_100.npz
                                                 tests/golden/mei_dict_std_40.npz  (std constraint, pixel range [0, 1], neurons [2, 0], 300 iterations,
                                                                                    zscore 0.5: the non-default arguments of mei.py:74-88)

`create_mask` inside the reference runs with the skimage stubs of oracle/ref_harness.py (`label` via scipy.ndimage, `convex_hull_image` restated):
the hull step of this fixture is as pinned as tests/golden/create_mask.npz is, everything else is the reference's own code.
"""
import contextlib
import io
import os

import numpy as np
import torch

from oracle.make_golden import OUT
from oracle.ref_harness import import_reference


def main():
    laminr = import_reference()
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    np.random.seed(0)
    model = laminr.neuron_models.simulated("demo1", img_res=[100, 100])
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        meis = laminr.get_mei_dict(model, [1, 100, 100], pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, required_img_norm=1.0)
    out = dict(
        meis=np.stack([np.asarray(meis[i]["mei"], np.float32) for i in range(3)]).astype(np.float16),  # 1e-3 relative is all the test asks of the images
        masks=np.stack([np.asarray(meis[i]["mask"], np.float32) for i in range(3)]),
        activations=np.array([meis[i]["activation"] for i in range(3)], np.float64),
        centers=np.array([[meis[i]["center_pos"]["x"], meis[i]["center_pos"]["y"]] for i in range(3)], np.float64),
    )
    np.savez_compressed(os.path.join(OUT, "mei_dict_100.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()}, out["activations"], out["centers"])

    torch.manual_seed(3)
    model = laminr.neuron_models.simulated("demo1", img_res=[40, 40])
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        meis = laminr.get_mei_dict(model, [1, 40, 40], pixel_value_lower_bound=0.0, pixel_value_upper_bound=1.0, neuron_idxs=[2, 0], required_img_std=0.05,
                                   zscore=0.5, step_size=5, num_iterations=300)
    assert sorted(meis) == [0, 1]  # keyed by enumeration index (mei.py:114)
    out = dict(
        meis=np.stack([np.asarray(meis[i]["mei"], np.float32) for i in range(2)]),
        masks=np.stack([np.asarray(meis[i]["mask"], np.float32) for i in range(2)]),
        activations=np.array([meis[i]["activation"] for i in range(2)], np.float64),
        centers=np.array([[meis[i]["center_pos"]["x"], meis[i]["center_pos"]["y"]] for i in range(2)], np.float64),
        rng_after=torch.get_rng_state().numpy(),
    )
    np.savez_compressed(os.path.join(OUT, "mei_dict_std_40.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()}, out["activations"], out["centers"], (out["masks"] > 0.5).sum(axis=(1, 2)))


if __name__ == "__main__":
    main()
