"""TEST INFRASTRUCTURE ONLY -- the reference's `state_dict()`s a checkpoint of a LAMINR run consists of, as ordered key lists with the sha256 of every
entry: `InvarianceManifold(...).template` of demo1 at 20x20 after `torch.manual_seed(0)`, with the coordinate transform of two targets initialised
(inr.py:226-262) and the RF shifts registered (:256-282), and the `neuron_models.simulated("demo1")` response model.  A checkpoint written by the
reference must load into the host mirror and vice versa: same names, order, shapes, values.

    python -m oracle.make_golden_state_dicts        ->  tests/golden/state_dicts_20.npz
"""
import hashlib
import os

import numpy as np
import torch

from oracle.make_golden import OUT, synthetic_meis
from oracle.ref_harness import import_reference


def digests(pkg, **im_kw):
    model = pkg.neuron_models.simulated("demo1", img_res=[20, 20])
    meis = synthetic_meis(model, 20, None)
    torch.manual_seed(0)
    im = pkg.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, **im_kw)
    t = im.template
    t.initialize_coordinate_transform(2, True, False, 0.1, True, True, False, True)
    rf = {k: np.array([v["center_pos"]["x"], v["center_pos"]["y"]], np.float32) for k, v in meis.items()}
    t.register_coords_shifts(torch.from_numpy(rf[0]), torch.from_numpy(np.array([rf[1], rf[2]])))
    out = {}
    for tag, sd in (("template", t.state_dict()), ("model", model.state_dict())):
        out[f"{tag}_keys"] = np.array(list(sd))
        out[f"{tag}_shapes"] = np.array([str(tuple(v.shape)) + str(v.dtype) for v in sd.values()])
        out[f"{tag}_sha"] = np.stack([np.frombuffer(hashlib.sha256(v.detach().cpu().contiguous().numpy().tobytes()).digest(), np.uint8) for v in sd.values()])
    return out


def main():
    out = digests(import_reference())
    np.savez_compressed(os.path.join(OUT, "state_dicts_20.npz"), **out)
    print({k: v.shape for k, v in out.items()}, list(out["template_keys"]))


if __name__ == "__main__":
    main()
