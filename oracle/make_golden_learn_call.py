"""TEST INFRASTRUCTURE ONLY -- golden output of a whole `InvarianceManifold.learn(0, ...)` call of the UNMODIFIED reference at 20x20, from the
constructor's own seeded initial network (nothing loaded): DataLoader seed draws, per-epoch jitter streams, Adam steps, the per-epoch evaluation on
the un-jittered grid, the final getter.  Two calls are recorded:

  short : steps_per_epoch=5, num_max_epochs=4            (20 steps; the control logic only counts)
  long  : steps_per_epoch=10, num_max_epochs=14          (140 steps; epochs > min_epochs=10 run the reg_scale decay rule, inv_temp_learn_match.py:210-237)

generated in the build container (needs /root/reference; see oracle/ref_harness.py).

    python -m oracle.make_golden_learn_call                ->  tests/golden/learn_call_20.npz
    python -m oracle.make_golden_learn_call --selfcheck    ->  prints how far the unmodified reference lands from that golden when only its BLAS thread
                                                               count changes (1 and 3 threads instead of all cores): the tolerance basis of the long call
                                                               in tests/test_gpu_api.py::test_learn_call_vs_reference_call
"""
import contextlib
import io
import os

import numpy as np
import torch

from oracle.make_golden import OUT, synthetic_meis, template_arrays
from oracle.ref_harness import import_reference


def run(laminr, res, tag, **kw):
    model = laminr.neuron_models.simulated("demo1", img_res=[res, res])
    meis = synthetic_meis(model, res, torch.Generator().manual_seed(res))
    torch.manual_seed(0)
    im = laminr.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    torch.manual_seed(11)
    with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
        imgs, acts = im.learn(0, **kw)
    d = {f"{tag}_{k}": v for k, v in template_arrays(im.template).items() if k[0] in "Wb"}
    d.update({f"{tag}_imgs": np.asarray(imgs, np.float32), f"{tag}_acts": np.asarray(acts, np.float32), f"{tag}_rng_after": torch.get_rng_state().numpy()})
    return d, meis


def main():
    laminr = import_reference()
    torch.set_num_threads(os.cpu_count())
    res = 20
    d, meis = run(laminr, res, "short", steps_per_epoch=5, num_max_epochs=4)
    d2, _ = run(laminr, res, "long", steps_per_epoch=10, num_max_epochs=14)
    d.update(d2)
    d.update(masks=np.stack([meis[k]["mask"] for k in range(3)]), centers=np.array([[meis[k]["center_pos"]["x"], meis[k]["center_pos"]["y"]] for k in range(3)]),
             mei_acts=np.array([meis[k]["activation"] for k in range(3)]))
    np.savez_compressed(os.path.join(OUT, "learn_call_20.npz"), **d)
    print({k: d[k].shape for k in d if "imgs" in k or "acts" in k}, d["short_acts"].mean(), d["long_acts"].mean())


def selfcheck():
    laminr = import_reference()
    g = np.load(os.path.join(OUT, "learn_call_20.npz"))
    rel = lambda a, b: float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))
    for threads in (1, 3):
        torch.set_num_threads(threads)
        for tag, kw in (("short", dict(steps_per_epoch=5, num_max_epochs=4)), ("long", dict(steps_per_epoch=10, num_max_epochs=14))):
            d, _ = run(laminr, 20, tag, **kw)
            print(f"threads {threads} {tag}: weights {max(rel(d[f'{tag}_W{l}'], g[f'{tag}_W{l}']) for l in range(5)):.2e}, images "
                  f"{rel(d[f'{tag}_imgs'], g[f'{tag}_imgs']):.2e}, activations {rel(d[f'{tag}_acts'], g[f'{tag}_acts']):.2e}, mean activation "
                  f"{d[f'{tag}_acts'].mean():.4f} (golden {g[f'{tag}_acts'].mean():.4f})")


if __name__ == "__main__":
    import sys
    selfcheck() if "--selfcheck" in sys.argv else main()
