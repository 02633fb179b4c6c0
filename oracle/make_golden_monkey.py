"""TEST INFRASTRUCTURE ONLY -- golden outputs of the UNMODIFIED reference's macaque-V1 model classes, constructed offline with
random initial weights (the published checkpoints need a download), generated in the build container (needs /root/reference).

    python -m oracle.make_golden_monkey

  tests/golden/monkey_pointpool.npz   per-session state_dict (seed 3, moved off its trivial init by monkey_oracle.perturb_state(., 7)); responses
                                      of keyless_pointpool_monkey_model (458 neurons) to 3 images, the image gradient of sum(gw * responses), and the
                                      per-session model's responses for the first session
  tests/golden/monkey_gaussian.npz    the same for get_gaussian_monkey_model(seed 5) / perturb 9 / keyless_gaussian_monkey_model
  tests/golden/monkey_ensemble.npz    ensamble_keyless_{pointpool,gaussian}_monkey_model over three checkpoints written from seeds 3,4,5 / perturb 8,9,10:
                                      responses to 2 images and the image gradient (inputs and outputs only; the tests rebuild the checkpoints)
"""
import os
import tempfile
import warnings

import numpy as np
import torch

from oracle.make_golden import OUT
from oracle.monkey_oracle import perturb_state, sd_to_numpy
from oracle.ref_harness import import_reference

ENSEMBLE_SEEDS = ((3, 8), (4, 9), (5, 10))


def images(n, seed):
    return torch.randn(n, 1, 93, 93, generator=torch.Generator().manual_seed(seed)) * 0.8


def upstream(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 458, generator=g) * (torch.rand(n, 458, generator=g) < 0.2)


def resp_and_grad(model, x, gw):
    x = x.clone().requires_grad_(True)
    y = model(x)
    (gx,) = torch.autograd.grad((y * gw).sum(), x)
    return y.detach().numpy(), gx.numpy()


def write_checkpoints(mm_get, folder, tag):
    paths = []
    for i, (seed, pseed) in enumerate(ENSEMBLE_SEEDS):
        m = perturb_state(mm_get(seed=seed), pseed)
        paths.append(os.path.join(folder, f"{tag}_{i}.pt"))
        torch.save(m.state_dict(), paths[-1])
    return paths


def main():
    warnings.simplefilter("ignore")
    import_reference()
    from laminr.neuron_models import monkey_models as mm

    x, gw = images(3, 21), upstream(3, 22)
    for tag, get, keyless, seed, pseed in (("pointpool", mm.get_pointpool_monkey_model, mm.keyless_pointpool_monkey_model, 3, 7),
                                           ("gaussian", mm.get_gaussian_monkey_model, mm.keyless_gaussian_monkey_model, 5, 9)):
        m = perturb_state(get(seed=seed), pseed)
        sd = sd_to_numpy(m.state_dict())
        torch.manual_seed(1)  # the merged readout draws its own initial positions at construction
        km = keyless(m)
        y, gx = resp_and_grad(km, x, gw)
        key0 = list(m.readout.keys())[0]
        with torch.no_grad():
            y_session = m.eval()(x, data_key=key0).numpy()
        out = {f"sd/{k}": v for k, v in sd.items()}
        out.update(x=x.numpy(), gw=gw.numpy(), y=y, gx=gx, y_session=y_session, session=np.array(key0), seed=np.int64(seed), perturb=np.int64(pseed))
        if tag == "gaussian":
            out["merged_mu"] = km.readout._mu.detach().numpy()
        np.savez_compressed(os.path.join(OUT, f"monkey_{tag}.npz"), **out)
        print(tag, "y", y.min(), y.mean(), y.max(), "|gx|", np.abs(gx).max())

    xe, gwe = images(2, 31), upstream(2, 32)
    out = dict(x=xe.numpy(), gw=gwe.numpy(), seeds=np.array(ENSEMBLE_SEEDS))
    with tempfile.TemporaryDirectory() as tmp:
        for tag, get, ens in (("pointpool", mm.get_pointpool_monkey_model, mm.ensamble_keyless_pointpool_monkey_model),
                              ("gaussian", mm.get_gaussian_monkey_model, mm.ensamble_keyless_gaussian_monkey_model)):
            paths = write_checkpoints(get, tmp, tag)
            np.random.seed(0)  # the ensemble builds its members with numpy-drawn seeds before loading the checkpoints
            torch.manual_seed(2)
            e = ens(paths)
            out[f"y_{tag}"], out[f"gx_{tag}"] = resp_and_grad(e, xe, gwe)
            print("ensemble", tag, out[f"y_{tag}"].mean())
    np.savez_compressed(os.path.join(OUT, "monkey_ensemble.npz"), **out)


if __name__ == "__main__":
    main()
