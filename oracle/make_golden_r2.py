"""TEST INFRASTRUCTURE ONLY -- golden outputs of the UNMODIFIED reference at the sizes BASELINE.json's configs name (round 2), generated in the
build container (needs /root/reference; see oracle/ref_harness.py).  Nothing here is imported by the product.

    python -m oracle.make_golden_r2 [case ...]

      match_100    -> tests/golden/match_step_100.npz   one match hot step (inv_temp_learn_match.py:352-365) on demo1 targets [1, 2] at 100x100
                      tests/golden/match_call_100.npz   the whole `match([1, 2], num_epochs=5)` call at 100x100 (BASELINE configs[1])
      learn_late   -> tests/golden/learn_late_100.npz   a learn hot step LATE in the trajectory: the reference's own loop (jittered grids, Adam) run for
                                                        600 steps at 100x100, then one step's loss / images / parameter gradients from that state -- near
                                                        convergence the gradient is a small residual of two competing terms (precision probe)
      conv_200     -> tests/golden/conv_learn_match_200.npz   BASELINE configs[3] at full size: random-init Stacked2dCore(1->32, kernels 9/7/7, 3 layers) +
                                                        FullGaussian2d(128 neurons) at 200x200, one learn hot step and one match hot step
      config3      -> tests/golden/config3_sample_100.npz     BASELINE configs[2]/[4]: an 8-neuron sample of the 1024-neuron synthetic population
                                                        (SURVEY 8d config 3): `get_mei_dict` of the sample and `match` of neuron 0 -> the 7 others

Images are stored sub-sampled (every `sub`-th pixel) or as float16 where the tests ask 1e-3 of them; everything else is float32 / float64.
"""
import contextlib
import io
import os
import sys
import time
import warnings

import numpy as np
import torch

from oracle.make_golden import OUT, export_convcore_arrays, npy, synthetic_meis, template_arrays
from oracle.ref_harness import import_reference

N = 20
CONFIG3_SAMPLE = [0, 1, 2, 100, 500, 777, 1000, 1023]  # population indices of the 8-neuron sample (neuron 0 = the template)


def _quiet():
    return contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO())


def _fixed_grid():
    return (torch.linspace(0, 2 * np.pi, N + 1)[:-1] + np.pi / N + 0.0123).reshape(N, 1)


def _match_step(laminr, im, model, meis, targets, grid):
    """One match hot step through the reference's own forward (inv_temp_learn_match.py:352-365) from fixed affine parameters."""
    im.template_neuron_idx, im.target_neuron_idxs = 0, targets
    t = im.template
    for p in t.func.parameters():
        p.grad = None
    D = len(targets)
    t.initialize_coordinate_transform(D, True, False, 0.1, True, True, False, True)
    rf = {k: np.array([v["center_pos"]["x"], v["center_pos"]["y"]], np.float32) for k, v in meis.items()}
    t.register_coords_shifts(torch.from_numpy(rf[0]), torch.from_numpy(np.array([rf[k] for k in targets])))
    aff = t.coordinate_transform.transforms.Affine
    with torch.no_grad():
        aff.angles.copy_(torch.tensor([0.7, 4.1]))
        aff.scalings.copy_(torch.tensor([[1.1, 0.9], [0.8, 1.2]]))
        aff.shears.copy_(torch.tensor([[0.05, -0.02], [-0.1, 0.08]]))
        aff.translation.copy_(torch.tensor([[[0.03, -0.04]], [[-0.05, 0.02]]]))
    t.train()
    img_pre, img_post, _acts, _ = im.forward(grid, t, im.img_transforms, model, return_template=False, other_neurons_loc_in_list=np.arange(D),
                                             other_neurons_loc_in_model=targets)
    rel = _acts[range(D), :, range(D)]
    mei_acts = torch.tensor([meis[k]["activation"] for k in targets])
    acts = rel.mean(dim=1) / mei_acts
    loss = -acts.mean()
    loss.backward()
    return dict(loss=npy(loss), acts=npy(acts), acts_dn=npy(rel), img_pre=npy(img_pre), img_post=npy(img_post), angles=npy(aff.angles),
                scalings=npy(aff.scalings), shears=npy(aff.shears), translation=npy(aff.translation)[:, 0], d_angles=npy(aff.angles.grad),
                d_scalings=npy(aff.scalings.grad), d_shears=npy(aff.shears.grad), d_translation=npy(aff.translation.grad)[:, 0], template_xy=rf[0],
                target_xy=np.array([rf[k] for k in targets]), mei_acts=npy(mei_acts), targets=np.array(targets))


def case_match_100(laminr):
    res, sub = 100, 7
    model = laminr.neuron_models.simulated("demo1", img_res=[res, res])
    g = torch.Generator().manual_seed(res)
    meis = synthetic_meis(model, res, g)
    torch.manual_seed(0)
    im = laminr.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    with torch.no_grad():
        for p in im.template.func.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.1)
    d = template_arrays(im.template)
    grid = _fixed_grid()
    m = _match_step(laminr, im, model, meis, [1, 2], grid)
    m["img_pre"] = m["img_pre"].reshape(2, N, -1)[:, :, ::sub]
    m["img_post"] = m["img_post"].reshape(2, N, -1)[:, :, ::sub]
    d.update(m, grid=npy(grid)[:, 0], sub=np.int64(sub))
    np.savez_compressed(os.path.join(OUT, "match_step_100.npz"), **d)
    print("match_step_100", d["loss"], d["acts"], d["d_angles"])
    # the whole call, from seeds alone
    from oracle.make_golden_match_call import run
    c = run(laminr, res, [1, 2])
    c["imgs"] = c["imgs"].astype(np.float16)  # the test asks 1e-2 of the final images
    np.savez_compressed(os.path.join(OUT, "match_call_100.npz"), **c)
    print("match_call_100", c["angles"], c["acts"].mean(axis=1))


def case_learn_late(laminr, steps=600):
    from laminr.contrastive_loss import SimCLROnGrid
    from laminr.utils.general import SingleNeuronModel
    from laminr.utils.mask import apply_mask
    from laminr.utils.training import JitteringGridDatamodule

    res, sub = 100, 37
    model = laminr.neuron_models.simulated("demo1", img_res=[res, res])
    meis = synthetic_meis(model, res, None)
    torch.manual_seed(0)
    im = laminr.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    dm = JitteringGridDatamodule(1, N, 50)
    reg = SimCLROnGrid(1, N, 0.1, 0.3, True)
    opt = torch.optim.Adam(im.template.func.parameters(), lr=1e-3)
    snm = SingleNeuronModel(model, 0)
    mask = torch.from_numpy(meis[0]["mask"])
    torch.manual_seed(123)

    def objective(gr):
        img_pre, img_post, _acts, _ = im.forward(gr, im.template, im.img_transforms, snm, return_template=True)
        acts = _acts / meis[0]["activation"]
        sim = reg.reg_term(apply_mask(img_post, mask, -1.0, 1.0))
        return img_pre, img_post, acts, sim, -acts.mean() - 2 * sim

    t0, done, means = time.time(), 0, []
    while done < steps:  # the reference's loop (:185-208): 50 jittered grids per epoch, reg_scale 2 (the scheduler starts after epoch 10 and did not
        for gr in dm.train_dataloader():  # fire in the reference's quick-start run before epoch 12 either)
            _, _, acts, _, loss = objective(gr)
            opt.zero_grad()
            loss.backward()
            opt.step()
            done += 1
            means.append(acts.mean().item())
            if done >= steps:
                break
    print(f"learn_late: {steps} steps in {time.time() - t0:.0f} s, mean activation {means[0]:.3f} -> {means[-1]:.3f}")
    d = template_arrays(im.template)  # the state AFTER `steps` Adam steps: the input of the pinned step
    grid = _fixed_grid()
    opt.zero_grad()
    img_pre, img_post, acts, sim, loss = objective(grid)
    loss.backward()
    d.update(grid=npy(grid)[:, 0], mask=meis[0]["mask"], loss=npy(loss), sim=npy(sim), acts=npy(acts).reshape(-1), sub=np.int64(sub),
             img_pre=npy(img_pre).reshape(N, -1)[:, ::sub], img_post=npy(img_post).reshape(N, -1)[:, ::sub], reg_scale=np.float32(2.0),
             mei_act=np.float64(meis[0]["activation"]), steps=np.int64(steps), train_means=np.array(means, np.float32))
    for l in range(5):
        d[f"dW{l}"] = npy(getattr(im.template.func, f"layer{l}").weight.grad)
        d[f"db{l}"] = npy(getattr(im.template.func, f"layer{l}").bias.grad)
    np.savez_compressed(os.path.join(OUT, "learn_late_100.npz"), **d)
    print("learn_late_100", d["loss"], d["acts"].mean(), float(np.sqrt(sum((d[f'dW{l}'] ** 2).sum() for l in range(5)))))


def conv_meis(mu, res, n):
    """Synthetic meis_dict for a Gaussian-readout encoder: Gaussian-bump RF masks at the readout positions (an *input fixture*)."""
    yy, xx = np.meshgrid(np.arange(res) + 0.5, np.arange(res) + 0.5, indexing="ij")
    out = {}
    for i in range(n):
        cx, cy = (mu[i, 0] + 1) / 2 * (res - 1) + 0.5, (mu[i, 1] + 1) / 2 * (res - 1) + 0.5
        mask = np.exp(-0.5 * (((xx - cx) / (0.08 * res)) ** 2 + ((yy - cy) / (0.08 * res)) ** 2)).astype(np.float32)
        out[i] = dict(mei=np.zeros((1, res, res), np.float32), activation=1.0 + 0.05 * i, mask=mask, center_pos=dict(x=float(cx), y=float(cy)))
    return out


def case_conv_200(laminr):
    from laminr.contrastive_loss import SimCLROnGrid
    from laminr.utils.general import SingleNeuronModel
    from laminr.utils.mask import apply_mask
    from neuralpredictors.layers.cores.conv2d import Stacked2dCore
    from neuralpredictors.layers.encoders.firing_rate import FiringRateEncoder
    from neuralpredictors.layers.readouts.gaussian import FullGaussian2d

    res, hc, n, sub = 200, 32, 128, 53
    torch.manual_seed(0)  # constructors in this order under seed 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        core = Stacked2dCore(input_channels=1, hidden_channels=hc, input_kern=9, hidden_kern=7, layers=3, stack=-1, pad_input=True)
        ro = FullGaussian2d(in_shape=(hc, res, res), outdims=n, bias=True)
    model = FiringRateEncoder(core, ro).eval()
    meis = conv_meis(npy(ro._mu).reshape(-1, 2).clip(-1, 1), res, 3)
    norm = 0.05 * res
    torch.manual_seed(0)
    im = laminr.InvarianceManifold(model, meis, required_img_norm=norm, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for p in im.template.func.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.1)
    grid = _fixed_grid()
    reg = SimCLROnGrid(1, N, 0.1, 0.3, True)
    t0 = time.time()
    img_pre, img_post, _acts, _ = im.forward(grid, im.template, im.img_transforms, SingleNeuronModel(model, 0), return_template=True)
    acts = _acts / meis[0]["activation"]
    sim = reg.reg_term(apply_mask(img_post, torch.from_numpy(meis[0]["mask"]), -1.0, 1.0))
    loss = -acts.mean() - 2 * sim
    loss.backward()
    d = template_arrays(im.template)
    d.update({"m_" + k: v for k, v in export_convcore_arrays(model).items()})
    d.update(grid=npy(grid)[:, 0], mask=meis[0]["mask"], loss=npy(loss), sim=npy(sim), acts=npy(acts).reshape(-1), sub=np.int64(sub),
             img_pre=npy(img_pre).reshape(N, -1)[:, ::sub], img_post=npy(img_post).reshape(N, -1)[:, ::sub], reg_scale=np.float32(2.0),
             mei_act=np.float64(meis[0]["activation"]), norm=np.float64(norm), res=np.int64(res),
             centers=np.array([[meis[i]["center_pos"]["x"], meis[i]["center_pos"]["y"]] for i in range(3)]),
             mei_acts_all=np.array([meis[i]["activation"] for i in range(3)]))
    for l in range(5):
        d[f"dW{l}"] = npy(getattr(im.template.func, f"layer{l}").weight.grad)
        d[f"db{l}"] = npy(getattr(im.template.func, f"layer{l}").bias.grad)
    print(f"conv_200 learn step {time.time() - t0:.1f} s: loss {d['loss']}, acts {d['acts'][:3]}")
    t0 = time.time()
    m = _match_step(laminr, im, model, meis, [1, 2], grid)
    for k in ("img_pre", "img_post"):
        m[k] = m[k].reshape(2, N, -1)[:, :, ::sub]
    m.pop("img_pre")
    d.update({("match_" + k if k in ("loss", "acts", "acts_dn", "img_post") else k): v for k, v in m.items()})
    np.savez_compressed(os.path.join(OUT, "conv_learn_match_200.npz"), **d)
    print(f"conv_200 match step {time.time() - t0:.1f} s: loss {d['match_loss']}, d_angles {d['d_angles']}")


def case_config3(laminr):
    from laminr.neuron_models.simulated_models import phase_invariant_neuron_generator
    from laminr.neuron_models.utils.simulated_model import ArbitraryMultiNeuronModel, random_transformation_matrices

    res, n_pop, sub = 100, 1024, 5
    np.random.seed(0)  # matrices first, then the locations, from the NumPy generator
    mats = random_transformation_matrices(n_pop)
    locs = np.random.uniform(-0.35, 0.35, size=(n_pop, 2))
    neurons = [phase_invariant_neuron_generator(locs[i], [res, res], transformation_matrix=None if i == 0 else mats[i], num_filters=20)[0] for i in CONFIG3_SAMPLE]
    model = ArbitraryMultiNeuronModel(neurons)
    n = len(CONFIG3_SAMPLE)
    import hashlib
    bank_sha = np.array([hashlib.sha256(np.ascontiguousarray(m.filters.numpy())).hexdigest() for m in neurons])
    torch.manual_seed(0)
    t0 = time.time()
    with _quiet()[0], _quiet()[1]:
        meis = laminr.get_mei_dict(model, [1, res, res], pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, required_img_norm=1.0)
    rng_after_meis = torch.get_rng_state().numpy()
    print(f"config3: get_mei_dict of {n} neurons {time.time() - t0:.0f} s, activations {[round(meis[i]['activation'], 6) for i in range(n)]}")
    torch.manual_seed(0)
    im = laminr.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    g = torch.Generator().manual_seed(res)
    with torch.no_grad():
        for p in im.template.func.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.1)
    im.template_neuron_idx = 0
    d = template_arrays(im.template)
    torch.manual_seed(5)
    t0 = time.time()
    with _quiet()[0], _quiet()[1]:
        imgs, acts = im.match(list(range(1, n)), num_epochs=5)
    print(f"config3: match of {n - 1} targets, 5 epochs: {time.time() - t0:.0f} s, acts {np.asarray(acts).mean(axis=1)}")
    aff = im.template.coordinate_transform.transforms.Affine
    d.update(sample=np.array(CONFIG3_SAMPLE), n_pop=np.int64(n_pop), bank_sha=bank_sha, sub=np.int64(sub),
             meis=np.stack([np.asarray(meis[i]["mei"], np.float32) for i in range(n)]).astype(np.float16),
             masks=np.stack([np.asarray(meis[i]["mask"], np.float32) for i in range(n)]),  # float32: their sums set the initial scalings of match
             mask_sums=np.array([np.asarray(meis[i]["mask"], np.float64).sum() for i in range(n)]),
             mask_counts=np.array([(np.asarray(meis[i]["mask"]) > 0.5).sum() for i in range(n)]),
             activations=np.array([meis[i]["activation"] for i in range(n)], np.float64),
             centers=np.array([[meis[i]["center_pos"]["x"], meis[i]["center_pos"]["y"]] for i in range(n)], np.float64), rng_after_meis=rng_after_meis,
             imgs=np.asarray(imgs, np.float32).reshape(n - 1, N, -1)[:, :, ::sub].astype(np.float16), acts=np.asarray(acts, np.float32), angles=npy(aff.angles),
             scalings=npy(aff.scalings), shears=npy(aff.shears), translation=npy(aff.translation)[:, 0], rng_after=torch.get_rng_state().numpy())
    np.savez_compressed(os.path.join(OUT, "config3_sample_100.npz"), **d)


CASES = dict(match_100=case_match_100, learn_late=case_learn_late, conv_200=case_conv_200, config3=case_config3)


def main():
    laminr = import_reference()
    torch.set_num_threads(os.cpu_count())
    for name in (sys.argv[1:] or list(CASES)):
        t0 = time.time()
        CASES[name](laminr)
        print(f"[{name}] {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
