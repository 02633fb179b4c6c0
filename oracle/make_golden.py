"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference (sinzlab/laminr).

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
The fixtures are committed; the GPU box never runs this script.  Every array is an output of the reference's own
modules (torch CPU, float32) for seeded inputs that are stored alongside, so the numpy oracle and the CUDA path
can both be replayed on identical inputs.
"""
import hashlib
import os

import numpy as np
import torch

from oracle.ref_harness import import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def npy(t):
    return t.detach().cpu().numpy().copy()


def template_arrays(tmpl):
    sd = tmpl.state_dict()
    L = len([k for k in sd if k.startswith("func.layer") and k.endswith("weight")])
    out = {f"W{l}": npy(sd[f"func.layer{l}.weight"]) for l in range(L)}
    out.update({f"b{l}": npy(sd[f"func.layer{l}.bias"]) for l in range(L)})
    out["B"], out["auxB"], out["coords"] = npy(sd["B"]), npy(sd["auxB"]), npy(sd["coords"]).reshape(-1, 2)
    return out


def make_template(laminr, img_res, channels=1, seed=0):
    from laminr.inr import INRTemplates
    from torch import nn

    torch.manual_seed(seed)
    return INRTemplates(
        img_res=img_res, num_templates=1, out_channels=channels, widths=[50] * 4,
        positional_encoding_projection_scale=10, positional_encoding_dim=50, aux_positional_encoding_dim=50,
        aux_positional_encoding_projection_scale=0.1, periodic_invariance=True, nonlinearity=nn.Tanh,
        final_nonlinearity=nn.Tanh, weights_scale=0.1, bias=True,
    )


def case_inr(laminr):
    """INRTemplates.forward(return_template=True) + autograd grads for a seeded upstream gradient."""
    for name, res, C, N in (("inr_12x10", (12, 10), 1, 5), ("inr_9x16_c2", (9, 16), 2, 7)):
        t = make_template(laminr, res, C, seed=1)
        g = torch.Generator().manual_seed(7)
        with torch.no_grad():  # make the net less trivial than the 0.1-scale init
            for p in t.func.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.15)
        grid = torch.rand(N, 1, generator=g) * 2 * np.pi
        out = t(aux=grid, return_template=True)
        up = torch.randn(out.shape, generator=g)
        (out * up).sum().backward()
        d = template_arrays(t)
        d.update(grid=npy(grid)[:, 0], out=npy(out), upstream=npy(up))
        for l in range(5):
            d[f"dW{l}"] = npy(getattr(t.func, f"layer{l}").weight.grad)
            d[f"db{l}"] = npy(getattr(t.func, f"layer{l}").bias.grad)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)


def case_match_inr(laminr):
    """INRTemplates.forward(return_template=False) through CoordinateTransform(only_affine) + affine grads."""
    res, D, N = (14, 14), 3, 4
    t = make_template(laminr, res, 1, seed=2)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for p in t.func.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.15)
    t.initialize_coordinate_transform(D, True, False, 0.1, True, True, False, True)
    aff = t.coordinate_transform.transforms.Affine
    with torch.no_grad():
        aff.angles.copy_(torch.rand(D, generator=g) * 2 * np.pi)
        aff.scalings.copy_(0.5 + torch.rand(D, 2, generator=g))
        aff.shears.copy_((torch.rand(D, 2, generator=g) - 0.5) * 0.3)
        aff.translation.copy_((torch.rand(D, 1, 2, generator=g) - 0.5) * 0.2)
    template_xy = torch.tensor([8.3, 5.1])
    target_xy = torch.tensor([[4.0, 9.5], [7.2, 6.6], [10.1, 3.3]])
    t.register_coords_shifts(template_xy, target_xy)
    grid = torch.rand(N, 1, generator=g) * 2 * np.pi
    t.train()
    out = t(aux=grid, return_template=False)
    up = torch.randn(out.shape, generator=g)
    (out * up).sum().backward()
    coords_t = t.apply_coordinate_transform(t.coords, False, False)
    d = template_arrays(t)
    d.update(
        grid=npy(grid)[:, 0], out=npy(out), upstream=npy(up), coords_transformed=npy(coords_t),
        angles=npy(aff.angles), scalings=npy(aff.scalings), shears=npy(aff.shears), translation=npy(aff.translation)[:, 0],
        d_angles=npy(aff.angles.grad), d_scalings=npy(aff.scalings.grad), d_shears=npy(aff.shears.grad),
        d_translation=npy(aff.translation.grad)[:, 0], template_xy=npy(template_xy), target_xy=npy(target_xy),
        shift_to=npy(t.coords_shift_to_center), tc=npy(t.template_center_in_coords_space),
        shift_from=npy(t.coords_shift_from_center),
    )
    np.savez_compressed(os.path.join(OUT, "match_inr_14x14.npz"), **d)


def case_constrain(laminr):
    from laminr.utils.img_transforms import FixEnergyNormClip, StandardizeClip

    g = torch.Generator().manual_seed(3)
    d = {}
    for tag, mod, kw in (
        ("norm", FixEnergyNormClip, dict(norm=3.0, mean=0.25, pixel_min=-0.5, pixel_max=1.0)),
        ("std", StandardizeClip, dict(std=0.4, mean=0.25, pixel_min=-0.5, pixel_max=1.0)),
    ):
        x = (torch.randn(6, 2, 7, 9, generator=g) * 0.3).requires_grad_()
        y = mod(**kw)(x)
        up = torch.randn(y.shape, generator=g)
        (y * up).sum().backward()
        d.update({f"{tag}_x": npy(x), f"{tag}_y": npy(y), f"{tag}_up": npy(up), f"{tag}_gx": npy(x.grad)})
    np.savez_compressed(os.path.join(OUT, "constrain.npz"), **d)


def case_simresp(laminr):
    d = {}
    for res in (16, 24):
        model = laminr.neuron_models.simulated("demo1", img_res=[res, res])
        g = torch.Generator().manual_seed(res)
        x = (torch.randn(5, 1, res, res, generator=g) * 0.2).requires_grad_()
        y = model(x)
        up = torch.randn(y.shape, generator=g)
        (y * up).sum().backward()
        for i, m in enumerate(model.neuron_models_list):
            d[f"bank{i}_{res}"] = npy(m.filters)
        d.update({f"x_{res}": npy(x), f"y_{res}": npy(y), f"up_{res}": npy(up), f"gx_{res}": npy(x.grad)})
    # two-channel input: the same filter is summed against every channel (simulated_model.py:14)
    model = laminr.neuron_models.simulated("demo1", img_res=[16, 16])
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(3, 2, 16, 16, generator=g) * 0.2).requires_grad_()
    y = model(x)
    y.sum().backward()
    d.update(x_c2=npy(x), y_c2=npy(y), gx_c2=npy(x.grad))
    np.savez_compressed(os.path.join(OUT, "simresp.npz"), **d)


def case_banks_100(laminr):
    """sha256 of the demo1 float32 banks at 100x100 / 200x200 (+ a few sampled values)."""
    d = {}
    for res in (100, 200):
        model = laminr.neuron_models.simulated("demo1", img_res=[res, res])
        for i, m in enumerate(model.neuron_models_list):
            f = npy(m.filters)
            d[f"sha_{res}_{i}"] = np.frombuffer(hashlib.sha256(f.tobytes()).digest(), np.uint8)
            d[f"shape_{res}_{i}"] = np.array(f.shape)
            d[f"sample_{res}_{i}"] = f.reshape(-1)[:: 997].copy()
    # population recipe of config 3: first 4 neurons only
    from laminr.neuron_models.utils.simulated_model import random_transformation_matrices
    from laminr.neuron_models.simulated_models import phase_invariant_neuron_generator

    np.random.seed(0)
    mats = random_transformation_matrices(4)
    locs = np.random.uniform(-0.35, 0.35, size=(4, 2))
    pop = []
    for i in range(4):
        m, _ = phase_invariant_neuron_generator(locs[i], [100, 100], transformation_matrix=None if i == 0 else mats[i], num_filters=20)
        pop.append(npy(m.filters))
    pop = np.stack(pop)
    d["pop4_sha"] = np.frombuffer(hashlib.sha256(pop.tobytes()).digest(), np.uint8)
    d["pop4_sample"] = pop.reshape(-1)[:: 997].copy()
    np.savez_compressed(os.path.join(OUT, "banks.npz"), **d)


def case_simclr(laminr):
    from laminr.contrastive_loss import SimCLROnGrid
    from laminr.utils.mask import apply_mask

    d = {}
    g = torch.Generator().manual_seed(9)
    for tag, n, periodic, pm in (("p20", 20, True, (-1.0, 1.0)), ("np7", 7, False, (0.0, 2.0)), ("p10", 10, True, (-2.0, 1.0))):
        reg = SimCLROnGrid(1, n, 0.1 if n != 7 else 0.3, 0.3, periodic)
        x = (torch.randn(n, 1, 8, 8, generator=g) * 0.5 + (pm[0] + pm[1]) / 2).requires_grad_()
        mask = torch.rand(8, 8, generator=g)
        val = reg.reg_term(apply_mask(x, mask, pm[0], pm[1]))
        val.backward()
        d.update({
            f"{tag}_x": npy(x), f"{tag}_mask": npy(mask), f"{tag}_reg": npy(val), f"{tag}_gx": npy(x.grad),
            f"{tag}_pos": npy(reg.flat_neighbor_mask_pos), f"{tag}_neg": npy(reg.flat_neighbor_mask_neg),
            f"{tag}_pm": np.array(pm, np.float32),
        })
    np.savez_compressed(os.path.join(OUT, "simclr.npz"), **d)


def case_training_utils(laminr):
    from laminr.utils.training import JitteringGridDatamodule, ParamReduceOnPlateau, ImprovementChecker, check_activity_requirements

    d = {}
    for n in (20, 7, 32):
        dm = JitteringGridDatamodule(1, n, 3)
        d[f"grid_{n}"] = npy(dm.grid)[:, 0]
    torch.manual_seed(4)
    dm = JitteringGridDatamodule(1, 20, 5)
    state = torch.get_rng_state()
    batches = torch.stack([b for b in dm.train_dataloader()])
    torch.set_rng_state(state)
    d["jitter_u01"] = npy(torch.rand([5, 1]))[:, 0]
    d["jitter_batches"] = npy(batches)[:, :, 0]
    rng = np.random.RandomState(0)
    seq = np.cumsum(rng.randn(60) * 0.01).astype(np.float32) + 0.5
    sch = ParamReduceOnPlateau(factor=0.8, patience=5, threshold=0.005, mode="max")
    val, vals = 2.0, []
    for s in seq:
        val = sch.step(torch.tensor(s), val)
        vals.append(val)
    d["plateau_seq"], d["plateau_vals"] = seq, np.array(vals)
    chk = ImprovementChecker(patience=4, ignore_diff_smaller_than=1e-3)
    d["improve_flags"] = np.array([chk.is_increasing(float(s)) for s in seq])
    acts = torch.tensor(rng.rand(8, 20).astype(np.float32) * np.linspace(0.005, 0.03, 8, dtype=np.float32)[:, None] + np.linspace(0.979, 0.9895, 8, dtype=np.float32)[:, None])
    d["req_acts"] = npy(acts)
    d["req_pass"] = np.array([check_activity_requirements(a, dict(avg=0.99, std=1.0, necessary_min=0.98)) for a in acts])
    np.savez_compressed(os.path.join(OUT, "training_utils.npz"), **d)


def case_adam(laminr):
    g = torch.Generator().manual_seed(21)
    p = torch.nn.Parameter(torch.randn(257, generator=g))
    opt = torch.optim.Adam([p], lr=1e-3)
    d = dict(p0=npy(p))
    grads, ps = [], []
    for _ in range(5):
        gr = torch.randn(257, generator=g) * 0.01
        p.grad = gr.clone()
        opt.step()
        grads.append(npy(gr))
        ps.append(npy(p))
    d.update(grads=np.stack(grads), ps=np.stack(ps))
    np.savez_compressed(os.path.join(OUT, "adam.npz"), **d)


def case_mei(laminr):
    """featurevis.gradient_ascent through generate_mei's recipe (mei.py:33-56) on demo1 @16x16, 60 iterations."""
    import featurevis
    import featurevis.ops as ops
    from featurevis.utils import Compose
    from laminr.utils.general import SingleNeuronModel

    d = {}
    model = laminr.neuron_models.simulated("demo1", img_res=[16, 16])
    for tag, post0 in (("norm", ops.ChangeNorm(norm=1.0, mean=0.0)), ("std", ops.ChangeStats(std=0.1, mean=0.0))):
        for idx in range(3):
            torch.manual_seed(100 + idx)
            x0 = torch.randn(1, 1, 16, 16)
            post = Compose([post0, ops.ClipRange(-1.0, 1.0)])
            xi = post(x0)
            x, fevals, _ = featurevis.gradient_ascent(SingleNeuronModel(model, idx), xi, step_size=10, num_iterations=60, post_update=post, print_iters=1001)
            d.update({f"{tag}_x0_{idx}": npy(x0), f"{tag}_x_{idx}": npy(x), f"{tag}_fevals_{idx}": np.array(fevals)})
    np.savez_compressed(os.path.join(OUT, "mei.npz"), **d)


def synthetic_meis(model, res, g):
    """A synthetic meis_dict (Gaussian-bump RF masks at the filter centres) -- an *input fixture*, not a result."""
    yy, xx = np.meshgrid(np.arange(res) + 0.5, np.arange(res) + 0.5, indexing="ij")
    centres = [(0.6 * res, 0.6 * res), (0.4 * res, 0.4 * res), (0.4 * res, 0.6 * res)]  # (x, y)
    out = {}
    for i, (cx, cy) in enumerate(centres):
        mask = np.exp(-0.5 * (((xx - cx) / (0.16 * res)) ** 2 + ((yy - cy) / (0.16 * res)) ** 2)).astype(np.float32)
        out[i] = dict(mei=np.zeros((1, res, res), np.float32), activation=1.0000009536743164, mask=mask, center_pos=dict(x=cx, y=cy))
    return out


def case_learn_match_steps(laminr):
    """One learn hot step (inv_temp_learn_match.py:187-207) and one match hot step (:352-365), via the reference's
    own InvarianceManifold.forward, on demo1 at 20x20 and (learn only, sub-sampled) 100x100."""
    from laminr.contrastive_loss import SimCLROnGrid
    from laminr.utils.mask import apply_mask
    from laminr.utils.general import SingleNeuronModel

    for res in (20, 100):
        model = laminr.neuron_models.simulated("demo1", img_res=[res, res])
        g = torch.Generator().manual_seed(res)
        meis = synthetic_meis(model, res, g)
        torch.manual_seed(0)
        im = laminr.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
        with torch.no_grad():  # move away from the near-constant init so every branch of the backward is exercised
            for p in im.template.func.parameters():
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
        N = 20
        grid = (torch.linspace(0, 2 * np.pi, N + 1)[:-1] + np.pi / N + 0.0123).reshape(N, 1)
        reg = SimCLROnGrid(1, N, 0.1, 0.3, True)
        img_pre, img_post, _acts, _ = im.forward(grid, im.template, im.img_transforms, SingleNeuronModel(model, 0), return_template=True)
        acts = _acts / meis[0]["activation"]
        sim = reg.reg_term(apply_mask(img_post, torch.from_numpy(meis[0]["mask"]), -1.0, 1.0))
        loss = -acts.mean() - 2 * sim
        loss.backward()
        d = template_arrays(im.template)
        sub = slice(None) if res == 20 else slice(None, None, 37)
        d.update(
            grid=npy(grid)[:, 0], mask=meis[0]["mask"], loss=npy(loss), sim=npy(sim), acts=npy(acts).reshape(-1),
            img_pre=npy(img_pre).reshape(N, -1)[:, sub], img_post=npy(img_post).reshape(N, -1)[:, sub], reg_scale=np.float32(2.0),
            mei_act=np.float64(meis[0]["activation"]),
        )
        for l in range(5):
            d[f"dW{l}"] = npy(getattr(im.template.func, f"layer{l}").weight.grad)
            d[f"db{l}"] = npy(getattr(im.template.func, f"layer{l}").bias.grad)
        np.savez_compressed(os.path.join(OUT, f"learn_step_{res}.npz"), **d)

        if res != 20:
            continue
        # ---- the same learn step with the Gaussian-blur gradient preconditioner (learn(gaussian_blur_sigma=1.0): a backward hook on img_post,
        #      inv_temp_learn_match.py:158,524-525); `scipy.signal.gaussian` is aliased to scipy.signal.windows.gaussian (gone from scipy >= 1.13)
        import scipy.signal
        import scipy.signal.windows

        if not hasattr(scipy.signal, "gaussian"):
            scipy.signal.gaussian = scipy.signal.windows.gaussian
        import featurevis.ops as fops

        for p_ in im.template.func.parameters():
            p_.grad = None
        _, img_post_b, _acts_b, _ = im.forward(grid, im.template, im.img_transforms, SingleNeuronModel(model, 0), gb=fops.GaussianBlur(1.0), return_template=True)
        loss_b = -(_acts_b / meis[0]["activation"]).mean() - 2 * reg.reg_term(apply_mask(img_post_b, torch.from_numpy(meis[0]["mask"]), -1.0, 1.0))
        loss_b.backward()
        db_ = dict(loss=npy(loss_b))
        for l in range(5):
            db_[f"dW{l}"] = npy(getattr(im.template.func, f"layer{l}").weight.grad)
            db_[f"db{l}"] = npy(getattr(im.template.func, f"layer{l}").bias.grad)
        np.savez_compressed(os.path.join(OUT, "learn_step_20_gb.npz"), **db_)
        # ---- match step on targets [1, 2]
        targets = [1, 2]
        im.template_neuron_idx = 0
        im.target_neuron_idxs = targets
        t = im.template
        for p in t.func.parameters():
            p.grad = None
        t.initialize_coordinate_transform(2, True, False, 0.1, True, True, False, True)
        rf = {k: np.array([v["center_pos"]["x"], v["center_pos"]["y"]], np.float32) for k, v in meis.items()}
        t.register_coords_shifts(torch.from_numpy(rf[0]), torch.from_numpy(np.array([rf[k] for k in targets])))
        aff = t.coordinate_transform.transforms.Affine
        with torch.no_grad():
            aff.angles.copy_(torch.tensor([0.7, 4.1]))
            aff.scalings.copy_(torch.tensor([[1.1, 0.9], [0.8, 1.2]]))
            aff.shears.copy_(torch.tensor([[0.05, -0.02], [-0.1, 0.08]]))
            aff.translation.copy_(torch.tensor([[[0.03, -0.04]], [[-0.05, 0.02]]]))
        t.train()
        loc = np.arange(2)
        img_pre, img_post, _acts, _ = im.forward(grid, t, im.img_transforms, model, return_template=False, other_neurons_loc_in_list=loc, other_neurons_loc_in_model=targets)
        rel = _acts[range(2), :, range(2)]
        mei_acts = torch.tensor([meis[k]["activation"] for k in targets])
        acts = rel.mean(dim=1) / mei_acts
        loss = -acts.mean()
        loss.backward()
        d = template_arrays(t)
        d.update(
            grid=npy(grid)[:, 0], loss=npy(loss), acts=npy(acts), acts_dn=npy(rel), img_pre=npy(img_pre), img_post=npy(img_post),
            angles=npy(aff.angles), scalings=npy(aff.scalings), shears=npy(aff.shears), translation=npy(aff.translation)[:, 0],
            d_angles=npy(aff.angles.grad), d_scalings=npy(aff.scalings.grad), d_shears=npy(aff.shears.grad),
            d_translation=npy(aff.translation.grad)[:, 0], template_xy=rf[0], target_xy=np.array([rf[k] for k in targets]),
            mei_acts=npy(mei_acts), targets=np.array(targets),
        )
        np.savez_compressed(os.path.join(OUT, "match_step_20.npz"), **d)


def case_trajectory(laminr):
    """A short `learn` trajectory (30 Adam steps, demo1 @20x20) replaying inv_temp_learn_match.py:185-208 with
    the reference modules: per-step loss / mean act / sim, and the final parameters."""
    from laminr.contrastive_loss import SimCLROnGrid
    from laminr.utils.mask import apply_mask
    from laminr.utils.general import SingleNeuronModel
    from laminr.utils.training import JitteringGridDatamodule

    res = 20
    model = laminr.neuron_models.simulated("demo1", img_res=[res, res])
    meis = synthetic_meis(model, res, None)
    torch.manual_seed(0)
    im = laminr.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    d = {"init_" + k: v for k, v in template_arrays(im.template).items()}
    dm = JitteringGridDatamodule(1, 20, 30)
    reg = SimCLROnGrid(1, 20, 0.1, 0.3, True)
    opt = torch.optim.Adam(im.template.func.parameters(), lr=1e-3)
    snm = SingleNeuronModel(model, 0)
    mask = torch.from_numpy(meis[0]["mask"])
    torch.manual_seed(123)
    state = torch.get_rng_state()
    grids = [b for b in dm.train_dataloader()]
    torch.set_rng_state(state)
    d["jitter_u01"] = npy(torch.rand([30, 1]))[:, 0]
    losses, means, sims = [], [], []
    for gr in grids:
        img_pre, img_post, _acts, _ = im.forward(gr, im.template, im.img_transforms, snm, return_template=True)
        acts = _acts / meis[0]["activation"]
        sim = reg.reg_term(apply_mask(img_post, mask, -1.0, 1.0))
        loss = -acts.mean() - 2 * sim
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item()), means.append(acts.mean().item()), sims.append(sim.item())
    d.update({"final_" + k: v for k, v in template_arrays(im.template).items()})
    d.update(losses=np.array(losses), means=np.array(means), sims=np.array(sims), mask=meis[0]["mask"], grids=npy(torch.stack(grids))[:, :, 0])
    np.savez_compressed(os.path.join(OUT, "learn_traj_20.npz"), **d)


def export_convcore_arrays(model):
    """Flat arrays of a FiringRateEncoder(Stacked2dCore, FullGaussian2d) in eval mode (the layout oracle.convcore_oracle.model_from_arrays reads)."""
    core, ro = model.core, model.readout
    d = dict(n_layers=np.int64(len(core.features)), xs=np.float64(core.elu_xshift), ys=np.float64(core.elu_yshift), offset=np.float64(model.offset),
             align_corners=np.bool_(ro.align_corners), mu=npy(ro.mu).reshape(-1, 2), features=npy(ro.features).reshape(ro.features.shape[1], -1))
    if ro.bias is not None:
        d["bias"] = npy(ro.bias)
    for l, feat in enumerate(core.features):
        d[f"w{l}"] = npy(feat.conv.weight)
        if feat.conv.bias is not None:
            d[f"b{l}"] = npy(feat.conv.bias)
        if hasattr(feat, "norm"):
            d[f"bn{l}_mean"], d[f"bn{l}_var"] = npy(feat.norm.running_mean), npy(feat.norm.running_var)
            d[f"bn{l}_gamma"], d[f"bn{l}_beta"], d[f"bn{l}_eps"] = npy(feat.norm.weight), npy(feat.norm.bias), np.float64(feat.norm.eps)
        d[f"nonlin{l}"] = np.bool_(hasattr(feat, "nonlin"))
    return d


def case_convcore(laminr):
    """FiringRateEncoder(Stacked2dCore, FullGaussian2d).eval() from the reference's vendored neuralpredictors: responses of all neurons,
    and the input gradient of sum_b g_b * act[b, neuron_b] (what SingleNeuronModel + autograd give the learn / match / MEI loops)."""
    import warnings

    from neuralpredictors.layers.cores.conv2d import Stacked2dCore
    from neuralpredictors.layers.encoders.firing_rate import FiringRateEncoder
    from neuralpredictors.layers.readouts.gaussian import FullGaussian2d

    cfgs = (
        # name, C_in, hidden, k_in, k_hid, layers, (H, W), n_neurons, B, final_nonlin, elu_shift, offset, align_corners
        ("convcore_a", 1, 8, 9, 7, 3, (40, 40), 6, 4, True, (0, 0), 0.0, True),       # config-4 kernel sizes, small frame: cones clipped by every border
        ("convcore_b", 2, 4, 5, 3, 2, (24, 30), 5, 3, False, (0.3, 0.1), 0.2, True),  # 2 input channels, non-square, linear last layer, shifted ELU
        ("convcore_c", 1, 12, 3, 3, 4, (17, 13), 4, 2, True, (0, 0), -0.1, False),    # 4 layers, align_corners=False
    )
    for name, cin, hc, kin, khid, L, (H, W), n, B, fnl, shift, off, ac in cfgs:
        torch.manual_seed(11)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            core = Stacked2dCore(input_channels=cin, hidden_channels=hc, input_kern=kin, hidden_kern=khid, layers=L, stack=-1, pad_input=True,
                                 final_nonlinearity=fnl, elu_shift=shift)
            ro = FullGaussian2d(in_shape=(hc, H, W), outdims=n, bias=True, align_corners=ac)
        g = torch.Generator().manual_seed(5)
        with torch.no_grad():  # a trained-looking state: non-trivial batch-norm statistics, biases, feature vectors and positions
            for feat in core.features:
                feat.norm.running_mean.copy_(torch.randn(hc, generator=g) * 0.3)
                feat.norm.running_var.copy_(torch.rand(hc, generator=g) * 1.5 + 0.25)
                feat.norm.weight.copy_(torch.rand(hc, generator=g) + 0.5)
                feat.norm.bias.copy_(torch.randn(hc, generator=g) * 0.2)
                feat.conv.weight.mul_(2.0)
                if feat.conv.bias is not None:
                    feat.conv.bias.copy_(torch.randn(hc, generator=g) * 0.1)
            ro._features.copy_(torch.randn(ro._features.shape, generator=g) * 0.5)
            ro.bias.copy_(torch.randn(n, generator=g) * 0.2)
            mu = torch.rand(1, n, 1, 2, generator=g) * 2 - 1
            mu[0, 0, 0] = torch.tensor([1.0, -1.0])     # exactly on two borders
            mu[0, 1, 0] = torch.tensor([-0.97, 0.99])   # cone clipped by two borders
            mu[0, 2, 0] = torch.tensor([1.3, 0.2])      # outside: clamped by sample_grid
            ro._mu.copy_(mu)
        model = FiringRateEncoder(core, ro, elu_offset=off).eval()
        x = (torch.randn(B, cin, H, W, generator=g) * 0.5).requires_grad_(True)
        neuron_of_image = torch.tensor([(3 * b + 1) % n for b in range(B)])
        neuron_of_image[0] = 0
        g_act = torch.randn(B, generator=g)
        acts = model(x)
        (acts[torch.arange(B), neuron_of_image] * g_act).sum().backward()
        d = export_convcore_arrays(model)   # after the forward: sample_grid clamps mu in place (gaussian.py:408)
        d.update(x=npy(x), acts=npy(acts), neuron_of_image=neuron_of_image.numpy(), g_act=npy(g_act), g_x=npy(x.grad))
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **d)


def case_conv_learn_match(laminr):
    """One learn hot step and one match hot step of the reference's InvarianceManifold with a conv-core + Gaussian-readout response model
    (FiringRateEncoder(Stacked2dCore, FullGaussian2d).eval(), BASELINE config 4 at a small size)."""
    import warnings

    from laminr.contrastive_loss import SimCLROnGrid
    from laminr.utils.general import SingleNeuronModel
    from laminr.utils.mask import apply_mask
    from neuralpredictors.layers.cores.conv2d import Stacked2dCore
    from neuralpredictors.layers.encoders.firing_rate import FiringRateEncoder
    from neuralpredictors.layers.readouts.gaussian import FullGaussian2d

    res, hc, n = 24, 8, 3
    torch.manual_seed(21)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        core = Stacked2dCore(input_channels=1, hidden_channels=hc, input_kern=5, hidden_kern=3, layers=3, stack=-1, pad_input=True)
        ro = FullGaussian2d(in_shape=(hc, res, res), outdims=n, bias=True)
    g = torch.Generator().manual_seed(9)
    with torch.no_grad():
        for feat in core.features:
            feat.norm.running_mean.copy_(torch.randn(hc, generator=g) * 0.2)
            feat.norm.running_var.copy_(torch.rand(hc, generator=g) + 0.5)
            feat.norm.weight.copy_(torch.rand(hc, generator=g) + 0.5)
            feat.norm.bias.copy_(torch.randn(hc, generator=g) * 0.2)
            feat.conv.weight.mul_(3.0)
        ro._features.copy_(torch.randn(ro._features.shape, generator=g))
        ro._mu.copy_(torch.tensor([[0.2, 0.2], [-0.2, -0.2], [-0.2, 0.2]]).reshape(1, n, 1, 2))  # (x, y): the synthetic RF centres below
    model = FiringRateEncoder(core, ro).eval()
    meis = synthetic_meis(model, res, g)
    for i in meis:
        meis[i]["activation"] = 1.25 + 0.1 * i
    torch.manual_seed(0)
    im = laminr.InvarianceManifold(model, meis, required_img_norm=2.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    with torch.no_grad():
        for p in im.template.func.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.1)
    N = 20
    grid = (torch.linspace(0, 2 * np.pi, N + 1)[:-1] + np.pi / N + 0.0123).reshape(N, 1)
    reg = SimCLROnGrid(1, N, 0.1, 0.3, True)
    img_pre, img_post, _acts, _ = im.forward(grid, im.template, im.img_transforms, SingleNeuronModel(model, 0), return_template=True)
    acts = _acts / meis[0]["activation"]
    sim = reg.reg_term(apply_mask(img_post, torch.from_numpy(meis[0]["mask"]), -1.0, 1.0))
    loss = -acts.mean() - 2 * sim
    loss.backward()
    d = template_arrays(im.template)
    d.update({"m_" + k: v for k, v in export_convcore_arrays(model).items()})
    d.update(grid=npy(grid)[:, 0], mask=meis[0]["mask"], loss=npy(loss), sim=npy(sim), acts=npy(acts).reshape(-1), img_pre=npy(img_pre).reshape(N, -1),
             img_post=npy(img_post).reshape(N, -1), reg_scale=np.float32(2.0), mei_act=np.float64(meis[0]["activation"]), norm=np.float64(2.0))
    for l in range(5):
        d[f"dW{l}"] = npy(getattr(im.template.func, f"layer{l}").weight.grad)
        d[f"db{l}"] = npy(getattr(im.template.func, f"layer{l}").bias.grad)
    # ---- match step on targets [1, 2]
    targets = [1, 2]
    im.template_neuron_idx, im.target_neuron_idxs = 0, targets
    t = im.template
    for p in t.func.parameters():
        p.grad = None
    t.initialize_coordinate_transform(2, True, False, 0.1, True, True, False, True)
    rf = {k: np.array([v["center_pos"]["x"], v["center_pos"]["y"]], np.float32) for k, v in meis.items()}
    t.register_coords_shifts(torch.from_numpy(rf[0]), torch.from_numpy(np.array([rf[k] for k in targets])))
    aff = t.coordinate_transform.transforms.Affine
    with torch.no_grad():
        aff.angles.copy_(torch.tensor([0.7, 4.1]))
        aff.scalings.copy_(torch.tensor([[1.1, 0.9], [0.8, 1.2]]))
        aff.shears.copy_(torch.tensor([[0.05, -0.02], [-0.1, 0.08]]))
        aff.translation.copy_(torch.tensor([[[0.03, -0.04]], [[-0.05, 0.02]]]))
    t.train()
    loc = np.arange(2)
    img_pre, img_post, _acts, _ = im.forward(grid, t, im.img_transforms, model, return_template=False, other_neurons_loc_in_list=loc,
                                             other_neurons_loc_in_model=targets)
    rel = _acts[range(2), :, range(2)]
    mei_acts = torch.tensor([meis[k]["activation"] for k in targets])
    acts = rel.mean(dim=1) / mei_acts
    loss = -acts.mean()
    loss.backward()
    d.update(match_loss=npy(loss), match_acts=npy(acts), match_acts_dn=npy(rel), match_img_post=npy(img_post), angles=npy(aff.angles),
             scalings=npy(aff.scalings), shears=npy(aff.shears), translation=npy(aff.translation)[:, 0], d_angles=npy(aff.angles.grad),
             d_scalings=npy(aff.scalings.grad), d_shears=npy(aff.shears.grad), d_translation=npy(aff.translation.grad)[:, 0], template_xy=rf[0],
             target_xy=np.array([rf[k] for k in targets]), mei_acts=npy(mei_acts), targets=np.array(targets))
    np.savez_compressed(os.path.join(OUT, "conv_learn_match_24.npz"), **d)


def case_create_mask(laminr):
    """The reference's own create_mask (mask.py:25-64; skimage's label / convex_hull_image through the import stubs of ref_harness) on
    seeded MEI-like images: 16 at 40x40, 4 at 100x100, 2 at 37x53 (non-square), thresholds 0.3 (get_mei_dict's) and 1.5 (the default)."""
    from laminr.utils.mask import create_mask
    from oracle.mask_oracle import synthetic_meis

    d = {}
    for tag, (n, H, W, seed) in dict(a=(16, 40, 40, 1), b=(4, 100, 100, 2), c=(2, 37, 53, 3)).items():
        meis = synthetic_meis(n, H, W, seed)
        for zt in (0.3, 1.5):
            masks, cents = [], []
            for i in range(n):
                m, c = create_mask(torch.from_numpy(meis[i])[None], zscore_thresh=zt, return_mask_centroid=True)
                masks.append(npy(m)), cents.append([c["y"], c["x"]])
            d[f"{tag}_mask_z{zt}"] = np.stack(masks)
            d[f"{tag}_centre_z{zt}"] = np.array(cents, np.float64)
        d[f"{tag}_shape_seed"] = np.array([n, H, W, seed])
    np.savez_compressed(os.path.join(OUT, "create_mask.npz"), **d)


def case_gaussian_blur(laminr):
    """featurevis.ops.GaussianBlur of the reference (ops.py:378-423) on seeded images.  The reference calls `scipy.signal.gaussian`, which
    scipy >= 1.13 only provides as `scipy.signal.windows.gaussian`: the alias is restored here (nothing else is patched)."""
    import scipy.signal
    import scipy.signal.windows

    if not hasattr(scipy.signal, "gaussian"):
        scipy.signal.gaussian = scipy.signal.windows.gaussian
    import featurevis.ops as fops

    d = {}
    rng = np.random.RandomState(5)
    cases = dict(a=((3, 1, 20, 24), 1.0, 4, "reflect"), b=((2, 2, 17, 9), (1.5, 0.6), 4, "reflect"), c=((2, 1, 12, 12), 2.0, 3, "replicate"),
                 e=((1, 1, 30, 30), 0.4, 4, "constant"), f=((1, 1, 100, 100), 1.5, 4, "reflect"))
    for tag, (shape, sigma, trunc, mode) in cases.items():
        x = rng.randn(*shape).astype(np.float32)
        y = fops.GaussianBlur(sigma, truncate=trunc, pad_mode=mode)(torch.from_numpy(x))
        d[f"{tag}_x"], d[f"{tag}_y"] = x, npy(y)
        d[f"{tag}_cfg"] = np.array([sigma[0] if isinstance(sigma, tuple) else sigma, sigma[1] if isinstance(sigma, tuple) else sigma, trunc,
                                    {"constant": 0, "reflect": 1, "replicate": 2}[mode]], np.float64)
    # decay: sigma + decay_factor * (iteration - 1)  (ops.py:403-406)
    x = rng.randn(1, 1, 16, 16).astype(np.float32)
    d["g_x"], d["g_y"] = x, npy(fops.GaussianBlur(1.0, decay_factor=0.25)(torch.from_numpy(x), iteration=3))
    np.savez_compressed(os.path.join(OUT, "gaussian_blur.npz"), **d)


def main():
    os.makedirs(OUT, exist_ok=True)
    laminr = import_reference()
    torch.set_num_threads(os.cpu_count())
    import sys
    only = set(sys.argv[1:])
    for fn in (case_inr, case_match_inr, case_constrain, case_simresp, case_banks_100, case_simclr, case_training_utils,
               case_adam, case_mei, case_learn_match_steps, case_trajectory, case_convcore, case_conv_learn_match, case_create_mask, case_gaussian_blur):
        if only and fn.__name__ not in only:
            continue
        fn(laminr)
        print("wrote", fn.__name__)


if __name__ == "__main__":
    main()
