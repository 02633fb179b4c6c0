"""GPU parity at the API level: the drop-in modules / fused steppers / public functions of laminr_b200 against the golden
vectors of the unmodified reference and the numpy oracle."""
import numpy as np
import pytest
import torch

from tests._util import golden, relerr, split_flat

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import laminr_b200
    return laminr_b200


def _meis(res, mask0=None, act=1.0000009536743164):
    yy, xx = np.meshgrid(np.arange(res) + 0.5, np.arange(res) + 0.5, indexing="ij")
    out = {}
    for i, (cx, cy) in enumerate([(0.6 * res, 0.6 * res), (0.4 * res, 0.4 * res), (0.4 * res, 0.6 * res)]):
        mask = np.exp(-0.5 * (((xx - cx) / (0.16 * res)) ** 2 + ((yy - cy) / (0.16 * res)) ** 2)).astype(np.float32)
        out[i] = dict(mei=np.zeros((1, res, res), np.float32), activation=act, mask=mask, center_pos=dict(x=cx, y=cy))
    if mask0 is not None:
        out[0]["mask"] = mask0
    return out


def _manifold(L, res, d=None, prefix="", **kw):
    model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
    torch.manual_seed(0)
    im = L.InvarianceManifold(model, _meis(res, None if d is None else d.get("mask")), required_img_norm=1.0, pixel_value_lower_bound=-1.0,
                              pixel_value_upper_bound=1.0, **kw)
    if d is not None:
        sd = {f"func.layer{l}.weight": torch.from_numpy(d[f"{prefix}W{l}"]) for l in range(5)}
        sd.update({f"func.layer{l}.bias": torch.from_numpy(d[f"{prefix}b{l}"]) for l in range(5)})
        sd.update(B=torch.from_numpy(d[f"{prefix}B"]), auxB=torch.from_numpy(d[f"{prefix}auxB"]))
        im.template.load_state_dict(sd, strict=False)
    return model, im


def test_autograd_learn_step_matches_reference(L):
    """The generic (autograd) composition of the drop-in modules reproduces the reference's loss and parameter gradients."""
    from laminr_b200.contrastive_loss import SimCLROnGrid
    from laminr_b200.utils.general import SingleNeuronModel
    d = golden("learn_step_20")
    model, im = _manifold(L, 20, d)
    grid = torch.from_numpy(d["grid"]).reshape(-1, 1).cuda()
    reg = SimCLROnGrid(1, 20, 0.1, 0.3, True).cuda()
    img_pre, img_post, _acts, _ = im.forward(grid, im.template, im.img_transforms, SingleNeuronModel(model, 0), return_template=True)
    assert img_pre.shape == (20, 1, 20, 20) and _acts.shape == (20, 1)
    acts = _acts / float(d["mei_act"])
    sim = reg.masked_reg_term(img_post, torch.from_numpy(d["mask"]).cuda(), -1.0, 1.0)
    loss = -acts.mean() - 2 * sim
    loss.backward()
    assert abs(loss.item() - d["loss"]) < 3e-5 * abs(d["loss"])
    assert relerr(img_post.detach().cpu().numpy().reshape(20, -1), d["img_post"]) < 3e-5
    for l in range(5):
        lay = getattr(im.template.func, f"layer{l}")
        assert relerr(lay.weight.grad.cpu().numpy(), d[f"dW{l}"]) < 3e-3, l
        assert relerr(lay.bias.grad.cpu().numpy(), d[f"db{l}"]) < 3e-3, l
    # the all-neuron forward of the packed model agrees with per-neuron evaluation
    full = model(img_post.detach())
    assert full.shape == (20, 3)
    assert torch.allclose(full[:, 0], _acts[:, 0].detach(), rtol=1e-6, atol=1e-7)
    # reg_term on pre-masked images == masked_reg_term
    from laminr_b200.utils.mask import apply_mask
    sim2 = reg.reg_term(apply_mask(img_post.detach(), torch.from_numpy(d["mask"]).cuda(), -1.0, 1.0))
    assert abs(sim2.item() - sim.item()) < 1e-6


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_learn_trajectory_matches_reference(L, use_graph):
    """30 Adam steps of the fused stepper from the reference's initial network and jitter stream: per-step loss within 1e-3
    relative (north star), final parameters within 1e-2 relative."""
    from laminr_b200.contrastive_loss import SimCLROnGrid
    from laminr_b200.fused import LearnStepper
    d = golden("learn_traj_20")
    model, im = _manifold(L, 20, d, prefix="init_")
    reg = SimCLROnGrid(1, 20, 0.1, 0.3, True).cuda()
    st = LearnStepper(im.template, im.img_transforms, model, 0, d["mask"], 1.0000009536743164, reg, -1.0, 1.0, 20, lr=1e-3, reg_scale=2.0,
                      use_graph=use_graph)
    grids = torch.from_numpy(d["grids"]).cuda()
    losses, means = [], []
    for i in range(30):
        st.step(grids[i])
        losses.append(st.loss().item())
        means.append(st.last_acts().mean().item())
    np.testing.assert_allclose(losses, d["losses"], rtol=1e-3)
    np.testing.assert_allclose(means, d["means"], rtol=1e-3)
    assert max(abs(a - b) / abs(b) for a, b in zip(losses, d["losses"])) < 2e-4  # what fp32 achieves
    assert st.step_count[0].item() == 30
    dW, db = split_flat(st.flat.cpu().numpy(), d, prefix="final_")
    for l in range(5):
        assert relerr(dW[l], d[f"final_W{l}"]) < 1e-2, l
        assert relerr(dW[l], d[f"final_W{l}"]) < 2e-3, l
    # the nn.Parameters are views of the flat vector the kernels updated
    assert torch.equal(im.template.func.layer1.weight.detach().reshape(-1), st.flat[50 * 201 + 50:50 * 201 + 50 + 2500])


def test_host_fed_step_equals_device_fed_step(L):
    """LearnStepper.step with a HOST row (pinned memory read by the preparation kernel, loss written into pinned memory by the objective kernel,
    one graph launch) walks the same trajectory, bit for bit, as the device-fed step; sync_loss() returns the step's loss."""
    from laminr_b200.contrastive_loss import SimCLROnGrid
    from laminr_b200.fused import LearnStepper
    d = golden("learn_traj_20")
    grids = torch.from_numpy(d["grids"])
    flats, losses = [], []
    for host in (False, True):
        model, im = _manifold(L, 20, d, prefix="init_")
        reg = SimCLROnGrid(1, 20, 0.1, 0.3, True).cuda()
        st = LearnStepper(im.template, im.img_transforms, model, 0, d["mask"], 1.0000009536743164, reg, -1.0, 1.0, 20, lr=1e-3, reg_scale=2.0,
                          precision="bf16x3", use_graph=True)
        ls = []
        for i in range(6):
            if host:
                st.step(grids[i])  # CPU tensor
                ls.append(st.sync_loss())
            else:
                st.step(grids[i].cuda())
                ls.append(st.loss().item())
        flats.append(st.flat.detach().cpu().clone()), losses.append(ls)
        assert st.step_count[0].item() == 6
    assert losses[0] == losses[1]
    assert torch.equal(flats[0], flats[1])
    # mixing the two kinds of step on one stepper
    st.step(grids[6].cuda())
    st.step(grids[7])
    assert np.isfinite(st.sync_loss()) and st.step_count[0].item() == 8


def test_learn_api_fused_equals_generic(L):
    """InvarianceManifold.learn: the fused path and the generic autograd path (user-style nn.Module model) walk the same trajectory."""
    class PlainModel(torch.nn.Module):  # a user-supplied response model: stock PyTorch ops, same maths as the simulated population
        def __init__(self, banks):
            super().__init__()
            self.banks = torch.nn.ParameterList([torch.nn.Parameter(b, requires_grad=False) for b in banks])

        def forward(self, x):
            outs = []
            for b in self.banks:
                r = torch.einsum("bchw,fhw->bf", x, b).max(dim=1)[0]
                outs.append((torch.nn.functional.elu(r) + 1) / 2 + 1e-6)
            return torch.stack(outs, 1)

    res = 16
    results = []
    for kind in ("fused", "generic"):
        model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
        if kind == "generic":
            model = PlainModel([m.filters.clone() for m in model.neuron_models_list]).cuda()
        torch.manual_seed(3)
        im = L.InvarianceManifold(model, _meis(res), required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
        imgs, acts = im.learn(0, steps_per_epoch=4, num_max_epochs=3, min_epochs=0)
        assert imgs.shape == (20, 1, res, res) and acts.shape == (20,)
        assert im.learn_steps_taken == 12
        results.append((imgs, acts, im.template.flat_params().cpu().numpy()))
    assert relerr(results[0][2], results[1][2]) < 1e-4
    assert relerr(results[0][0], results[1][0]) < 1e-3
    assert relerr(results[0][1], results[1][1]) < 1e-4
    # FixEnergyNorm known answer: ||img - mean|| == norm where nothing was clipped
    n = np.linalg.norm(results[0][0].reshape(20, -1), axis=1)
    assert np.all(n <= 1.0 + 1e-5)


def test_match_api_fused_equals_generic_and_reference_step(L):
    d = golden("match_step_20")
    outs = []
    for fused in (True, False):
        model, im = _manifold(L, 20, d)
        if not fused:
            im._fused_ok = lambda gaussian_blur=None: False
        im.template_neuron_idx = 0
        torch.manual_seed(5)
        imgs, acts = im.match([1, 2], num_epochs=6, patience=15)
        assert imgs.shape == (2, 20, 1, 20, 20) and acts.shape == (2, 20)
        aff = im.template.coordinate_transform.transforms.Affine
        outs.append((imgs, acts, [p.detach().cpu().numpy().copy() for p in (aff.angles, aff.scalings, aff.shears, aff.translation)]))
        assert im.match_steps_taken == 6
    for a, b in zip(outs[0][2], outs[1][2]):
        np.testing.assert_allclose(a, b, rtol=2e-4, atol=2e-6)
    assert relerr(outs[0][0], outs[1][0]) < 2e-3
    assert relerr(outs[0][1], outs[1][1]) < 1e-4
    # scalings were initialised to 1 / (mask-size fraction) on both axes (inv_temp_learn_match.py:591,619)
    m = _meis(20)
    frac = np.array([m[k]["mask"].sum() / m[0]["mask"].sum() for k in (1, 2)])
    assert np.allclose(outs[0][2][1], (1 / frac)[:, None], rtol=0.02)  # six Adam steps of 1e-3 later


@pytest.mark.parametrize("kw,epochs", [(dict(num_epochs=4, patience=15), 4), (dict(num_epochs=50, patience=1, ignore_diff_smaller_than=1e9), 2)])
def test_match_leaves_cpu_generator_like_the_reference_loop(L, kw, epochs):
    """The fused match loop draws the next epoch's jitter while the GPU still runs the current one; whether the loop ends by exhausting num_epochs
    or by the stop rule (:368-375), the CPU generator must end where the reference's loop (= the generic path) leaves it."""
    d = golden("match_step_20")
    states, finals = [], []
    for fused in (True, False):
        model, im = _manifold(L, 20, d)
        if not fused:
            im._fused_ok = lambda gaussian_blur=None: False
        im.template_neuron_idx = 0
        torch.manual_seed(11)
        imgs, acts = im.match([1, 2], save_results_every_n_epochs=1, **kw)
        assert im.match_epochs_run == epochs and im.match_steps_taken == epochs
        assert imgs.shape[0] == epochs + 1  # one snapshot per epoch that ran + the final one (a snapshot taken for an undone epoch is dropped)
        states.append(torch.get_rng_state().clone())
        aff = im.template.coordinate_transform.transforms.Affine
        finals.append([p.detach().cpu().numpy().copy() for p in (aff.angles, aff.scalings, aff.shears, aff.translation)] + [np.asarray(acts)])
    assert torch.equal(states[0], states[1])
    # the fused loop runs one epoch ahead of its stop rule and rolls it back: the final transform must be the sequential loop's, not one Adam step
    # (1e-3 per scalar) further
    for a, b in zip(finals[0][:4], finals[1][:4]):
        np.testing.assert_allclose(a, b, rtol=2e-4, atol=5e-5)
    np.testing.assert_allclose(finals[0][4], finals[1][4], rtol=2e-4, atol=2e-5)


def test_match_step_in_target_blocks_equals_one_pass(L, monkeypatch):
    """The match training step walks the targets in blocks when the kept activations of all targets exceed the memory budget (BASELINE config 3
    on few GPUs); forced here with one target per block: same parameters and results as all targets in one pass."""
    d = golden("match_step_20")
    outs = []
    for block in (None, "1"):
        if block is None:
            monkeypatch.delenv("LMR_MATCH_BLOCK", raising=False)
        else:
            monkeypatch.setenv("LMR_MATCH_BLOCK", block)
        model, im = _manifold(L, 20, d)
        im.template_neuron_idx = 0
        torch.manual_seed(5)
        imgs, acts = im.match([1, 2], num_epochs=5, patience=15)
        assert im._match_stepper.block == (2 if block is None else 1)
        aff = im.template.coordinate_transform.transforms.Affine
        outs.append((imgs, acts, [p.detach().cpu().numpy().copy() for p in (aff.angles, aff.scalings, aff.shears, aff.translation)]))
    for a, b in zip(outs[0][2], outs[1][2]):
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-7)
    assert relerr(outs[0][0], outs[1][0]) < 1e-6 and relerr(outs[0][1], outs[1][1]) < 1e-6


def test_batched_init_sweep_equals_candidate_loop(L):
    """MatchStepper.sweep (all candidates x targets as virtual targets of batched launches) == one candidate per pass on the stepper's
    own buffers == the reference's loop (set the angle / translation, evaluate all targets; inv_temp_learn_match.py:620-641)."""
    d = golden("match_step_20")
    model, im = _manifold(L, 20, d)
    im.template_neuron_idx = 0
    torch.manual_seed(3)
    im.match([1, 2], num_epochs=1, patience=15)
    st = im._match_stepper
    grid = im._grid_for(20, "cuda")
    aff = im.template.coordinate_transform.transforms.Affine
    angles = np.linspace(0, 2 * np.pi, 8)[:-1].astype(np.float32)
    tx = np.linspace(-0.1, 0.1, 7).astype(np.float32)
    ty = tx[::-1].copy()
    for with_t in (False, True):
        cand = (tx, ty) if with_t else (None, None)
        a = st.sweep(grid, angles, *cand, max_virtual=128).cpu().numpy()
        b = st.sweep(grid, angles, *cand, max_virtual=1).cpu().numpy()
        saved = [p.detach().clone() for p in (aff.angles, aff.translation)]
        c = np.zeros_like(a)
        with torch.no_grad():
            for k in range(len(angles)):
                aff.angles.data.fill_(float(angles[k]))
                if with_t:
                    aff.translation.data[:, 0, 0] = float(tx[k])
                    aff.translation.data[:, 0, 1] = float(ty[k])
                c[k] = st.evaluate(grid).mean(dim=1).cpu().numpy()
            aff.angles.data.copy_(saved[0]), aff.translation.data.copy_(saved[1])
        assert a.shape == (7, 2) and np.ptp(c, axis=0).min() > 1e-3  # the candidates matter
        np.testing.assert_allclose(a, c, rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(b, c, rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("with_t", [False, True])
def test_init_sweep_vs_reference_golden(L, with_t):
    """Match initialisation sweep on the GPU (MatchStepper.sweep through initialize_with_best_angle_and_translation) against the table and the
    final choice of the UNMODIFIED reference function (tests/golden/init_sweep_20.npz, oracle/make_golden_sweep.py): angle only (the `match` default,
    60 candidates) and angle + translation (60 x 11 x 11 candidates)."""
    from laminr_b200.fused import MatchStepper
    from laminr_b200.utils.training import JitteringGridDatamodule
    d = golden("init_sweep_20")
    res, targets, tag = 20, [1, 2], "at" if with_t else "a"
    model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
    xy = np.concatenate([d["template_xy"][None], d["target_xy"]])
    masks = np.concatenate([d["template_mask"][None], d["target_masks"]])
    meis = {i: dict(mei=np.zeros((1, res, res), np.float32), activation=1.0 if i == 0 else float(d["mei_acts"][i - 1]), mask=masks[i],
                    center_pos=dict(x=float(xy[i, 0]), y=float(xy[i, 1]))) for i in range(3)}
    torch.manual_seed(0)
    im = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    sd = {f"func.layer{l}.weight": torch.from_numpy(d[f"W{l}"]) for l in range(5)}
    sd.update({f"func.layer{l}.bias": torch.from_numpy(d[f"b{l}"]) for l in range(5)})
    sd.update(B=torch.from_numpy(d["B"]), auxB=torch.from_numpy(d["auxB"]))
    im.template.load_state_dict(sd, strict=False)
    im.template_neuron_idx, im.target_neuron_idxs = 0, targets
    t = im.template
    t.initialize_coordinate_transform(2, True, False, 0.1, True, True, False, True)
    t.register_coords_shifts(torch.from_numpy(d["template_xy"]).cuda(), torch.from_numpy(d["target_xy"]).cuda())
    aff = t.coordinate_transform.transforms.Affine
    with torch.no_grad():
        aff.angles.copy_(torch.from_numpy(d["angles0"])), aff.scalings.copy_(torch.from_numpy(d["scalings0"]))
        aff.shears.copy_(torch.from_numpy(d["shears0"])), aff.translation.copy_(torch.from_numpy(d["translation0"])[:, None, :])
    st = MatchStepper(t, im.img_transforms, model, targets, d["mei_acts"].astype(np.float32), 20, precision=im.precision)
    dl = JitteringGridDatamodule(num_invariances=1, grid_points_per_dim=20, steps_per_epoch=1)
    im.initialize_with_best_angle_and_translation(t, model, im.img_transforms, dl, np.arange(2), targets, template_rf_mask=d["template_mask"],
                                                  others_rf_mask=d["target_masks"], rotate_angle_and_scale=True, find_best_translation=with_t,
                                                  _stepper=st)
    np.testing.assert_array_equal(aff.angles.detach().cpu().numpy(), d[f"angles_{tag}"])
    np.testing.assert_allclose(aff.scalings.detach().cpu().numpy(), d[f"scalings_{tag}"], rtol=1e-6)
    np.testing.assert_allclose(aff.translation.detach().cpu().numpy()[:, 0], d[f"translation_{tag}"], rtol=0, atol=1e-7)
    # the table itself (angle-only candidates at the scalings the sweep uses, the translation at its initial value)
    if not with_t:
        angles = np.linspace(0, 2 * np.pi, 61)[:-1].astype(np.float32)
        got = st.sweep(dl.grid.cuda(), angles).cpu().numpy()
        assert relerr(got, d["table_a"][:, 0, 0]) < 1e-4


def test_match_call_vs_reference_call(L):
    """A whole `match([1, 2], num_epochs=5)` call (transform initialisation, 60-angle sweep, five jittered Adam epochs, final snapshot) against the
    same call of the UNMODIFIED reference with the same seeds (tests/golden/match_call_20.npz, oracle/make_golden_match_call.py): final affine
    parameters, images and activations within the north star's tolerances, and the CPU generator left in the same state."""
    d = golden("match_call_20")
    res = 20
    model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
    meis = {i: dict(mei=np.zeros((1, res, res), np.float32), activation=float(d["mei_acts"][i]), mask=d["masks"][i],
                    center_pos=dict(x=float(d["centers"][i, 0]), y=float(d["centers"][i, 1]))) for i in range(3)}
    torch.manual_seed(0)
    im = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    sd = {f"func.layer{l}.weight": torch.from_numpy(d[f"W{l}"]) for l in range(5)}
    sd.update({f"func.layer{l}.bias": torch.from_numpy(d[f"b{l}"]) for l in range(5)})
    sd.update(B=torch.from_numpy(d["B"]), auxB=torch.from_numpy(d["auxB"]))
    im.template.load_state_dict(sd, strict=False)
    im.template_neuron_idx = 0
    torch.manual_seed(5)
    imgs, acts = im.match([1, 2], num_epochs=5)
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(d["rng_after"]))
    aff = im.template.coordinate_transform.transforms.Affine
    np.testing.assert_allclose(aff.angles.detach().cpu().numpy(), d["angles"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(aff.scalings.detach().cpu().numpy(), d["scalings"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(aff.shears.detach().cpu().numpy(), d["shears"], rtol=0, atol=2e-4)        # five Adam steps of 1e-3 from 0
    np.testing.assert_allclose(aff.translation.detach().cpu().numpy()[:, 0], d["translation"], rtol=0, atol=2e-4)
    assert imgs.shape == d["imgs"].shape and acts.shape == d["acts"].shape
    assert relerr(imgs, d["imgs"]) < 1e-2 and relerr(acts, d["acts"]) < 1e-2  # north star: final images and activations within 1e-2 relative


def test_learn_call_vs_reference_call(L, capsys):
    """Whole `learn(0, ...)` calls from the constructor's own seeded initial network (nothing loaded) against the same calls of the UNMODIFIED reference
    with the same seeds (tests/golden/learn_call_20.npz, oracle/make_golden_learn_call.py).

    short: 20 steps in 4 epochs -- final network, images and activations within the north star's 1e-2 relative, CPU generator in the same state.
    long : 140 steps in 14 epochs (the epochs past min_epochs=10 run the reg_scale scheduler and return the activations of the un-jittered evaluation
           pass) -- generator state, step count and shapes exactly; the network only to the reference's OWN reproducibility: the unmodified reference
           rerun with 1 or 3 BLAS threads instead of 8 (summation order is the only change) lands 3.7e-2 from this golden in the weights, 0.23..0.36 in
           the images and 0.9086 / 0.8862 instead of 0.9239 in the mean activation (Adam turns rounding-level differences of near-zero gradient
           entries into lr-sized steps); measured here on B200: 2.8e-2 in the weights, mean activation 0.91."""
    d = golden("learn_call_20")
    res = 20
    for tag, kw in (("short", dict(steps_per_epoch=5, num_max_epochs=4)), ("long", dict(steps_per_epoch=10, num_max_epochs=14))):
        model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
        meis = {i: dict(mei=np.zeros((1, res, res), np.float32), activation=float(d["mei_acts"][i]), mask=d["masks"][i],
                        center_pos=dict(x=float(d["centers"][i, 0]), y=float(d["centers"][i, 1]))) for i in range(3)}
        torch.manual_seed(0)
        im = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
        torch.manual_seed(11)
        imgs, acts = im.learn(0, **kw)
        assert torch.equal(torch.get_rng_state(), torch.from_numpy(d[f"{tag}_rng_after"])), tag
        assert im.learn_steps_taken == kw["steps_per_epoch"] * kw["num_max_epochs"]
        assert imgs.shape == d[f"{tag}_imgs"].shape and acts.shape == d[f"{tag}_acts"].shape
        sd = im.template.state_dict()
        errs = [relerr(sd[f"func.layer{l}.weight"].cpu().numpy(), d[f"{tag}_W{l}"]) for l in range(5)]
        e_img, e_act = relerr(imgs, d[f"{tag}_imgs"]), relerr(acts, d[f"{tag}_acts"])
        with capsys.disabled():
            print(f"\n[learn call {tag}] weight rel err {max(errs):.2e}, images {e_img:.2e}, activations {e_act:.2e}, mean activation {float(np.mean(acts)):.4f}")
        if tag == "short":
            assert max(errs) < 1e-2, (tag, errs)
            assert e_img < 1e-2 and e_act < 1e-2, (tag, e_img, e_act)  # north star: final images and activations within 1e-2 relative
        else:
            assert max(errs) < 1e-1, (tag, errs)
            # the reference itself: 0.015 (3 threads), 0.038 (1 thread) away from its 8-thread run; B200 runs of this repo: 0.013 (round 1), 0.081 (round 2,
            # after the backward started taking y from the forward's record instead of rebuilding it: a last-bit change) -- the trajectory is chaotic;
            # what bounds the arithmetic near convergence is tests/test_gpu_round2.py::test_learn_step_late_in_the_trajectory_vs_reference
            assert abs(float(np.mean(acts)) - float(d["long_acts"].mean())) < 0.15 and 0.75 < float(np.mean(acts)) < 1.0


def test_get_mei_dict_vs_reference_call_100(L):
    """`get_mei_dict` on demo1 at 100x100 (BASELINE configs[0], first call of the quick start) against the output of the UNMODIFIED reference call with
    the same seeds (tests/golden/mei_dict_100.npz, oracle/make_golden_meis.py): activations, MEI images, RF masks and their centroids."""
    d = golden("mei_dict_100")
    torch.manual_seed(0)
    np.random.seed(0)
    model = L.neuron_models.simulated("demo1", img_res=[100, 100]).cuda()
    meis = L.get_mei_dict(model, [1, 100, 100], pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, required_img_norm=1.0)
    assert sorted(meis) == [0, 1, 2]
    for i in range(3):
        assert abs(float(meis[i]["activation"]) - d["activations"][i]) < 2e-6
        assert relerr(np.asarray(meis[i]["mei"], np.float32), d["meis"][i].astype(np.float32)) < 3e-3  # float16 fixture (5e-4) + 1000 ascent steps
        m, want = np.asarray(meis[i]["mask"], np.float32), d["masks"][i]
        assert m.shape == want.shape == (100, 100)
        assert (m > 0.5).sum() == (want > 0.5).sum()
        assert np.abs(m - want).mean() < 1e-4 and np.abs(m - want).max() < 0.35  # at most a hull-boundary pixel apart
        assert abs(meis[i]["center_pos"]["x"] - d["centers"][i, 0]) < 0.05 and abs(meis[i]["center_pos"]["y"] - d["centers"][i, 1]) < 0.05


def test_get_mei_dict_known_answer(L):
    """demo1 filters are unit-norm and images are constrained to norm 1, so the MEI activation is (elu(1)+1)/2 + 1e-6
    (the reference measures 1.0000009536743164 for all three neurons)."""
    model = L.neuron_models.simulated("demo1", img_res=[100, 100]).cuda()
    torch.manual_seed(0)
    np.random.seed(0)
    meis = L.get_mei_dict(model, [1, 100, 100], -1.0, 1.0, required_img_norm=1.0)
    assert sorted(meis) == [0, 1, 2]
    for k, v in meis.items():
        assert v["mei"].shape == (1, 100, 100) and v["mask"].shape == (100, 100)
        assert abs(v["activation"] - 1.0000009536743164) < 2e-6, v["activation"]
        assert abs(np.linalg.norm(v["mei"]) - 1.0) < 1e-5
    # RF centres land on the filter centres of demo1: loc (0.2,0.2) -> pixel (60,60); loc (-0.2,-0.2) under the global 45-degree
    # transform sits at T^-1 loc = (-0.283, 0) -> pixel (35.9, 50); neuron 2's three Gabors are centred near loc (-0.2,0.2) -> (40,60)
    for k, (x, y) in {0: (60, 60), 1: (35.9, 50), 2: (40, 60)}.items():
        assert abs(meis[k]["center_pos"]["x"] - x) < 4 and abs(meis[k]["center_pos"]["y"] - y) < 4, (k, meis[k]["center_pos"])
    # RF mask sizes measured on the unmodified reference for the same seed: 1033 / 1035 / 1621 px
    for k, want in {0: 1033, 1: 1035, 2: 1621}.items():
        assert abs(meis[k]["mask"].sum() - want) < 0.03 * want, (k, meis[k]["mask"].sum())


def test_generic_mei_path_matches_batched_kernel(L):
    """featurevis.gradient_ascent (generic autograd path) == lmr_mei_ascent on the same start image."""
    from laminr_b200 import featurevis
    from laminr_b200.featurevis import ops as fops
    from laminr_b200.utils.general import SingleNeuronModel
    from laminr_b200.utils.mei import batched_simulated_meis
    d = golden("mei")
    model = L.neuron_models.simulated("demo1", img_res=[16, 16]).cuda()
    x0 = torch.from_numpy(d["norm_x0_1"]).cuda()
    post = featurevis.Compose([fops.ChangeNorm(norm=1.0, mean=0.0), fops.ClipRange(-1.0, 1.0)])
    x, fevals, _ = featurevis.gradient_ascent(SingleNeuronModel(model, 1), post(x0), step_size=10, num_iterations=60, post_update=post, print_iters=1001)
    assert len(fevals) == 61
    np.testing.assert_allclose(fevals, d["norm_fevals_1"], rtol=5e-5)
    xb, fb = batched_simulated_meis(model, [1], x0, -1.0, 1.0, norm=1.0, step_size=10, num_iterations=60)
    np.testing.assert_allclose(fb[0].cpu().numpy(), fevals, rtol=2e-5)
    assert relerr(xb.cpu().numpy(), x.cpu().numpy()) < 1e-4


def test_state_dict_round_trip_and_getters(L):
    model, im = _manifold(L, 16)
    sd = {k: v.clone() for k, v in im.template.state_dict().items()}
    imgs = im.get_images_from_learned_template(n_images=7)
    assert imgs.shape == (7, 1, 16, 16)
    model2, im2 = _manifold(L, 16)
    with torch.no_grad():
        for p in im2.template.parameters():
            p.add_(1.0)
    im2.template.load_state_dict(sd)
    np.testing.assert_array_equal(im2.get_images_from_learned_template(n_images=7), imgs)
    im.template_neuron_idx = 0
    im.match([2], num_epochs=2)
    out = im.get_images_from_matched_template(2, n_images=5)
    assert out.shape == (5, 1, 16, 16)
    with pytest.raises(ValueError):
        im.get_images_from_matched_template(1)
    with pytest.raises(ValueError):
        im.get_images_from_matched_template("2")
