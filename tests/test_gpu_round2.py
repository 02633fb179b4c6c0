"""GPU regression tests of round 2: the advisor's findings (sweep tail batch, zero-epoch calls, stale weight packs)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import laminr_b200
    return laminr_b200


def _pop_meis(res, n, seed=0):
    yy, xx = np.meshgrid(np.arange(res) + 0.5, np.arange(res) + 0.5, indexing="ij")
    rs = np.random.RandomState(seed)
    out = {}
    for i in range(n):
        cx, cy = rs.uniform(0.35, 0.65, 2) * res
        mask = np.exp(-0.5 * (((xx - cx) / (0.15 * res)) ** 2 + ((yy - cy) / (0.15 * res)) ** 2)).astype(np.float32)
        out[i] = dict(mei=np.zeros((1, res, res), np.float32), activation=1.0, mask=mask, center_pos=dict(x=float(cx), y=float(cy)))
    return out


@pytest.mark.parametrize("res,D,translation", [(24, 8, False), (24, 13, False), (40, 7, False), (32, 9, True)])
def test_sweep_tail_batch_has_workspace(L, res, D, translation):
    """The batched init sweep's last, smaller candidate batch calls the simulated response with fewer groups than the batch it was sized for; the
    partials' chunk size shrinks with the group count, so the workspace need is NOT monotonic in it (these (res, D) combinations failed with
    'simresp_fwd: workspace too small').  The batched table must equal the candidate-by-candidate loop."""
    from laminr_b200.fused import MatchStepper
    pop = L.neuron_models.simulated_population(D + 1, img_res=[res, res], seed=3).cuda()
    meis = _pop_meis(res, D + 1)
    torch.manual_seed(0)
    im = L.InvarianceManifold(pop, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    t = im.template
    t.initialize_coordinate_transform(D, True, False, 0.1, True, True, False, True)
    rfs = torch.tensor([[meis[i]["center_pos"]["x"], meis[i]["center_pos"]["y"]] for i in range(D + 1)], dtype=torch.float32, device="cuda")
    t.register_coords_shifts(rfs[0], rfs[1:])
    ms = MatchStepper(t, im.img_transforms, pop, list(range(1, D + 1)), np.ones(D, np.float32), 20, use_graph=False)
    grid = im._grid_for(20, "cuda")
    angles = np.linspace(0, 2 * np.pi, 61)[:-1].astype(np.float32)
    if translation:
        tr = np.linspace(-0.1, 0.1, 3).astype(np.float32)
        ii, jj, kk = np.meshgrid(np.arange(60), np.arange(3), np.arange(3), indexing="ij")
        cand = (angles[ii.ravel()], tr[jj.ravel()], tr[kk.ravel()])
    else:
        cand = (angles, None, None)
    batched = ms.sweep(grid, *cand).cpu().numpy()          # several candidates per pass incl. a smaller tail batch
    loop = ms.sweep(grid, *cand, max_virtual=1).cpu().numpy()  # one candidate per pass on the stepper's own buffers
    assert batched.shape == (len(cand[0]), D)
    np.testing.assert_allclose(batched, loop, rtol=2e-5, atol=2e-6)


def test_match_with_default_sweep_at_the_failing_size(L):
    """Whole match() with the default rotate_angle_and_scale=True at 24x24 with 8 targets (crashed in the sweep's tail batch)."""
    res, D = 24, 8
    pop = L.neuron_models.simulated_population(D + 1, img_res=[res, res], seed=3).cuda()
    torch.manual_seed(0)
    im = L.InvarianceManifold(pop, _pop_meis(res, D + 1), required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    im.template_neuron_idx = 0
    imgs, acts = im.match(list(range(1, D + 1)), num_epochs=3)
    assert imgs.shape == (D, 20, 1, res, res) and acts.shape == (D, 20) and np.isfinite(acts).all()


def test_zero_epoch_calls_return_the_snapshot(L):
    """learn(num_max_epochs=0) / match(num_epochs=0) return the snapshot of the untouched template like the reference (they raised NameError)."""
    res = 20
    model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
    torch.manual_seed(0)
    im = L.InvarianceManifold(model, _pop_meis(res, 3), required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    imgs, acts = im.learn(0, num_max_epochs=0)
    assert imgs.shape == (20, 1, res, res) and acts.shape == (20,) and im.learn_epochs_run == 0 and im.learn_steps_taken == 0
    np.testing.assert_allclose(imgs, im.get_images_from_learned_template(20), rtol=1e-5, atol=1e-6)
    imgs, acts = im.match([1, 2], num_epochs=0)
    assert imgs.shape == (2, 20, 1, res, res) and acts.shape == (2, 20) and im.match_epochs_run == 0 and im.match_steps_taken == 0


def test_population_reads_loaded_banks(L):
    """load_state_dict into a simulated population must reach the packed bank the kernels read (it kept the old filters)."""
    res = 20
    a = L.neuron_models.simulated_population(4, img_res=[res, res], seed=1).cuda()
    b = L.neuron_models.simulated_population(4, img_res=[res, res], seed=2).cuda()
    x = torch.randn(5, 1, res, res, device="cuda")
    want = b(x)
    assert not torch.allclose(a(x), want)
    a.load_state_dict(b.state_dict())
    assert torch.equal(a(x), want)


def test_conv_encoder_rebuilds_its_weight_pack(L):
    """The conv encoder's permuted weight pack follows load_state_dict, in-place edits and .to() without a manual refresh_cone()."""
    a = L.neuron_models.conv_core_gaussian_model((1, 24, 24), n_neurons=4, hidden_channels=8, input_kern=5, hidden_kern=3, layers=2, seed=0).cuda()
    b = L.neuron_models.conv_core_gaussian_model((1, 24, 24), n_neurons=4, hidden_channels=8, input_kern=5, hidden_kern=3, layers=2, seed=1).cuda()
    x = torch.randn(3, 1, 24, 24, device="cuda")
    ya, want = a(x), b(x)
    assert not torch.allclose(ya, want)
    a.load_state_dict(b.state_dict())
    assert torch.equal(a(x), want)
    with torch.no_grad():
        a.readout.bias.add_(1.0)
        b.readout.bias.add_(1.0)
    got, want2 = a(x), b(x)
    assert torch.equal(got, want2) and not torch.equal(want2, want)


# ----------------------------------------------------------------------------------------------------------------------------------------
# the BASELINE configs at their own sizes against the UNMODIFIED reference (oracle/make_golden_r2.py)
# ----------------------------------------------------------------------------------------------------------------------------------------
from tests._util import golden, relerr, split_flat  # noqa: E402


def _load_template(im, d):
    sd = {f"func.layer{l}.weight": torch.from_numpy(d[f"W{l}"]) for l in range(5)}
    sd.update({f"func.layer{l}.bias": torch.from_numpy(d[f"b{l}"]) for l in range(5)})
    sd.update(B=torch.from_numpy(d["B"]), auxB=torch.from_numpy(d["auxB"]))
    im.template.load_state_dict(sd, strict=False)


def _demo_meis(res, act=1.0000009536743164):
    yy, xx = np.meshgrid(np.arange(res) + 0.5, np.arange(res) + 0.5, indexing="ij")
    out = {}
    for i, (cx, cy) in enumerate([(0.6 * res, 0.6 * res), (0.4 * res, 0.4 * res), (0.4 * res, 0.6 * res)]):
        mask = np.exp(-0.5 * (((xx - cx) / (0.16 * res)) ** 2 + ((yy - cy) / (0.16 * res)) ** 2)).astype(np.float32)
        out[i] = dict(mei=np.zeros((1, res, res), np.float32), activation=act, mask=mask, center_pos=dict(x=cx, y=cy))
    return out


def _set_affine(t, d, meis, targets):
    D = len(targets)
    t.initialize_coordinate_transform(D, True, False, 0.1, True, True, False, True)
    rf = {k: np.array([v["center_pos"]["x"], v["center_pos"]["y"]], np.float32) for k, v in meis.items()}
    t.register_coords_shifts(torch.from_numpy(rf[0]).cuda(), torch.from_numpy(np.array([rf[k] for k in targets])).cuda())
    aff = t.coordinate_transform.transforms.Affine
    with torch.no_grad():
        aff.angles.copy_(torch.from_numpy(d["angles"])), aff.scalings.copy_(torch.from_numpy(d["scalings"]))
        aff.shears.copy_(torch.from_numpy(d["shears"])), aff.translation.copy_(torch.from_numpy(d["translation"]).reshape(D, 1, 2))


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_match_step_100_vs_reference(L, precision):
    """BASELINE configs[1] at its own size: one match hot step on demo1 targets [1, 2] at 100x100 (tests/golden/match_step_100.npz): loss 5e-5,
    activations 5e-5, images 1e-4, gradients of the 7 affine scalars 5e-3 -- for the fp32 and the tcgen05 split-bf16 path alike."""
    from laminr_b200.fused import MatchStepper
    d = golden("match_step_100")
    res, sub = 100, int(d["sub"])
    model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
    meis = _demo_meis(res)
    torch.manual_seed(0)
    im = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=precision)
    _load_template(im, d)
    targets = [int(k) for k in d["targets"]]
    _set_affine(im.template, d, meis, targets)
    ms = MatchStepper(im.template, im.img_transforms, model, targets, d["mei_acts"], 20, lr=1e-3, precision=precision, use_graph=False)
    before = [p.detach().clone() for p in ms.params]
    ms.step(torch.from_numpy(d["grid"]).cuda())
    torch.cuda.synchronize()
    acts = ms.act.cpu().numpy()
    assert relerr(acts, d["acts_dn"]) < 5e-5
    loss = -(acts.mean(axis=1) / d["mei_acts"]).mean()
    assert abs(loss - float(d["loss"])) < 5e-5 * abs(float(d["loss"]))
    assert relerr(ms.img_pre.cpu().numpy().reshape(2, 20, -1)[:, :, ::sub], d["img_pre"]) < 1e-4
    assert relerr(ms.z.cpu().numpy().reshape(2, 20, -1)[:, :, ::sub], d["img_post"]) < 1e-4
    for got, key in zip(ms.grads, ("d_angles", "d_scalings", "d_shears", "d_translation")):
        assert relerr(got.cpu().numpy().reshape(d[key].shape), d[key]) < 5e-3, key
    for p, p0, g in zip(ms.params, before, ms.grads):  # first Adam step: every scalar moves by lr against its gradient's sign
        np.testing.assert_allclose((p.detach() - p0).cpu().numpy(), -1e-3 * np.sign(g.cpu().numpy()), rtol=0, atol=2e-6)


def test_match_call_100_vs_reference_call(L):
    """BASELINE configs[1]: the whole `match([1, 2], num_epochs=5)` call at 100x100 from seeds alone against the unmodified reference's call
    (tests/golden/match_call_100.npz): chosen sweep angles and scalings exactly / 1e-4, sheared / translated parameters after five Adam steps, final
    images and activations within the north star's 1e-2, CPU generator in the same state."""
    d = golden("match_call_100")
    res = 100
    model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
    meis = {i: dict(mei=np.zeros((1, res, res), np.float32), activation=float(d["mei_acts"][i]), mask=d["masks"][i],
                    center_pos=dict(x=float(d["centers"][i, 0]), y=float(d["centers"][i, 1]))) for i in range(3)}
    torch.manual_seed(0)
    im = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    _load_template(im, d)
    im.template_neuron_idx = 0
    torch.manual_seed(5)
    imgs, acts = im.match([1, 2], num_epochs=5)
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(d["rng_after"]))
    aff = im.template.coordinate_transform.transforms.Affine
    np.testing.assert_allclose(aff.angles.detach().cpu().numpy(), d["angles"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(aff.scalings.detach().cpu().numpy(), d["scalings"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(aff.shears.detach().cpu().numpy(), d["shears"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(aff.translation.detach().cpu().numpy()[:, 0], d["translation"], rtol=0, atol=2e-4)
    assert imgs.shape == d["imgs"].shape and acts.shape == d["acts"].shape
    assert relerr(imgs, d["imgs"].astype(np.float32)) < 1e-2 and relerr(acts, d["acts"]) < 1e-2


def _cosine(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


@pytest.mark.parametrize("precision,grad_tol,cos_min", [("fp32", 5e-3, 0.99998), ("bf16x3", 5e-3, 0.99998), ("bf16x3_fast", 1.5e-1, 0.99)])
def test_learn_step_late_in_the_trajectory_vs_reference(L, precision, grad_tol, cos_min):
    """The learn hot step from the network the reference's own loop reaches after 600 Adam steps at 100x100 (mean activation 0.91;
    tests/golden/learn_late_100.npz).  Near convergence the parameter gradient is a small residual of two competing terms -- SURVEY 7 measured
    rel-L2 2.1 / cosine -0.93 for plain bf16 operands and 1.9e-3 / 1.00000 for the error-compensated split.  Pinned per precision: loss 5e-5 (1e-3
    for the MUFU-tanh variant), images 1e-4 (2e-3), whole-gradient rel-L2 and cosine as parametrised; the split-bf16 tcgen05 path must be as
    close to the fp32 reference as the fp32 CUDA-core path is."""
    from laminr_b200.contrastive_loss import SimCLROnGrid
    from laminr_b200.fused import LearnStepper
    d = golden("learn_late_100")
    res, sub = 100, int(d["sub"])
    model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
    meis = _demo_meis(res)
    np.testing.assert_allclose(meis[0]["mask"], d["mask"], atol=1e-7)
    torch.manual_seed(0)
    im = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=precision)
    _load_template(im, d)
    reg = SimCLROnGrid(1, 20, 0.1, 0.3, True).cuda()
    st = LearnStepper(im.template, im.img_transforms, model, 0, d["mask"], float(d["mei_act"]), reg, -1.0, 1.0, 20, lr=1e-3, reg_scale=2.0,
                      precision=precision, use_graph=False)
    st.step(torch.from_numpy(d["grid"]).cuda())
    torch.cuda.synchronize()
    fast = precision == "bf16x3_fast"
    assert abs(st.loss().item() - float(d["loss"])) < (1e-3 if fast else 5e-5) * abs(float(d["loss"]))
    assert relerr(st.last_acts().cpu().numpy(), d["acts"]) < (1e-3 if fast else 5e-5)
    assert relerr(st.z.cpu().numpy().reshape(20, -1)[:, ::sub], d["img_post"]) < (2e-3 if fast else 1e-4)
    want = np.concatenate([np.concatenate([d[f"dW{l}"].ravel(), d[f"db{l}"].ravel()]) for l in range(5)])
    got = st.g_flat.cpu().numpy()
    print(f"late-trajectory gradient [{precision}]: rel-L2 {relerr(got, want):.2e}, cosine {_cosine(got, want):.6f}")
    assert relerr(got, want) < grad_tol and _cosine(got, want) > cos_min


def _conv_model_from_golden(L, d, prefix="m_"):
    """The product's encoder mirror with the golden's (reference-constructed) arrays loaded."""
    hc, n = d[prefix + "w0"].shape[0], d[prefix + "mu"].shape[0]
    res = int(d["res"])
    model = L.neuron_models.conv_core_gaussian_model((1, res, res), n_neurons=n, hidden_channels=hc, input_kern=d[prefix + "w0"].shape[-1],
                                                     hidden_kern=d[prefix + "w1"].shape[-1], layers=int(d[prefix + "n_layers"]), seed=0)
    t = lambda k: torch.from_numpy(np.asarray(d[prefix + k]))
    with torch.no_grad():
        for l, feat in enumerate(model.core.features):
            feat.conv.weight.copy_(t(f"w{l}"))
            if feat.conv.bias is not None:
                feat.conv.bias.copy_(t(f"b{l}"))
            feat.norm.running_mean.copy_(t(f"bn{l}_mean")), feat.norm.running_var.copy_(t(f"bn{l}_var"))
            feat.norm.weight.copy_(t(f"bn{l}_gamma")), feat.norm.bias.copy_(t(f"bn{l}_beta"))
        model.readout._mu.copy_(t("mu").reshape(model.readout._mu.shape))
        model.readout._features.copy_(t("features").reshape(model.readout._features.shape))
        model.readout.bias.copy_(t("bias"))
    return model.cuda().eval()


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_config4_learn_and_match_step_200_vs_reference(L, precision):
    """BASELINE configs[3] at full size against the unmodified reference (tests/golden/conv_learn_match_200.npz): Stacked2dCore(1->32, 9/7/7) +
    FullGaussian2d(128) at 200x200, learn hot step (loss 1e-4, activations 5e-5, images 1e-4, parameter gradients 5e-3) and match hot step
    (activations 5e-5, images 1e-4, affine gradients 1e-2: at random init they are ~1e-5 in magnitude, differences of nearly cancelling terms)."""
    from laminr_b200.contrastive_loss import SimCLROnGrid
    from laminr_b200.fused import LearnStepper, MatchStepper
    d = golden("conv_learn_match_200")
    res, sub = int(d["res"]), int(d["sub"])
    model = _conv_model_from_golden(L, d)
    # the constructor mirror consumes the generator like the reference's: with seed 0 it IS this model already
    fresh = L.neuron_models.conv_core_gaussian_model((1, res, res), n_neurons=128, seed=0)
    assert torch.equal(fresh.core.features[0].conv.weight, model.core.features[0].conv.weight.cpu())
    assert torch.equal(fresh.readout._mu.clamp(-1, 1), model.readout._mu.cpu())
    yy, xx = np.meshgrid(np.arange(res) + 0.5, np.arange(res) + 0.5, indexing="ij")
    meis = {}
    for i in range(3):
        cx, cy = d["centers"][i]
        mask = np.exp(-0.5 * (((xx - cx) / (0.08 * res)) ** 2 + ((yy - cy) / (0.08 * res)) ** 2)).astype(np.float32)
        meis[i] = dict(mei=np.zeros((1, res, res), np.float32), activation=float(d["mei_acts_all"][i]), mask=mask, center_pos=dict(x=float(cx), y=float(cy)))
    np.testing.assert_allclose(meis[0]["mask"], d["mask"], atol=1e-6)
    torch.manual_seed(0)
    im = L.InvarianceManifold(model, meis, required_img_norm=float(d["norm"]), pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=precision)
    _load_template(im, d)
    reg = SimCLROnGrid(1, 20, 0.1, 0.3, True).cuda()
    st = LearnStepper(im.template, im.img_transforms, model, 0, d["mask"], float(d["mei_act"]), reg, -1.0, 1.0, 20, lr=1e-3, reg_scale=2.0,
                      precision=precision, use_graph=False)
    flat0 = st.flat.clone()
    grid = torch.from_numpy(d["grid"]).cuda()
    st.step(grid)
    torch.cuda.synchronize()
    assert abs(st.loss().item() - float(d["loss"])) < 1e-4 * abs(float(d["loss"]))
    assert relerr(st.last_acts().cpu().numpy(), d["acts"]) < 5e-5
    assert relerr(st.z.cpu().numpy().reshape(20, -1)[:, ::sub], d["img_post"]) < 1e-4
    dW, db = split_flat(st.g_flat.cpu().numpy(), d)
    for l in range(5):
        assert relerr(dW[l], d[f"dW{l}"]) < 5e-3, l
        assert relerr(db[l], d[f"db{l}"]) < 5e-3, l
    st.flat.copy_(flat0)
    targets = [int(k) for k in d["targets"]]
    _set_affine(im.template, d, meis, targets)
    ms = MatchStepper(im.template, im.img_transforms, model, targets, d["mei_acts"], 20, lr=1e-3, precision=precision, use_graph=False)
    ms.step(grid)
    torch.cuda.synchronize()
    assert relerr(ms.act.cpu().numpy(), d["match_acts_dn"]) < 5e-5
    assert relerr(ms.z.cpu().numpy().reshape(2, 20, -1)[:, :, ::sub], d["match_img_post"]) < 1e-4
    for got, key in zip(ms.grads, ("d_angles", "d_scalings", "d_shears", "d_translation")):
        print(key, got.cpu().numpy().ravel(), d[key].ravel())
        assert relerr(got.cpu().numpy().reshape(d[key].shape), d[key]) < 1e-2, key


def test_config3_sample_meis_and_match_vs_reference(L):
    """BASELINE configs[2] / [4] on an 8-neuron sample of the 1024-neuron synthetic population (tests/golden/config3_sample_100.npz): `get_mei_dict`
    of the sample (activations 2e-6, MEI images 3e-3 against the float16 fixture, RF masks to the pixel above 0.5 and 1e-3 in the mask sum, centroids
    0.05 px, generator state) and `match(neuron 0 -> the 7 others, num_epochs=5)` with the reference's meis_dict as the input (parameters, final
    images 1e-2, activations 1e-2, generator state)."""
    d = golden("config3_sample_100")
    sample, res = [int(i) for i in d["sample"]], 100
    n = len(sample)
    pop = L.neuron_models.simulated_population(int(d["n_pop"]), img_res=[res, res], seed=0, only=sample).cuda()
    torch.manual_seed(0)
    meis = L.get_mei_dict(pop, [1, res, res], pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, required_img_norm=1.0)
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(d["rng_after_meis"]))
    assert sorted(meis) == list(range(n))
    for k in range(n):
        assert abs(meis[k]["activation"] - float(d["activations"][k])) < 2e-6, k
        assert relerr(meis[k]["mei"], d["meis"][k].astype(np.float32)) < 3e-3, k
        want_mask = d["masks"][k].astype(np.float32)
        assert ((meis[k]["mask"] > 0.5) == (want_mask > 0.5)).all() and int((meis[k]["mask"] > 0.5).sum()) == int(d["mask_counts"][k]), k
        assert abs(float(np.asarray(meis[k]["mask"], np.float64).sum()) - float(d["mask_sums"][k])) < 1e-3 * float(d["mask_sums"][k]), k
        assert abs(meis[k]["center_pos"]["x"] - d["centers"][k, 0]) < 0.05 and abs(meis[k]["center_pos"]["y"] - d["centers"][k, 1]) < 0.05, k
    # match with the REFERENCE's dict as the shared input fixture (mask sums set the initial scalings, centres the coordinate shifts)
    ref_meis = {k: dict(mei=d["meis"][k].astype(np.float32), activation=float(d["activations"][k]), mask=d["masks"][k].astype(np.float32),
                        center_pos=dict(x=float(d["centers"][k, 0]), y=float(d["centers"][k, 1]))) for k in range(n)}
    torch.manual_seed(0)
    im = L.InvarianceManifold(pop, ref_meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    _load_template(im, d)
    im.template_neuron_idx = 0
    torch.manual_seed(5)
    imgs, acts = im.match(list(range(1, n)), num_epochs=5)
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(d["rng_after"]))
    aff = im.template.coordinate_transform.transforms.Affine
    np.testing.assert_allclose(aff.angles.detach().cpu().numpy(), d["angles"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(aff.scalings.detach().cpu().numpy(), d["scalings"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(aff.shears.detach().cpu().numpy(), d["shears"], rtol=0, atol=5e-4)
    np.testing.assert_allclose(aff.translation.detach().cpu().numpy()[:, 0], d["translation"], rtol=0, atol=5e-4)
    sub = int(d["sub"])
    assert relerr(imgs.reshape(n - 1, 20, -1)[:, :, ::sub], d["imgs"].astype(np.float32)) < 1e-2 and relerr(acts, d["acts"]) < 1e-2


def test_torch_library_operators_opcheck(L):
    """torch.library.opcheck of the registered operators on small inputs: schema, fake kernel vs real kernel, autograd registration."""
    from torch.library import opcheck
    torch.manual_seed(0)
    x = torch.randn(3, 1, 6, 5, device="cuda", requires_grad=True)
    opcheck(torch.ops.laminr_b200.constrain, (x, 0, 1.0, 0.0, True, -1.0, True, 1.0, 1e-12), test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    pop = L.neuron_models.simulated_population(4, img_res=[5, 6], seed=1, num_filters=3).cuda()
    idx = torch.tensor([2, 0], dtype=torch.int32, device="cuda")
    opcheck(torch.ops.laminr_b200.simresp, (x, pop.packed_filters, pop.nfilt, idx, 1e-6), test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    # gradients through the operators == the raw C-ABI calls
    act, rmax, argmax = torch.ops.laminr_b200.simresp(x, pop.packed_filters, pop.nfilt, idx, 1e-6)
    g = torch.randn_like(act)
    (gx,) = torch.autograd.grad(act, x, g)
    from laminr_b200 import ops
    want = torch.empty_like(x)
    ops.raw_simresp_bwd(g.contiguous(), rmax, argmax, 0, 3, 2, 1, 30, pop.packed_filters, idx, want)
    assert torch.equal(gx, want)


@pytest.mark.parametrize("widths,precision", [([32, 20, 50], "fp32"), ([24] * 4, "fp32"), ([40, 16, 50, 33], "bf16x3"), ([50, 50, 8], "fp32")])
def test_narrow_widths_forward_backward_vs_oracle(L, widths, precision):
    """INR networks narrower than the kernels' 50 (the reference's `widths=` argument) run zero-padded on the same kernels: images and parameter
    gradients against the numpy oracle evaluated on the TRUE shapes (fp32: 1e-5 / 2e-3; split-bf16: 1e-4 / 5e-3), gradients arrive in the
    parameters' own shapes, and three fused Adam steps leave the padding exactly zero."""
    from torch import nn
    from laminr_b200.inr import INRTemplates
    from oracle import laminr_oracle as O
    H, W, N = 9, 11, 20
    torch.manual_seed(1)
    t = INRTemplates((H, W), widths=widths, periodic_invariance=True, positional_encoding_dim=50, positional_encoding_projection_scale=10,
                     aux_positional_encoding_dim=50, aux_positional_encoding_projection_scale=0.1, final_nonlinearity=nn.Tanh, bias=True, weights_scale=0.1,
                     precision=precision).cuda()
    with torch.no_grad():
        for p in t.func.parameters():
            p.add_(torch.randn_like(p) * 0.1)
    grid = (torch.linspace(0, 2 * np.pi, N + 1)[:-1] + 0.3).reshape(N, 1).cuda()
    img = t(aux=grid, return_template=True)
    g = torch.randn_like(img)
    (img * g).sum().backward()
    Ws = [getattr(t.func, f"layer{i}").weight.detach().cpu().numpy() for i in range(len(widths) + 1)]
    bs = [getattr(t.func, f"layer{i}").bias.detach().cpu().numpy() for i in range(len(widths) + 1)]
    coords = t.coords.reshape(-1, 2).cpu().numpy()
    want, cache = O.inr_forward(Ws, bs, t.B.cpu().numpy(), t.auxB.cpu().numpy(), grid[:, 0].cpu().numpy(), coords[None], (H, W), keep=True)
    tol_i, tol_g = (1e-5, 2e-3) if precision == "fp32" else (1e-4, 5e-3)
    assert relerr(img.detach().cpu().numpy().reshape(want.shape), want) < tol_i
    grads = O.inr_backward(Ws, t.B.cpu().numpy(), cache, g.detach().cpu().numpy().reshape(want.shape), want_weights=True)
    dW, db = grads["dW"], grads["db"]
    for i in range(len(widths) + 1):
        lay = getattr(t.func, f"layer{i}")
        assert lay.weight.grad.shape == lay.weight.shape and relerr(lay.weight.grad.cpu().numpy(), dW[i]) < tol_g, i
        assert relerr(lay.bias.grad.cpu().numpy(), db[i]) < tol_g, i
    # the padding of the flat vector stays zero under the fused Adam step
    flat = t.flat_params()
    live = torch.zeros_like(flat, dtype=torch.bool)
    for p in t.func.parameters():
        live.as_strided(p.shape, p.stride(), p.storage_offset() - flat.storage_offset()).fill_(True)
    assert float(flat[~live].abs().max()) == 0.0
    from laminr_b200 import ops
    m, v, step, gp = torch.zeros_like(flat), torch.zeros_like(flat), torch.zeros(2, dtype=torch.int32, device="cuda"), torch.empty_like(flat)
    cflat = t.coords.reshape(-1, 2).contiguous()
    for _ in range(3):
        ws = torch.empty(ops.forward_workspace_bytes(t._desc, N, H * W, 1, precision != "fp32"), dtype=torch.uint8, device="cuda")
        ops.raw_inr_forward(t._desc, flat, t.B, t.auxB, grid.reshape(-1), cflat, None, None, precision=precision, workspace=ws, keep_activations=True)
        ops.raw_inr_backward_adam(t._desc, flat, t.B, t.auxB, grid.reshape(-1), cflat, None, g.reshape(1, N, 1, H * W).contiguous(), gp, m, v, step,
                                  precision=precision, forward_workspace=ws, keep_activations=True)
    assert float(flat[~live].abs().max()) == 0.0 and float(gp[~live].abs().max()) == 0.0 and float(m[~live].abs().max()) == 0.0


@pytest.mark.parametrize("mode_name,shape", [("norm", (128, 1, 200, 200)), ("std", (5, 3, 50, 37)), ("norm", (3, 1, 17, 9)), ("std", (40, 1, 100, 100))])
def test_fused_ascent_update_is_bit_identical_to_the_two_launch_path(L, mode_name, shape, monkeypatch):
    """lmr_ascent_update: one CTA per image with the image in shared memory (default when it fits) against the statistics + apply launches
    (LMR_ASCENT_FUSED=0): same chunk partials, same reduction trees, same combine -> identical bits; and against a float64 evaluation."""
    from laminr_b200 import _lib, ops
    torch.manual_seed(3)
    x0 = torch.randn(shape, device="cuda") * 0.3
    g = torch.randn(shape, device="cuda") * 1e-3
    g[:, :, : shape[2] // 2] = 0.0  # (a cone's gradient is zero outside its window)
    mode = _lib.MODE_NORM if mode_name == "norm" else _lib.MODE_STD
    outs = []
    for fused in ("1", "0"):
        monkeypatch.setenv("LMR_ASCENT_FUSED", fused)
        x = x0.clone()
        ops.raw_ascent_update(mode, x, g, 10.0, 7.5, 0.0, -1.0, 1.0, eps=1e-9)
        outs.append(x)
    assert torch.equal(outs[0], outs[1])
    v = (x0.double() + 10.0 * g.double()).reshape(shape[0], -1)
    if mode_name == "norm":
        want = v / (v.norm(dim=1, keepdim=True) + 1e-9) * 7.5
    else:
        want = 7.5 * (v - v.mean(dim=1, keepdim=True)) / (v.std(dim=1, keepdim=True) + 1e-9)
    want = want.clamp(-1.0, 1.0).reshape(shape)
    assert (outs[0].double() - want).abs().max().item() < 2e-6
