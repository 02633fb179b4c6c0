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
