"""Macaque-V1 response models: squeeze-excitation core with depth-separable dilated convolutions + point-pooled / Gaussian
readout, keyless 458-neuron wrappers, ensembles.  Goldens come from the reference's own classes constructed offline (oracle/make_golden_monkey.py).

CPU tier: the oracle restatement and the host-side mirror (construction, RNG draw order, state_dict names) against the goldens.
GPU tier: the full-frame kernels (csrc/fullframe.cu) through the mirrored modules against the same goldens, plus oracle comparisons on other inputs.
"""
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import monkey_oracle as MO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL_ACT, TOL_GRAD = 2e-5, 1e-4  # fp32 sums of 576 (first layer) / 81 / 32 terms in a different order from torch's CPU convolution


def golden(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False))


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _mirror(tag, d=None, seed=None, pseed=None):
    from laminr_b200.neuron_models import monkey_models as mm
    get = mm.get_pointpool_monkey_model if tag == "pointpool" else mm.get_gaussian_monkey_model
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = get(seed=int(d["seed"]) if seed is None else seed)
    return MO.perturb_state(m, int(d["perturb"]) if pseed is None else pseed)


def _keyless(tag, m, seed=1):
    from laminr_b200.neuron_models import monkey_models as mm
    torch.manual_seed(seed)
    return (mm.keyless_pointpool_monkey_model if tag == "pointpool" else mm.keyless_gaussian_monkey_model)(m)


def _oracle_resp_grad(fn, x, gw):
    x = torch.as_tensor(x).clone().requires_grad_(True)
    y = fn(x)
    (gx,) = torch.autograd.grad((y * torch.as_tensor(gw)).sum(), x)
    return y.detach().numpy(), gx.numpy()


# ----------------------------------------------------------------------------------------------------------------------
# CPU tier
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["pointpool", "gaussian"])
def test_mirror_reproduces_reference_state(tag):
    """Same seed -> the reference's initial weights, bit for bit, under the reference's state_dict names (the fixture stores the reference's)."""
    d = golden("monkey_" + tag)
    want = MO.sd_from_numpy(d, "sd/")
    got = _mirror(tag, d).state_dict()
    assert list(got.keys()) == list(want.keys())
    for k in want:
        assert torch.equal(got[k], want[k]), k
    if tag == "gaussian":  # the merged readout keeps its own construction-time positions (reference quirk, see keyless_gaussian_monkey_model._merge)
        km = _keyless(tag, _mirror(tag, d))
        assert np.array_equal(km.readout._mu.detach().numpy(), d["merged_mu"])
        assert km.readout._features.shape == (1, 32, 1, 458)


@pytest.mark.parametrize("tag", ["pointpool", "gaussian"])
def test_oracle_vs_reference_golden(tag):
    d = golden("monkey_" + tag)
    sd = MO.sd_from_numpy(d, "sd/")
    if tag == "pointpool":
        fn = lambda x: MO.keyless_pointpool(sd, x)
    else:
        fn = lambda x: MO.keyless_gaussian(sd, x, torch.as_tensor(d["merged_mu"]))
    y, gx = _oracle_resp_grad(fn, d["x"], d["gw"])
    assert y.shape == (3, 458)
    assert relerr(y, d["y"]) < 1e-5 and relerr(gx, d["gx"]) < 1e-5
    if tag == "pointpool":
        with torch.no_grad():
            ys = MO.session_pointpool(sd, torch.as_tensor(d["x"]), str(d["session"])).numpy()
        assert relerr(ys, d["y_session"]) < 1e-5


def _ensemble_members(tag, tmp_path, seeds):
    """The three checkpoints the golden script wrote, rebuilt with the mirror (same seeds, same perturbation)."""
    paths = []
    for i, (seed, pseed) in enumerate(seeds):
        m = _mirror(tag, seed=int(seed), pseed=int(pseed))
        paths.append(str(tmp_path / f"{tag}_{i}.pt"))
        torch.save(m.state_dict(), paths[-1])
    return paths


def _build_ensemble(tag, paths):
    from laminr_b200.neuron_models import monkey_models as mm
    np.random.seed(0)
    torch.manual_seed(2)
    return mm.monkey_v1(tag, model_paths=paths)


@pytest.mark.parametrize("tag", ["pointpool", "gaussian"])
def test_ensemble_oracle_vs_reference_golden(tag, tmp_path):
    d = golden("monkey_ensemble")
    e = _build_ensemble(tag, _ensemble_members(tag, tmp_path, d["seeds"]))
    assert len(e.models) == 3 and not e.training

    def fn(x):
        ys = []
        for km in e.models:
            # members' parameters through the oracle: core from the member, merged readout from the keyless wrapper
            sd = {"core." + k: v for k, v in km.core.state_dict().items()}
            if tag == "pointpool":
                y = MO.point_readout(MO.core_forward(sd, x), km.readout.grid.detach(), km.readout.features.detach(), km.readout.bias.detach())
            else:
                y = MO.point_readout(MO.core_forward(sd, x), km.readout._mu.detach(), km.readout._features.detach(), km.readout.bias.detach())
            ys.append(torch.nn.functional.elu(y) + 1)
        return torch.stack(ys).mean(0)

    y, gx = _oracle_resp_grad(fn, d["x"], d["gw"])
    assert relerr(y, d[f"y_{tag}"]) < 1e-5 and relerr(gx, d[f"gx_{tag}"]) < 1e-5


def test_monkey_v1_download_is_the_only_thing_missing():
    from laminr_b200.neuron_models import monkey_models as mm, monkey_v1
    with pytest.raises(NotImplementedError, match="Hugging Face"):
        monkey_v1()
    with pytest.raises(ValueError):
        monkey_v1("other")
    assert mm.N_NEURONS == 458 and len(mm.data_info) == 32
    e = monkey_v1("pointpool", pretrained=False)
    assert len(e.models) == 3 and e.models[0].readout.features.shape == (1, 96, 1, 458)
    with pytest.raises(RuntimeError):  # CPU tensors are refused: no stock-PyTorch path
        e(torch.zeros(1, 1, 93, 93))
    with pytest.raises(NotImplementedError):
        e.models[0].core(torch.zeros(1, 1, 93, 93))


# ----------------------------------------------------------------------------------------------------------------------
# GPU tier
# ----------------------------------------------------------------------------------------------------------------------
def _gpu_resp_grad(model, x, gw):
    x = torch.as_tensor(x).cuda().requires_grad_(True)
    y = model(x)
    (gx,) = torch.autograd.grad((y * torch.as_tensor(gw).cuda()).sum(), x)
    return y.detach().cpu().numpy(), gx.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["pointpool", "gaussian"])
def test_keyless_model_vs_reference_golden(tag):
    d = golden("monkey_" + tag)
    m = _mirror(tag, d)
    km = _keyless(tag, m).cuda()
    y, gx = _gpu_resp_grad(km, d["x"], d["gw"])
    assert relerr(y, d["y"]) < TOL_ACT, relerr(y, d["y"])
    assert relerr(gx, d["gx"]) < TOL_GRAD, relerr(gx, d["gx"])
    if tag == "pointpool":  # the per-session model the checkpoints are state_dicts of
        with torch.no_grad():
            ys = m.cuda().eval()(torch.as_tensor(d["x"]).cuda(), data_key=str(d["session"]))
        assert relerr(ys.cpu().numpy(), d["y_session"]) < TOL_ACT


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["pointpool", "gaussian"])
def test_ensemble_vs_reference_golden(tag, tmp_path):
    d = golden("monkey_ensemble")
    e = _build_ensemble(tag, _ensemble_members(tag, tmp_path, d["seeds"])).cuda()
    y, gx = _gpu_resp_grad(e, d["x"], d["gw"])
    assert relerr(y, d[f"y_{tag}"]) < TOL_ACT
    assert relerr(gx, d[f"gx_{tag}"]) < TOL_GRAD


@pytest.mark.gpu
def test_single_neuron_gradient_and_batch_sizes_vs_oracle():
    """What the LAMINR loops do with the model: one neuron's response, gradient back to a batch of 20 images; odd batch sizes; repeated backward."""
    d = golden("monkey_pointpool")
    sd = MO.sd_from_numpy(d, "sd/")
    km = _keyless("pointpool", _mirror("pointpool", d)).cuda()
    g = torch.Generator().manual_seed(5)
    for B, n in ((20, 17), (1, 457), (7, 0)):
        x = torch.randn(B, 1, 93, 93, generator=g)
        gw = torch.zeros(B, 458)
        gw[:, n] = torch.randn(B, generator=g)
        want_y, want_g = _oracle_resp_grad(lambda t: MO.keyless_pointpool(sd, t), x.numpy(), gw.numpy())
        y, gx = _gpu_resp_grad(km, x.numpy(), gw.numpy())
        assert relerr(y, want_y) < TOL_ACT and relerr(gx, want_g) < TOL_GRAD
    # two graphs alive at once: the second forward overwrites the kept layer outputs, the first backward must notice
    xa, xb = (torch.randn(2, 1, 93, 93, generator=g).cuda().requires_grad_(True) for _ in range(2))
    ya, yb = km(xa), km(xb)
    ya[:, 3].sum().backward()
    yb[:, 3].sum().backward()
    for xt in (xa, xb):
        gwn = np.zeros((2, 458), np.float32)
        gwn[:, 3] = 1
        _, want_g = _oracle_resp_grad(lambda t: MO.keyless_pointpool(sd, t), xt.detach().cpu().numpy(), gwn)
        assert relerr(xt.grad.cpu().numpy(), want_g) < TOL_GRAD


@pytest.mark.gpu
def test_state_changes_rebuild_the_pack():
    d = golden("monkey_gaussian")
    km = _keyless("gaussian", _mirror("gaussian", d)).cuda()
    x = torch.as_tensor(d["x"]).cuda()
    y0 = km(x)
    with torch.no_grad():
        km.readout.bias.add_(0.5)
    y1 = km(x)
    assert (y1 - y0).abs().max() > 0.05
    km.train()
    with pytest.raises(NotImplementedError):
        km(x)


class _OracleResponse(torch.nn.Module):
    """The CPU oracle behind the nn.Module interface the LAMINR loops call (device tensors in and out, autograd through both copies)."""

    def __init__(self, sd):
        super().__init__()
        self.sd, self.anchor = sd, torch.nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return MO.keyless_pointpool(self.sd, x.cpu()).to(x.device)


@pytest.mark.gpu
def test_laminr_loops_on_a_monkey_model_vs_oracle_model():
    """get_mei_dict -> InvarianceManifold.learn -> match (generic nn.Module path) on the 93 x 93 point-pooled model: the full-frame kernels and the
    CPU oracle as the response model give the same MEIs, template images and activations."""
    import contextlib
    import io
    from laminr_b200 import InvarianceManifold, get_mei_dict
    d = golden("monkey_pointpool")
    km = _keyless("pointpool", _mirror("pointpool", d)).cuda()
    om = _OracleResponse(MO.sd_from_numpy(d, "sd/")).cuda()
    out = {}
    for name, model in (("kernels", km), ("oracle", om)):
        torch.manual_seed(0)
        np.random.seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            meis = get_mei_dict(model, [1, 93, 93], -2.0, 2.0, neuron_idxs=[5, 40], required_img_norm=12.0, num_iterations=15)
            im = InvarianceManifold(model, meis, required_img_norm=12.0, pixel_value_lower_bound=-2.0, pixel_value_upper_bound=2.0)
            imgs, acts = im.learn(0, steps_per_epoch=3, num_max_epochs=1, min_epochs=1, grid_points_per_dim=6)
            m_imgs, m_acts = im.match([1], num_epochs=1, steps_per_epoch=2, grid_points_per_dim=6)
        out[name] = (meis, imgs, acts, m_imgs, m_acts)
    a, b = out["kernels"], out["oracle"]
    for i in (0, 1):
        assert relerr(a[0][i]["mei"], b[0][i]["mei"]) < 1e-3
        assert abs(a[0][i]["activation"] - b[0][i]["activation"]) < 1e-3 * abs(b[0][i]["activation"])
        assert np.array_equal(a[0][i]["mask"], b[0][i]["mask"])
    assert relerr(a[1], b[1]) < 1e-3 and relerr(a[2], b[2]) < 1e-3
    assert relerr(a[3], b[3]) < 1e-3 and relerr(a[4], b[4]) < 1e-3
