"""GPU parity of the stacked-conv core + Gaussian readout response kernel (lmr_convcore_response) through the C ABI:
against the golden vectors produced by the reference's vendored neuralpredictors modules and against the numpy oracle."""
import warnings

import numpy as np
import pytest
import torch

from tests._util import golden, relerr

pytestmark = pytest.mark.gpu

TOL_ACT = 2e-5   # float32 kernel vs float32 torch reference (different summation order)
TOL_GRAD = 1e-4


def _cone_from_golden(d):
    from laminr_b200 import ops
    dev = "cuda"
    L = int(d["n_layers"])
    layers = []
    for l in range(L):
        t = lambda k: torch.as_tensor(d[k]).to(dev)
        bn = (t(f"bn{l}_mean"), t(f"bn{l}_var"), t(f"bn{l}_gamma"), t(f"bn{l}_beta"), float(d[f"bn{l}_eps"])) if f"bn{l}_mean" in d else None
        layers.append(dict(weight=t(f"w{l}"), conv_bias=t(f"b{l}") if f"b{l}" in d else None, bn=bn, nonlin=bool(d[f"nonlin{l}"])))
    H, W = d["x"].shape[2:]
    return ops.ConvCone(layers, torch.as_tensor(d["mu"]).to(dev), torch.as_tensor(d["features"]).to(dev), torch.as_tensor(d["bias"]).to(dev), H, W,
                        xshift=float(d["xs"]), yshift=float(d["ys"]), offset=float(d["offset"]), align_corners=bool(d["align_corners"]))


@pytest.mark.parametrize("name", ["convcore_a", "convcore_b", "convcore_c"])
def test_cone_response_and_gradient_vs_reference(name):
    from laminr_b200 import ops
    d = golden(name)
    cone = _cone_from_golden(d)
    x = torch.as_tensor(d["x"]).cuda()
    B, C, H, W = x.shape
    nog = torch.as_tensor(d["neuron_of_image"], dtype=torch.int32).cuda()
    g_act = torch.as_tensor(d["g_act"]).cuda()
    # one group per image (group stride = one image), each with its own neuron: the match-stage layout
    g_img = torch.full_like(x, 7.0)  # overwritten everywhere (accumulate = 0), zero outside the cone
    act = ops.raw_convcore_response(cone, x, C * H * W, 1, B, nog, g_act=g_act.view(B, 1), g_img=g_img, accumulate=False)
    want = d["acts"][np.arange(B), d["neuron_of_image"]]
    assert relerr(act.cpu().numpy().ravel(), want) < TOL_ACT
    assert relerr(g_img.cpu().numpy(), d["g_x"]) < TOL_GRAD
    assert np.array_equal(g_img.cpu().numpy() == 0, d["g_x"] == 0) or relerr(g_img.cpu().numpy(), d["g_x"]) < TOL_GRAD
    # accumulate adds on top
    base = torch.randn_like(x)
    acc = base.clone()
    ops.raw_convcore_response(cone, x, C * H * W, 1, B, nog, g_act=g_act.view(B, 1), g_img=acc, accumulate=True)
    assert relerr((acc - base).cpu().numpy(), d["g_x"]) < 2 * TOL_GRAD
    # every neuron on every image (stride 0, forward only) == the reference's full model output
    all_n = torch.arange(d["acts"].shape[1], dtype=torch.int32, device="cuda")
    act_all = ops.raw_convcore_response(cone, x, 0, B, all_n.numel(), all_n)
    assert relerr(act_all.t().cpu().numpy(), d["acts"]) < TOL_ACT
    # one group, NB images, one neuron (the learn-stage layout); g_act = None means 1
    n0 = int(d["neuron_of_image"][0])
    g1 = torch.empty_like(x)
    a1 = ops.raw_convcore_response(cone, x, 0, B, 1, torch.tensor([n0], dtype=torch.int32, device="cuda"), g_img=g1, accumulate=False)
    assert relerr(a1.cpu().numpy().ravel(), d["acts"][:, n0]) < TOL_ACT
    from oracle import convcore_oracle as CO
    _, g_want, _ = CO.response_and_grad(d["x"], CO.model_from_arrays(d), [n0] * B)
    assert relerr(g1.cpu().numpy(), g_want) < TOL_GRAD


@pytest.mark.parametrize("cin,hidden,kernels", [(3, 6, (11, 3)), (1, 5, (5, 13)), (4, 8, (1, 7, 3)), (2, 4, (9, 9)), (4, 32, (9, 9, 9))])
def test_cone_odd_shapes_vs_oracle(cin, hidden, kernels):
    """Shapes the reference goldens do not hold: kernel sizes without a compile-time instantiation (1, 11, 13: the sliding-window loop),
    channel counts that run 2- and 1-channel blocks (unpadded, bounds-checked gradient maps) and a 4-channel image (padded first gradient
    map), and a cone too large for padded gradient maps (4 -> 32 channels, 9 / 9 / 9: 238 KB padded, falls back to plain rows), against the
    full-frame numpy oracle on random parameters."""
    from laminr_b200 import ops
    from oracle import convcore_oracle as CO
    rs = np.random.RandomState(cin * 100 + hidden)
    H, W, B, n = 31, 37, 6, 5
    d = dict(n_layers=len(kernels), xs=0.0, ys=0.0, offset=0.1, align_corners=True)
    for l, k in enumerate(kernels):
        ci = cin if l == 0 else hidden
        d[f"w{l}"] = (rs.randn(hidden, ci, k, k) / np.sqrt(ci * k * k)).astype(np.float32)
        d[f"bn{l}_mean"], d[f"bn{l}_var"] = (0.1 * rs.randn(hidden)).astype(np.float32), rs.uniform(0.5, 1.5, hidden).astype(np.float32)
        d[f"bn{l}_gamma"], d[f"bn{l}_beta"], d[f"bn{l}_eps"] = rs.uniform(0.5, 1.5, hidden).astype(np.float32), (0.1 * rs.randn(hidden)).astype(np.float32), 1e-5
        d[f"nonlin{l}"] = True
    d["mu"] = rs.uniform(-0.95, 0.95, (n, 2)).astype(np.float32)
    d["mu"][0] = (-0.999, 0.999)  # a cone clipped by two borders
    d["features"] = rs.randn(hidden, n).astype(np.float32)
    d["bias"] = (0.1 * rs.randn(n)).astype(np.float32)
    x = rs.randn(B, cin, H, W).astype(np.float32)
    neurons = rs.randint(0, n, B)
    g_act = rs.randn(B).astype(np.float32)
    d["x"] = x
    cone = _cone_from_golden(d)
    xg = torch.as_tensor(x).cuda()
    g_img = torch.full_like(xg, 3.0)
    act = ops.raw_convcore_response(cone, xg, cin * H * W, 1, B, torch.as_tensor(neurons, dtype=torch.int32).cuda(),
                                    g_act=torch.as_tensor(g_act).cuda().view(B, 1), g_img=g_img, accumulate=False)
    a_want, g_want, _ = CO.response_and_grad(x, CO.model_from_arrays(d), list(neurons), g_act=g_act)
    assert relerr(act.cpu().numpy().ravel(), a_want) < TOL_ACT
    assert relerr(g_img.cpu().numpy(), g_want) < TOL_GRAD
    # the learn-stage layout (one neuron, every image: a cluster of CTAs per image)
    g1 = torch.empty_like(xg)
    a1 = ops.raw_convcore_response(cone, xg, 0, B, 1, torch.tensor([1], dtype=torch.int32, device="cuda"), g_img=g1, accumulate=False)
    a_want, g_want, _ = CO.response_and_grad(x, CO.model_from_arrays(d), [1] * B)
    assert relerr(a1.cpu().numpy().ravel(), a_want) < TOL_ACT
    assert relerr(g1.cpu().numpy(), g_want) < TOL_GRAD


def test_cone_window_only_gradient_write():
    """accumulate = 2: the kernel overwrites the receptive-field window and leaves the rest of the gradient image alone (the MEI ascent keeps it
    at zero); inside the window the values are those of the whole-image write."""
    from laminr_b200 import ops
    d = golden("convcore_a")
    cone = _cone_from_golden(d)
    x = torch.as_tensor(d["x"]).cuda()
    B, C, H, W = x.shape
    nog = torch.as_tensor(d["neuron_of_image"], dtype=torch.int32).cuda()
    full = torch.empty_like(x)
    a0 = ops.raw_convcore_response(cone, x, C * H * W, 1, B, nog, g_img=full, accumulate=0)
    win = torch.full_like(x, 5.0)
    a2 = ops.raw_convcore_response(cone, x, C * H * W, 1, B, nog, g_img=win, accumulate=2)
    assert torch.equal(a0, a2)
    touched = win != 5.0
    assert touched.any() and not touched.all()
    assert torch.equal(win[touched], full[touched])
    assert torch.all(full[~touched] == 0)


def _load_mirror(d):
    """The host mirror (laminr_b200.neuron_models) with the golden fixture's parameters loaded through state_dict names."""
    from laminr_b200.neuron_models import FiringRateEncoder, FullGaussian2d, Stacked2dCore
    L, w0 = int(d["n_layers"]), d["w0"]
    hid_k = d["w1"].shape[-1] if L > 1 else 3
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        core = Stacked2dCore(input_channels=w0.shape[1], hidden_channels=w0.shape[0], input_kern=w0.shape[-1], hidden_kern=hid_k, layers=L, stack=-1,
                             pad_input=True, final_nonlinearity=bool(d[f"nonlin{L - 1}"]), elu_shift=(float(d["xs"]), float(d["ys"])))
        H, W = d["x"].shape[2:]
        ro = FullGaussian2d(in_shape=(w0.shape[0], H, W), outdims=d["mu"].shape[0], bias=True, align_corners=bool(d["align_corners"]))
    sd = {}
    for l in range(L):
        sd[f"core.features.layer{l}.conv.weight"] = torch.as_tensor(d[f"w{l}"])
        if f"b{l}" in d:
            sd[f"core.features.layer{l}.conv.bias"] = torch.as_tensor(d[f"b{l}"])
        sd[f"core.features.layer{l}.norm.running_mean"] = torch.as_tensor(d[f"bn{l}_mean"])
        sd[f"core.features.layer{l}.norm.running_var"] = torch.as_tensor(d[f"bn{l}_var"])
        sd[f"core.features.layer{l}.norm.weight"] = torch.as_tensor(d[f"bn{l}_gamma"])
        sd[f"core.features.layer{l}.norm.bias"] = torch.as_tensor(d[f"bn{l}_beta"])
    sd["readout._mu"] = torch.as_tensor(d["mu"]).reshape(1, -1, 1, 2)
    sd["readout._features"] = torch.as_tensor(d["features"]).reshape(1, w0.shape[0], 1, -1)
    sd["readout.bias"] = torch.as_tensor(d["bias"])
    model = FiringRateEncoder(core, ro, elu_offset=float(d["offset"]))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith(("num_batches_tracked", "sigma", "laplace.filter")) for k in missing), (missing, unexpected)
    return model.cuda().eval()


@pytest.mark.parametrize("name", ["convcore_a", "convcore_b"])
def test_module_forward_and_autograd(name):
    """The drop-in modules: model(x) for all neurons and autograd through SingleNeuronModel-style selection."""
    d = golden(name)
    model = _load_mirror(d)
    x = torch.as_tensor(d["x"]).cuda().requires_grad_(True)
    acts = model(x)
    assert relerr(acts.detach().cpu().numpy(), d["acts"]) < TOL_ACT
    B = x.shape[0]
    sel = torch.as_tensor(d["neuron_of_image"]).cuda()
    (acts[torch.arange(B, device="cuda"), sel] * torch.as_tensor(d["g_act"]).cuda()).sum().backward()
    assert relerr(x.grad.cpu().numpy(), d["g_x"]) < TOL_GRAD
    with pytest.raises(NotImplementedError):
        model.train()
        model.refresh_cone()
        model(x)
    with pytest.raises(RuntimeError):
        model.eval()
        model.refresh_cone()
        model(x.detach().cpu())


def test_config4_full_size_against_oracle():
    """BASELINE config 4 shapes (1 -> 32 channels, kernels 9/7/7, 200x200, 128 neurons): cone kernel vs the full-frame numpy oracle."""
    from laminr_b200 import ops
    from laminr_b200.neuron_models import conv_core_gaussian_model
    from oracle import convcore_oracle as CO
    model = conv_core_gaussian_model((1, 200, 200), n_neurons=128, seed=0)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():  # spread the readout positions over the frame (init is +-0.1) and make batch norm non-trivial
        model.readout._mu.copy_(torch.rand(1, 128, 1, 2, generator=g) * 2 - 1)
        for feat in model.core.features:
            feat.norm.running_mean.copy_(torch.randn(32, generator=g) * 0.1)
            feat.norm.running_var.copy_(torch.rand(32, generator=g) + 0.5)
    model = model.cuda().eval()
    cone = model.cone()
    assert cone.cone_edge == 22
    idx = [0, 5, 127]
    x = torch.randn(3, 1, 200, 200, generator=g) * 0.1
    xd = x.cuda()
    nog = torch.tensor(idx, dtype=torch.int32, device="cuda")
    g_img = torch.empty_like(xd)
    act = ops.raw_convcore_response(cone, xd, 200 * 200, 1, 3, nog, g_img=g_img, accumulate=False)
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    layers = []
    for l in range(3):
        p = f"core.features.layer{l}."
        layers.append(dict(w=sd[p + "conv.weight"], b=sd.get(p + "conv.bias"), nonlin=True, xs=0.0, ys=0.0,
                           bn=(sd[p + "norm.running_mean"], sd[p + "norm.running_var"], sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)))
    om = dict(layers=layers, mu=sd["readout._mu"].reshape(-1, 2), features=sd["readout._features"].reshape(32, 128), bias=sd["readout.bias"], offset=0.0,
              align_corners=True)
    want_act, want_g, _ = CO.response_and_grad(x.numpy(), om, idx)
    assert relerr(act.cpu().numpy().ravel(), want_act) < TOL_ACT
    assert relerr(g_img.cpu().numpy(), want_g) < TOL_GRAD
    nz = (g_img != 0).sum(dim=(1, 2, 3)).cpu().numpy()
    assert (nz <= 22 * 22).all() and (nz > 0).all()


# ----------------------------------------------------------------------------------------------------------------------
# the learn / match hot steps and the MEI ascent driven by the conv-core encoder
# ----------------------------------------------------------------------------------------------------------------------
def _mirror_from_prefixed(d, prefix="m_"):
    sub = {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}
    sub["x"] = np.zeros((1, sub["w0"].shape[1], 24, 24), np.float32)  # only the frame size is read from it
    return _load_mirror(sub)


def _conv_manifold(d, precision):
    import laminr_b200 as L
    model = _mirror_from_prefixed(d)
    res = 24
    yy, xx = np.meshgrid(np.arange(res) + 0.5, np.arange(res) + 0.5, indexing="ij")
    meis = {}
    for i, (cx, cy) in enumerate([(0.6 * res, 0.6 * res), (0.4 * res, 0.4 * res), (0.4 * res, 0.6 * res)]):
        mask = np.exp(-0.5 * (((xx - cx) / (0.16 * res)) ** 2 + ((yy - cy) / (0.16 * res)) ** 2)).astype(np.float32)
        meis[i] = dict(mei=np.zeros((1, res, res), np.float32), activation=1.25 + 0.1 * i, mask=mask, center_pos=dict(x=cx, y=cy))
    np.testing.assert_allclose(meis[0]["mask"], d["mask"], atol=1e-7)
    torch.manual_seed(0)
    im = L.InvarianceManifold(model, meis, required_img_norm=float(d["norm"]), pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=precision)
    sd = {f"func.layer{l}.weight": torch.from_numpy(d[f"W{l}"]) for l in range(5)}
    sd.update({f"func.layer{l}.bias": torch.from_numpy(d[f"b{l}"]) for l in range(5)})
    sd.update(B=torch.from_numpy(d["B"]), auxB=torch.from_numpy(d["auxB"]))
    im.template.load_state_dict(sd, strict=False)
    return model, meis, im


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_conv_fused_learn_and_match_step_vs_reference(precision, use_graph):
    from laminr_b200.contrastive_loss import SimCLROnGrid
    from laminr_b200.fused import LearnStepper, MatchStepper
    d = golden("conv_learn_match_24")
    model, meis, im = _conv_manifold(d, precision)
    assert im._fused_ok()
    reg = SimCLROnGrid(1, 20, 0.1, 0.3, True).cuda()
    st = LearnStepper(im.template, im.img_transforms, model, 0, d["mask"], float(d["mei_act"]), reg, -1.0, 1.0, 20, lr=1e-3, reg_scale=2.0,
                      precision=precision, use_graph=use_graph)
    flat0 = st.flat.clone()
    grid = torch.from_numpy(d["grid"]).cuda()
    for rep in range(3 if use_graph else 1):  # eager step, capture, replay: restore the parameters in between so that every step is the same step
        st.flat.copy_(flat0)
        st.m.zero_(), st.v.zero_(), st.step_count.zero_()
        st.step(grid)
    torch.cuda.synchronize()
    assert abs(st.loss().item() - d["loss"]) < 1e-4 * abs(d["loss"])
    assert relerr(st.last_acts().cpu().numpy(), d["acts"]) < 5e-5
    assert relerr(st.z.cpu().numpy().reshape(20, -1), d["img_post"]) < 1e-4
    from tests._util import split_flat
    dW, db = split_flat(st.g_flat.cpu().numpy(), d)
    for l in range(5):
        assert relerr(dW[l], d[f"dW{l}"]) < 5e-3, l
        assert relerr(db[l], d[f"db{l}"]) < 5e-3, l
    # ---- match step on targets [1, 2]
    st.flat.copy_(flat0)
    targets = [int(k) for k in d["targets"]]
    t = im.template
    t.initialize_coordinate_transform(2, True, False, 0.1, True, True, False, True)
    rf = {k: np.array([v["center_pos"]["x"], v["center_pos"]["y"]], np.float32) for k, v in meis.items()}
    t.register_coords_shifts(torch.from_numpy(rf[0]).cuda(), torch.from_numpy(np.array([rf[k] for k in targets])).cuda())
    aff = t.coordinate_transform.transforms.Affine
    with torch.no_grad():
        aff.angles.copy_(torch.from_numpy(d["angles"])), aff.scalings.copy_(torch.from_numpy(d["scalings"]))
        aff.shears.copy_(torch.from_numpy(d["shears"])), aff.translation.copy_(torch.from_numpy(d["translation"]).reshape(2, 1, 2))
    ms = MatchStepper(t, im.img_transforms, model, targets, d["mei_acts"], 20, lr=1e-3, precision=precision, use_graph=False)
    ms.step(grid)
    torch.cuda.synchronize()
    assert relerr(ms.act.cpu().numpy(), d["match_acts_dn"]) < 5e-5
    assert relerr(ms.z.cpu().numpy(), d["match_img_post"]) < 1e-4
    for got, key in zip(ms.grads, ("d_angles", "d_scalings", "d_shears", "d_translation")):
        assert relerr(got.cpu().numpy().reshape(d[key].shape), d[key]) < 5e-3, key


def test_conv_generic_autograd_path_vs_reference():
    """The drop-in modules composed by stock autograd (what a user-written loop would do) reproduce the reference's learn-step gradients."""
    from laminr_b200.contrastive_loss import SimCLROnGrid
    from laminr_b200.utils.general import SingleNeuronModel
    d = golden("conv_learn_match_24")
    model, meis, im = _conv_manifold(d, "fp32")
    grid = torch.from_numpy(d["grid"]).reshape(-1, 1).cuda()
    reg = SimCLROnGrid(1, 20, 0.1, 0.3, True).cuda()
    _, img_post, _acts, _ = im.forward(grid, im.template, im.img_transforms, SingleNeuronModel(model, 0), return_template=True)
    acts = _acts / float(d["mei_act"])
    loss = -acts.mean() - 2 * reg.masked_reg_term(img_post, torch.from_numpy(d["mask"]).cuda(), -1.0, 1.0)
    loss.backward()
    assert abs(loss.item() - d["loss"]) < 5e-5 * abs(d["loss"])
    for l in range(5):
        assert relerr(getattr(im.template.func, f"layer{l}").weight.grad.cpu().numpy(), d[f"dW{l}"]) < 3e-3, l


def test_conv_batched_mei_ascent_vs_oracle():
    """get_mei_dict's batched ascent for the conv encoder == the oracle's per-neuron loop (x <- clip(renorm(x + lr*g)))."""
    from laminr_b200.utils.mei import batched_conv_meis
    from oracle import convcore_oracle as CO
    from oracle import laminr_oracle as O
    d = golden("convcore_a")
    model = _load_mirror(d)
    om = CO.model_from_arrays(d)
    n, iters = d["mu"].shape[0], 12
    rng = np.random.RandomState(4)
    x0 = rng.randn(n, 1, 40, 40).astype(np.float32)
    for kw in (dict(norm=6.0), dict(std=0.2)):
        meis, fev = batched_conv_meis(model, list(range(n)), torch.from_numpy(x0).cuda(), -1.0, 1.0, step_size=10, num_iterations=iters, **kw)
        post = (lambda x: np.clip(O.change_norm(x, kw["norm"], 0.0), -1, 1)) if "norm" in kw else (lambda x: np.clip(O.change_stats(x, kw["std"], 0.0), -1, 1))
        x = post(x0.copy())
        want_f = []
        for _ in range(iters):
            act, g, _ = CO.response_and_grad(x, om, list(range(n)))
            want_f.append(act)
            x = post((x + np.float32(10) * g.astype(np.float32)).astype(np.float32)).astype(np.float32)
        want_f.append(CO.response_and_grad(x, om, list(range(n)))[0])
        assert fev.shape == (n, iters + 1)
        assert relerr(fev.cpu().numpy(), np.stack(want_f, 1)) < 1e-4
        assert relerr(meis.cpu().numpy(), x) < 1e-4


def test_all_neuron_forward_at_config4_size_equals_cone_path():
    """`model(x)` (all 128 neurons, one full-frame evaluation of the core: csrc/fullframe.cu) against the single-neuron cone kernel
    (csrc/convcone.cu, pinned by the reference goldens up to 200x200) on BASELINE config 4 shapes, responses and image gradients."""
    from laminr_b200.neuron_models import conv_core_gaussian_model
    model = conv_core_gaussian_model((1, 200, 200), n_neurons=128, seed=1)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        model.readout._mu.copy_(torch.rand(1, 128, 1, 2, generator=g) * 2 - 1)
        model.readout._features.copy_(torch.randn(1, 32, 1, 128, generator=g) * 0.2)
        model.readout.bias.copy_(torch.randn(128, generator=g) * 0.1)
        for feat in model.core.features:
            feat.norm.running_mean.copy_(torch.randn(32, generator=g) * 0.1)
            feat.norm.running_var.copy_(torch.rand(32, generator=g) + 0.5)
    model = model.cuda().eval()
    x = (torch.randn(3, 1, 200, 200, generator=g) * 0.2).cuda().requires_grad_(True)
    idx = [0, 17, 64, 127]
    full = model(x)
    assert full.shape == (3, 128)
    cone = model.forward_neurons(x, idx)
    assert relerr(full[:, idx].detach().cpu().numpy(), cone.detach().cpu().numpy()) < TOL_ACT
    w = torch.randn(3, 4, generator=g).cuda()
    (g_full,) = torch.autograd.grad((full[:, idx] * w).sum(), x, retain_graph=True)  # `full` is differentiated again below
    (g_cone,) = torch.autograd.grad((cone * w).sum(), x)
    assert relerr(g_full.cpu().numpy(), g_cone.cpu().numpy()) < TOL_GRAD
    # dense upstream gradient over all neurons: against float64 stock ops on the host
    import torch.nn.functional as F
    gw = torch.randn(3, 128, generator=g)
    (g_all,) = torch.autograd.grad((full * gw.cuda()).sum(), x)
    xd = x.detach().cpu().double().requires_grad_(True)
    h = xd
    for feat in model.core.features:
        n = feat.norm
        h = F.conv2d(h, feat.conv.weight.detach().cpu().double(), padding=feat.conv.padding)
        h = F.elu(F.batch_norm(h, n.running_mean.cpu().double(), n.running_var.cpu().double(), n.weight.detach().cpu().double(), n.bias.detach().cpu().double(),
                               False, 0.0, n.eps))
    ro = model.readout
    y = F.grid_sample(h, ro._mu.detach().cpu().double().expand(3, -1, 1, 2), align_corners=True).squeeze(-1)
    y = F.elu((y * ro._features.detach().cpu().double().view(1, 32, 128)).sum(1) + ro.bias.detach().cpu().double()) + 1
    (g_want,) = torch.autograd.grad((y * gw.double()).sum(), xd)
    assert relerr(full.detach().cpu().numpy(), y.detach().numpy()) < TOL_ACT
    assert relerr(g_all.cpu().numpy(), g_want.numpy()) < TOL_GRAD
