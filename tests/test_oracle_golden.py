"""CPU: the numpy oracle (oracle/laminr_oracle.py) against golden vectors produced by the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle; the GPU parity tests then compare CUDA against the oracle."""
import hashlib

import numpy as np
import pytest

from oracle import laminr_oracle as O
from tests._util import golden, net_params, relerr

TOL = 2e-5  # float32 restatement vs float32 torch reference (different summation orders)


@pytest.mark.parametrize("name,res", [("inr_12x10", (12, 10)), ("inr_9x16_c2", (9, 16))])
def test_inr_forward_backward(name, res):
    d = golden(name)
    W, b = net_params(d)
    # torch.linspace vectorises (base + step*lane), so the scalar restatement may differ by 1 ulp; coords are an INPUT
    # of the kernels (the product builds the buffer with torch.linspace exactly like inr.py:194-198)
    np.testing.assert_allclose(O.pixel_coords(res), d["coords"], rtol=0, atol=6e-8)
    img, cache = O.inr_forward(W, b, d["B"], d["auxB"], d["grid"], d["coords"][None], res, keep=True)
    assert relerr(img[0], d["out"]) < TOL
    g = O.inr_backward(W, d["B"], cache, d["upstream"][None])
    for l in range(5):
        assert relerr(g["dW"][l], d[f"dW{l}"]) < 5e-5, l
        assert relerr(g["db"][l], d[f"db{l}"]) < 5e-5, l
    # float64 restatement agrees too (tighter): rules out a same-precision coincidence
    W64, b64 = [w.astype(np.float64) for w in W], [x.astype(np.float64) for x in b]
    img64 = O.inr_forward(W64, b64, d["B"].astype(np.float64), d["auxB"].astype(np.float64), d["grid"].astype(np.float64),
                          d["coords"][None].astype(np.float64), res)
    assert relerr(img64[0], d["out"]) < TOL


def test_match_coords_and_affine_grads():
    d = golden("match_inr_14x14")
    W, b = net_params(d)
    res = (14, 14)
    shift_to, tc, shift_from = O.coords_shifts(d["template_xy"], d["target_xy"], res)
    np.testing.assert_allclose(shift_to, d["shift_to"], atol=1e-6)
    np.testing.assert_allclose(tc, d["tc"], atol=1e-6)
    np.testing.assert_allclose(shift_from, d["shift_from"], atol=1e-6)
    M, _ = O.affine_matrices(d["angles"], d["scalings"], d["shears"])
    cdp = O.match_coords(d["coords"], M, d["translation"], tc, shift_to, shift_from)
    np.testing.assert_allclose(cdp, d["coords_transformed"], atol=2e-6)
    assert (np.abs(cdp) == 1).any(), "fixture should exercise the clamp"
    img, cache = O.inr_forward(W, b, d["B"], d["auxB"], d["grid"], cdp, res, keep=True)
    assert relerr(img, d["out"]) < TOL
    G = O.inr_backward(W, d["B"], cache, d["upstream"], want_weights=False, want_coords=True)["dcoords"]
    da, ds, dsh, dt = O.affine_backward(G, d["coords"], tc, d["angles"], d["scalings"], d["shears"])
    for got, key in ((da, "d_angles"), (ds, "d_scalings"), (dsh, "d_shears"), (dt, "d_translation")):
        assert relerr(got, d[key]) < 2e-4, key


def test_constrain():
    d = golden("constrain")
    y, c = O.constrain_norm_fwd(d["norm_x"], 3.0, 0.25, -0.5, 1.0)
    assert relerr(y, d["norm_y"]) < 1e-6
    assert relerr(O.constrain_norm_bwd(d["norm_up"], c, 3.0, -0.5, 1.0), d["norm_gx"]) < TOL
    y, c = O.constrain_std_fwd(d["std_x"], 0.4, 0.25, -0.5, 1.0)
    assert relerr(y, d["std_y"]) < 1e-6
    assert relerr(O.constrain_std_bwd(d["std_up"], c, 0.4, -0.5, 1.0), d["std_gx"]) < TOL


def test_simresp_and_banks():
    d = golden("simresp")
    for res in (16, 24):
        banks = O.demo1_banks([res, res])
        for i in range(3):
            np.testing.assert_array_equal(banks[i], d[f"bank{i}_{res}"])  # bit-identical float32 banks
        x = d[f"x_{res}"]
        assert relerr(O.multi_resp_fwd(x, banks), d[f"y_{res}"]) < 1e-6
        gx = np.zeros_like(x)
        for i in range(3):
            _, c = O.simresp_fwd(x, banks[i])
            gx += O.simresp_bwd(d[f"up_{res}"][:, i], c, x.shape, banks[i])
        assert relerr(gx, d[f"gx_{res}"]) < 1e-5
    banks = O.demo1_banks([16, 16])
    assert relerr(O.multi_resp_fwd(d["x_c2"], banks), d["y_c2"]) < 1e-6
    gx = sum(O.simresp_bwd(np.ones(3, np.float32), O.simresp_fwd(d["x_c2"], f)[1], d["x_c2"].shape, f) for f in banks)
    assert relerr(gx, d["gx_c2"]) < 1e-5


def test_banks_full_resolution_hashes():
    d = golden("banks")
    for res in (100, 200):
        for i, f in enumerate(O.demo1_banks([res, res])):
            assert tuple(d[f"shape_{res}_{i}"]) == f.shape
            assert hashlib.sha256(f.tobytes()).digest() == d[f"sha_{res}_{i}"].tobytes(), (res, i)
    # known-answer: unit-norm filters
    for f in O.demo1_banks([100, 100]):
        np.testing.assert_allclose(np.linalg.norm(f.reshape(len(f), -1).astype(np.float64), axis=1), 1.0, atol=1e-6)


def test_simclr():
    d = golden("simclr")
    for tag, n, periodic, nb in (("p20", 20, True, 0.1), ("np7", 7, False, 0.3), ("p10", 10, True, 0.1)):
        pos, neg = O.pos_neg_masks(n, nb, periodic)
        np.testing.assert_array_equal(pos, d[f"{tag}_pos"])
        np.testing.assert_array_equal(neg, d[f"{tag}_neg"])
        pmin, pmax = map(float, d[f"{tag}_pm"])
        x, mask = d[f"{tag}_x"], d[f"{tag}_mask"]
        xm = O.apply_mask(x, mask, pmin, pmax)
        reg, c = O.simclr_reg_fwd(xm, pos, neg, 0.3)
        assert abs(reg - d[f"{tag}_reg"]) < 1e-6 * max(1, abs(d[f"{tag}_reg"]))
        gx = O.simclr_reg_bwd(c, pos, neg, 0.3).reshape(x.shape) * mask
        assert relerr(gx, d[f"{tag}_gx"]) < 5e-5, tag
    pos, neg = O.pos_neg_masks(20, 0.1, True)  # known-answer: 4 positives / 15 negatives per row
    assert (pos.sum(1) == 4).all() and (neg.sum(1) == 15).all()


def test_training_utils():
    d = golden("training_utils")
    for n in (20, 7, 32):
        np.testing.assert_allclose(O.latent_grid(n), d[f"grid_{n}"], rtol=0, atol=5e-7)  # 1 ulp: torch.linspace vectorisation
    np.testing.assert_allclose(O.jittered_grids(O.latent_grid(20), d["jitter_u01"], 2 * np.pi / 20), d["jitter_batches"], atol=5e-7)
    sch = O.ParamReduceOnPlateau(factor=0.8, patience=5, threshold=0.005, mode="max")
    val, vals = 2.0, []
    for s in d["plateau_seq"]:
        val = sch.step(s, val)
        vals.append(val)
    np.testing.assert_array_equal(np.array(vals), d["plateau_vals"])
    chk = O.ImprovementChecker(patience=4, ignore_diff_smaller_than=1e-3)
    np.testing.assert_array_equal(np.array([chk.is_increasing(float(s)) for s in d["plateau_seq"]]), d["improve_flags"])
    req = dict(avg=0.99, std=1.0, necessary_min=0.98)
    np.testing.assert_array_equal(np.array([O.check_activity_requirements(a, req) for a in d["req_acts"]]), d["req_pass"])
    assert d["req_pass"].any() and not d["req_pass"].all()


def test_adam():
    d = golden("adam")
    p, m, v = d["p0"], np.zeros_like(d["p0"]), np.zeros_like(d["p0"])
    for i, g in enumerate(d["grads"]):
        p, m, v = O.adam_step(p, g, m, v, i + 1)
        np.testing.assert_allclose(p, d["ps"][i], rtol=0, atol=2e-7)


def test_mei_ascent():
    d = golden("mei")
    banks = O.demo1_banks([16, 16])
    for tag, kw in (("norm", dict(norm=1.0)), ("std", dict(std=0.1))):
        for idx in range(3):
            x, fevals = O.mei_ascent(d[f"{tag}_x0_{idx}"], banks[idx], 10, 60, -1.0, 1.0, **kw)
            assert len(fevals) == 61  # known-answer: num_iterations + 1 (featurevis/core.py:122-126)
            np.testing.assert_allclose(fevals, d[f"{tag}_fevals_{idx}"], rtol=2e-5)
            assert relerr(x, d[f"{tag}_x_{idx}"]) < 1e-4, (tag, idx)


@pytest.mark.parametrize("res", [20, 100])
def test_learn_step(res):
    d = golden(f"learn_step_{res}")
    W, b = net_params(d)
    pos, neg = O.pos_neg_masks(20, 0.1, True)
    banks = O.demo1_banks([res, res])
    out = O.learn_step(W, b, d["B"], d["auxB"], d["grid"], d["coords"], (res, res), banks[0], d["mask"], float(d["mei_act"]),
                       2.0, 1.0, -1.0, 1.0, pos, neg, 0.3)
    assert abs(out["loss"] - d["loss"]) < 1e-3 * abs(d["loss"])  # north-star tolerance ...
    assert abs(out["loss"] - d["loss"]) < 2e-5 * abs(d["loss"])  # ... and what fp32 actually achieves
    assert abs(out["sim"] - d["sim"]) < 2e-5 * abs(d["sim"])
    assert relerr(out["acts"], d["acts"]) < 1e-5
    sub = slice(None) if res == 20 else slice(None, None, 37)
    assert relerr(out["img_pre"].reshape(20, -1)[:, sub], d["img_pre"]) < TOL
    assert relerr(out["img_post"].reshape(20, -1)[:, sub], d["img_post"]) < TOL
    for l in range(5):
        assert relerr(out["dW"][l], d[f"dW{l}"]) < 2e-3, l  # gradients are small residuals; fp32 summation-order noise
        assert relerr(out["db"][l], d[f"db{l}"]) < 2e-3, l


def test_learn_step_grads_fp64_close():
    """fp64 oracle vs the fp32 reference gradients: the fp32 restatement error above is rounding, not maths."""
    d = golden("learn_step_20")
    f64 = lambda a: np.asarray(a, np.float64)
    W, b = net_params(d)
    pos, neg = O.pos_neg_masks(20, 0.1, True, np.float64)
    bank = f64(O.demo1_banks([20, 20])[0])
    out = O.learn_step([f64(w) for w in W], [f64(x) for x in b], f64(d["B"]), f64(d["auxB"]), f64(d["grid"]), f64(d["coords"]),
                       (20, 20), bank, f64(d["mask"]), float(d["mei_act"]), 2.0, 1.0, -1.0, 1.0, pos, neg, 0.3)
    for l in range(5):
        assert relerr(out["dW"][l], d[f"dW{l}"]) < 2e-3, l


def test_match_step():
    d = golden("match_step_20")
    W, b = net_params(d)
    res = (20, 20)
    banks = O.demo1_banks(list(res))
    shifts = O.coords_shifts(d["template_xy"], d["target_xy"], res)
    aff = dict(angles=d["angles"], scalings=d["scalings"], shears=d["shears"], translation=d["translation"])
    out = O.match_step(W, b, d["B"], d["auxB"], d["grid"], d["coords"], res, [banks[k] for k in d["targets"]],
                       d["mei_acts"].astype(np.float32), aff, shifts, 1.0, -1.0, 1.0)
    assert abs(out["loss"] - d["loss"]) < 2e-5 * abs(d["loss"])
    assert relerr(out["acts_dn"], d["acts_dn"]) < 1e-5
    assert relerr(out["img_pre"], d["img_pre"]) < TOL
    assert relerr(out["img_post"], d["img_post"]) < TOL
    for k in ("d_angles", "d_scalings", "d_shears", "d_translation"):
        assert relerr(out[k], d[k]) < 1e-3, k


def test_learn_trajectory():
    """30 Adam steps replayed with the oracle: per-step loss within 1e-3 rel (north-star), final params close."""
    d = golden("learn_traj_20")
    W = [d[f"init_W{l}"] for l in range(5)]
    b = [d[f"init_b{l}"] for l in range(5)]
    mW, vW = [np.zeros_like(w) for w in W], [np.zeros_like(w) for w in W]
    mb, vb = [np.zeros_like(x) for x in b], [np.zeros_like(x) for x in b]
    pos, neg = O.pos_neg_masks(20, 0.1, True)
    bank = O.demo1_banks([20, 20])[0]
    grids = O.jittered_grids(O.latent_grid(20), d["jitter_u01"], 2 * np.pi / 20)
    np.testing.assert_allclose(grids, d["grids"], atol=5e-7)
    for i in range(30):
        out = O.learn_step(W, b, d["init_B"], d["init_auxB"], d["grids"][i], d["init_coords"], (20, 20), bank, d["mask"],
                           1.0000009536743164, 2.0, 1.0, -1.0, 1.0, pos, neg, 0.3)
        assert abs(out["loss"] - d["losses"][i]) < 1e-3 * abs(d["losses"][i]), i
        for l in range(5):
            W[l], mW[l], vW[l] = O.adam_step(W[l], out["dW"][l], mW[l], vW[l], i + 1)
            b[l], mb[l], vb[l] = O.adam_step(b[l], out["db"][l], mb[l], vb[l], i + 1)
    for l in range(5):
        assert relerr(W[l], d[f"final_W{l}"]) < 1e-2, l


@pytest.mark.parametrize("name", ["convcore_a", "convcore_b", "convcore_c"])
def test_convcore_oracle_against_reference(name):
    """Stacked-conv core + Gaussian readout + firing-rate encoder (row a-9): all-neuron responses and the input gradient of the
    selected-neuron responses, as produced by the reference's vendored neuralpredictors modules."""
    from oracle import convcore_oracle as CO

    d = golden(name)
    model = CO.model_from_arrays(d)
    acts_all = CO.model_forward(d["x"], model)
    assert relerr(acts_all, d["acts"]) < TOL
    act, g_x, acts_all2 = CO.response_and_grad(d["x"], model, d["neuron_of_image"], d["g_act"])
    B = d["x"].shape[0]
    assert relerr(act, d["acts"][np.arange(B), d["neuron_of_image"]]) < TOL
    assert relerr(g_x, d["g_x"]) < 5e-5
    np.testing.assert_allclose(acts_all2, acts_all)
    # the gradient is confined to the receptive-field cone of the selected neuron (what the CUDA path exploits)
    k_total = sum(L["w"].shape[-1] - 1 for L in model["layers"]) + 2
    for b in range(B):
        nz = np.argwhere(np.abs(d["g_x"][b]).sum(0) > 0)
        assert nz[:, 0].max() - nz[:, 0].min() + 1 <= k_total and nz[:, 1].max() - nz[:, 1].min() + 1 <= k_total


def _conv_response(model, neuron):
    from oracle import convcore_oracle as CO

    def response(z, g_act):
        act, g_z, _ = CO.response_and_grad(z, model, [neuron] * len(z), g_act)
        return act, g_z
    return response


def test_conv_learn_and_match_step():
    """Learn / match hot steps of the reference's InvarianceManifold driven by a conv-core + Gaussian-readout encoder (config 4, small)."""
    from oracle import convcore_oracle as CO

    d = golden("conv_learn_match_24")
    W, b = net_params(d)
    res = (24, 24)
    model = CO.model_from_arrays(d, prefix="m_")
    pos, neg = O.pos_neg_masks(20, 0.1, True)
    out = O.learn_step(W, b, d["B"], d["auxB"], d["grid"], d["coords"], res, None, d["mask"], float(d["mei_act"]), 2.0, float(d["norm"]), -1.0, 1.0,
                       pos, neg, 0.3, response=_conv_response(model, 0))
    assert abs(out["loss"] - d["loss"]) < 2e-5 * abs(d["loss"])
    assert relerr(out["acts"], d["acts"]) < 1e-5
    assert relerr(out["img_post"].reshape(20, -1), d["img_post"]) < TOL
    for l in range(5):
        assert relerr(out["dW"][l], d[f"dW{l}"]) < 2e-3, l
        assert relerr(out["db"][l], d[f"db{l}"]) < 2e-3, l
    shifts = O.coords_shifts(d["template_xy"], d["target_xy"], res)
    aff = dict(angles=d["angles"], scalings=d["scalings"], shears=d["shears"], translation=d["translation"])
    out = O.match_step(W, b, d["B"], d["auxB"], d["grid"], d["coords"], res, None, d["mei_acts"].astype(np.float32), aff, shifts, float(d["norm"]), -1.0, 1.0,
                       responses=[_conv_response(model, int(k)) for k in d["targets"]])
    assert abs(out["loss"] - d["match_loss"]) < 2e-5 * abs(d["match_loss"])
    assert relerr(out["acts_dn"], d["match_acts_dn"]) < 1e-5
    assert relerr(out["img_post"], d["match_img_post"]) < TOL
    for k in ("d_angles", "d_scalings", "d_shears", "d_translation"):
        assert relerr(out[k], d[k]) < 1e-3, k


def test_torch_port_conv_response_matches_reference():
    """The torch-CPU port used as bench.py's CPU baseline for config 4 (oracle/torch_port.py) reproduces the reference's learn step."""
    import torch
    from oracle.torch_port import ConvResponsePort, LearnPort

    d = golden("conv_learn_match_24")
    W, b = net_params(d)
    pos, neg = O.pos_neg_masks(20, 0.1, True)
    port = LearnPort(W, b, d["B"], d["auxB"], d["coords"], None, d["mask"], float(d["mei_act"]), pos, neg, norm=float(d["norm"]), res=(24, 24),
                     response=ConvResponsePort(d, 0, prefix="m_"))
    loss, acts, _ = port.loss(d["grid"])
    assert abs(float(loss) - d["loss"]) < 2e-5 * abs(d["loss"])
    assert relerr(acts.detach().numpy(), d["acts"]) < 1e-5
    loss.backward()
    for l in range(5):
        assert relerr(port.W[l].grad.numpy(), d[f"dW{l}"]) < 1e-3, l


def test_init_sweep_table_and_choice():
    """Match initialisation sweep (inv_temp_learn_match.py:573-650): the oracle's no-grad forward reproduces the table of mean activations the
    UNMODIFIED reference function builds (recorded by oracle/make_golden_sweep.py), and the reference's final choice is the first maximum of that
    table in (angle, tx, ty) order with the scalings set to template-mask / target-mask size on both axes."""
    d = golden("init_sweep_20")
    W, b = net_params(d)
    res = (20, 20)
    banks = [O.demo1_banks(list(res))[k] for k in d["targets"]]
    shift_to, tc, shift_from = O.coords_shifts(d["template_xy"], d["target_xy"], res)
    angles = np.linspace(0, 2 * np.pi, 61)[:-1].astype(np.float32)
    trs = np.linspace(-0.1, 0.1, 11).astype(np.float32)
    inv = (1.0 / (d["target_masks"].sum(axis=(1, 2)) / d["template_mask"].sum())).astype(np.float32)
    for tag in ("a", "at"):
        np.testing.assert_allclose(d[f"scalings_{tag}"], np.repeat(inv[:, None], 2, axis=1), rtol=1e-6)

    def mean_acts(angle, trans):
        M, _ = O.affine_matrices(np.full(2, angle, np.float32), d["scalings_a"], d["shears0"])
        cdp = O.match_coords(d["coords"], M, trans, tc, shift_to, shift_from, True)
        img = O.inr_forward(W, b, d["B"], d["auxB"], d["grid"], cdp, res)
        D, N = img.shape[:2]
        z, _ = O.constrain_norm_fwd(img.reshape((D * N,) + img.shape[2:]), 1.0, 0.0, -1.0, 1.0)
        zs = z.reshape(img.shape)
        return np.array([O.simresp_fwd(zs[k], banks[k])[0].mean() for k in range(D)])

    # angle only (the `match` default): all 60 candidates; the translation keeps its (noisy) initial value
    got = np.stack([mean_acts(a, d["translation0"]) for a in angles])
    assert relerr(got, d["table_a"][:, 0, 0]) < 2e-5
    for k in range(2):
        assert angles[np.argmax(got[:, k])] == d["angles_a"][k] == angles[np.argmax(d["table_a"][:, 0, 0, k])]
    np.testing.assert_array_equal(d["translation_a"], d["translation0"])
    # angle + translation: a sub-grid of the 60 x 11 x 11 candidates, and the reference's choice == first maximum of ITS table
    for i in range(0, 60, 10):
        for j in (0, 5, 10):
            for k in (0, 5, 10):
                tr = np.tile(np.array([trs[j], trs[k]], np.float32), (2, 1))
                np.testing.assert_allclose(mean_acts(angles[i], tr), d["table_at"][i, j, k], rtol=3e-5)
    best = np.argmax(d["table_at"].reshape(-1, 2), axis=0)
    bi, bj, bk = np.unravel_index(best, (60, 11, 11))
    np.testing.assert_array_equal(d["angles_at"], angles[bi])
    np.testing.assert_array_equal(d["translation_at"], np.stack([trs[bj], trs[bk]], axis=1))


def test_learn_call_replayed_with_the_oracle_and_the_host_stream():
    """The reference's whole `learn(0, steps_per_epoch=5, num_max_epochs=4)` call (tests/golden/learn_call_20.npz, oracle/make_golden_learn_call.py)
    replayed from seeds alone: the product's CPU-side constructor (initial network, inr.py:49-51 draw order) and jittered-grid stream
    (training.py:18-53) feed the numpy oracle's learn step + Adam; final network, images, activations and the generator state must be the call's."""
    import torch
    import laminr_b200 as L
    from laminr_b200.utils.training import JitteringGridDatamodule
    d = golden("learn_call_20")
    res = 20
    meis = {i: dict(mei=np.zeros((1, res, res), np.float32), activation=float(d["mei_acts"][i]), mask=d["masks"][i],
                    center_pos=dict(x=float(d["centers"][i, 0]), y=float(d["centers"][i, 1]))) for i in range(3)}
    torch.manual_seed(0)
    im = L.InvarianceManifold(L.neuron_models.simulated("demo1", img_res=[res, res]), meis, device="cpu", required_img_norm=1.0,
                              pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    sd = {k: v.numpy().copy() for k, v in im.template.state_dict().items()}
    W, b = [sd[f"func.layer{l}.weight"] for l in range(5)], [sd[f"func.layer{l}.bias"] for l in range(5)]
    coords = sd["coords"].reshape(-1, 2)
    mW, vW = [np.zeros_like(w) for w in W], [np.zeros_like(w) for w in W]
    mb, vb = [np.zeros_like(x) for x in b], [np.zeros_like(x) for x in b]
    pos, neg = O.pos_neg_masks(20, 0.1, True)
    bank = O.demo1_banks([res, res])[0]
    torch.manual_seed(11)
    dm = JitteringGridDatamodule(num_invariances=1, grid_points_per_dim=20, steps_per_epoch=5)
    dm.train_dataloader()  # the reference builds (and discards) one loader before the loop: one jitter draw (inv_temp_learn_match.py:137)
    step = 0
    for _ in range(4):
        for g in dm.train_dataloader():
            out = O.learn_step(W, b, sd["B"], sd["auxB"], g[:, 0].numpy(), coords, (res, res), bank, d["masks"][0], float(d["mei_acts"][0]), 2.0, 1.0,
                               -1.0, 1.0, pos, neg, 0.3)
            step += 1
            for l in range(5):
                W[l], mW[l], vW[l] = O.adam_step(W[l], out["dW"][l], mW[l], vW[l], step)
                b[l], mb[l], vb[l] = O.adam_step(b[l], out["db"][l], mb[l], vb[l], step)
    assert step == 20
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(d["short_rng_after"]))
    for l in range(5):
        assert relerr(W[l], d[f"short_W{l}"]) < 2e-3, l
        assert relerr(b[l], d[f"short_b{l}"]) < 2e-2, l     # biases start at 0: twenty lr-sized steps, sign-ambiguous entries weigh more
    final = O.learn_step(W, b, sd["B"], sd["auxB"], O.latent_grid(20), coords, (res, res), bank, d["masks"][0], float(d["mei_acts"][0]), 2.0, 1.0, -1.0,
                         1.0, pos, neg, 0.3)
    assert relerr(final["img_post"].reshape(d["short_imgs"].shape), d["short_imgs"]) < 1e-2
    assert relerr(final["acts"], d["short_acts"]) < 1e-2


_MATCH_VARIANTS = {
    "default": dict(),
    "default_100": dict(),
    "uniform_noshear": dict(uniform_scale_coordinate_transformation=True, allow_shear_coordinate_transformation=False),
    "noscale_noclamp": dict(allow_scale_coordinate_transformation=False, coordinate_transform_clamp_boundaries=False),
    "nosweep_3steps": dict(rotate_angle_and_scale=False, steps_per_epoch=3),
}


@pytest.mark.parametrize("tag", list(_MATCH_VARIANTS))
def test_match_call_replayed_with_the_oracle_and_the_host_stream(tag):
    """Whole `match([1, 2], num_epochs=5, ...)` calls of the unmodified reference (tests/golden/match_call_20.npz, match_call_variants_20.npz;
    oracle/make_golden_match_call.py) replayed from the seed alone: the 60-angle initialisation sweep (:573-650), the product's jittered-grid stream,
    the oracle's match step and Adam over exactly the scalars the flags leave trainable (:257-278, coordinate_transforms.py:226-262), the final
    un-jittered snapshot; the CPU generator must end where the call leaves it."""
    import torch
    from laminr_b200.utils.training import JitteringGridDatamodule
    kw = _MATCH_VARIANTS[tag]
    d = golden("match_call_100" if tag == "default_100" else "match_call_20" if tag == "default" else "match_call_variants_20")
    pre = "" if tag.startswith("default") else tag + "_"
    res, targets, D = ((100, 100) if tag == "default_100" else (20, 20)), [1, 2], 2   # default_100: BASELINE configs[1] at its own size
    W, b = [d[f"W{l}"] for l in range(5)], [d[f"b{l}"] for l in range(5)]
    banks = [O.demo1_banks(list(res))[k] for k in targets]
    mei_acts = d["mei_acts"][targets]
    shifts = O.coords_shifts(d["centers"][0].astype(np.float32), d["centers"][targets].astype(np.float32), res)
    clamp = kw.get("coordinate_transform_clamp_boundaries", True)
    uniform = kw.get("uniform_scale_coordinate_transformation", False)
    P = dict(angles=np.zeros(D, np.float32), scalings=np.ones((D, 1 if uniform else 2), np.float32), shears=np.zeros((D, 2), np.float32),
             translation=np.zeros((D, 2), np.float32))
    trainable = ["angles", "translation"] + (["scalings"] if kw.get("allow_scale_coordinate_transformation", True) else []) + \
                (["shears"] if kw.get("allow_shear_coordinate_transformation", True) else [])
    full = lambda: dict(P, scalings=np.repeat(P["scalings"], 2, axis=1) if uniform else P["scalings"])
    grid0 = O.latent_grid(20)

    def forward(grid):
        M, _ = O.affine_matrices(P["angles"], full()["scalings"], P["shears"])
        cdp = O.match_coords(d["coords"], M, P["translation"], shifts[1], shifts[0], shifts[2], clamp)
        img = O.inr_forward(W, b, d["B"], d["auxB"], grid, cdp, res)
        z, _ = O.constrain_norm_fwd(img.reshape((D * 20,) + img.shape[2:]), 1.0, 0.0, -1.0, 1.0)
        zs = z.reshape(img.shape)
        return zs, np.stack([O.simresp_fwd(zs[k], banks[k])[0] for k in range(D)])

    torch.manual_seed(5)
    dm = JitteringGridDatamodule(num_invariances=1, grid_points_per_dim=20, steps_per_epoch=kw.get("steps_per_epoch", 1))
    if kw.get("rotate_angle_and_scale", True):
        inv = (1.0 / (d["masks"][targets].sum(axis=(1, 2)) / d["masks"][0].sum())).astype(np.float32)
        P["scalings"] = np.ones_like(P["scalings"]) * inv[:, None]
        angles = np.linspace(0, 2 * np.pi, 61)[:-1].astype(np.float32)
        table = np.zeros((60, D))
        for i, a in enumerate(angles):
            P["angles"] = np.full(D, a, np.float32)
            table[i] = forward(grid0)[1].mean(axis=1)
        P["angles"] = angles[np.argmax(table, axis=0)]
    m, v = {k: np.zeros_like(P[k]) for k in trainable}, {k: np.zeros_like(P[k]) for k in trainable}
    step = 0
    for _ in range(5):
        for g in dm.train_dataloader():
            out = O.match_step(W, b, d["B"], d["auxB"], g[:, 0].numpy(), d["coords"], res, banks, mei_acts, full(), shifts, 1.0, -1.0, 1.0, clamp=clamp)
            grads = dict(angles=out["d_angles"], shears=out["d_shears"], translation=out["d_translation"],
                         scalings=out["d_scalings"].sum(axis=1, keepdims=True) if uniform else out["d_scalings"])
            step += 1
            for k in trainable:
                P[k], m[k], v[k] = O.adam_step(P[k], grads[k].astype(np.float32), m[k], v[k], step)
    assert step == 5 * kw.get("steps_per_epoch", 1)
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(d[pre + "rng_after"]))
    for k in ("angles", "scalings", "shears", "translation"):
        assert P[k].shape == d[pre + k].shape, k
        np.testing.assert_allclose(P[k], d[pre + k], rtol=1e-4, atol=5e-4, err_msg=k)   # the GPU test's tolerances
    imgs, acts = forward(grid0)
    assert relerr(imgs.reshape(d[pre + "imgs"].shape), d[pre + "imgs"].astype(np.float32)) < 1e-2 and relerr(acts, d[pre + "acts"]) < 1e-2


def test_get_mei_dict_std_call_replayed_with_the_oracle():
    """The reference's `get_mei_dict(model, [1, 40, 40], 0, 1, neuron_idxs=[2, 0], required_img_std=0.05, zscore=0.5, step_size=5,
    num_iterations=300)` (tests/golden/mei_dict_std_40.npz, oracle/make_golden_meis.py) replayed from the seed: one CPU draw per listed neuron in list
    order (mei.py:33), + mean, ChangeStats + ClipRange ascent (featurevis/ops.py:294-307,441-459), create_mask at the given z-score (mask.py:25-64)."""
    import torch
    from oracle import mask_oracle as MO
    d = golden("mei_dict_std_40")
    banks = O.demo1_banks([40, 40])
    torch.manual_seed(3)
    for k, neuron in enumerate([2, 0]):
        x0 = torch.randn(1, 1, 40, 40, dtype=torch.float32).numpy() + np.float32(0.5)
        x, fevals = O.mei_ascent(x0, banks[neuron], 5.0, 300, 0.0, 1.0, std=0.05)
        assert len(fevals) == 301 and abs(fevals[-1] - d["activations"][k]) < 2e-5 * d["activations"][k]
        assert relerr(x[0], d["meis"][k]) < 1e-4
        mask, centre = MO.create_mask(d["meis"][k][0], zscore_thresh=0.5)[:2]
        assert np.abs(np.asarray(mask, np.float32) - d["masks"][k]).max() < 1e-6
        assert abs(centre["x"] - d["centers"][k, 0]) < 1e-9 and abs(centre["y"] - d["centers"][k, 1]) < 1e-9
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(d["rng_after"]))


def test_std_constraint_calls_with_snapshots_replayed_with_the_oracle():
    """`InvarianceManifold(required_img_std=0.1)`: whole `learn(0, steps_per_epoch=5, num_max_epochs=4, save_results_every_n_epochs=2)` and, continuing
    on the same generator, `match([1, 2], num_epochs=4, save_results_every_n_epochs=2)` of the unmodified reference (tests/golden/std_calls_20.npz,
    oracle/make_golden_std_calls.py) replayed from seeds alone: StandardizeClip in both loops, snapshots at epochs 0, 2 and the end."""
    import torch
    import laminr_b200 as L
    from laminr_b200.utils.training import JitteringGridDatamodule
    d = golden("std_calls_20")
    res, targets, D = (20, 20), [1, 2], 2
    meis = {i: dict(mei=np.zeros((1,) + res, np.float32), activation=float(d["mei_acts"][i]), mask=d["masks"][i],
                    center_pos=dict(x=float(d["centers"][i, 0]), y=float(d["centers"][i, 1]))) for i in range(3)}
    torch.manual_seed(0)
    im = L.InvarianceManifold(L.neuron_models.simulated("demo1", img_res=list(res)), meis, device="cpu", required_img_std=0.1,
                              pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    sd = {k: v.numpy().copy() for k, v in im.template.state_dict().items()}
    W, b = [sd[f"func.layer{l}.weight"] for l in range(5)], [sd[f"func.layer{l}.bias"] for l in range(5)]
    coords, Bm, auxB = sd["coords"].reshape(-1, 2), sd["B"], sd["auxB"]
    banks = O.demo1_banks(list(res))
    pos, neg = O.pos_neg_masks(20, 0.1, True)
    grid0 = O.latent_grid(20)
    mei_act = float(d["mei_acts"][0])

    # ---- learn
    def learn_fwd(grid):
        return O.learn_step(W, b, Bm, auxB, grid, coords, res, banks[0], d["masks"][0], mei_act, 2.0, None, -1.0, 1.0, pos, neg, 0.3, std=0.1)

    mW, vW = [np.zeros_like(w) for w in W], [np.zeros_like(w) for w in W]
    mb, vb = [np.zeros_like(x) for x in b], [np.zeros_like(x) for x in b]
    torch.manual_seed(21)
    dm = JitteringGridDatamodule(num_invariances=1, grid_points_per_dim=20, steps_per_epoch=5)
    dm.train_dataloader()  # built and discarded by the reference (:137)
    snaps, step = [], 0
    for epoch in range(4):
        if epoch % 2 == 0:
            snaps.append(learn_fwd(grid0))
        for g in dm.train_dataloader():
            out = learn_fwd(g[:, 0].numpy())
            step += 1
            for l in range(5):
                W[l], mW[l], vW[l] = O.adam_step(W[l], out["dW"][l], mW[l], vW[l], step)
                b[l], mb[l], vb[l] = O.adam_step(b[l], out["db"][l], mb[l], vb[l], step)
    snaps.append(learn_fwd(grid0))
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(d["rng_after_learn"]))
    assert d["learn_imgs"].shape == (3, 20, 1, 20, 20) and d["learn_acts"].shape == (3, 20)
    for k, s in enumerate(snaps):
        assert relerr(s["img_post"].reshape(20, 1, 20, 20), d["learn_imgs"][k]) < 1e-2, k
        assert relerr(s["acts"] * mei_act, d["learn_acts"][k]) < 1e-2, k          # the getter returns the raw activations (:399)
    for l in range(5):
        assert relerr(W[l], d[f"learned_W{l}"]) < 2e-3, l

    # ---- match, from the learned template, on the same generator
    mei_acts = d["mei_acts"][targets]
    shifts = O.coords_shifts(d["centers"][0].astype(np.float32), d["centers"][targets].astype(np.float32), res)
    P = dict(angles=np.zeros(D, np.float32), scalings=np.ones((D, 2), np.float32), shears=np.zeros((D, 2), np.float32), translation=np.zeros((D, 2), np.float32))

    def match_fwd(grid):
        M, _ = O.affine_matrices(P["angles"], P["scalings"], P["shears"])
        img = O.inr_forward(W, b, Bm, auxB, grid, O.match_coords(coords, M, P["translation"], shifts[1], shifts[0], shifts[2], True), res)
        z, _ = O.constrain_std_fwd(img.reshape((D * 20,) + img.shape[2:]), 0.1, 0.0, -1.0, 1.0)
        zs = z.reshape(img.shape)
        return zs, np.stack([O.simresp_fwd(zs[k], banks[targets[k]])[0] for k in range(D)])

    dm = JitteringGridDatamodule(num_invariances=1, grid_points_per_dim=20, steps_per_epoch=1)
    inv = (1.0 / (d["masks"][targets].sum(axis=(1, 2)) / d["masks"][0].sum())).astype(np.float32)
    P["scalings"] = np.ones_like(P["scalings"]) * inv[:, None]
    angles = np.linspace(0, 2 * np.pi, 61)[:-1].astype(np.float32)
    table = np.zeros((60, D))
    for i, a in enumerate(angles):
        P["angles"] = np.full(D, a, np.float32)
        table[i] = match_fwd(grid0)[1].mean(axis=1)
    P["angles"] = angles[np.argmax(table, axis=0)]
    m, v = {k: np.zeros_like(P[k]) for k in P}, {k: np.zeros_like(P[k]) for k in P}
    msnaps, step = [], 0
    for epoch in range(4):
        if epoch % 2 == 0:
            msnaps.append(match_fwd(grid0))
        for g in dm.train_dataloader():
            out = O.match_step(W, b, Bm, auxB, g[:, 0].numpy(), coords, res, [banks[t] for t in targets], mei_acts, P, shifts, None, -1.0, 1.0, std=0.1)
            step += 1
            for k in P:
                P[k], m[k], v[k] = O.adam_step(P[k], out["d_" + k].astype(np.float32), m[k], v[k], step)
    msnaps.append(match_fwd(grid0))
    assert torch.equal(torch.get_rng_state(), torch.from_numpy(d["rng_after_match"]))
    assert d["match_imgs"].shape == (3, 2, 20, 1, 20, 20) and d["match_acts"].shape == (3, 2, 20)
    for k, (zs, acts) in enumerate(msnaps):
        assert relerr(zs.reshape(2, 20, 1, 20, 20), d["match_imgs"][k]) < 1e-2 and relerr(acts, d["match_acts"][k]) < 1e-2, k
    for k in P:
        np.testing.assert_allclose(P[k], d[k], rtol=1e-4, atol=5e-4, err_msg=k)


# ----------------------------------------------------------------------------------------------------------------------------------------
# round 2: the BASELINE configs at their own sizes (oracle/make_golden_r2.py)
# ----------------------------------------------------------------------------------------------------------------------------------------
def test_match_step_100():
    """One match hot step on demo1 targets [1, 2] at 100x100 (BASELINE configs[1]); images stored every `sub`-th pixel."""
    d = golden("match_step_100")
    W, b = net_params(d)
    res, sub = (100, 100), int(d["sub"])
    banks = O.demo1_banks(list(res))
    shifts = O.coords_shifts(d["template_xy"], d["target_xy"], res)
    aff = dict(angles=d["angles"], scalings=d["scalings"], shears=d["shears"], translation=d["translation"])
    out = O.match_step(W, b, d["B"], d["auxB"], d["grid"], d["coords"], res, [banks[k] for k in d["targets"]], d["mei_acts"].astype(np.float32), aff,
                       shifts, 1.0, -1.0, 1.0)
    assert abs(out["loss"] - d["loss"]) < 2e-5 * abs(d["loss"])
    assert relerr(out["acts_dn"], d["acts_dn"]) < 1e-5
    assert relerr(out["img_pre"].reshape(2, 20, -1)[:, :, ::sub], d["img_pre"]) < TOL
    assert relerr(out["img_post"].reshape(2, 20, -1)[:, :, ::sub], d["img_post"]) < TOL
    for k in ("d_angles", "d_scalings", "d_shears", "d_translation"):
        assert relerr(out[k], d[k]) < 1e-3, k


def _cosine(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


def test_learn_step_late_in_the_trajectory():
    """A learn hot step from the network the reference's own loop reaches after 600 Adam steps at 100x100 (mean activation 0.91): near
    convergence the parameter gradient is a small residual of the activation and the contrastive term (plain bf16 operands flip its
    sign there), so this is where arithmetic precision shows.  fp32 oracle vs the fp32 reference: loss 2e-5, gradient rel-L2 / cosine as
    asserted; the fp64 oracle shows what remains is the reference's own fp32 rounding."""
    d = golden("learn_late_100")
    assert int(d["steps"]) == 600 and 0.85 < float(d["acts"].mean()) < 0.95
    W, b = net_params(d)
    res, sub = (100, 100), int(d["sub"])
    pos, neg = O.pos_neg_masks(20, 0.1, True)
    bank = O.demo1_banks(list(res))[0]
    out = O.learn_step(W, b, d["B"], d["auxB"], d["grid"], d["coords"], res, bank, d["mask"], float(d["mei_act"]), 2.0, 1.0, -1.0, 1.0, pos, neg, 0.3)
    assert abs(out["loss"] - d["loss"]) < 2e-5 * abs(d["loss"])
    assert abs(out["sim"] - d["sim"]) < 5e-5 * abs(d["sim"])
    assert relerr(out["acts"], d["acts"]) < 1e-5
    assert relerr(out["img_post"].reshape(20, -1)[:, ::sub], d["img_post"]) < TOL
    got = np.concatenate([np.concatenate([out["dW"][l].ravel(), out["db"][l].ravel()]) for l in range(5)])
    want = np.concatenate([np.concatenate([d[f"dW{l}"].ravel(), d[f"db{l}"].ravel()]) for l in range(5)])
    assert relerr(got, want) < 5e-3 and _cosine(got, want) > 0.99998
    f64 = lambda a: np.asarray(a, np.float64)
    p64, n64 = O.pos_neg_masks(20, 0.1, True, np.float64)
    o64 = O.learn_step([f64(w) for w in W], [f64(x) for x in b], f64(d["B"]), f64(d["auxB"]), f64(d["grid"]), f64(d["coords"]), res, f64(bank),
                       f64(d["mask"]), float(d["mei_act"]), 2.0, 1.0, -1.0, 1.0, p64, n64, 0.3)
    g64 = np.concatenate([np.concatenate([o64["dW"][l].ravel(), o64["db"][l].ravel()]) for l in range(5)])
    assert relerr(g64, want) < 5e-3 and _cosine(g64, want) > 0.99998


def _conv_response_torch(arrays, neuron, prefix):
    """Response + image gradient of one neuron of the conv encoder through the torch-CPU port (full-frame conv2d: seconds at 200x200, where the numpy
    full-frame oracle needs minutes; the port is pinned against the reference in test_torch_port_conv_response_matches_reference)."""
    import torch
    from oracle.torch_port import ConvResponsePort
    port = ConvResponsePort(arrays, neuron, prefix=prefix)

    def response(z, g_act):
        x = torch.from_numpy(np.ascontiguousarray(z, np.float32)).requires_grad_(True)
        act = port(x).reshape(-1)
        (act * torch.from_numpy(np.asarray(g_act, np.float32).reshape(-1))).sum().backward()
        return act.detach().numpy(), x.grad.numpy()
    return response


def test_conv_learn_and_match_step_200():
    """BASELINE configs[3] at full size: random-init Stacked2dCore(1->32, kernels 9/7/7) + FullGaussian2d(128) at 200x200, one learn hot step and one
    match hot step of the reference; numpy INR / constraint / contrastive oracle around the torch-port response."""
    d = golden("conv_learn_match_200")
    W, b = net_params(d)
    res, sub = (200, 200), int(d["sub"])
    assert d["m_w0"].shape == (32, 1, 9, 9) and d["m_w2"].shape == (32, 32, 7, 7) and d["m_mu"].shape == (128, 2)
    pos, neg = O.pos_neg_masks(20, 0.1, True)
    out = O.learn_step(W, b, d["B"], d["auxB"], d["grid"], d["coords"], res, None, d["mask"], float(d["mei_act"]), 2.0, float(d["norm"]), -1.0, 1.0,
                       pos, neg, 0.3, response=_conv_response_torch(d, 0, "m_"))
    assert abs(out["loss"] - d["loss"]) < 2e-5 * abs(d["loss"])
    assert relerr(out["acts"], d["acts"]) < 1e-5
    assert relerr(out["img_post"].reshape(20, -1)[:, ::sub], d["img_post"]) < TOL
    for l in range(5):
        assert relerr(out["dW"][l], d[f"dW{l}"]) < 2e-3, l
        assert relerr(out["db"][l], d[f"db{l}"]) < 2e-3, l
    shifts = O.coords_shifts(d["template_xy"], d["target_xy"], res)
    aff = dict(angles=d["angles"], scalings=d["scalings"], shears=d["shears"], translation=d["translation"])
    out = O.match_step(W, b, d["B"], d["auxB"], d["grid"], d["coords"], res, None, d["mei_acts"].astype(np.float32), aff, shifts, float(d["norm"]), -1.0, 1.0,
                       responses=[_conv_response_torch(d, int(k), "m_") for k in d["targets"]])
    assert abs(out["loss"] - d["match_loss"]) < 2e-5 * abs(d["match_loss"])
    assert relerr(out["acts_dn"], d["match_acts_dn"]) < 1e-5
    assert relerr(out["img_post"].reshape(2, 20, -1)[:, :, ::sub], d["match_img_post"]) < TOL
    for k in ("d_angles", "d_scalings", "d_shears", "d_translation"):
        assert relerr(out[k], d[k]) < 2e-3, k


def test_config3_sample_banks_and_meis():
    """The 8-neuron sample of the 1024-neuron synthetic population (BASELINE configs[2] / [4]): the product's population builder reproduces the
    reference generators' filter banks bit for bit (sha256), `only=` selects rows of the SAME population, and the oracle's MEI ascent reaches the
    reference's MEI activations and images for two of the neurons."""
    import hashlib
    import laminr_b200 as L
    d = golden("config3_sample_100")
    sample = [int(i) for i in d["sample"]]
    pop = L.neuron_models.simulated_population(int(d["n_pop"]), img_res=[100, 100], seed=0, only=sample)
    banks = pop.packed_filters.reshape(len(sample), 20, 100, 100).numpy()
    for k in range(len(sample)):
        assert hashlib.sha256(np.ascontiguousarray(banks[k])).hexdigest() == str(d["bank_sha"][k]), sample[k]
    # `only` keeps the random stream of the full population: the first three neurons equal those of a smaller `only` list and of a full 3-neuron... no:
    # the locations are drawn AFTER all n matrices, so only the same n gives the same neurons
    sub = L.neuron_models.simulated_population(int(d["n_pop"]), img_res=[100, 100], seed=0, only=sample[3:5])
    assert np.array_equal(sub.packed_filters.numpy(), pop.packed_filters[3:5].numpy())
    import torch
    torch.manual_seed(0)
    torch.randn(1, 1, 100, 100)  # neuron_idxs=None: get_mei_dict first draws one image to count the model's outputs (mei.py:92-95)
    x0 = [torch.randn(1, 1, 100, 100).numpy() for _ in sample]  # ... then one starting image per neuron, in order (mei.py:33)
    for k in (0, 5):
        x, fev = O.mei_ascent(x0[k], banks[k], 10.0, 1000, -1.0, 1.0, norm=1.0)
        assert abs(fev[-1] - float(d["activations"][k])) < 2e-6
        assert relerr(x.reshape(100, 100), d["meis"][k].astype(np.float32).reshape(100, 100)) < 3e-3  # float16 fixture
