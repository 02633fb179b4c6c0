"""GPU parity: every C-ABI kernel (through laminr_b200.ops -> ctypes -> liblaminr_b200.so) against the numpy oracle
and the golden vectors of the unmodified reference.  fp32 path; tolerances are float32 summation-order noise."""
import numpy as np
import pytest
import torch

from oracle import laminr_oracle as O
from tests._util import golden, net_params, relerr, flat_params, split_flat

pytestmark = pytest.mark.gpu

TOL = 3e-5


@pytest.fixture(scope="module")
def ops():
    from laminr_b200 import ops as _ops, _lib
    assert _lib.load().lmr_check_device() == 0, _lib.load().lmr_last_error()
    return _ops


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def desc_for(ops, d, C):
    return ops.make_desc(50, 4, d["B"].shape[1], d["auxB"].shape[1], C)


@pytest.mark.parametrize("name,res,C", [("inr_12x10", (12, 10), 1), ("inr_9x16_c2", (9, 16), 2)])
def test_inr_forward_backward_golden(ops, name, res, C):
    d = golden(name)
    desc = desc_for(ops, d, C)
    params = dev(flat_params(d))
    img = ops.raw_inr_forward(desc, params, dev(d["B"]), dev(d["auxB"]), dev(d["grid"]), dev(d["coords"]), None, None)
    N = len(d["grid"])
    got = img.cpu().numpy().reshape(N, C, *res)
    assert relerr(got, d["out"]) < TOL
    g_params, _ = ops.raw_inr_backward(desc, params, dev(d["B"]), dev(d["auxB"]), dev(d["grid"]), dev(d["coords"]), None, None,
                                       dev(d["upstream"]).reshape(1, N, C, -1))
    dW, db = split_flat(g_params.cpu().numpy(), d)
    for l in range(5):
        ref = d[f"dW{l}"].copy()
        assert relerr(dW[l], ref) < 1e-4, (l, relerr(dW[l], ref))
        assert relerr(db[l], d[f"db{l}"]) < 1e-4, l
    assert np.all(dW[0][:, 0] == 0)  # template input is the constant 0


def test_inr_forward_many_latents(ops):
    """N > 128 latents (get_images_from_learned_template(n_images=...)): several latent chunks per pixel."""
    d = golden("inr_12x10")
    W, b = net_params(d)
    grid = np.linspace(0, 2 * np.pi, 150, endpoint=False).astype(np.float32)
    want = O.inr_forward(W, b, d["B"], d["auxB"], grid, d["coords"][None], (12, 10))
    img = ops.raw_inr_forward(desc_for(ops, d, 1), dev(flat_params(d)), dev(d["B"]), dev(d["auxB"]), dev(grid), dev(d["coords"]), None, None)
    assert relerr(img.cpu().numpy().reshape(want.shape), want) < TOL


def _affine(ops, d):
    return ops.AffineArgs(dev(d["angles"]), dev(d["scalings"]), dev(d["shears"]), dev(d["translation"]), tc=dev(d["tc"]),
                          shift_to=dev(d["shift_to"]), shift_from=dev(d["shift_from"]), clamp=True)


def test_match_inr_golden(ops):
    d = golden("match_inr_14x14")
    desc = desc_for(ops, d, 1)
    params = dev(flat_params(d))
    aff = _affine(ops, d)
    args = (desc, params, dev(d["B"]), dev(d["auxB"]), dev(d["grid"]), dev(d["coords"]), None, aff)
    img = ops.raw_inr_forward(*args)
    assert relerr(img.cpu().numpy().reshape(d["out"].shape), d["out"]) < TOL
    D, N = d["out"].shape[:2]
    g_params, g_aff = ops.raw_inr_backward(*args, dev(d["upstream"]).reshape(D, N, 1, -1), want_params=True, want_affine=True)
    for got, key in zip(g_aff, ("d_angles", "d_scalings", "d_shears", "d_translation")):
        assert relerr(got.cpu().numpy().reshape(d[key].shape), d[key]) < 3e-4, key
    # weight gradients in match mode against the oracle (the reference accumulates them too, inv_temp_learn_match.py:365)
    W, b = net_params(d)
    M, _ = O.affine_matrices(d["angles"], d["scalings"], d["shears"])
    cdp = O.match_coords(d["coords"], M, d["translation"], d["tc"], d["shift_to"], d["shift_from"])
    _, cache = O.inr_forward(W, b, d["B"], d["auxB"], d["grid"], cdp, (14, 14), keep=True)
    want = O.inr_backward(W, d["B"], cache, d["upstream"])
    dW, db = split_flat(g_params.cpu().numpy(), d)
    for l in range(5):
        assert relerr(dW[l], want["dW"][l]) < 1e-4, l
        assert relerr(db[l], want["db"][l]) < 1e-4, l


@pytest.mark.parametrize("tag,mode,target", [("norm", 0, 3.0), ("std", 1, 0.4)])
def test_constrain_golden(ops, tag, mode, target):
    d = golden("constrain")
    x = dev(d[f"{tag}_x"])
    z, stats = ops.raw_constrain_fwd(mode, x, target, 0.25, -0.5, 1.0, 1e-12)
    assert relerr(z.cpu().numpy(), d[f"{tag}_y"]) < 2e-6
    gx = ops.raw_constrain_bwd(mode, x, dev(d[f"{tag}_up"]), stats, target, 0.25, -0.5, 1.0, 1e-12)
    assert relerr(gx.cpu().numpy(), d[f"{tag}_gx"]) < TOL


@pytest.mark.parametrize("n", [7, 2049, 40000])
def test_constrain_sizes_vs_oracle(ops, n):
    """chunk boundaries (2048 elements per CTA) and the 200x200 image size; one-sided / absent bounds."""
    rng = np.random.RandomState(n)
    x = (rng.randn(3, n) * 0.3 + 0.1).astype(np.float32)
    for pmin, pmax in ((-0.2, 0.4), (None, 0.3), (None, None)):
        for mode, fwd, bwd in ((0, O.constrain_norm_fwd, O.constrain_norm_bwd), (1, O.constrain_std_fwd, O.constrain_std_bwd)):
            want, cache = fwd(x, 2.0, 0.05, pmin, pmax)
            z, stats = ops.raw_constrain_fwd(mode, dev(x), 2.0, 0.05, pmin, pmax, 1e-12)
            assert relerr(z.cpu().numpy(), want) < 1e-5
            up = rng.randn(*x.shape).astype(np.float32)
            gx = ops.raw_constrain_bwd(mode, dev(x), dev(up), stats, 2.0, 0.05, pmin, pmax, 1e-12)
            want_g = bwd(up, cache, 2.0, pmin, pmax)
            assert relerr(gx.cpu().numpy(), want_g) < 2e-4 or np.abs(gx.cpu().numpy() - want_g).max() < 1e-5


def _pack_banks(banks):
    Fmax = max(len(b) for b in banks)
    packed = np.stack([np.concatenate([b, np.repeat(b[:1], Fmax - len(b), 0)]) for b in banks]).astype(np.float32)
    return packed.reshape(len(banks), Fmax, -1), np.array([len(b) for b in banks], np.int32)


@pytest.mark.parametrize("res", [16, 24])
def test_simresp_golden(ops, res):
    d = golden("simresp")
    banks = [d[f"bank{i}_{res}"] for i in range(3)]
    packed, nfilt = _pack_banks(banks)
    x = dev(d[f"x_{res}"])
    Bn = x.shape[0]
    nog = dev(np.arange(3), torch.int32)
    act, rmax, argmax = ops.raw_simresp_fwd(x, 0, Bn, 3, 1, res * res, dev(packed), dev(nfilt, torch.int32), nog, 1e-6)
    assert relerr(act.cpu().numpy().T, d[f"y_{res}"]) < 2e-6
    g_img = torch.zeros_like(x)
    ops.raw_simresp_bwd(dev(np.ascontiguousarray(d[f"up_{res}"].T)), rmax, argmax, 0, Bn, 3, 1, res * res, dev(packed), nog, g_img)
    assert relerr(g_img.cpu().numpy(), d[f"gx_{res}"]) < 1e-5


def test_simresp_two_channels_and_own_neuron(ops):
    d = golden("simresp")
    banks = [d[f"bank{i}_16"] for i in range(3)]
    packed, nfilt = _pack_banks(banks)
    x = dev(d["x_c2"])
    nog = dev(np.arange(3), torch.int32)
    act, rmax, argmax = ops.raw_simresp_fwd(x, 0, 3, 3, 2, 256, dev(packed), dev(nfilt, torch.int32), nog, 1e-6)
    assert relerr(act.cpu().numpy().T, d["y_c2"]) < 2e-6
    g_img = torch.empty_like(x)
    ops.raw_simresp_bwd(torch.ones_like(act), rmax, argmax, 0, 3, 3, 2, 256, dev(packed), nog, g_img)
    assert relerr(g_img.cpu().numpy(), d["gx_c2"]) < 1e-5
    # own-neuron grouping (match / MEI): group g sees its own block of images and only neuron order[g]
    order = np.array([2, 0, 1], np.int32)
    xs = np.stack([d["x_16"][:4] * (1 + 0.1 * g) for g in range(3)])  # [G=3, NB=4, 1, 16, 16]
    act, rmax, argmax = ops.raw_simresp_fwd(dev(xs), 4 * 256, 4, 3, 1, 256, dev(packed), dev(nfilt, torch.int32), dev(order, torch.int32), 1e-6)
    up = np.random.RandomState(0).randn(3, 4).astype(np.float32)
    g_img = torch.empty(xs.shape, device="cuda")
    ops.raw_simresp_bwd(dev(up), rmax, argmax, 4 * 256, 4, 3, 1, 256, dev(packed), dev(order, torch.int32), g_img)
    for g in range(3):
        want, cache = O.simresp_fwd(xs[g], banks[order[g]])
        assert relerr(act[g].cpu().numpy(), want) < 2e-6
        assert relerr(g_img[g].cpu().numpy(), O.simresp_bwd(up[g], cache, xs[g].shape, banks[order[g]])) < 1e-5


@pytest.mark.parametrize("tag,n", [("p20", 20), ("np7", 7), ("p10", 10)])
def test_simclr_golden(ops, tag, n):
    d = golden("simclr")
    pmin, pmax = map(float, d[f"{tag}_pm"])
    x = dev(d[f"{tag}_x"])
    reg, K = ops.raw_simclr_fwd(x, dev(d[f"{tag}_mask"]), n, 1, 64, pmin, pmax, dev(d[f"{tag}_pos"]), dev(d[f"{tag}_neg"]), 0.3, 1e-6)
    assert abs(reg.item() - d[f"{tag}_reg"]) < 2e-6 * max(1.0, abs(d[f"{tag}_reg"]))
    g = torch.empty_like(x)
    ops.raw_simclr_bwd(x, dev(d[f"{tag}_mask"]), K, torch.ones(1, device="cuda"), 1.0, n, 1, 64, pmin, pmax, g)
    assert relerr(g.cpu().numpy(), d[f"{tag}_gx"]) < 1e-4, relerr(g.cpu().numpy(), d[f"{tag}_gx"])


def test_adam_golden(ops):
    d = golden("adam")
    p = dev(d["p0"])
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(2, dtype=torch.int32, device="cuda")
    for i, g in enumerate(d["grads"]):
        ops.raw_adam_step(p, dev(g), m, v, step)
        np.testing.assert_allclose(p.cpu().numpy(), d["ps"][i], rtol=0, atol=3e-7)
    assert step.tolist() == [5, 0]


def test_adam_multi_equals_separate_steps(ops):
    """lmr_adam_step_multi (the match step's one launch over angles / scalings / shears / translation) == lmr_adam_step per tensor, bit for bit:
    parameters, both moments and every group's own step word, over several steps and ragged sizes (one group spans several blocks)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    sizes = [2, 4, 4, 4, 1023, 700]
    mk = lambda: [torch.randn(n, device="cuda", generator=g) for n in sizes]
    p_a = mk()
    p_b = [x.clone() for x in p_a]
    st_a = [dict(m=torch.zeros_like(x), v=torch.zeros_like(x), s=torch.zeros(2, dtype=torch.int32, device="cuda")) for x in p_a]
    st_b = [dict(m=torch.zeros_like(x), v=torch.zeros_like(x), s=torch.zeros(2, dtype=torch.int32, device="cuda")) for x in p_b]
    st_b[2]["s"][0] = 7  # groups keep their own step counts
    st_a[2]["s"][0] = 7
    for it in range(4):
        grads = mk()
        for x, gr, st in zip(p_a, grads, st_a):
            ops.raw_adam_step(x, gr, st["m"], st["v"], st["s"], lr=1e-2)
        ops.raw_adam_step_multi([(x, gr, st["m"], st["v"], st["s"]) for x, gr, st in zip(p_b, grads, st_b)], lr=1e-2)
    for x, y, sa, sb in zip(p_a, p_b, st_a, st_b):
        assert torch.equal(x, y) and torch.equal(sa["m"], sb["m"]) and torch.equal(sa["v"], sb["v"])
        assert sa["s"].tolist() == sb["s"].tolist() and sb["s"][1].item() == 0
    assert st_b[2]["s"][0].item() == 11 and st_b[0]["s"][0].item() == 4


@pytest.mark.parametrize("shape", [(3, 20, 1, 10, 10), (40, 20, 1, 100, 101), (7, 5, 3, 333, 77)])
def test_to_host_numpy_equals_cpu_copy(ops, shape):
    """The pipelined device-to-host copy of result blocks (pinned ring + worker threads) returns exactly `t.cpu().numpy()`; small tensors take the
    plain path, the larger ones cross several chunks with a ragged tail; a non-contiguous view is handled too."""
    x = torch.randn(*shape, device="cuda")
    for t in (x, x.transpose(0, 1)):
        got = ops.to_host_numpy(t, chunk_bytes=1 << 20)
        assert got.shape == tuple(t.shape) and got.dtype == np.float32
        assert np.array_equal(got, t.cpu().numpy())


@pytest.mark.parametrize("D,N", [(1, 20), (2, 20), (37, 7), (1024, 20)])
def test_match_epoch_stats_vs_host_reductions(ops, D, N):
    """lmr_match_epoch_stats == the reference's host-driven reductions of the match loop (inv_temp_learn_match.py:361-362,368-371):
    acts_d = mean_n act[d,n] / mei_act_d, then their sum / min / max (float64 on the host as the yardstick)."""
    rs = np.random.RandomState(D + N)
    act = rs.rand(D, N).astype(np.float32) + 0.05
    mei = (0.5 + rs.rand(D)).astype(np.float32)
    out = torch.empty(3, device="cuda")
    ops.raw_match_epoch_stats(dev(act), dev(mei), out)
    a = act.astype(np.float64).mean(axis=1) / mei
    np.testing.assert_allclose(out.cpu().numpy(), [a.sum(), a.min(), a.max()], rtol=2e-6)


@pytest.mark.parametrize("tag,mode,target", [("norm", 0, 1.0), ("std", 1, 0.1)])
def test_mei_golden(ops, tag, mode, target):
    d = golden("mei")
    gs = golden("simresp")
    packed, nfilt = _pack_banks([gs[f"bank{i}_16"] for i in range(3)])
    x = dev(np.concatenate([d[f"{tag}_x0_{i}"] for i in range(3)]))
    fevals = ops.raw_mei_ascent(x, dev(packed), dev(nfilt, torch.int32), dev(np.arange(3), torch.int32), 10.0, 60, mode, target, 0.0, -1.0, 1.0)
    for i in range(3):
        np.testing.assert_allclose(fevals[i].cpu().numpy(), d[f"{tag}_fevals_{i}"], rtol=5e-5)
        assert relerr(x[i].cpu().numpy(), d[f"{tag}_x_{i}"][0]) < 2e-4


def _learn_step_gpu(ops, d, res, reg_scale=2.0):
    """inv_temp_learn_match.py:187-207 composed from the raw C-ABI calls."""
    N, P = 20, res * res
    desc = desc_for(ops, d, 1)
    params, B, auxB, grid, coords = dev(flat_params(d)), dev(d["B"]), dev(d["auxB"]), dev(d["grid"]), dev(d["coords"])
    bank = O.demo1_banks([res, res])[0]
    packed, nfilt = _pack_banks([bank])
    packed, nfilt, nog = dev(packed), dev(nfilt, torch.int32), torch.zeros(1, dtype=torch.int32, device="cuda")
    pos, neg = (dev(a) for a in O.pos_neg_masks(N, 0.1, True))
    mask = dev(d["mask"])
    mei_act = float(d["mei_act"])
    img_pre = ops.raw_inr_forward(desc, params, B, auxB, grid, coords, None, None).reshape(N, 1, res, res)
    z, stats = ops.raw_constrain_fwd(0, img_pre, 1.0, 0.0, -1.0, 1.0, 1e-12)
    act, rmax, argmax = ops.raw_simresp_fwd(z, 0, N, 1, 1, P, packed, nfilt, nog, 1e-6)
    reg, K = ops.raw_simclr_fwd(z, mask, N, 1, P, -1.0, 1.0, pos, neg, 0.3, 1e-6)
    acts = act[0] / mei_act
    loss = -acts.mean() - reg_scale * reg[0]
    g_z = torch.empty_like(z)
    ops.raw_simresp_bwd(torch.full((1, N), -1.0 / (N * mei_act), device="cuda"), rmax, argmax, 0, N, 1, 1, P, packed, nog, g_z)
    ops.raw_simclr_bwd(z, mask, K, torch.ones(1, device="cuda"), -reg_scale, N, 1, P, -1.0, 1.0, g_z, accumulate=True)
    g_x = ops.raw_constrain_bwd(0, img_pre, g_z, stats, 1.0, 0.0, -1.0, 1.0, 1e-12)
    g_params, _ = ops.raw_inr_backward(desc, params, B, auxB, grid, coords, None, None, g_x.reshape(1, N, 1, P))
    return dict(loss=loss.item(), sim=reg.item(), acts=acts.cpu().numpy(), img_pre=img_pre.cpu().numpy(), img_post=z.cpu().numpy(),
                g_params=g_params.cpu().numpy(), g_x=g_x.cpu().numpy())


@pytest.mark.parametrize("res", [20, 100])
def test_learn_step_golden(ops, res):
    d = golden(f"learn_step_{res}")
    out = _learn_step_gpu(ops, d, res)
    assert abs(out["loss"] - d["loss"]) < 1e-3 * abs(d["loss"])   # north-star per-step tolerance
    assert abs(out["loss"] - d["loss"]) < 3e-5 * abs(d["loss"])   # what the fp32 path achieves
    assert abs(out["sim"] - d["sim"]) < 3e-5 * abs(d["sim"])
    assert relerr(out["acts"], d["acts"]) < 1e-5
    sub = slice(None) if res == 20 else slice(None, None, 37)
    assert relerr(out["img_pre"].reshape(20, -1)[:, sub], d["img_pre"]) < TOL
    assert relerr(out["img_post"].reshape(20, -1)[:, sub], d["img_post"]) < TOL
    dW, db = split_flat(out["g_params"], d)
    for l in range(5):
        assert relerr(dW[l], d[f"dW{l}"]) < 3e-3, (l, relerr(dW[l], d[f"dW{l}"]))
        assert relerr(db[l], d[f"db{l}"]) < 3e-3, l
    # and against the fp64 oracle on the same inputs (tighter statement about the maths)
    f64 = lambda a: np.asarray(a, np.float64)
    W, b = net_params(d)
    pos, neg = O.pos_neg_masks(20, 0.1, True, np.float64)
    want = O.learn_step([f64(w) for w in W], [f64(x) for x in b], f64(d["B"]), f64(d["auxB"]), f64(d["grid"]), f64(d["coords"]), (res, res),
                        f64(O.demo1_banks([res, res])[0]), f64(d["mask"]), float(d["mei_act"]), 2.0, 1.0, -1.0, 1.0, pos, neg, 0.3)
    for l in range(5):
        assert relerr(dW[l], want["dW"][l]) < 2e-3, (l, relerr(dW[l], want["dW"][l]))


def test_match_step_golden(ops):
    d = golden("match_step_20")
    res, N, D, P = 20, 20, 2, 400
    desc = desc_for(ops, d, 1)
    params, B, auxB, grid, coords = dev(flat_params(d)), dev(d["B"]), dev(d["auxB"]), dev(d["grid"]), dev(d["coords"])
    banks = O.demo1_banks([res, res])
    packed, nfilt = _pack_banks(banks)
    packed, nfilt = dev(packed), dev(nfilt, torch.int32)
    nog = dev(d["targets"], torch.int32)
    shift_to, tc, shift_from = O.coords_shifts(d["template_xy"], d["target_xy"], (res, res))
    aff = ops.AffineArgs(dev(d["angles"]), dev(d["scalings"]), dev(d["shears"]), dev(d["translation"]), tc=dev(tc), shift_to=dev(shift_to),
                         shift_from=dev(shift_from), clamp=True)
    img_pre = ops.raw_inr_forward(desc, params, B, auxB, grid, coords, None, aff)  # [D,N,1,P]
    flat = img_pre.reshape(D * N, 1, res, res)
    z, stats = ops.raw_constrain_fwd(0, flat, 1.0, 0.0, -1.0, 1.0, 1e-12)
    act, rmax, argmax = ops.raw_simresp_fwd(z, N * P, N, D, 1, P, packed, nfilt, nog, 1e-6)
    mei_acts = dev(d["mei_acts"].astype(np.float32))
    acts = act.mean(1) / mei_acts
    loss = -acts.mean()
    assert abs(loss.item() - d["loss"]) < 3e-5 * abs(d["loss"])
    assert relerr(act.cpu().numpy(), d["acts_dn"]) < 1e-5
    assert relerr(img_pre.cpu().numpy().reshape(d["img_pre"].shape), d["img_pre"]) < TOL
    assert relerr(z.cpu().numpy().reshape(d["img_post"].shape), d["img_post"]) < TOL
    g_act = (-1.0 / (N * D * mei_acts))[:, None].expand(D, N).contiguous()
    g_z = torch.empty_like(z)
    ops.raw_simresp_bwd(g_act, rmax, argmax, N * P, N, D, 1, P, packed, nog, g_z)
    g_x = ops.raw_constrain_bwd(0, flat, g_z, stats, 1.0, 0.0, -1.0, 1.0, 1e-12)
    _, g_aff = ops.raw_inr_backward(desc, params, B, auxB, grid, coords, None, aff, g_x.reshape(D, N, 1, P), want_params=False, want_affine=True)
    for got, key in zip(g_aff, ("d_angles", "d_scalings", "d_shears", "d_translation")):
        assert relerr(got.cpu().numpy().reshape(d[key].shape), d[key]) < 2e-3, (key, relerr(got.cpu().numpy().reshape(d[key].shape), d[key]))


@pytest.mark.parametrize("res", [20, 100])
@pytest.mark.parametrize("mode,target", [(0, 1.0), (1, 0.05)])
def test_fused_learn_objective_equals_unfused_chain(ops, res, mode, target):
    """lmr_learn_objective (one persistent launch, software grid barriers) == constrain + simresp + simclr forward/backward chain; replayed
    several times on the same workspace (the barrier state must survive back-to-back launches)."""
    d = golden(f"learn_step_{res}")
    N, P = 20, res * res
    rng = np.random.RandomState(3)
    x = dev(d["img_pre"] if res == 20 else rng.randn(N, 1, res, res).astype(np.float32) * 0.3).reshape(N, 1, res, res)
    bank = O.demo1_banks([res, res])[2]  # 60 filters
    packed, nfilt = _pack_banks([bank])
    packed, nfilt, nog = dev(packed), dev(nfilt, torch.int32), torch.zeros(1, dtype=torch.int32, device="cuda")
    pos, neg = (dev(a) for a in O.pos_neg_masks(N, 0.1, True))
    mask, mei_act, reg_scale = dev(d["mask"]), 1.07, 1.6
    z, stats = ops.raw_constrain_fwd(mode, x, target, 0.0, -1.0, 1.0, 1e-12)
    act, rmax, argmax = ops.raw_simresp_fwd(z, 0, N, 1, 1, P, packed, nfilt, nog, 1e-6)
    reg, K = ops.raw_simclr_fwd(z, mask, N, 1, P, -1.0, 1.0, pos, neg, 0.3, 1e-6)
    loss = -(act[0] / mei_act).mean() - reg_scale * reg[0]
    g_z = torch.empty_like(z)
    ops.raw_simresp_bwd(torch.full((1, N), -1.0 / (N * mei_act), device="cuda"), rmax, argmax, 0, N, 1, 1, P, packed, nog, g_z)
    ops.raw_simclr_bwd(z, mask, K, torch.ones(1, device="cuda"), -reg_scale, N, 1, P, -1.0, 1.0, g_z, accumulate=True)
    g_x = ops.raw_constrain_bwd(mode, x, g_z, stats, target, 0.0, -1.0, 1.0, 1e-12)
    ws = ops.learn_objective_workspace(N, 1, P, bank.shape[0], "cuda")
    assert ws is not None
    z2, gx2 = torch.empty_like(z), torch.empty_like(z)
    act2, reg2, loss2 = torch.empty(N, device="cuda"), torch.empty(1, device="cuda"), torch.empty(1, device="cuda")
    g_reg = torch.full((1,), -reg_scale, device="cuda")
    for _ in range(3):
        gx2.zero_()
        ops.raw_learn_objective(mode, x, target, 0.0, -1.0, 1.0, 1e-12, dev(bank.reshape(len(bank), -1)), 1e-6, mei_act, mask, pos, neg, 0.3, 1e-6, g_reg,
                                z2, act2, reg2, loss2, gx2, ws)
        torch.cuda.synchronize()
        assert relerr(z2.cpu().numpy(), z.cpu().numpy()) < 1e-6
        assert relerr(act2.cpu().numpy(), act[0].cpu().numpy()) < 1e-6
        assert abs(reg2.item() - reg.item()) < 2e-5 * abs(reg.item())
        assert abs(loss2.item() - loss.item()) < 2e-5 * abs(loss.item())
        assert relerr(gx2.cpu().numpy(), g_x.cpu().numpy()) < 2e-4, relerr(gx2.cpu().numpy(), g_x.cpu().numpy())
