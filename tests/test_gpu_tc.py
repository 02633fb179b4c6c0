"""GPU parity of the tcgen05 tensor-core INR path (split-bf16 x3, fp32 accumulate in TMEM) against the golden vectors of
the unmodified reference.  Tolerances are stated per precision mode:
  bf16x3       ex2/rcp tanh, ~2^-16 relative operand error  -> fp32-class: images 1e-4, gradients 2e-3 (same as the fp32 path)
  bf16x3_fast  MUFU tanh.approx.f32 (2^-11)                 -> images 2e-3, per-step loss 1e-3 (north star), gradients 3e-2
"""
import numpy as np
import pytest
import torch

from oracle import laminr_oracle as O
from tests._util import golden, net_params, relerr, flat_params, split_flat

pytestmark = pytest.mark.gpu

IMG_TOL = {"bf16x3": 1e-4, "bf16x3_fast": 2e-3}
GRAD_TOL = {"bf16x3": 2e-3, "bf16x3_fast": 3e-2}
MODES = ["bf16x3", "bf16x3_fast"]


@pytest.fixture(scope="module")
def ops():
    from laminr_b200 import ops as _ops, _lib
    assert _lib.load().lmr_check_device() == 0
    return _ops


def dev(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).cuda()


def desc_for(ops, d, C):
    return ops.make_desc(50, 4, d["B"].shape[1], d["auxB"].shape[1], C)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name,res,C", [("inr_12x10", (12, 10), 1), ("inr_9x16_c2", (9, 16), 2)])
def test_tc_forward_backward_golden(ops, name, res, C, mode):
    d = golden(name)
    desc = desc_for(ops, d, C)
    params = dev(flat_params(d))
    N = len(d["grid"])
    args = (desc, params, dev(d["B"]), dev(d["auxB"]), dev(d["grid"]), dev(d["coords"]), None, None)
    img = ops.raw_inr_forward(*args, precision=mode)
    assert relerr(img.cpu().numpy().reshape(N, C, *res), d["out"]) < IMG_TOL[mode]
    g_params, _ = ops.raw_inr_backward(*args, dev(d["upstream"]).reshape(1, N, C, -1), precision=mode)
    dW, db = split_flat(g_params.cpu().numpy(), d)
    for l in range(5):
        assert relerr(dW[l], d[f"dW{l}"]) < GRAD_TOL[mode], (l, relerr(dW[l], d[f"dW{l}"]))
        assert relerr(db[l], d[f"db{l}"]) < GRAD_TOL[mode], (l, relerr(db[l], d[f"db{l}"]))


@pytest.mark.parametrize("mode", MODES)
def test_tc_many_latents_and_match(ops, mode):
    d = golden("inr_12x10")
    W, b = net_params(d)
    grid = np.linspace(0, 2 * np.pi, 150, endpoint=False).astype(np.float32)
    want = O.inr_forward(W, b, d["B"], d["auxB"], grid, d["coords"][None], (12, 10))
    img = ops.raw_inr_forward(desc_for(ops, d, 1), dev(flat_params(d)), dev(d["B"]), dev(d["auxB"]), dev(grid), dev(d["coords"]), None, None, precision=mode)
    assert relerr(img.cpu().numpy().reshape(want.shape), want) < IMG_TOL[mode]
    d = golden("match_inr_14x14")
    aff = ops.AffineArgs(dev(d["angles"]), dev(d["scalings"]), dev(d["shears"]), dev(d["translation"]), tc=dev(d["tc"]), shift_to=dev(d["shift_to"]),
                         shift_from=dev(d["shift_from"]), clamp=True)
    args = (desc_for(ops, d, 1), dev(flat_params(d)), dev(d["B"]), dev(d["auxB"]), dev(d["grid"]), dev(d["coords"]), None, aff)
    img = ops.raw_inr_forward(*args, precision=mode)
    assert relerr(img.cpu().numpy().reshape(d["out"].shape), d["out"]) < IMG_TOL[mode]
    D, N = d["out"].shape[:2]
    g_params, g_aff = ops.raw_inr_backward(*args, dev(d["upstream"]).reshape(D, N, 1, -1), want_params=True, want_affine=True, precision=mode)
    for got, key in zip(g_aff, ("d_angles", "d_scalings", "d_shears", "d_translation")):
        assert relerr(got.cpu().numpy().reshape(d[key].shape), d[key]) < GRAD_TOL[mode], (key, relerr(got.cpu().numpy().reshape(d[key].shape), d[key]))
    _, g_aff2 = ops.raw_inr_backward(*args, dev(d["upstream"]).reshape(D, N, 1, -1), want_params=False, want_affine=True, precision=mode)
    for a, b_ in zip(g_aff, g_aff2):
        assert torch.equal(a, b_)  # the affine gradient does not depend on whether weight gradients were requested


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("res", [20, 100])
def test_tc_learn_step_golden(ops, res, mode):
    from tests.test_gpu_ops import _learn_step_gpu
    import laminr_b200.ops as lops
    d = golden(f"learn_step_{res}")
    lops.set_default_precision(mode)
    try:
        out = _learn_step_gpu(ops, d, res)
    finally:
        lops.set_default_precision("fp32")
    assert abs(out["loss"] - d["loss"]) < 1e-3 * abs(d["loss"])  # north-star per-step tolerance, both modes
    if mode == "bf16x3":
        assert abs(out["loss"] - d["loss"]) < 5e-5 * abs(d["loss"])
    sub = slice(None) if res == 20 else slice(None, None, 37)
    assert relerr(out["img_post"].reshape(20, -1)[:, sub], d["img_post"]) < IMG_TOL[mode]
    dW, db = split_flat(out["g_params"], d)
    tol = 5e-3 if mode == "bf16x3" else 5e-2  # the parameter gradient is a small residual of competing terms
    for l in range(5):
        assert relerr(dW[l], d[f"dW{l}"]) < tol, (l, relerr(dW[l], d[f"dW{l}"]))


@pytest.mark.parametrize("mode", MODES)
def test_tc_learn_trajectory(ops, mode):
    """30 Adam steps with the tensor-core INR: per-step loss within 1e-3 relative of the reference, final parameters within 1e-2."""
    import laminr_b200 as L
    from laminr_b200.contrastive_loss import SimCLROnGrid
    from laminr_b200.fused import LearnStepper
    from tests.test_gpu_api import _manifold
    d = golden("learn_traj_20")
    model, im = _manifold(L, 20, d, prefix="init_", precision=mode)
    reg = SimCLROnGrid(1, 20, 0.1, 0.3, True).cuda()
    st = LearnStepper(im.template, im.img_transforms, model, 0, d["mask"], 1.0000009536743164, reg, -1.0, 1.0, 20, lr=1e-3, reg_scale=2.0,
                      precision=mode, use_graph=True)
    grids = torch.from_numpy(d["grids"]).cuda()
    losses = []
    for i in range(30):
        st.step(grids[i])
        losses.append(st.loss().item())
    np.testing.assert_allclose(losses, d["losses"], rtol=1e-3)
    dW, db = split_flat(st.flat.cpu().numpy(), d, prefix="final_")
    for l in range(5):
        assert relerr(dW[l], d[f"final_W{l}"]) < 1e-2, (l, relerr(dW[l], d[f"final_W{l}"]))


@pytest.mark.parametrize("mode", MODES)
def test_tc_kept_activations_backward(ops, mode):
    """LMR_PREC_KEEP_ACTIVATIONS: the forward keeps the operand tiles of h_1..h_3 in its workspace and the backward loads them (bulk async copies)
    instead of recomputing the forward.  Same gradients as the recompute path and as the reference: learn (D = 1, weight gradients, 1 and 2
    channels, several tiles per CTA) and match (D = 2, affine gradients, with and without weight gradients)."""
    for name, res, C in (("inr_12x10", (12, 10), 1), ("inr_9x16_c2", (9, 16), 2)):
        d = golden(name)
        desc = desc_for(ops, d, C)
        args = (desc, dev(flat_params(d)), dev(d["B"]), dev(d["auxB"]), dev(d["grid"]), dev(d["coords"]), None, None)
        N, P = len(d["grid"]), res[0] * res[1]
        ws = torch.empty(ops.forward_workspace_bytes(desc, N, P, 1, True), dtype=torch.uint8, device="cuda")
        img = ops.raw_inr_forward(*args, precision=mode, workspace=ws, keep_activations=True)
        assert relerr(img.cpu().numpy().reshape(N, C, *res), d["out"]) < IMG_TOL[mode]
        up = dev(d["upstream"]).reshape(1, N, C, -1)
        g_keep, _ = ops.raw_inr_backward(*args, up, precision=mode, forward_workspace=ws, keep_activations=True)
        g_rec, _ = ops.raw_inr_backward(*args, up, precision=mode)
        assert relerr(g_keep.cpu().numpy(), g_rec.cpu().numpy()) < 1e-5
        gw, gb = split_flat(g_keep.cpu().numpy(), d)
        for l in range(5):
            assert relerr(gw[l], d[f"dW{l}"]) < GRAD_TOL[mode] and relerr(gb[l], d[f"db{l}"]) < GRAD_TOL[mode], l
    # a bigger image set: 45 x 41 pixels x 20 latents = 308 tiles (2-3 per CTA: the rolling prefetch of the next tile's activations is exercised)
    d = golden("inr_12x10")
    rng = np.random.RandomState(11)
    res = (45, 41)
    yy, xx = np.meshgrid(np.linspace(-1, 1, res[0]), np.linspace(-1, 1, res[1]), indexing="ij")
    coords = dev(np.stack([yy, xx], -1).reshape(-1, 2).astype(np.float32))
    grid = dev(np.linspace(0.1, 6.2, 20).astype(np.float32))
    desc = desc_for(ops, d, 1)
    args = (desc, dev(flat_params(d)), dev(d["B"]), dev(d["auxB"]), grid, coords, None, None)
    P = res[0] * res[1]
    ws = torch.empty(ops.forward_workspace_bytes(desc, 20, P, 1, True), dtype=torch.uint8, device="cuda")
    up = dev(rng.randn(1, 20, 1, P).astype(np.float32))
    for _ in range(2):  # twice on the same workspace
        img_k = ops.raw_inr_forward(*args, precision=mode, workspace=ws, keep_activations=True)
        g_keep, _ = ops.raw_inr_backward(*args, up, precision=mode, forward_workspace=ws, keep_activations=True)
    img_r = ops.raw_inr_forward(*args, precision=mode)
    g_rec, _ = ops.raw_inr_backward(*args, up, precision=mode)
    assert torch.equal(img_k, img_r) and relerr(g_keep.cpu().numpy(), g_rec.cpu().numpy()) < 1e-5
    # match: two targets with an affine transform
    d = golden("match_inr_14x14")
    aff = ops.AffineArgs(dev(d["angles"]), dev(d["scalings"]), dev(d["shears"]), dev(d["translation"]), tc=dev(d["tc"]), shift_to=dev(d["shift_to"]),
                         shift_from=dev(d["shift_from"]), clamp=True)
    desc = desc_for(ops, d, 1)
    args = (desc, dev(flat_params(d)), dev(d["B"]), dev(d["auxB"]), dev(d["grid"]), dev(d["coords"]), None, aff)
    D, N = d["out"].shape[:2]
    ws = torch.empty(ops.forward_workspace_bytes(desc, N, d["coords"].shape[0], D, True), dtype=torch.uint8, device="cuda")
    img = ops.raw_inr_forward(*args, precision=mode, workspace=ws, keep_activations=True)
    assert relerr(img.cpu().numpy().reshape(d["out"].shape), d["out"]) < IMG_TOL[mode]
    up = dev(d["upstream"]).reshape(D, N, 1, -1)
    for want_params in (True, False):
        _, g_aff = ops.raw_inr_backward(*args, up, want_params=want_params, want_affine=True, precision=mode, forward_workspace=ws, keep_activations=True)
        for got, key in zip(g_aff, ("d_angles", "d_scalings", "d_shears", "d_translation")):
            assert relerr(got.cpu().numpy().reshape(d[key].shape), d[key]) < GRAD_TOL[mode], key


def test_two_stream_backward_variant_parity():
    """The two-stream 64-row backward kernel (development switch LMR_BWD_MODE=2, csrc/inr_tc.cu inr_bwd2_kernel) passes the same parity tests as the
    default one-stream kernel.  The switch is read once per process, so the tests of this file are re-run in a child process."""
    import os
    import subprocess
    import sys

    if os.environ.get("LMR_BWD_MODE"):
        pytest.skip("already running under a backward-variant switch")
    env = dict(os.environ, LMR_BWD_MODE="2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"], env=env, cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("res", [(12, 10), (100, 100)])
def test_reused_coordinate_table_is_bit_identical(ops, mode, res):
    """LMR_PREC_REUSE_COORD_TABLE: the learn loop's second and later forwards take the pixel grid's sin/cos table from the workspace of the previous
    call instead of evaluating it again (the grid never changes; parameters and latents do): same images, bit for bit, after a parameter change."""
    d = golden("inr_12x10")
    desc = desc_for(ops, d, 1)
    P = res[0] * res[1]
    coords = dev(O.pixel_coords(res))
    ws = torch.empty(ops.forward_workspace_bytes(desc, len(d["grid"]), P, 1, True), dtype=torch.uint8, device="cuda")
    p0 = dev(flat_params(d))
    g = torch.Generator(device="cuda").manual_seed(3)
    p1 = p0 + 0.01 * torch.randn(p0.shape, device="cuda", generator=g)
    grid1 = dev(d["grid"]) + 0.05
    fwd = lambda p, gr, **kw: ops.raw_inr_forward(desc, p, dev(d["B"]), dev(d["auxB"]), gr, coords, None, None, precision=mode, **kw)
    want = fwd(p1, grid1).clone()                                  # fresh workspace: evaluates the table
    fwd(p0, dev(d["grid"]), workspace=ws, keep_activations=True)   # first call on `ws` writes the table
    got = fwd(p1, grid1, workspace=ws, keep_activations=True, reuse_coord_table=True)
    assert torch.equal(got, want)
    got2 = fwd(p1, grid1, workspace=ws, reuse_coord_table=True)    # the evaluation forward (no kept tiles) shares the table
    assert torch.equal(got2, want)
