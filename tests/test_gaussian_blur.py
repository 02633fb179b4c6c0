"""featurevis.ops.GaussianBlur (featurevis/ops.py:378-423), the optional gradient preconditioner of get_mei_dict(gb=) and learn(gaussian_blur_sigma=).
CPU: the numpy restatement against outputs of the reference class (tests/golden/gaussian_blur.npz).  GPU: lmr_gaussian_blur through the C ABI."""
import numpy as np
import pytest
import torch

from oracle import laminr_oracle as O
from tests._util import golden, relerr

MODES = {0: "constant", 1: "reflect", 2: "replicate"}
TAGS = ["a", "b", "c", "e", "f"]


def _cfg(d, tag):
    sy, sx, trunc, mode = d[f"{tag}_cfg"]
    return (float(sy), float(sx)), int(trunc), MODES[int(mode)]


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_matches_reference(tag):
    d = golden("gaussian_blur")
    sigma, trunc, mode = _cfg(d, tag)
    assert relerr(O.gaussian_blur(d[f"{tag}_x"], sigma, trunc, mode), d[f"{tag}_y"]) < 1e-6
    assert relerr(O.gaussian_blur(d["g_x"], 1.0 + 0.25 * 2), d["g_y"]) < 1e-6  # decayed sigma at iteration 3


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_gaussian_blur_golden(tag):
    from laminr_b200.featurevis.ops import GaussianBlur

    d = golden("gaussian_blur")
    sigma, trunc, mode = _cfg(d, tag)
    x = torch.from_numpy(d[f"{tag}_x"]).cuda()
    y = GaussianBlur(sigma, truncate=trunc, pad_mode=mode)(x)
    assert y.shape == x.shape and relerr(y.cpu().numpy(), d[f"{tag}_y"]) < 1e-6
    if tag == "a":
        yg = GaussianBlur(1.0, decay_factor=0.25)(torch.from_numpy(d["g_x"]).cuda(), iteration=3)
        assert relerr(yg.cpu().numpy(), d["g_y"]) < 1e-6
        with pytest.raises(Exception, match="reflect padding"):
            GaussianBlur(5.0)(torch.zeros(1, 1, 8, 8, device="cuda"))  # padding 20 >= size 8: torch raises too
        with pytest.raises(ValueError, match="pad_mode"):
            GaussianBlur(1.0, pad_mode="circular")(x)


@pytest.mark.gpu
def test_gpu_mei_with_gradient_blur_runs_and_matches_oracle_loop():
    """generate_mei's recipe with gb= (mei.py:44-56): blurred-gradient ascent through featurevis.gradient_ascent + our operators vs the numpy loop."""
    import laminr_b200 as L
    from laminr_b200 import featurevis
    from laminr_b200.featurevis import ops as fops
    from laminr_b200.featurevis.utils import Compose
    from laminr_b200.utils.general import SingleNeuronModel

    res = 16
    model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
    bank = O.demo1_banks([res, res])[0]
    torch.manual_seed(3)
    x0 = torch.randn(1, 1, res, res)
    post = Compose([fops.ChangeNorm(norm=1.0, mean=0.0), fops.ClipRange(-1.0, 1.0)])
    xi = post(x0.cuda())
    x, fevals, _ = featurevis.gradient_ascent(SingleNeuronModel(model, 0), xi, step_size=10, num_iterations=25, post_update=post,
                                              gradient_f=fops.GaussianBlur(1.0), print_iters=1001)
    # numpy loop: f, grad (oracle), blur (oracle), SGD step, post-update
    xn = O.change_norm(x0.numpy(), 1.0, 0.0).clip(-1, 1)
    fe = []
    for _ in range(25):
        act, cache = O.simresp_fwd(xn, bank)
        fe.append(float(act[0]))
        g = O.simresp_bwd(np.ones(1, np.float32), cache, xn.shape, bank)
        xn = O.change_norm(xn + 10 * O.gaussian_blur(g.astype(np.float32), 1.0), 1.0, 0.0).clip(-1, 1)
    fe.append(float(O.simresp_fwd(xn, bank)[0][0]))
    assert relerr(np.array(fevals), np.array(fe)) < 1e-4 and relerr(x.cpu().numpy(), xn) < 1e-3


@pytest.mark.gpu
def test_gpu_learn_step_with_gradient_blur_matches_reference():
    """learn(gaussian_blur_sigma=1.0)'s step: the blur is a backward hook on img_post (inv_temp_learn_match.py:524-525); loss and parameter
    gradients of the generic (autograd) composition against the reference's (tests/golden/learn_step_20_gb.npz)."""
    import laminr_b200 as L
    from laminr_b200.contrastive_loss import SimCLROnGrid
    from laminr_b200.utils.general import SingleNeuronModel
    from tests.test_gpu_api import _manifold

    d, dg = golden("learn_step_20"), golden("learn_step_20_gb")
    model, im = _manifold(L, 20, d)
    grid = torch.from_numpy(d["grid"]).reshape(-1, 1).cuda()
    reg = SimCLROnGrid(1, 20, 0.1, 0.3, True).cuda()
    _, img_post, _acts, _ = im.forward(grid, im.template, im.img_transforms, SingleNeuronModel(model, 0), gb=im._blur(1.0), return_template=True)
    loss = -(_acts / float(d["mei_act"])).mean() - 2 * reg.masked_reg_term(img_post, torch.from_numpy(d["mask"]).cuda(), -1.0, 1.0)
    loss.backward()
    assert abs(loss.item() - float(dg["loss"])) < 3e-5 * abs(float(dg["loss"]))
    for l in range(5):
        lay = getattr(im.template.func, f"layer{l}")
        assert relerr(lay.weight.grad.cpu().numpy(), dg[f"dW{l}"]) < 2e-3, l
        assert relerr(lay.bias.grad.cpu().numpy(), dg[f"db{l}"]) < 2e-3, l
    # and it differs from the un-blurred gradient (the hook is live)
    assert relerr(im.template.func.layer1.weight.grad.cpu().numpy(), d["dW1"]) > 0.05
