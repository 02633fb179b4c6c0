"""Times the full-frame response kernels (csrc/fullframe.cu) on the macaque-V1 point-pooled model (random init) and, next to it, the same
network on stock torch / cuDNN ops -- the library baseline on the same GPU.  CUDA events, 20 repetitions after warm-up.

    python tools/time_fullframe.py [B ...]
"""
import sys
import warnings

import torch
import torch.nn.functional as F

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from laminr_b200.neuron_models import monkey_models as mm  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def torch_forward(km, x):
    """the same eval-mode network on torch ops (cuDNN convolutions, grid_sample)"""
    for feat in km.core.features:
        if hasattr(feat, "conv"):
            x = F.conv2d(x, feat.conv.weight)
        else:
            ds = feat.ds_conv
            x = F.conv2d(x, ds.in_depth_conv.weight)
            x = F.conv2d(x, ds.spatial_conv.weight, padding=ds.spatial_conv.padding, dilation=ds.spatial_conv.dilation, groups=x.shape[1])
            x = F.conv2d(x, ds.out_depth_conv.weight)
        n = feat.norm
        x = F.elu(F.batch_norm(x, n.running_mean, n.running_var, n.weight, n.bias, False, 0.0, n.eps))
        if hasattr(feat, "seg_ex_block"):
            se = feat.seg_ex_block.se
            s = torch.sigmoid(se[3](torch.relu(se[1](x.flatten(2).mean(-1)))))
            x = x * s[:, :, None, None]
    ro = km.readout
    g = ro.grid.expand(x.shape[0], -1, 1, 2)
    pools = [F.grid_sample(x, g, align_corners=True)]
    for _ in range(ro.pool_steps):
        x = F.avg_pool2d(x, ro.pool_kern, stride=ro.pool_kern)
        pools.append(F.grid_sample(x, g, align_corners=True))
    y = (torch.cat(pools, 1).squeeze(-1) * ro.features.view(1, -1, ro.outdims)).sum(1) + ro.bias
    return F.elu(y) + 1


def main():
    warnings.simplefilter("ignore")
    km = mm.keyless_pointpool_monkey_model(mm.get_pointpool_monkey_model(seed=0)).cuda()
    for p in km.parameters():
        p.requires_grad_(False)
    for B in [int(a) for a in sys.argv[1:]] or [1, 20, 64]:
        x = torch.randn(B, 1, 93, 93, device="cuda")

        def fwd(model_fn):
            with torch.no_grad():
                return model_fn(x)

        def fwd_bwd(model_fn):
            xr = x.clone().requires_grad_(True)
            model_fn(xr)[:, 7].sum().backward()
            return xr.grad

        ours, lib = (lambda t: km(t)), (lambda t: torch_forward(km, t))
        err = (fwd(ours) - fwd(lib)).abs().max().item()
        gerr = (fwd_bwd(ours) - fwd_bwd(lib)).abs().max().item() / fwd_bwd(lib).abs().max().item()
        print(f"B={B:3d}  forward: kernels {timed(lambda: fwd(ours)) * 1e3:8.1f} us   torch/cuDNN {timed(lambda: fwd(lib)) * 1e3:8.1f} us   |  "
              f"forward+image gradient: kernels {timed(lambda: fwd_bwd(ours)) * 1e3:8.1f} us   torch/cuDNN {timed(lambda: fwd_bwd(lib)) * 1e3:8.1f} us   "
              f"(max |diff| {err:.2e}, grad rel {gerr:.2e})")


if __name__ == "__main__":
    main()
