"""Debug: affine gradients of lmr_inr_backward at sizes where every CTA walks several tiles: fp32 vs bf16x3 (recompute) vs bf16x3 (kept activations)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from laminr_b200 import ops

torch.manual_seed(0)
dev = "cuda"
for res, D in ((20, 2), (40, 2), (60, 2), (100, 1), (100, 2), (100, 5)):
    N, P = 20, res * res
    desc = ops.make_desc(50, 4, 50, 50, 1)
    rng = np.random.RandomState(0)
    W = [rng.randn(50, 201).astype(np.float32) * 0.15] + [rng.randn(50, 50).astype(np.float32) * 0.2 for _ in range(3)] + [rng.randn(1, 50).astype(np.float32) * 0.2]
    b = [rng.randn(w.shape[0]).astype(np.float32) * 0.05 for w in W]
    flat = torch.from_numpy(np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(W, b)])).to(dev)
    B = torch.from_numpy(rng.randn(2, 50).astype(np.float32) * 10).to(dev)
    auxB = torch.from_numpy(rng.randn(2, 50).astype(np.float32) * 0.1).to(dev)
    grid = torch.linspace(0, 2 * np.pi, N + 1)[:-1].to(dev) + 0.1
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, res), torch.linspace(-1, 1, res), indexing="ij")
    coords = torch.stack([yy, xx], -1).reshape(-1, 2).contiguous().to(dev)
    t = lambda a: torch.from_numpy(np.asarray(a, np.float32)).to(dev)
    aff = ops.AffineArgs(t(rng.uniform(0, 6, D)), t(rng.uniform(0.8, 1.2, (D, 2))), t(rng.uniform(-0.1, 0.1, (D, 2))), t(rng.uniform(-0.05, 0.05, (D, 2))),
                         tc=t([0.2, 0.2]), shift_to=t([0.2, 0.2]), shift_from=t(rng.uniform(-0.2, 0.2, (D, 2))), clamp=True)
    g = torch.from_numpy(rng.randn(D, N, 1, P).astype(np.float32)).to(dev)
    outs = {}
    for name, prec, keep in (("fp32", "fp32", False), ("tc", "bf16x3", False), ("tc_keep", "bf16x3", True)):
        ws = torch.empty(ops.forward_workspace_bytes(desc, N, P, D, keep), dtype=torch.uint8, device=dev)
        img = ops.raw_inr_forward(desc, flat, B, auxB, grid, coords, None, aff, precision=prec, workspace=ws, keep_activations=keep)
        _, ga = ops.raw_inr_backward(desc, flat, B, auxB, grid, coords, None, aff, g, want_params=False, want_affine=True, precision=prec,
                                     forward_workspace=ws, keep_activations=keep)
        torch.cuda.synchronize()
        outs[name] = (img.cpu().numpy(), [x.cpu().numpy().ravel() for x in ga])
    ref = outs["fp32"]
    for name in ("tc", "tc_keep"):
        o = outs[name]
        rel = lambda a, b_: float(np.linalg.norm(a - b_) / (np.linalg.norm(b_) + 1e-30))
        print(f"res {res} D {D} {name}: img {rel(o[0], ref[0]):.1e}  grads " + " ".join(f"{rel(x, y):.1e}" for x, y in zip(o[1], ref[1])), flush=True)
    # also the parameter gradient in the multi-target case
    for name, prec, keep in (("fp32", "fp32", False), ("tc_keep", "bf16x3", True)):
        ws = torch.empty(ops.forward_workspace_bytes(desc, N, P, D, keep), dtype=torch.uint8, device=dev)
        ops.raw_inr_forward(desc, flat, B, auxB, grid, coords, None, aff, precision=prec, workspace=ws, keep_activations=keep)
        gp, _ = ops.raw_inr_backward(desc, flat, B, auxB, grid, coords, None, aff, g, want_params=True, want_affine=False, precision=prec,
                                     forward_workspace=ws, keep_activations=keep)
        torch.cuda.synchronize()
        outs["p_" + name] = gp.cpu().numpy()
    print(f"   param grads keep vs fp32: {float(np.linalg.norm(outs['p_tc_keep'] - outs['p_fp32']) / np.linalg.norm(outs['p_fp32'])):.1e}", flush=True)
