"""The fused learn tail (lmr_inr_backward_adam) against the separate launches it replaces: gradients, parameters and Adam moments must be bit-identical."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from laminr_b200 import ops

for res in (20, 37, 100):
    dev, N, P = "cuda", 20, res * res
    desc = ops.make_desc(50, 4, 50, 50, 1)
    rng = np.random.RandomState(res)
    W = [rng.randn(50, 201).astype(np.float32) * 0.15] + [rng.randn(50, 50).astype(np.float32) * 0.2 for _ in range(3)] + [rng.randn(1, 50).astype(np.float32) * 0.2]
    b = [rng.randn(w.shape[0]).astype(np.float32) * 0.05 for w in W]
    flat0 = torch.from_numpy(np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(W, b)])).to(dev)
    B = torch.from_numpy(rng.randn(2, 50).astype(np.float32) * 10).to(dev)
    auxB = torch.from_numpy(rng.randn(2, 50).astype(np.float32) * 0.1).to(dev)
    grid = torch.linspace(0, 2 * np.pi, N + 1)[:-1].to(dev) + 0.1
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, res), torch.linspace(-1, 1, res), indexing="ij")
    coords = torch.stack([yy, xx], -1).reshape(-1, 2).contiguous().to(dev)
    outs = []
    for fused in (False, True):
        flat = flat0.clone()
        m, v, step = torch.zeros_like(flat), torch.zeros_like(flat), torch.zeros(2, dtype=torch.int32, device=dev)
        gp = torch.empty_like(flat)
        ws = torch.empty(ops.forward_workspace_bytes(desc, N, P, 1, True), dtype=torch.uint8, device=dev)
        for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):  # three steps: the step counter, the moments and the re-armed barrier word all carry over
            g = torch.from_numpy(np.random.RandomState(100 + it).randn(1, N, 1, P).astype(np.float32)).to(dev)
            ops.raw_inr_forward(desc, flat, B, auxB, grid, coords, None, None, precision="bf16x3", workspace=ws, keep_activations=True)
            if fused:
                ops.raw_inr_backward_adam(desc, flat, B, auxB, grid, coords, None, g, gp, m, v, step, lr=1e-3, precision="bf16x3", forward_workspace=ws,
                                          keep_activations=True)
            else:
                ops.raw_inr_backward(desc, flat, B, auxB, grid, coords, None, None, g, precision="bf16x3", g_params=gp, forward_workspace=ws, keep_activations=True)
                ops.raw_adam_step(flat, gp, m, v, step, lr=1e-3)
        torch.cuda.synchronize()
        outs.append([x.cpu().numpy().copy() for x in (gp, flat, m, v, step)])
    same = [np.array_equal(a, b_) for a, b_ in zip(*outs)]
    if not all(same):
        for name, a, b_ in zip(("grad", "params", "m", "v"), *[o[:4] for o in outs]):
            d = np.abs(a.astype(np.float64) - b_.astype(np.float64))
            bad = np.nonzero(d > 0)[0]
            print(f"   {name}: {bad.size} differing entries, max abs diff {d.max():.3e} (max |value| {np.abs(a).max():.3e}), first indices {bad[:8].tolist()}, last {bad[-4:].tolist()}")
    print(f"res {res}: grad / params / m / v / step identical: {same}; step = {outs[1][4].tolist()}", flush=True)
    assert all(same)
print("fused tail ok")
