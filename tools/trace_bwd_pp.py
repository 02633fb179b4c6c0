"""LMR_TC_TRACE=1 [LMR_PP_MODE=m] python tools/trace_bwd_pp.py [res]: one eager forward (keeping activations) + backward at res x res; the library prints
CTA 0's time line of the ping-pong backward kernel on stderr (debug path of csrc/inr_tc.cu)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from laminr_b200 import ops

res = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev, N, P = "cuda", 20, res * res
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
g = torch.from_numpy(rng.randn(1, N, 1, P).astype(np.float32)).to(dev)
ws = torch.empty(ops.forward_workspace_bytes(desc, N, P, 1, True), dtype=torch.uint8, device=dev)
for rep in range(2):
    ops.raw_inr_forward(desc, flat, B, auxB, grid, coords, None, None, precision="bf16x3", workspace=ws, keep_activations=True)
    if rep == 1:
        print("---- second call (warm) ----", file=sys.stderr, flush=True)
    ops.raw_inr_backward(desc, flat, B, auxB, grid, coords, None, None, g, precision="bf16x3", forward_workspace=ws, keep_activations=True)
    torch.cuda.synchronize()
