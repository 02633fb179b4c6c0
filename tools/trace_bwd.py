"""Debug helper: one tensor-core INR backward at 100x100 with LMR_TC_TRACE=1 (prints CTA 0's pipeline time line)."""
import os, sys
os.environ["LMR_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laminr_b200 import ops
res, N = 100, 20
P = res * res
rng = np.random.RandomState(0)
W = [rng.randn(50, 201).astype(np.float32) * 0.1] + [rng.randn(50, 50).astype(np.float32) * 0.1 for _ in range(3)] + [rng.randn(1, 50).astype(np.float32) * 0.1]
b = [np.zeros(w.shape[0], np.float32) for w in W]
flat = torch.tensor(np.concatenate([np.concatenate([w.ravel(), x.ravel()]) for w, x in zip(W, b)])).cuda()
B, auxB = torch.tensor(rng.randn(2, 50).astype(np.float32) * 10).cuda(), torch.tensor(rng.randn(2, 50).astype(np.float32) * 0.1).cuda()
grid = torch.linspace(0, 6, N).cuda()
yy, xx = np.meshgrid(np.linspace(-1, 1, res), np.linspace(-1, 1, res), indexing="ij")
coords = torch.tensor(np.stack([yy, xx], -1).reshape(-1, 2).astype(np.float32)).cuda()
desc = ops.make_desc(50, 4, 50, 50, 1)
g = torch.randn(1, N, 1, P, device="cuda")
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
for _ in range(2):
    ops.raw_inr_backward(desc, flat, B, auxB, grid, coords, None, None, g, precision=prec)
torch.cuda.synchronize()
