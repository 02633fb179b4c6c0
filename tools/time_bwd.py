"""Times lmr_inr_forward (keeping activations) and lmr_inr_backward_after_forward of the learn step (and the match step with --match D) and checks the
gradient against the fp32 CUDA-core path.  Environment switches are read by the library once per process: LMR_BWD_PP, LMR_BWD_REVERSE."""
import argparse
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from laminr_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("--res", type=int, default=100)
ap.add_argument("--match", type=int, default=0)
ap.add_argument("--reps", type=int, default=50)
args = ap.parse_args()
dev, res, N, P = "cuda", args.res, 20, args.res * args.res
D = max(1, args.match)
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
aff = None
if args.match:
    aff = ops.AffineArgs(t(rng.uniform(0, 6, D)), t(rng.uniform(0.8, 1.2, (D, 2))), t(rng.uniform(-0.1, 0.1, (D, 2))), t(rng.uniform(-0.05, 0.05, (D, 2))),
                         tc=t([0.2, 0.2]), shift_to=t([0.2, 0.2]), shift_from=t(rng.uniform(-0.2, 0.2, (D, 2))), clamp=True)
g = torch.from_numpy(rng.randn(D, N, 1, P).astype(np.float32)).to(dev)
kw = dict(want_params=not args.match, want_affine=bool(args.match))


def run(prec, keep):
    ws = torch.empty(ops.forward_workspace_bytes(desc, N, P, D, keep), dtype=torch.uint8, device=dev)
    img = ops.raw_inr_forward(desc, flat, B, auxB, grid, coords, None, aff, precision=prec, workspace=ws, keep_activations=keep)
    gp, ga = ops.raw_inr_backward(desc, flat, B, auxB, grid, coords, None, aff, g, precision=prec, forward_workspace=ws, keep_activations=keep, **kw)
    torch.cuda.synchronize()
    out = gp if gp is not None else torch.cat([x.reshape(-1) for x in ga])
    return ws, img, out


_, img0, ref = run("fp32", False)
ws, img, got = run("bf16x3", True)
rel = lambda a_, b_: float((a_ - b_).norm() / b_.norm())
print(f"res {res} D {D}: img rel {rel(img, img0):.2e}  grad rel {rel(got, ref):.2e}", flush=True)
img_buf = torch.empty_like(img)
wsb = torch.empty(int(_lib.load().lmr_inr_backward_workspace_bytes(ctypes.byref(desc), N, P, D)), dtype=torch.uint8, device=dev)
gpar = torch.empty_like(flat) if not args.match else None
for name, fn in (("fwd_keep", lambda: ops.raw_inr_forward(desc, flat, B, auxB, grid, coords, None, aff, out=img_buf, precision="bf16x3", workspace=ws, keep_activations=True)),
                 ("bwd", lambda: ops.raw_inr_backward(desc, flat, B, auxB, grid, coords, None, aff, g, precision="bf16x3", workspace=wsb, g_params=gpar,
                                                      forward_workspace=ws, keep_activations=True, **kw))):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(gr, stream=side):
        fn()
    torch.cuda.synchronize()
    a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_.record()
    for _ in range(args.reps):
        gr.replay()
    b_.record()
    torch.cuda.synchronize()
    print(f"   {name}: {a_.elapsed_time(b_) / args.reps * 1e3:.1f} us (launch group, back to back)", flush=True)
