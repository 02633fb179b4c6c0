"""Warm CUDA-event times of the unfused masked-contrastive launches (lmr_simclr_fwd = gram + finalize, lmr_simclr_bwd).  usage: python tools/simclr_time.py [res]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laminr_b200 import ops
from laminr_b200.contrastive_loss import SimCLROnGrid

res = int(sys.argv[1]) if len(sys.argv) > 1 else 200
N, P = 20, res * res
z = torch.randn(N, 1, res, res, device="cuda") * 0.01
mask = torch.rand(P, device="cuda")
_reg = SimCLROnGrid(1, N, 0.1, 0.3, True)
pos, neg = _reg.flat_neighbor_mask_pos.cuda(), _reg.flat_neighbor_mask_neg.cuda()
reg, K = torch.empty(1, device="cuda"), torch.empty(N, N, device="cuda")
from laminr_b200 import _lib
ws = torch.empty(_lib.load().lmr_simclr_workspace_bytes(N, 1, P), dtype=torch.uint8, device="cuda")
g = torch.empty_like(z)
g_reg = torch.full((1,), -2.0, device="cuda")


def t(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(gr, stream=s):
        for _ in range(reps):
            fn()
    torch.cuda.synchronize()
    a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


fwd = lambda: ops.raw_simclr_fwd(z, mask, N, 1, P, -1.0, 1.0, pos, neg, 0.3, 1e-6, reg=reg, K=K, workspace=ws)
bwd = lambda: ops.raw_simclr_bwd(z, mask, K, g_reg, 1.0, N, 1, P, -1.0, 1.0, g, accumulate=False)
print(f"res {res}: simclr_fwd (gram + finalize) {t(fwd):.1f} us, simclr_bwd {t(bwd):.1f} us")
