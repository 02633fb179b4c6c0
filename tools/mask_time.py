"""Times lmr_create_mask on a batch of MEI-like images against the scipy restatement on the host (one image at a time, as the reference does).
Developer tool: `oracle/` appears here only as the CPU baseline being timed beside the kernel (like bench.py's cpu_baseline leg), never on the product path."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from laminr_b200 import ops
from oracle import mask_oracle as MO

for (n, H, W) in ((1024, 100, 100), (128, 200, 200)):
    meis = MO.synthetic_meis(n, H, W, seed=7)
    x = torch.from_numpy(meis).cuda()
    for _ in range(3):
        ops.raw_create_mask(x, 0.3, 2, 1.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.raw_create_mask(x, 0.3, 2, 1.5)
    e1.record(); torch.cuda.synchronize()
    gpu_ms = e0.elapsed_time(e1) / 10
    t = time.time()
    for i in range(16):
        MO.create_mask(meis[i], 0.3)
    cpu_ms = (time.time() - t) / 16 * 1e3
    print(f"create_mask {n} x {H}x{W}: GPU {gpu_ms:.3f} ms per batch ({n / gpu_ms * 1e3:.0f} masks/s); host scipy {cpu_ms:.2f} ms per image ({1e3 / cpu_ms:.0f} masks/s)")
