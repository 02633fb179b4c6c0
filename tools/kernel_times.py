"""Per-kernel GPU durations of the learn step (warm, back to back) via torch.profiler/CUPTI.  usage: python tools/kernel_times.py [precision] [res]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import laminr_b200 as L
from laminr_b200.contrastive_loss import SimCLROnGrid
from laminr_b200.fused import LearnStepper
from bench import synthetic_meis

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
res = int(sys.argv[2]) if len(sys.argv) > 2 else 100
torch.manual_seed(0)
model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
meis = synthetic_meis(res)
im = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=prec)
reg = SimCLROnGrid(1, 20, 0.1, 0.3, True).cuda()
st = LearnStepper(im.template, im.img_transforms, model, 0, meis[0]["mask"], meis[0]["activation"], reg, -1.0, 1.0, 20, precision=prec, use_graph=False)
grid = torch.linspace(0.1, 6.2, 20).cuda()
for _ in range(5):
    st.step(grid)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(20):
        st.step(grid)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / max(e.count, 1)) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(c * t for _, c, t in rows) / 20
for k, c, t in sorted(rows, key=lambda r: -r[1] * r[2]):
    print(f"{k[:90]:90s} n/step={c / 20:5.1f} avg={t:8.2f} us  share={c * t / 20 / tot:.3f}")
print(f"sum of kernel time per step: {tot:.1f} us")
