"""Per-kernel GPU durations of the conv-encoder learn step at 200x200 (BASELINE config 4), warm, via torch.profiler/CUPTI."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import laminr_b200 as L
from laminr_b200.contrastive_loss import SimCLROnGrid
from laminr_b200.fused import LearnStepper
from bench import conv_meis

res = int(sys.argv[1]) if len(sys.argv) > 1 else 200
warnings.simplefilter("ignore")
model = L.neuron_models.conv_core_gaussian_model((1, res, res), n_neurons=128, seed=0).cuda().eval()
meis = conv_meis(model, res, 2)
torch.manual_seed(0)
im = L.InvarianceManifold(model, meis, required_img_norm=0.05 * res, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
reg = SimCLROnGrid(1, 20, 0.1, 0.3, True).cuda()
st = LearnStepper(im.template, im.img_transforms, model, 0, meis[0]["mask"], 1.0, reg, -1.0, 1.0, 20, precision=im.precision, use_graph=False)
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
