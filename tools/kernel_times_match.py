"""Per-kernel GPU durations of the match step (D targets) via torch.profiler.  usage: python tools/kernel_times_match.py [D] [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import laminr_b200 as L
from laminr_b200.fused import MatchStepper

D = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
res = 100
torch.manual_seed(0)
pop = L.neuron_models.simulated_population(D + 1, img_res=[res, res], seed=1).cuda()
yy, xx = np.meshgrid(np.arange(res) + 0.5, np.arange(res) + 0.5, indexing="ij")
rs = np.random.RandomState(0)
meis = {}
for i in range(D + 1):
    cx, cy = rs.uniform(0.3, 0.7, 2) * res
    mask = np.exp(-0.5 * (((xx - cx) / 15.0) ** 2 + ((yy - cy) / 15.0) ** 2)).astype(np.float32)
    meis[i] = dict(mei=np.zeros((1, res, res), np.float32), activation=1.0, mask=mask, center_pos=dict(x=cx, y=cy))
im = L.InvarianceManifold(pop, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=prec)
im.template.initialize_coordinate_transform(D, True, False, 0.1, True, True, False, True)
rfs = torch.tensor([[meis[i]["center_pos"]["x"], meis[i]["center_pos"]["y"]] for i in range(D + 1)], dtype=torch.float32, device="cuda")
im.template.register_coords_shifts(rfs[0], rfs[1:])
ms = MatchStepper(im.template, im.img_transforms, pop, list(range(1, D + 1)), np.ones(D, np.float32), 20, precision=prec, use_graph=False)
grid = torch.linspace(0.1, 6.2, 20).cuda()
for _ in range(3):
    ms.step(grid)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        ms.step(grid)
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / max(e.count, 1)) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(c * t for _, c, t in rows) / 5
for k, c, t in sorted(rows, key=lambda r: -r[1] * r[2])[:14]:
    print(f"{k[:90]:90s} n/step={c / 5:5.1f} avg={t:9.2f} us  share={c * t / 5 / tot:.3f}")
print(f"sum of kernel time per step: {tot:.1f} us  = {tot / D:.1f} us per target-step")
