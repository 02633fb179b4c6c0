"""Debug helper: a few fused learn steps of demo1 at 100x100 with LMR_TAIL_TRACE=1 (prints the phase time line of learn_tail_kernel: one block of each role)."""
import os, sys
os.environ["LMR_TAIL_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import laminr_b200 as L
from laminr_b200.contrastive_loss import SimCLROnGrid
from laminr_b200.fused import LearnStepper
from bench import synthetic_meis, N_LATENTS

res = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device("cuda", 0)
torch.manual_seed(0); np.random.seed(0)
model = L.neuron_models.simulated("demo1", img_res=[res, res]).to(dev)
meis = synthetic_meis(res)
im = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision="bf16x3")
reg = SimCLROnGrid(1, N_LATENTS, 0.1, 0.3, True).to(dev)
st = LearnStepper(im.template, im.img_transforms, model, 0, meis[0]["mask"], meis[0]["activation"], reg, -1.0, 1.0, N_LATENTS, lr=1e-3, reg_scale=2.0,
                  precision="bf16x3", use_graph=False)
grid = torch.linspace(0.1, 6.2, N_LATENTS).to(dev)
for i in range(6):
    st.step(grid)
torch.cuda.synchronize()
