"""ncu target: the cone kernel in the MEI configuration (128 (image, neuron) pairs, one CTA each, forward + image gradient)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, warnings
warnings.simplefilter("ignore")
from laminr_b200 import ops
from laminr_b200.neuron_models import conv_core_gaussian_model
m = conv_core_gaussian_model((1, 200, 200), n_neurons=128).cuda().eval()
cone = m.cone()
G, NB = 128, 1
x = torch.randn(G, NB, 1, 200, 200, device="cuda") * 0.1
nog = torch.arange(G, dtype=torch.int32, device="cuda")
g = torch.empty_like(x)
for _ in range(4):
    ops.raw_convcore_response(cone, x, NB * 40000, NB, G, nog, g_img=g, accumulate=False)
torch.cuda.synchronize()
