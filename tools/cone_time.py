import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, time, warnings
warnings.simplefilter("ignore")
from laminr_b200 import ops
from laminr_b200.neuron_models import conv_core_gaussian_model
m = conv_core_gaussian_model((1,200,200), n_neurons=128).cuda().eval()
cone = m.cone()
for (G, NB) in ((1,20),(127,20),(128,1)):
    x = torch.randn(G, NB, 1, 200, 200, device="cuda")*0.1
    nog = torch.arange(G, dtype=torch.int32, device="cuda")
    g = torch.empty_like(x)
    for with_grad in (False, True):
        f = lambda: ops.raw_convcore_response(cone, x, NB*40000, NB, G, nog, g_img=g if with_grad else None, accumulate=False)
        for _ in range(3): f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20): f()
        b.record(); torch.cuda.synchronize()
        print(f"G={G} NB={NB} grad={with_grad}: {a.elapsed_time(b)/20*1e3:.1f} us")
