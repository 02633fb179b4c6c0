"""Where the time of a CUDA-graph capture of the match step goes (host wall clock).  usage: python tools/capture_time.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import laminr_b200 as L
import bench
from laminr_b200.fused import MatchStepper

res = 100
model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
meis = bench.synthetic_meis(res)
im = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision="bf16x3")
im.template.initialize_coordinate_transform(2, True, False, 0.1, True, True, False, True)
rfs = torch.tensor([[meis[i]["center_pos"]["x"], meis[i]["center_pos"]["y"]] for i in range(3)], dtype=torch.float32, device="cuda")
im.template.register_coords_shifts(rfs[0], rfs[1:])
grid = torch.linspace(0.1, 6.2, 20).cuda()
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ms = MatchStepper(im.template, im.img_transforms, model, [1, 2], np.ones(2, np.float32), 20, precision="bf16x3", use_graph=True)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ms._train()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        ta = time.perf_counter(); g.capture_begin(); tb = time.perf_counter(); ms._train(); tc = time.perf_counter(); g.capture_end(); td = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    g.replay(); torch.cuda.synchronize(); t4 = time.perf_counter()
    g.replay(); torch.cuda.synchronize(); t5 = time.perf_counter()
    print(f"rep {rep}: construct {1e3*(t1-t0):.2f} ms, eager step {1e3*(t2-t1):.2f}, capture_begin {1e3*(tb-ta):.2f}, launches {1e3*(tc-tb):.2f}, capture_end {1e3*(td-tc):.2f}, "
          f"first replay {1e3*(t4-t3):.2f}, second replay {1e3*(t5-t4):.2f}")
