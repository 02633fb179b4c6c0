"""cProfile of the FIRST InvarianceManifold.learn(0) of a process (after CUDA context creation)"""
import cProfile, contextlib, io, os, pstats, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import laminr_b200 as L
import bench
res = 100
torch.manual_seed(0)
model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
meis = bench.synthetic_meis(res)
torch.cuda.synchronize()
c = dict(required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
for k in range(3):
  with contextlib.redirect_stderr(io.StringIO()):
    im = L.InvarianceManifold(model, meis, precision="bf16x3", **c)
    pr = cProfile.Profile()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pr.enable()
    im.learn(0, num_max_epochs=20)
    torch.cuda.synchronize()
    pr.disable()
    print(f"call {k}: {(time.perf_counter() - t0) * 1e3:.1f} ms")
  out = io.StringIO()
  pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(14)
  print(out.getvalue()[:3500])
