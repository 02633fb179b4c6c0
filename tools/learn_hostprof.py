"""Host-side profile (cProfile) of InvarianceManifold.learn(0) on the demo1 objects.  usage: python tools/learn_hostprof.py [epochs]"""
import cProfile, contextlib, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import laminr_b200 as L
import bench

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 40
res = 100
torch.manual_seed(0)
model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
meis = bench.synthetic_meis(res)
c = dict(required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
with contextlib.redirect_stderr(io.StringIO()):
    im = L.InvarianceManifold(model, meis, precision="bf16x3", **c)
    im.learn(0, num_max_epochs=3)
    for ep in (10, epochs):
        im = L.InvarianceManifold(model, meis, precision="bf16x3", **c)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        im.learn(0, num_max_epochs=ep)
        torch.cuda.synchronize()
        print(f"learn(0, {ep} epochs): {(time.perf_counter() - t0) * 1e3:.1f} ms, {im.learn_steps_taken} steps, {im.learn_epochs_run} epochs")
    im = L.InvarianceManifold(model, meis, precision="bf16x3", **c)
    pr = cProfile.Profile()
    pr.enable()
    im.learn(0, num_max_epochs=epochs)
    torch.cuda.synchronize()
    pr.disable()
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(22)
print(out.getvalue()[:5000])
