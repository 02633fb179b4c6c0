"""Host-side profile (cProfile) of a whole match() call on the demo1 objects (BASELINE configs[1]).  usage: python tools/match_hostprof.py [epochs]"""
import cProfile, contextlib, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import laminr_b200 as L
import bench

epochs = int(sys.argv[1]) if len(sys.argv) > 1 else 100
res = 100
torch.manual_seed(0)
model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
meis = bench.synthetic_meis(res)
im = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision="bf16x3")
im.template_neuron_idx = 0
with contextlib.redirect_stderr(io.StringIO()):
    im.match([1, 2], num_epochs=3, patience=10 ** 9)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    im.match([1, 2], num_epochs=epochs, patience=10 ** 9)
    torch.cuda.synchronize()
    print(f"match([1,2], {epochs} epochs): {(time.perf_counter() - t0) * 1e3:.1f} ms")
    t0 = time.perf_counter()
    im.match([1, 2], num_epochs=1, patience=10 ** 9)
    torch.cuda.synchronize()
    print(f"match([1,2], 1 epoch): {(time.perf_counter() - t0) * 1e3:.1f} ms  (fixed part: stepper construction, sweep, capture, snapshot)")
    pr = cProfile.Profile()
    pr.enable()
    im.match([1, 2], num_epochs=epochs, patience=10 ** 9)
    torch.cuda.synchronize()
    pr.disable()
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats(os.environ.get("LMR_PROF_SORT", "cumulative")).print_stats(28)
print(out.getvalue()[:6000])
