import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import laminr_b200 as L
from laminr_b200.utils.mei import batched_simulated_meis
for n in (33, 128):
    pop = L.neuron_models.simulated_population(n, img_res=[100, 100], seed=1).cuda()
    x0 = torch.randn(n, 1, 100, 100).cuda()
    batched_simulated_meis(pop, list(range(n)), x0, -1.0, 1.0, norm=1.0, num_iterations=10)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _, fev = batched_simulated_meis(pop, list(range(n)), x0, -1.0, 1.0, norm=1.0, num_iterations=1000)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    gb = (20 * 10000 * 4 + 3 * 10000 * 4) * 1000 * n / 1e9
    print(f"n={n}: {ms:.1f} ms  {n / ms * 1e3:.0f} MEIs/s  {gb / ms * 1e3:.0f} GB/s algorithmic  final act {fev[:, -1].mean().item():.6f}")
