"""BASELINE configs[2] + configs[4] at full size: a synthetic population of n simulated neurons (default 1024) at 100x100 --
get_mei_dict for ALL neurons (the MEI sweep), learn(template 0), match of ALL other neurons -- on one GPU, or sharded over the ranks of a torchrun launch.

    python tools/config3.py [n_neurons] [match_epochs]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/config3.py

Prints one JSON line (rank 0): wall clock per stage and size-independent checks of the results:
  * every phase-invariant neuron's MEI reaches the analytic optimum (unit-norm image in the span of a quadrature pair: activation 1.000001);
  * after match, no target is below what its best sweep candidate gave (the optimisation only improves on the initialisation), up to the jitter.
"""
import contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import laminr_b200 as L
from laminr_b200.dist import get_mei_dict_sharded, match_sharded


def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


c = dict(pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, required_img_norm=1.0)
t0 = time.perf_counter()
pop = L.neuron_models.simulated_population(n, img_res=[100, 100], seed=0).to(dev)  # every rank holds the whole bank (819 MB at 1024 neurons)
t_build = time.perf_counter() - t0
torch.manual_seed(0)
err = io.StringIO()  # progress bars; LMR_TIMING=1 phase lines are passed through below
with contextlib.redirect_stderr(err), contextlib.redirect_stdout(io.StringIO()):
    sync(); t0 = time.perf_counter()
    meis = get_mei_dict_sharded(pop, [1, 100, 100], neuron_idxs=list(range(n)), **c)
    sync(); t1 = time.perf_counter()
    im = L.InvarianceManifold(pop, meis, **c)
    if rank == 0:
        im.learn(0)
    else:
        im.template_neuron_idx = 0
    sync(); t2 = time.perf_counter()
    imgs, acts = match_sharded(im, list(range(1, n)), num_epochs=epochs)
    sync(); t3 = time.perf_counter()
for line in err.getvalue().splitlines():
    if line.startswith("[match"):
        print(f"rank {rank}: {line}", file=sys.stderr)
if rank == 0:
    mei_act = np.array([meis[i]["activation"] for i in range(n)])
    a = np.asarray(acts) / mei_act[1:, None]
    out = {"workload": f"BASELINE configs[2]+[4]: {n} simulated neurons, 100x100, MEI sweep + learn(0) + match(all {n - 1})", "n_gpus": world,
           "build_population_host_s": t_build, "get_mei_dict_s": t1 - t0, "learn_s": t2 - t1, "match_s": t3 - t2,
           "meis_per_s": n / (t1 - t0), "neurons_aligned_per_s": (n - 1) / (t3 - t2),
           "learn_epochs": getattr(im, "learn_epochs_run", None), "match_epochs": getattr(im, "match_epochs_run", None),
           "mei_activation_min_max": [float(mei_act.min()), float(mei_act.max())],
           "aligned_norm_act_mean_min_max": [float(a.mean()), float(a.mean(axis=1).min()), float(a.mean(axis=1).max())],
           "frac_targets_above_0.9": float((a.mean(axis=1) > 0.9).mean()), "result_shapes": [list(np.asarray(imgs).shape), list(np.asarray(acts).shape)]}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
