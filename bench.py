#!/usr/bin/env python
"""bench.py -- INR opt steps/sec of `InvarianceManifold.learn` on demo1 @ 100x100 (BASELINE.json metric), B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--res 100] [--precision fp32|bf16x3|bf16x3_fast]

One "step" is one pass of the learn hot path (reference laminr/inv_temp_learn_match.py:186-208): INR forward over
20 latents x HxW pixels, constraint projection, template-neuron response, masked contrastive term, backward, Adam.

  value    : steps/s with inputs resident in HBM.  Each step is timed with its own CUDA-event pair on the launching
             stream; L2 is flushed (256 MiB write) between timed steps, outside the event pairs.
  e2e      : steps/s through the host-facing call, the host waiting for every step's loss: LearnStepper.step(host row) + sync_loss() --
             the step's N latents travel host -> device (pinned memory read by the step's first kernel), its loss device -> host
             (written into pinned memory by the objective kernel); `copy_engine` = the same loop with cudaMemcpyAsync copies.
  roofline : the INR backward launch group (dominant), algorithmic FLOPs / CUDA-event time vs measured bf16 peak.
  cpu_baseline : the reference's CPU path (torch-CPU port of its op sequence, oracle/torch_port.py; or the reference
             itself when /root/reference is mounted) on a bounded sample of the same workload.
  api_learn: extra -- steps/s through the public call itself, InvarianceManifold.learn(0, num_max_epochs=20): stepper construction, graph
             capture, per-epoch control logic and the final snapshot to the host included (rank 0).
  match    : extra -- target-steps/s of the match hot step (:351-366), targets sharded over the N ranks (weak scaling, --match-targets per
             GPU); `aligned` = target neurons aligned/s of a whole match() call (60-angle sweep + --match-epochs epochs + results on the host);
             `mei` = MEIs/s of the batched ascent (BASELINE configs[4]) with a one-core numpy-port baseline; `config2` (N == 1) = BASELINE
             configs[1], demo1 match of targets [1, 2], whole call and per step, with the numpy-port CPU baseline of the match step.
  conv     : extra -- BASELINE configs[3]: learn / match / MEI with the conv-core + Gaussian-readout encoder at 200x200 and its CPU port baseline.
N > 1: learn does not shard (one template) -> N independent replicas ("replicas only"); value = N*K / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LATENTS = 20
TRAFFIC_BWD_KEEP = 165694720 + 3682048  # dram read + write of inr_bwd_pp_kernel (kept activations; profiles/r2g_ncu_learn_step.md)


def algorithmic_flops(res, N=N_LATENTS, h=50, pe=50, C=1):
    """separable layer 0, recomputation not counted."""
    P = res * res
    R = N * P
    fwd = 2 * P * 2 * pe + 2 * P * (2 * pe) * h + 2 * N * (2 * pe) * h + 3 * 2 * R * h * h + 2 * R * h * C
    bwd = 3 * 4 * R * h * h + 2 * 2 * R * h * C + 2 * P * (2 * pe) * h + 2 * N * (2 * pe) * h
    return fwd, bwd


def synthetic_meis(res):
    """Synthetic meis_dict: Gaussian-bump RF masks at the demo1 filter centres, MEI activation of a unit-norm optimum."""
    yy, xx = np.meshgrid(np.arange(res) + 0.5, np.arange(res) + 0.5, indexing="ij")
    out = {}
    for i, (cx, cy) in enumerate([(0.6 * res, 0.6 * res), (0.4 * res, 0.4 * res), (0.4 * res, 0.6 * res)]):
        mask = np.exp(-0.5 * (((xx - cx) / (0.16 * res)) ** 2 + ((yy - cy) / (0.16 * res)) ** 2)).astype(np.float32)
        out[i] = dict(mei=np.zeros((1, res, res), np.float32), activation=1.0000009536743164, mask=mask, center_pos=dict(x=cx, y=cy))
    return out


class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled DURING the timed region by an NVML polling thread (the recipe's clocks line:
    clocks.sm, clocks.max.sm, clocks_event_reasons.{hw_slowdown, hw_thermal_slowdown, sw_thermal_slowdown, sw_power_cap}).  Round 1 piped
    `nvidia-smi -lms` and terminated it: its block-buffered pipe never delivered a sample."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, period_s=0.002):
        import threading
        self.sm, self.smax, self.reasons, self.err = [], None, set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:  # CUDA_VISIBLE_DEVICES may renumber the devices: address the GPU by its UUID
                self.h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(index).uuid))
            except Exception:  # noqa: BLE001
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nv = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = f"NVML unavailable: {e}"
            return

        def poll():
            nv = self.nv
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop.is_set():
                try:
                    self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                    bits = int(get_reasons(self.h))
                    for name, bit in self.REASONS:
                        if bits & bit:
                            self.reasons.add(name)
                except Exception as e:  # noqa: BLE001
                    self.err = str(e)
                    return
                self._stop.wait(period_s)

        self._thread = threading.Thread(target=poll, daemon=True)
        self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smax, "reasons": sorted(self.reasons), "samples": len(self.sm),
               "source": "NVML polled every 2 ms during the timed steps"}
        if self.err:
            out["error"] = self.err
        return out


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline
# ----------------------------------------------------------------------------------------------------------------------
def workload_config(res):
    """`config` of the bench line: identical in both arms (the driver compares them)."""
    return {"workload": f"demo1 learn step, {res}x{res}x1, {N_LATENTS} latents, template neuron 0 (BASELINE configs[0] objects on B200)"}


def cpu_learn_steps(res, n_steps, warmup, threads, anomaly=True):
    """Times the reference's CPU learn step on the host cores -> (steps/s, seconds, kind, where).

    kind "reference": the UNMODIFIED reference (the offline pip install under baseline/_ref, else mounted sources) through its own public API and
    stock code path -- `InvarianceManifold.learn(0, steps_per_epoch=n_steps, num_max_epochs=1)`: one epoch of n_steps jittered steps (forward,
    loss, anomaly-mode backward, Adam; inv_temp_learn_match.py:185-208) plus the call's fixed tail (one un-jittered snapshot forward, :251), all timed.
    kind "port" (reference not installed): the torch-CPU port of the same op sequence, oracle/torch_port.py."""
    import contextlib
    import io

    import torch

    torch.set_num_threads(threads)
    from oracle import laminr_oracle as O

    meis = synthetic_meis(res)
    try:
        from oracle.ref_harness import import_reference, reference_available, reference_location
        if not reference_available(prefer_installed=True):
            raise ImportError("reference not installed")
        laminr = import_reference(prefer_installed=True)
        torch.manual_seed(0)
        np.random.seed(0)
        model = laminr.neuron_models.simulated("demo1", img_res=[res, res])
        im = laminr.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
        orig_anomaly = torch.autograd.set_detect_anomaly
        if not anomaly:  # (informational second number: the reference switches anomaly mode on inside its loop, :206)
            torch.autograd.set_detect_anomaly = lambda *a, **k: None
        try:
            with contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()):
                if warmup > 0:
                    im.learn(0, steps_per_epoch=warmup, num_max_epochs=1)
                t0 = time.perf_counter()
                im.learn(0, steps_per_epoch=n_steps, num_max_epochs=1)
                dt = time.perf_counter() - t0
        finally:
            torch.autograd.set_detect_anomaly = orig_anomaly
            orig_anomaly(False)  # the reference leaves anomaly mode on process-wide
        return n_steps / dt, dt, "reference", reference_location()
    except Exception as e:  # noqa: BLE001 -- any import problem -> the port
        why = f"{type(e).__name__}: {e}"
    from oracle.torch_port import LearnPort
    rng = np.random.RandomState(0)
    grid0 = O.latent_grid(N_LATENTS)
    sigma = 2 * np.pi / N_LATENTS
    grids = [grid0 + np.float32(rng.rand() * sigma - 0.5 * sigma) for _ in range(n_steps + warmup)]
    torch.manual_seed(0)
    h, pe, ape = 50, 50, 50
    W = [np.random.RandomState(1).randn(h, 1 + 2 * ape + 2 * pe).astype(np.float32) * 0.1] + \
        [np.random.RandomState(2 + l).randn(h, h).astype(np.float32) * 0.1 for l in range(3)] + \
        [np.random.RandomState(9).randn(1, h).astype(np.float32) * 0.1]
    b = [np.zeros(w.shape[0], np.float32) for w in W]
    B = np.random.RandomState(10).randn(2, pe).astype(np.float32) * 10
    auxB = np.random.RandomState(11).randn(2, ape).astype(np.float32) * 0.1
    pos, neg = O.pos_neg_masks(N_LATENTS, 0.1, True)
    port = LearnPort(W, b, B, auxB, O.pixel_coords((res, res)), O.demo1_banks([res, res])[0], meis[0]["mask"], meis[0]["activation"], pos, neg,
                     res=(res, res), anomaly_mode=anomaly)
    for g in grids[:warmup]:
        port.step(g)
    t0 = time.perf_counter()
    for g in grids[warmup:]:
        port.step(g)
    dt = time.perf_counter() - t0
    return n_steps / dt, dt, "port", f"oracle/torch_port.py ({why})"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 60 if args.res <= 100 else 15))  # bounded sample: a reference step is ~0.1-0.2 s (100^2) / ~0.9 s (200^2)
    warm = max(0, min(args.warmup, 10))
    sps, dt, kind, where = cpu_learn_steps(args.res, steps, warm, threads)
    what = (f"unmodified reference ({os.path.relpath(where, ROOT) if where.startswith(ROOT) else where}), public call InvarianceManifold.learn(0, "
            f"steps_per_epoch={steps}, num_max_epochs=1) incl. its final snapshot forward" if kind == "reference" else f"torch-CPU port of the reference op sequence: {where}")
    sample = f"{steps} learn steps of demo1 @{args.res}x{args.res} ({N_LATENTS} latents) in {dt:.2f} s on {threads} host threads; {what}; anomaly mode on as shipped"
    line = {
        "impl": "reference", "metric": "INR opt steps/sec (learn, 100x100)", "value": sps, "unit": "steps/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 / sps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.res),
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE config 4 (extra): random-init stacked-conv core + Gaussian readout, 200x200, 128 neurons
# ----------------------------------------------------------------------------------------------------------------------
def conv_meis(model, res, n):
    """Synthetic meis_dict for the conv encoder: Gaussian-bump RF masks at the readout positions."""
    mu = model.readout._mu.detach().cpu().numpy().reshape(-1, 2).clip(-1, 1)
    yy, xx = np.meshgrid(np.arange(res) + 0.5, np.arange(res) + 0.5, indexing="ij")
    out = {}
    for i in range(n):
        cx, cy = (mu[i, 0] + 1) / 2 * (res - 1) + 0.5, (mu[i, 1] + 1) / 2 * (res - 1) + 0.5
        mask = np.exp(-0.5 * (((xx - cx) / (0.08 * res)) ** 2 + ((yy - cy) / (0.08 * res)) ** 2)).astype(np.float32)
        out[i] = dict(mei=np.zeros((1, res, res), np.float32), activation=1.0, mask=mask, center_pos=dict(x=float(cx), y=float(cy)))
    return out


def bench_conv(args, dev, world, rank, barrier, flush):
    import contextlib
    import io
    import warnings

    import torch
    import torch.distributed as dist

    import laminr_b200 as L
    from laminr_b200.contrastive_loss import SimCLROnGrid
    from laminr_b200.fused import LearnStepper, MatchStepper
    from laminr_b200.utils.mei import batched_conv_meis
    from laminr_b200.utils.training import JitteringGridDatamodule

    res, n_neurons, K = args.conv_res, 128, max(3, min(args.steps, args.conv_steps))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = L.neuron_models.conv_core_gaussian_model((1, res, res), n_neurons=n_neurons, seed=0).to(dev).eval()
    Dloc = args.conv_targets
    meis = conv_meis(model, res, Dloc + 1)
    norm = 0.05 * res
    torch.manual_seed(0)
    im = L.InvarianceManifold(model, meis, required_img_norm=norm, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=args.precision)
    reg = SimCLROnGrid(1, N_LATENTS, 0.1, 0.3, True).to(dev)
    st = LearnStepper(im.template, im.img_transforms, model, 0, meis[0]["mask"], 1.0, reg, -1.0, 1.0, N_LATENTS, lr=1e-3, reg_scale=2.0,
                      precision=args.precision, use_graph=True)
    grids = torch.stack(list(JitteringGridDatamodule(1, N_LATENTS, K + 3).train_dataloader())).reshape(K + 3, N_LATENTS).to(dev)
    for i in range(3):
        st.step(grids[i])
    barrier()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    for i in range(K):
        flush.fill_(float(i))
        starts[i].record()
        st.step(grids[3 + i])
        ends[i].record()
    barrier()
    t_learn = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(starts, ends))], device=dev)
    # ---- match hot step: targets 1..Dloc per GPU
    targets = list(range(1, Dloc + 1))
    t = im.template
    t.initialize_coordinate_transform(Dloc, True, False, 0.1, True, True, False, True)
    rfs = torch.tensor([[meis[i]["center_pos"]["x"], meis[i]["center_pos"]["y"]] for i in range(Dloc + 1)], dtype=torch.float32, device=dev)
    t.register_coords_shifts(rfs[0], rfs[1:])
    ms = MatchStepper(t, im.img_transforms, model, targets, np.ones(Dloc, np.float32), N_LATENTS, total_targets=Dloc * world, precision=args.precision,
                      use_graph=True)
    mk = max(3, min(K, args.match_steps))
    for i in range(3):
        ms.step(grids[i])
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(mk):
        ms.step(grids[3 + (i % K)])
    b.record()
    barrier()
    t_match = torch.tensor([a.elapsed_time(b)], device=dev)
    # ---- MEI ascent for all 128 neurons of this rank's replica (1000 iterations, batched)
    x0 = torch.randn(n_neurons, 1, res, res).to(dev)
    batched_conv_meis(model, list(range(n_neurons)), x0, -1.0, 1.0, norm=norm, num_iterations=5)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _, fev = batched_conv_meis(model, list(range(n_neurons)), x0, -1.0, 1.0, norm=norm, num_iterations=1000)
    b.record()
    barrier()
    t_mei = torch.tensor([a.elapsed_time(b)], device=dev)
    tt = torch.cat([t_learn, t_match, t_mei])
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_learn, t_match, t_mei = tt.tolist()
    if rank != 0:
        return None
    fwd_f, bwd_f = algorithmic_flops(res)
    out = {"workload": f"BASELINE configs[3]: random-init Stacked2dCore(1->32, kernels 9/7/7, 3 layers) + FullGaussian2d({n_neurons}) encoder, {res}x{res}",
           "learn": {"metric": f"INR opt steps/sec (learn, {res}x{res}, conv-core response)", "value": world * K / (t_learn / 1e3), "unit": "steps/s", "steps": K,
                     "ms_per_step": t_learn / K, "algorithmic_gflop_inr": (fwd_f + bwd_f) / 1e9, "gpu_launches_per_step": st.kernels_per_step},
           "match": {"metric": "match target-steps/sec", "value": Dloc * world * mk / (t_match / 1e3), "unit": "target-steps/s", "targets_per_gpu": Dloc,
                     "steps": mk, "ms_per_step": t_match / mk},
           "mei": {"metric": "MEIs/sec (1000 ascent iterations each, batched)", "value": n_neurons * world / (t_mei / 1e3), "unit": "MEIs/s",
                   "neurons_per_gpu": n_neurons, "ms": t_mei, "final_activation_mean": float(fev[:, -1].mean().item())},
           "cpu_baseline": None}
    if world == 1 and not args.no_cpu_baseline:
        # the reference's op sequence on the host cores (full-frame convolutions, all neurons read out): torch-CPU port, 2 timed steps
        from laminr_b200.neuron_models.conv_models import export_arrays
        from oracle import laminr_oracle as O
        from oracle.torch_port import ConvResponsePort, LearnPort
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        sd = {k: v.detach().cpu().numpy() for k, v in im.template.state_dict().items()}
        W = [sd[f"func.layer{l}.weight"] for l in range(5)]
        bb = [sd[f"func.layer{l}.bias"] for l in range(5)]
        pos, neg = O.pos_neg_masks(N_LATENTS, 0.1, True)
        port = LearnPort(W, bb, sd["B"], sd["auxB"], sd["coords"].reshape(-1, 2), None, meis[0]["mask"], 1.0, pos, neg, norm=norm, res=(res, res),
                         anomaly_mode=True, response=ConvResponsePort(export_arrays(model), 0))
        g = grids.cpu().numpy()
        port.step(g[0])
        t0 = time.perf_counter()
        n_cpu = 2
        for i in range(n_cpu):
            port.step(g[1 + i])
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_cpu / dt, "unit": "steps/s", "cores": threads, "kind": "port",
                               "sample": f"{n_cpu} learn steps at {res}x{res} with the conv encoder ({dt:.1f} s), torch-CPU port of the reference op sequence, anomaly mode on as shipped"}
    return out


def cpu_mei(res, iters=200):
    """CPU port of the MEI gradient ascent (featurevis/core.py:54-136 as driven by laminr/utils/mei.py:33-59) for ONE neuron (the reference loops
    over neurons): numpy oracle, `iters` of the 1000 iterations timed and extrapolated."""
    from oracle import laminr_oracle as O

    bank = O.demo1_banks([res, res])[0]
    x0 = np.random.RandomState(0).randn(1, 1, res, res).astype(np.float32)
    O.mei_ascent(x0, bank, 10.0, 5, -1.0, 1.0, norm=1.0)
    t0 = time.perf_counter()
    O.mei_ascent(x0, bank, 10.0, iters, -1.0, 1.0, norm=1.0)
    dt = time.perf_counter() - t0
    return {"value": 1.0 / (dt * 1000 / iters), "unit": "MEIs/s", "cores": 1, "kind": "port",
            "sample": f"{iters} of 1000 ascent iterations of one demo1 neuron @{res}x{res} ({dt:.2f} s, extrapolated), numpy oracle"}


def cpu_match_steps(res, n_steps, warmup, threads):
    """CPU port of the match hot step (inv_temp_learn_match.py:352-365) for demo1 targets [1, 2]: the numpy oracle (BLAS threads), a bounded sample."""
    from oracle import laminr_oracle as O

    h, pe, ape, D = 50, 50, 50, 2
    W = [np.random.RandomState(1).randn(h, 1 + 2 * ape + 2 * pe).astype(np.float32) * 0.1] + \
        [np.random.RandomState(2 + l).randn(h, h).astype(np.float32) * 0.1 for l in range(3)] + [np.random.RandomState(9).randn(1, h).astype(np.float32) * 0.1]
    b = [np.zeros(w.shape[0], np.float32) for w in W]
    B = np.random.RandomState(10).randn(2, pe).astype(np.float32) * 10
    auxB = np.random.RandomState(11).randn(2, ape).astype(np.float32) * 0.1
    banks = O.demo1_banks([res, res])[1:3]
    meis = synthetic_meis(res)
    xy = lambda i: np.array([meis[i]["center_pos"]["x"], meis[i]["center_pos"]["y"]], np.float32)
    shifts = O.coords_shifts(xy(0), np.stack([xy(1), xy(2)]), (res, res))
    rs = np.random.RandomState(0)
    aff = dict(angles=rs.uniform(0, 2 * np.pi, D).astype(np.float32), scalings=np.ones((D, 2), np.float32), shears=np.zeros((D, 2), np.float32),
               translation=np.zeros((D, 2), np.float32))
    grid = O.latent_grid(N_LATENTS)
    coords = O.pixel_coords((res, res))
    run = lambda: O.match_step(W, b, B, auxB, grid, coords, (res, res), banks, np.ones(D, np.float32), aff, shifts, 1.0, -1.0, 1.0)
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(n_steps):
        run()
    dt = time.perf_counter() - t0
    return {"value": D * n_steps / dt, "unit": "target-steps/s", "cores": threads, "kind": "port",
            "sample": f"{n_steps} match steps of demo1 targets [1, 2] @{res}x{res} ({dt:.1f} s), numpy oracle (BLAS threads)"}


# ----------------------------------------------------------------------------------------------------------------------
# multi-GPU record: BASELINE configs[2] + configs[4] through the public sharded API (laminr_b200.dist)
# ----------------------------------------------------------------------------------------------------------------------
def bench_scale(args, dev, world, rank, barrier, template_state):
    """ONE shared synthetic population (`--pop` phase-invariant neurons at res x res), the same on every rank:
      * `get_mei_dict_sharded` of ALL neurons (1000 ascent iterations each, RF masks on the device, full dict gathered on every rank);
      * `match_sharded` of template 0 -> targets, STRONG (all pop-1 targets split over the ranks) and WEAK (--match-targets per GPU): template
        broadcast, 60-angle init sweep, a fixed number of epochs with the global 1/D loss scale and the global stop-rule collective every epoch
        (early stop disabled so that every N runs the same epochs), NCCL gather of the aligned images / activations and their copy to host memory
        on rank 0 -- all inside the timed region (CUDA events around the call, max over ranks);
      * sharded == unsharded: rank 0 re-runs `match` unsharded for a fixed subset of targets / `get_mei_dict` for the first neurons from the same
        generator state and compares with what the sharded calls returned.
    Returns the compact `scale_summary` (rank 0) -- printed as the LAST key of the JSON line."""
    import contextlib
    import io

    import torch
    import torch.distributed as dist

    import laminr_b200 as L
    from laminr_b200 import dist as LD

    res, n_pop, epochs = args.res, args.pop, args.match_epochs
    quiet = lambda: (contextlib.redirect_stderr(io.StringIO()), contextlib.redirect_stdout(io.StringIO()))

    def timed(fn):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, t.item() / 1e3

    t0 = time.perf_counter()
    pop = LD.simulated_population_sharded(n_pop, (res, res), seed=0, device=dev)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    kw = dict(required_img_norm=1.0)
    # ---- MEI sweep (BASELINE configs[4])
    with quiet()[0], quiet()[1]:
        LD.get_mei_dict_sharded(pop, [1, res, res], -1.0, 1.0, neuron_idxs=list(range(2 * world)), num_iterations=5, **kw)  # warm-up
        torch.manual_seed(0)
        meis, mei_s = timed(lambda: LD.get_mei_dict_sharded(pop, [1, res, res], -1.0, 1.0, **kw))
    # the ascent kernel alone on this rank's shard of the bank (N = 1: all `--pop` neurons, 819 MB at 1024 x 100x100 -> truly HBM-bound)
    from laminr_b200.utils.mei import batched_simulated_meis
    lo, hi = LD.shard_bounds(n_pop, world, rank)
    x0 = torch.randn(hi - lo, 1, res, res, device=dev)
    batched_simulated_meis(pop, list(range(lo, hi)), x0, -1.0, 1.0, norm=1.0, num_iterations=5)
    _, kern_s = timed(lambda: batched_simulated_meis(pop, list(range(lo, hi)), x0, -1.0, 1.0, norm=1.0, num_iterations=200))
    mei_kernel = {"neurons": hi - lo, "us_per_iteration": kern_s / 200 * 1e6,
                  "gbps": (20 * res * res * 4 + 3 * res * res * 4) * 200 * (hi - lo) / kern_s / 1e9}  # F*P*4 + 3*C*P*4 per neuron-iteration
    mei_diff = None
    if rank == 0:  # unsharded call for the first neurons from the same seed: the same initial images, so the same MEIs
        torch.manual_seed(0)
        with quiet()[0], quiet()[1]:
            few = L.get_mei_dict(pop, [1, res, res], -1.0, 1.0, neuron_idxs=[0, 1, 2, 3], **kw)
        mei_diff = max(float(np.abs(np.asarray(few[k]["mei"]) - np.asarray(meis[k]["mei"])).max()) for k in few)
    # ---- match (BASELINE configs[2]): template = neuron 0, learnt on rank 0 (the other ranks hold an untrained copy: overwritten by the broadcast)
    torch.manual_seed(0)
    im = L.InvarianceManifold(pop, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=args.precision)
    if rank == 0:
        im.template.load_state_dict(template_state, strict=False)
    im.template_neuron_idx = 0

    def run(targets, ep):
        state = torch.get_rng_state()
        with quiet()[0], quiet()[1]:
            (imgs, acts), secs = timed(lambda: LD.match_sharded(im, targets, num_epochs=ep, patience=10 ** 9))
        return imgs, acts, secs, state

    run(list(range(1, 1 + 2 * world)), 2)  # warm-up: NCCL channels, allocator, kernel attributes
    weak_t = list(range(1, 1 + min(args.match_targets * world, n_pop - 1)))
    _, _, weak_s, _ = run(weak_t, epochs)
    weak_epochs = int(im.match_epochs_run)
    strong_t = list(range(1, n_pop))
    imgs, acts, strong_s, state = run(strong_t, epochs)
    strong_epochs, steps = int(im.match_epochs_run), int(im.match_steps_taken)
    match_diff, act_mean = None, None
    if rank == 0:  # unsharded == sharded on a fixed subset (same generator state -> same jitter; same global 1/D loss scale)
        subset = sorted({1, 2, 3, n_pop // 3, n_pop // 2, (2 * n_pop) // 3, n_pop - 2, n_pop - 1} & set(strong_t))
        torch.set_rng_state(state)
        with quiet()[0], quiet()[1]:
            want_i, want_a = im.match(subset, num_epochs=epochs, patience=10 ** 9, _total_targets=len(strong_t))
        idx = [t - 1 for t in subset]
        match_diff = max(float(np.abs(want_i - imgs[idx]).max()), float(np.abs(want_a - acts[idx]).max()))
        act_mean = float(np.asarray(acts).mean())
    barrier()
    if rank != 0:
        return None
    r3 = lambda x: float(f"{x:.4g}")
    return {"n_gpus": world, "population": n_pop, "res": res, "epochs": epochs, "api": "dist.match_sharded / get_mei_dict_sharded, CUDA events, max over ranks",
            "match_strong": {"targets": len(strong_t), "neurons_per_s": r3(len(strong_t) / strong_s), "s": r3(strong_s), "epochs_run": strong_epochs,
                             "target_steps_per_s_incl_sweep_and_gather": r3(len(strong_t) * steps / strong_s)},
            "match_weak": {"targets_per_gpu": len(weak_t) // world, "neurons_per_s": r3(len(weak_t) / weak_s), "s": r3(weak_s), "epochs_run": weak_epochs},
            "mei": {"neurons": n_pop, "meis_per_s": r3(n_pop / mei_s), "s": r3(mei_s)},
            "max_abs_sharded_minus_unsharded": {"match_imgs_acts": match_diff, "meis": mei_diff}, "mean_act_norm": r3(act_mean),
            "population_build_s": r3(build_s), "mei_kernel": mei_kernel}


# ----------------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the laminr_b200 path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import laminr_b200 as L
    from laminr_b200.contrastive_loss import SimCLROnGrid
    from laminr_b200.fused import LearnStepper, MatchStepper
    from laminr_b200.utils.training import JitteringGridDatamodule

    res, K, Wm = args.res, args.steps, max(args.warmup, 3)
    torch.manual_seed(0)
    np.random.seed(0)
    model = L.neuron_models.simulated("demo1", img_res=[res, res]).to(dev)
    meis = synthetic_meis(res)
    im = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=args.precision)
    reg = SimCLROnGrid(1, N_LATENTS, 0.1, 0.3, True).to(dev)
    stepper = LearnStepper(im.template, im.img_transforms, model, 0, meis[0]["mask"], meis[0]["activation"], reg, -1.0, 1.0, N_LATENTS, lr=1e-3,
                           reg_scale=2.0, precision=args.precision, use_graph=True)
    dm = JitteringGridDatamodule(1, N_LATENTS, K + Wm)
    grids_host = torch.stack(list(dm.train_dataloader())).reshape(K + Wm, N_LATENTS).pin_memory()
    grids = grids_host.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs, per-step event pairs, L2 flushed between steps
    for i in range(Wm):
        stepper.step(grids[i])
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    for i in range(K):
        flush.fill_(float(i))
        starts[i].record()
        stepper.step(grids[Wm + i])
        ends[i].record()
    barrier()
    step_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    # back-to-back (no flush): what the training loop actually sees
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        stepper.step(grids[Wm + i])
    e1.record()
    barrier()
    b2b_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler is not None else None

    # ---- e2e: host grid -> step -> loss on the host, every step, through LearnStepper.step with a HOST row: the row is written into pinned memory
    #      that the step's first kernel reads over PCIe (80 bytes host -> device), the objective kernel writes the loss into pinned memory (4 bytes
    #      device -> host), sync_loss() waits for the step and returns the float.  `copy_engine` = the same loop with cudaMemcpyAsync H2D / D2H
    #      around the device-fed step (what round 1 and 2 a-f reported as e2e).
    for i in range(3):
        stepper.step(grids_host[i])
        stepper.sync_loss()
    barrier()
    t0 = time.perf_counter()
    e2e_loss = 0.0
    for i in range(K):
        stepper.step(grids_host[Wm + i])   # CPU tensor: host-fed step
        e2e_loss = stepper.sync_loss()     # the step's loss as a Python float
    e2e_s = time.perf_counter() - t0
    barrier()
    if not np.isfinite(e2e_loss):
        raise SystemExit("bench.py: the host-fed step returned a non-finite loss")
    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    t0 = time.perf_counter()
    for i in range(K):
        stepper.grid.copy_(grids_host[Wm + i], non_blocking=True)  # H2D of the step's jittered latents from pinned memory into the static input
        stepper.graph.replay()
        loss_host.copy_(stepper.loss().reshape(1), non_blocking=True)  # D2H of the step's loss
        torch.cuda.current_stream().synchronize()
    e2e_copy_s = time.perf_counter() - t0
    barrier()
    # extra, NOT the e2e value: the same copies every step, but the loss of step k is read (two pinned slots, one event each) after step k + 1 has
    # been queued, so the GPU does not wait for the host's round trip -- how a host loop that only logs the loss would be written
    loss2 = torch.zeros(2, dtype=torch.float32).pin_memory()
    done = [torch.cuda.Event(), torch.cuda.Event()]
    acc_loss = 0.0
    t0 = time.perf_counter()
    for i in range(K):
        stepper.grid.copy_(grids_host[Wm + i], non_blocking=True)
        stepper.graph.replay()
        loss2[i & 1:(i & 1) + 1].copy_(stepper.loss().reshape(1), non_blocking=True)
        done[i & 1].record()
        if i > 0:
            done[(i - 1) & 1].synchronize()
            acc_loss += float(loss2[(i - 1) & 1])
    done[(K - 1) & 1].synchronize()
    acc_loss += float(loss2[(K - 1) & 1])
    e2e_lag_s = time.perf_counter() - t0
    barrier()

    # ---- the public call itself (extra): InvarianceManifold.learn(0) for a fixed number of epochs -- stepper construction, graph capture, the
    #      per-epoch control logic (scheduler, evaluation forward, requirement check, progress bar) and the final snapshot to the host included
    api = None
    if rank == 0:
        import contextlib, io
        torch.manual_seed(0)
        im_api = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=args.precision)
        with contextlib.redirect_stderr(io.StringIO()):
            # the process's FIRST learn() also pays torch's one-time lazy loading of the reductions the control logic uses (mean / min / std /
            # stack: ~100 ms, tools/learn_firstcall.py); reported separately, the value is a later call on a fresh manifold object
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            im_api.learn(0, num_max_epochs=20)
            torch.cuda.synchronize()
            dt_first = time.perf_counter() - t0
            torch.manual_seed(0)
            im_api = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=args.precision)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            imgs_api, _ = im_api.learn(0, num_max_epochs=20)
            torch.cuda.synchronize()
            dt_api = time.perf_counter() - t0
        api = {"call": "InvarianceManifold.learn(0, num_max_epochs=20)", "value": im_api.learn_steps_taken / dt_api, "unit": "steps/s",
               "seconds": dt_api, "first_call_in_process_seconds": dt_first, "steps": int(im_api.learn_steps_taken),
               "epochs": int(im_api.learn_epochs_run), "result_shape": list(imgs_api.shape)}
    barrier()

    # ---- roofline of the dominant launch group (INR backward), timed alone with CUDA events
    from laminr_b200 import ops
    t = im.template
    gx = stepper.g_x.view(1, N_LATENTS, 1, res * res)
    keep = bool(getattr(stepper, "keep", False))  # as in the step: the forward keeps the hidden activations' operand tiles, the backward loads them
    fused_tail = keep and hasattr(stepper, "step_count")  # the step's own call: tile kernel + ONE cooperative tail launch (reductions + Adam)
    def bwd_only():
        if fused_tail:
            ops.raw_inr_backward_adam(t._desc, stepper.flat, t.B, t.auxB, stepper.grid, stepper.coords, None, gx, stepper.g_flat, stepper.m, stepper.v,
                                      stepper.step_count, lr=stepper.lr, precision=args.precision, workspace=stepper.ws_bwd, forward_workspace=stepper.ws_fwd,
                                      keep_activations=keep)
        else:
            ops.raw_inr_backward(t._desc, stepper.flat, t.B, t.auxB, stepper.grid, stepper.coords, None, None, gx, want_params=True,
                                 precision=args.precision, workspace=stepper.ws_bwd, g_params=stepper.g_flat,
                                 forward_workspace=stepper.ws_fwd if keep else None, keep_activations=keep)
    def fwd_only():
        ops.raw_inr_forward(t._desc, stepper.flat, t.B, t.auxB, stepper.grid, stepper.coords, None, None, out=stepper.img_pre.view(1, N_LATENTS, 1, res * res),
                            precision=args.precision, workspace=stepper.ws_fwd, keep_activations=keep)
    fwd_only()  # the backward below reads this forward's workspace
    def time_fn(fn, reps=30):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        # the launch group is replayed from a CUDA graph, as in the step: an eager replay from Python would time the host's launch calls
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for r in range(reps):
            flush.fill_(float(r))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / reps
    fwd_ms = time_fn(fwd_only)
    saved = [x.clone() for x in (stepper.flat, stepper.m, stepper.v, stepper.step_count)] if fused_tail else None  # the fused tail applies Adam: put the
    bwd_ms = time_fn(bwd_only)                                                                                    # template back afterwards
    if saved is not None:
        for dst, src in zip((stepper.flat, stepper.m, stepper.v, stepper.step_count), saved):
            dst.copy_(src)

    # ---- demo1 learn step at the second BASELINE resolution (200x200), back to back from the CUDA graph (rank 0)
    big = None
    if rank == 0 and args.res2 > 0:
        r2 = args.res2
        model2 = L.neuron_models.simulated("demo1", img_res=[r2, r2]).to(dev)
        meis2 = synthetic_meis(r2)
        torch.manual_seed(0)
        imb = L.InvarianceManifold(model2, meis2, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=args.precision)
        stb = LearnStepper(imb.template, imb.img_transforms, model2, 0, meis2[0]["mask"], meis2[0]["activation"], reg, -1.0, 1.0, N_LATENTS, lr=1e-3,
                           reg_scale=2.0, precision=args.precision, use_graph=True)
        nb = max(3, min(K, 50))
        for i in range(4):
            stb.step(grids[i])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(nb):
            stb.step(grids[Wm + (i % K)])
        b.record()
        torch.cuda.synchronize()
        big = {"res": r2, "ms_per_step": a.elapsed_time(b) / nb, "steps": nb, "launches_per_step": stb.kernels_per_step}
        del stb, imb, model2
        torch.cuda.empty_cache()
    barrier()

    # ---- match hot step (extra): targets sharded over ranks, weak scaling (fixed targets per GPU)
    match = None
    if args.match_targets > 0:
        Dloc = args.match_targets
        pop = L.neuron_models.simulated_population(Dloc + 1, img_res=[res, res], seed=1000 + rank).to(dev)
        yy, xx = np.meshgrid(np.arange(res) + 0.5, np.arange(res) + 0.5, indexing="ij")
        pmeis = {}
        rs = np.random.RandomState(rank)
        for i in range(Dloc + 1):
            cx, cy = rs.uniform(0.3, 0.7, 2) * res
            mask = np.exp(-0.5 * (((xx - cx) / (0.15 * res)) ** 2 + ((yy - cy) / (0.15 * res)) ** 2)).astype(np.float32)
            pmeis[i] = dict(mei=np.zeros((1, res, res), np.float32), activation=1.0000009536743164, mask=mask, center_pos=dict(x=cx, y=cy))
        im2 = L.InvarianceManifold(pop, pmeis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision=args.precision)
        im2.template.load_state_dict(im.template.state_dict())
        targets = list(range(1, Dloc + 1))
        im2.template.initialize_coordinate_transform(Dloc, True, False, 0.1, True, True, False, True)
        rfs = torch.tensor([[pmeis[i]["center_pos"]["x"], pmeis[i]["center_pos"]["y"]] for i in range(Dloc + 1)], dtype=torch.float32, device=dev)
        im2.template.register_coords_shifts(rfs[0], rfs[1:])
        aff = im2.template.coordinate_transform.transforms.Affine
        with torch.no_grad():
            aff.angles.copy_(torch.rand(Dloc, device=dev) * 2 * np.pi)
        ms = MatchStepper(im2.template, im2.img_transforms, pop, targets, np.ones(Dloc, np.float32), N_LATENTS, total_targets=Dloc * world,
                          precision=args.precision, use_graph=True)
        mk = max(3, min(K, args.match_steps))
        for i in range(3):
            ms.step(grids[i])
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(mk):
            ms.step(grids[Wm + (i % K)])
        b.record()
        barrier()
        mt = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        match = {"metric": "match target-steps/sec (one Adam step of one target's affine transform)", "value": Dloc * world * mk / (mt.item() / 1e3),
                 "unit": "target-steps/s", "targets_per_gpu": Dloc, "steps": mk, "ms_per_step": mt.item() / mk,
                 "algorithmic_gflop_per_target_step": 2 * algorithmic_flops(res)[0] / 1e9}
        # ---- whole match() call (BASELINE metric 2: target neurons aligned/sec): 60-angle init sweep + a FIXED number of epochs (so that the number is
        #      not hostage to the early-stop) + final snapshot and D2H of the aligned images; targets sharded over the ranks, global 1/D and stop rule
        im2.template_neuron_idx = 0
        epochs = args.match_epochs
        import contextlib, io
        if world == 1:  # (N > 1: the whole-call record goes through dist.match_sharded on one shared population, see scale_summary)
            with contextlib.redirect_stderr(io.StringIO()):
                # a process's first match() also pays one-time costs that are not the call's (torch's lazy kernel loading, the pinned staging ring,
                # pages of a fresh box: 0.19 s vs 0.57 s for the same call was seen): one 2-target, 1-epoch call first; the timed call still
                # builds its own stepper, sweeps, captures its graphs and copies its results
                im2.match(targets[:2], num_epochs=1, patience=10 ** 9)
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            prof = None
            if os.environ.get("LMR_BENCH_PROFILE"):  # debug: host profile of this call on stderr
                import cProfile
                prof = cProfile.Profile()
                prof.enable()
            with contextlib.redirect_stderr(io.StringIO()):  # tqdm bars
                imgs, acts_m = im2.match(targets, num_epochs=epochs, patience=10 ** 9)
            torch.cuda.synchronize()
            mt2 = time.perf_counter() - t0
            if prof is not None:
                import pstats
                prof.disable()
                pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(14)
            match["aligned"] = {"metric": "target neurons aligned/sec (whole match(): 60-angle sweep + epochs + result copy to host)", "value": Dloc / mt2,
                                "unit": "neurons/s", "epochs": int(im2.match_epochs_run), "seconds": mt2, "targets_total": Dloc, "result_shape": list(imgs.shape)}
        # ---- MEI sweep (BASELINE config 5): batched gradient ascent, 1000 iterations, neurons sharded over the ranks
        nmei = args.mei_neurons  # BASELINE configs[4]: 128 neurons per GPU
        if nmei > 0:
            from laminr_b200.utils.mei import batched_simulated_meis
            if nmei != Dloc + 1:
                pop = L.neuron_models.simulated_population(nmei, img_res=[res, res], seed=2000 + rank).to(dev)
            x0 = torch.randn(nmei, 1, res, res).to(dev)
            batched_simulated_meis(pop, list(range(nmei)), x0, -1.0, 1.0, norm=1.0, num_iterations=10)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _, fev = batched_simulated_meis(pop, list(range(nmei)), x0, -1.0, 1.0, norm=1.0, num_iterations=1000)
            b.record()
            barrier()
            mt3 = torch.tensor([a.elapsed_time(b)], device=dev)
            if world > 1:
                dist.all_reduce(mt3, op=dist.ReduceOp.MAX)
            mei_bytes = (20 * res * res * 4 + 3 * res * res * 4) * 1000 * nmei  # F*P*4 + 3*C*P*4 per neuron-iteration
            match["mei"] = {"metric": "MEIs/sec (1000 ascent iterations each, batched)", "value": nmei * world / (mt3.item() / 1e3), "unit": "MEIs/s",
                            "neurons_per_gpu": nmei, "ms": mt3.item(), "final_activation_mean": float(fev[:, -1].mean().item()),
                            "algorithmic_GBps_per_gpu": mei_bytes / (mt3.item() / 1e3) / 1e9,
                            "cpu_baseline": cpu_mei(res) if (world == 1 and not args.no_cpu_baseline) else None,
                            "note": "the per-GPU filter bank (neurons x 20 x HW x 4 B) is re-read every iteration; at 128 neurons x 100x100 it is 102 MB and mostly "
                                    "L2-resident (126 MB L2), so the algorithmic rate can exceed the HBM copy peak"}

    # ---- BASELINE configs[1] (extra, N == 1): demo1 match stage, template 0 -> targets [1, 2] at res x res on one B200, whole match() call
    if match is not None and world == 1:
        import contextlib, io
        im.template_neuron_idx = 0
        ep2 = 100
        with contextlib.redirect_stderr(io.StringIO()):
            im.match([1, 2], num_epochs=3, patience=10 ** 9)  # warm-up: allocations, graph capture
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            imgs2, _ = im.match([1, 2], num_epochs=ep2, patience=10 ** 9)
            torch.cuda.synchronize()
            dt2 = time.perf_counter() - t0
            ms2 = im._match_stepper
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(50):
                ms2.step(grids[Wm + (i % K)])
            b.record()
            torch.cuda.synchronize()
        match["config2"] = {"workload": f"BASELINE configs[1]: demo1, template 0 -> targets [1, 2], {res}x{res}, 1 GPU",
                            "aligned": {"value": 2 / dt2, "unit": "neurons/s", "seconds": dt2, "epochs": int(im.match_epochs_run),
                                        "note": "whole match(): 60-angle sweep + a fixed number of epochs (early stop disabled) + result copy to the host",
                                        "result_shape": list(imgs2.shape)},
                            "step": {"value": 2 * 50 / (a.elapsed_time(b) / 1e3), "unit": "target-steps/s", "ms_per_step": a.elapsed_time(b) / 50}}
        if not args.no_cpu_baseline:
            match["config2"]["cpu_baseline"] = cpu_match_steps(res, 2, 1, os.cpu_count() or 1)

    # ---- BASELINE config 4 (extra): conv-core encoder at 200x200
    conv = bench_conv(args, dev, world, rank, barrier, flush) if args.conv_res > 0 else None

    # ---- the multi-GPU record (every N, N = 1 included: the anchor of the scaling curve): config 3 through the public sharded API
    del flush
    torch.cuda.empty_cache()
    scale_summary = bench_scale(args, dev, world, rank, barrier, {k: v.detach().clone() for k, v in im.template.state_dict().items()
                                                                  if k.startswith("func.") or k in ("B", "auxB")}) if args.pop > 1 else None

    # ---- reduce over ranks (max time), CPU baseline on rank 0 at N == 1
    tt = torch.tensor([step_ms, b2b_ms, e2e_s * 1e3], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_ms, b2b_ms, e2e_ms = tt.tolist()
    if rank == 0:
        fwd_f, bwd_f = algorithmic_flops(res)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        peak_tf = peaks.get("bf16_tflops", 1590.0)  # burst figure: the launch group is timed alone
        achieved = bwd_f / (bwd_ms * 1e-3) / 1e12
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            n_cpu = 40 if res <= 100 else 12
            sps, dt, kind, where = cpu_learn_steps(res, n_cpu, 2, threads)
            sps_off, _, _, _ = cpu_learn_steps(res, max(4, n_cpu // 2), 1, threads, anomaly=False)
            cpu = {"value": sps, "unit": "steps/s", "cores": threads, "kind": kind, "value_anomaly_mode_off": sps_off,
                   "sample": f"{n_cpu} learn steps of demo1 @{res}x{res} ({dt:.1f} s), "
                             f"{'unmodified reference (baseline/_ref) through InvarianceManifold.learn(0, steps_per_epoch=%d, num_max_epochs=1)' % n_cpu if kind == 'reference' else 'torch-CPU port of the reference op sequence'}, anomaly mode on as shipped"}
        # every throughput as a fraction of its roofline (north star): tensor-bound lines against the SUSTAINED dense bf16 peak (kernels timed
        # inside long steps), HBM-bound lines against the measured copy bandwidth; algorithmic work per SURVEY 8d (recompute / x3 split / padding
        # not counted)
        peak_sus, peak_hbm = peaks.get("bf16_tflops_sustained", 1386.3), peaks.get("hbm_gbs", 6549.4)
        tf = lambda flop, ms: {"ms": ms, "tflops": flop / (ms * 1e-3) / 1e12, "frac": flop / (ms * 1e-3) / 1e12 / peak_sus}
        roofline_all = {"peaks": {"bf16_tflops_sustained": peak_sus, "hbm_gbs": peak_hbm, "source": "MEASURED_PEAKS.json" if peaks else "fallback"},
                        f"learn_step_{res}": dict(tf(fwd_f + bwd_f, b2b_ms / K), bound="tensor")}
        if big is not None:
            f2, b2 = algorithmic_flops(big["res"])
            roofline_all[f"learn_step_{big['res']}"] = dict(tf(f2 + b2, big["ms_per_step"]), bound="tensor", steps_per_s=1e3 / big["ms_per_step"])
        if match is not None:
            roofline_all[f"match_target_step_{res}"] = dict(tf(2 * fwd_f, match["ms_per_step"] / match["targets_per_gpu"]), bound="tensor")
            if "mei" in match:
                roofline_all[f"mei_{match['mei']['neurons_per_gpu']}_neurons_{res}"] = {
                    "bound": "hbm (bank mostly L2-resident at this size)", "gbps": match["mei"]["algorithmic_GBps_per_gpu"],
                    "frac": match["mei"]["algorithmic_GBps_per_gpu"] / peak_hbm}
        if conv is not None:
            fc, bc = algorithmic_flops(args.conv_res)
            roofline_all[f"conv_learn_step_{args.conv_res}"] = dict(tf(fc + bc, conv["learn"]["ms_per_step"]), bound="tensor (INR part; the cone kernel is fp32 CUDA-core)")
            roofline_all[f"conv_match_target_step_{args.conv_res}"] = dict(tf(2 * fc, conv["match"]["ms_per_step"] / conv["match"]["targets_per_gpu"]), bound="tensor (INR part)")
        if scale_summary is not None and scale_summary.get("mei_kernel"):
            mk_ = scale_summary.pop("mei_kernel")
            roofline_all[f"mei_{mk_['neurons']}_neurons_{res}"] = {"bound": "hbm", "gbps": mk_["gbps"], "frac": mk_["gbps"] / peak_hbm, "us_per_iteration": mk_["us_per_iteration"]}
        line = {
            "metric": "INR opt steps/sec (learn, 100x100)", "value": world * K / (step_ms / 1e3), "unit": "steps/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": step_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision in (None, "fp32") else args.precision,
            "data": "synthetic",
            "config": workload_config(res),
            "notes": {"precision": args.precision or "fp32", "l2": "flushed (256 MiB write) between timed steps, outside the event pairs",
                      "multi_gpu": "value: replicas only (learn has one template); the sharded match / MEI record is scale_summary", "cuda_graph": True},
            "back_to_back": {"value": world * K / (b2b_ms / 1e3), "unit": "steps/s", "ms_per_step": b2b_ms / K, "note": "K graph replays in one event pair, no L2 flush (the loop as it runs)"},
            "e2e": {"value": world * K / (e2e_ms / 1e3), "unit": "steps/s", "h2d_bytes_per_step": 4 * N_LATENTS, "d2h_bytes_per_step": 4,
                    "sync": "the host waits for every step's loss before it queues the next step",
                    "path": "LearnStepper.step(host row) + sync_loss(): pinned memory read / written by the step's kernels, one graph launch per step",
                    "copy_engine": {"value": K / e2e_copy_s, "unit": "steps/s per GPU (this rank)",
                                    "note": "same loop with cudaMemcpyAsync H2D / D2H around the device-fed step"},
                    "loss_read_one_step_late": {"value": K / e2e_lag_s, "unit": "steps/s per GPU (this rank)", "mean_loss": acc_loss / K,
                                                "note": "extra: same H2D + D2H every step, the loss of step k read while step k + 1 runs"}},
            "gpu_launches": int(stepper.kernels_per_step * K),
            "roofline": {"bound": "tensor", "kernel": ("lmr_inr_backward_adam launch group: inr_bwd_pp_kernel (dominant) + learn_tail_kernel (layer-0 weight gradient, partial reductions, Adam)"
                                                       if fused_tail else "lmr_inr_backward launch group (tile kernel + layer-0 weight gradient and partial reduction launches)"), "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf,
                         # dram__bytes_read.sum + dram__bytes_write.sum of the tile kernel, one `ncu --set full` capture of this workload:
                         # recompute variant 2.995 MB (profiles/r1b_ncu_inr_bwd_tc.md); kept-activation variant: the 160 MB of operand tiles it loads
                         # (profiles/r2g_ncu_learn_step.md)
                         "traffic": (TRAFFIC_BWD_KEEP if keep else 2995456) if (res == 100 and args.precision != "fp32") else None,
                         "traffic_unit": "bytes per launch (ncu, DRAM)", "kept_activations": keep,
                         "peak_source": "measured (MEASURED_PEAKS.json bf16 burst)" if peaks else "fallback",
                         # what the low fraction is made of: the split-bf16 operands cost three tensor-core MACs per algorithmic MAC (the
                         # tensor pipe therefore works at 3x `achieved`), and ncu has the pipe active for 29 % of the tile kernel
                         # (profiles/r2g_ncu_learn_step.md: sm__pipe_tensor_cycles_active, cold-cache replay; 32 % of the warm kernel time)
                         "operand_split": "bf16x3" if args.precision != "fp32" else None,
                         "issued_tflops": 3.0 * achieved if args.precision != "fp32" else None,
                         "tensor_pipe_active_pct_ncu": 29.0 if (res == 100 and args.precision == "bf16x3" and keep) else None,
                         "algorithmic_gflop": bwd_f / 1e9, "ms": bwd_ms, "forward": {"algorithmic_gflop": fwd_f / 1e9, "ms": fwd_ms, "achieved": fwd_f / (fwd_ms * 1e-3) / 1e12}},
            "cpu_baseline": cpu, "clocks": clocks, "api_learn": api, "roofline_all": roofline_all, "match": match, "conv": conv,
            "scale_summary": scale_summary,  # LAST key on purpose: it must survive a truncated tail of this line at every N
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--res", type=int, default=100)
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16x3_fast"],
                    help="INR hidden-layer arithmetic: bf16x3 = tcgen05 split-bf16 x3 (fp32-class, default), fp32 = CUDA cores")
    ap.add_argument("--match-targets", type=int, default=32, help="targets per GPU for the extra match measurement (0 = skip)")
    ap.add_argument("--match-steps", type=int, default=20)
    ap.add_argument("--match-epochs", type=int, default=30, help="epochs of the whole-match() measurement (fixed; the early stop is disabled)")
    ap.add_argument("--mei-neurons", type=int, default=128, help="neurons per GPU in the MEI sweep measurement (BASELINE configs[4]: 128 per GPU)")
    ap.add_argument("--res2", type=int, default=200, help="second resolution of the demo1 learn step (BASELINE: 100x100 and 200x200; 0 = skip)")
    ap.add_argument("--pop", type=int, default=1024, help="neurons of the shared synthetic population of the multi-GPU record (BASELINE configs[2]; 0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--conv-res", type=int, default=200, help="resolution of the extra conv-core (BASELINE config 4) measurement (0 = skip)")
    ap.add_argument("--conv-steps", type=int, default=100)
    ap.add_argument("--conv-targets", type=int, default=8, help="match targets per GPU in the conv-core measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
