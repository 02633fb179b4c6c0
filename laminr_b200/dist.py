"""Multi-GPU sharding of the match stage and the MEI sweep: one process per GPU (`torchrun`), `torch.distributed` over NCCL.

The reference has no distributed code; what shards naturally:
  * `InvarianceManifold.match` -- target neuron d's 7 affine scalars interact with other targets only through the `1/D` scale of
    the loss (laminr/inv_temp_learn_match.py:363) and the GLOBAL early-stop mean (:368).  Each rank aligns a contiguous block of
    targets; the loss keeps the global 1/D (so per-parameter gradients, and Adam's epsilon, see exactly the unsharded values) and the
    stop rule sees the all-reduced (sum, min, max) of the normalised activations -- one 3-float all-reduce per epoch.  Every rank then
    runs the same number of epochs, draws the same jitter (rank 0's CPU RNG state is broadcast), and the sharded result equals the
    unsharded one per target.
  * `get_mei_dict` -- fully independent per neuron (laminr/utils/mei.py:98-119).
  * `learn` does NOT shard (one template, one 20-image batch): template learning runs on one GPU and the template is broadcast.
Collectives: broadcast of the template state (~72 KB) before, gather of images / activations after.  Nothing on the per-step path.

The plumbing functions (`shard_bounds`, `broadcast_module_state`, `make_stop_reduce`, `gather_ragged`) take plain tensors and any
backend, so they are covered by world_size-2 gloo tests on CPU; the compute itself is CUDA-only.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, rank):
    """Contiguous block [lo, hi) of rank `rank`: the first n % world ranks get one extra item."""
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def broadcast_module_state(module, src=0):
    """In-place broadcast of every parameter and buffer of `module` from rank `src` (the template INR after `learn`)."""
    world, _ = _world()
    if world == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


def broadcast_template(template, src=0):
    """Broadcast of what `learn` produced: the INR network `template.func` and the fixed buffers B / auxB / coords, in a member-independent
    order.  The coordinate transform and the RF-shift buffers are NOT sent: every rank rebuilds them in `match` (a rank that has already run an
    unsharded `match` owns members the others lack, and a per-member broadcast of the whole module would then hang on mismatched collectives)."""
    world, _ = _world()
    if world == 1:
        return
    for t in list(template.func.parameters()) + [template.B, template.auxB, template.coords]:
        dist.broadcast(t.data, src=src)


def broadcast_rng_state(src=0, device=None):
    """Every rank continues with rank `src`'s CPU generator state: the jittered latent grids (laminr/utils/training.py:42-46) and MEI
    starting images (laminr/utils/mei.py:33) are drawn on the CPU generator, so the ranks must agree on it."""
    world, rank = _world()
    if world == 1:
        return
    state = torch.get_rng_state()
    buf = state.clone().to(device) if (device is not None and dist.get_backend() == "nccl") else state.clone()
    dist.broadcast(buf, src=src)
    torch.set_rng_state(buf.cpu())


def make_stop_reduce(device=None):
    """-> f(sum, min, max) all-reducing the three floats over the ranks (SUM, MIN, MAX): the global stop rule of `match` (:368-375).
    f.begin(t3) / f.end(handle) split the call for a device-resident 3-element tensor: `begin` queues the collective and the copy to pinned host
    memory behind the work queued so far and returns at once; `end` waits for that copy and combines -- the match loop calls `end` one epoch
    later, so neither the collective nor the copy ever stalls the GPU."""
    world, _ = _world()
    if world == 1:
        return None

    def combine(allr):
        allr = allr.double()
        return float(allr[:, 0].sum()), float(allr[:, 1].min()), float(allr[:, 2].max())

    def reduce(a_sum, a_min=None, a_max=None):
        # ONE collective per epoch: every rank gets every rank's (sum, min, max) and combines them in rank order (deterministic).
        # Called with three host floats, or with ONE 3-element tensor (the fused stepper's device-resident epoch statistics: the collective then
        # runs device to device and the epoch costs a single copy to the host).
        if torch.is_tensor(a_sum) and a_min is None:
            t = a_sum.detach().reshape(3)
        else:
            t = torch.tensor([a_sum, a_min, a_max], dtype=torch.float64, device=device)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return combine(torch.stack(parts).cpu())

    def begin(t3):
        t = t3.detach().reshape(3)
        if dist.get_backend() != "nccl":  # gloo (CPU tests of the plumbing; GPU tests on a one-GPU box): host tensors, blocking
            parts = [torch.empty(3, dtype=t.dtype) for _ in range(world)]
            dist.all_gather(parts, t.cpu())
            return torch.stack(parts), None
        full = torch.empty((world, 3), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, t.contiguous())
        host = torch.empty((world, 3), dtype=t.dtype).pin_memory()
        host.copy_(full, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return host, ev, full  # (`full` is kept alive until the copy has run)

    def end(handle):
        if handle[1] is not None:
            handle[1].synchronize()
        return combine(handle[0])

    reduce.begin, reduce.end = begin, end
    return reduce


def gather_ragged(local, n_total, dst=0):
    """Concatenates per-rank blocks along dim 0 (block sizes from `shard_bounds`) on rank `dst` (returns None elsewhere), or on EVERY rank with
    `dst=None`.  `local` is this rank's block (may be empty) of a tensor whose dim-0 length over all ranks is `n_total`."""
    world, rank = _world()
    if world == 1:
        return local
    sizes = [shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world)]
    nmax = max(sizes)
    if local.shape[0] == nmax:
        pad = local.contiguous()
    else:
        pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    if dist.get_backend() == "nccl":
        # ONE all-gather on the communicator's established channels.  (Point-to-point sends to rank `dst` open new peer connections on first use:
        # measured 1.65 s for 7 x 102 MB on an 8 x B200 box, against ~3 ms of NVLink time; the all-gather costs every rank a transient copy of the
        # result, 0.8 GB for 1023 targets at 100x100, which only `dst` keeps.)
        full = torch.empty((world * nmax,) + tuple(pad.shape[1:]), dtype=pad.dtype, device=pad.device)
        dist.all_gather_into_tensor(full, pad)
        if dst is not None and rank != dst:
            return None
        if all(sz == nmax for sz in sizes):
            return full
        return torch.cat([full[r * nmax: r * nmax + sizes[r]] for r in range(world)], dim=0)
    # gloo (CPU tests of the plumbing) gathers host tensors only
    dev = pad.device
    if dst is None:
        host = [torch.empty_like(pad, device="cpu") for _ in range(world)]
        dist.all_gather(host, pad.cpu())
    else:
        host = [torch.empty_like(pad, device="cpu") for _ in range(world)] if rank == dst else None
        dist.gather(pad.cpu(), gather_list=host, dst=dst)
        if rank != dst:
            return None
    return torch.cat([host[r][: sizes[r]].to(dev) for r in range(world)], dim=0)


_SAME_HOST = {}
_SHM_COUNT = [0]


def _same_host():
    """Do all ranks of the default group run on one machine (one NVSwitch box)?  Decided once per process group from the host names."""
    world, _ = _world()
    key = (id(dist.group.WORLD), world)
    if key not in _SAME_HOST:
        import socket
        names = [None] * world
        dist.all_gather_object(names, socket.gethostname())
        _SAME_HOST[key] = len(set(names)) == 1
    return _SAME_HOST[key]


class HostGather:
    """The per-rank blocks of a device tensor (dim 0 split by `shard_bounds`) as ONE NumPy array on rank 0, in two collective phases:

      g = HostGather(n_total, row_shape, dtype)   BEFORE the compute: ranks on one machine share a POSIX shared-memory segment of the result's size
                                                  (created by rank 0), and every rank starts touching the pages of ITS slice in the background
      arr = g.finish(local)                       AFTER it: every rank copies its own block device -> host straight into its slice -- eight PCIe
                                                  links side by side, into pages that are already resident, and no gathered copy on any device.
                                                  rank 0 gets the array (it owns the mapping; the segment's name is gone by then), the others None

    (The single-rank copy of the gathered 0.8 GB block was 56-75 ms of a 0.3-0.8 s sharded match call on 8 GPUs; filling a fresh segment from
    eight processes at the END of the call was worse, 150 ms: first touch of shared-memory pages does not scale across processes -- hence the early,
    background touch.)  Ranks on several machines: `finish` falls back to an NCCL gather to rank 0 and one copy."""

    def __init__(self, n_total, row_shape, dtype):
        import mmap
        import os
        from .ops import HostBuffer
        world, rank = _world()
        self.n_total, self.row_shape = int(n_total), tuple(int(x) for x in row_shape)
        self.np_dtype = torch.empty(0, dtype=dtype).numpy().dtype
        self.shared = world > 1 and _same_host()
        self.arr = self.mm = self.buf = None
        if not self.shared:
            return
        row_elems = int(np.prod(self.row_shape, dtype=np.int64))
        total = max(1, self.n_total * row_elems * self.np_dtype.itemsize)
        path = [None]
        if rank == 0:
            _SHM_COUNT[0] += 1
            path[0] = f"/dev/shm/laminr_b200_{os.getpid()}_{_SHM_COUNT[0]}"
            fd = os.open(path[0], os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
            os.ftruncate(fd, total)
        dist.broadcast_object_list(path, src=0)
        if rank != 0:
            fd = os.open(path[0], os.O_RDWR)
        self.path = path[0]
        self.mm = mmap.mmap(fd, total)
        os.close(fd)
        self.arr = np.frombuffer(self.mm, dtype=self.np_dtype, count=self.n_total * row_elems).reshape((self.n_total,) + self.row_shape)
        lo, hi = shard_bounds(self.n_total, world, rank)
        self.lo, self.hi = lo, hi
        self.buf = HostBuffer(0, self.np_dtype, flat=self.arr[lo:hi].reshape(-1)) if hi > lo else None

    def finish(self, local):
        import os
        from .ops import to_host_numpy
        world, rank = _world()
        if world == 1:
            return to_host_numpy(local)
        if not self.shared:
            full = gather_ragged(local, self.n_total)
            return None if rank != 0 else to_host_numpy(full)
        if self.buf is not None:
            to_host_numpy(local, out=self.buf.wait())
        dist.barrier()
        arr, self.arr, self.buf = self.arr, None, None
        if rank != 0:
            del arr
            self.mm.close()
            return None
        os.unlink(self.path)  # the name goes; the pages live as long as rank 0's mapping (owned by the returned array)
        return arr


def gather_to_host(local, n_total):
    """HostGather in one call (no early page touch): the blocks of all ranks as one NumPy array on rank 0, None elsewhere."""
    return HostGather(n_total, local.shape[1:], local.dtype).finish(local)


def simulated_population_sharded(n_neurons, img_res=(100, 100), seed=0, num_filters=20, device=None):
    """`neuron_models.simulated_population(n_neurons, ...)` built in parallel: every rank generates the Gabor banks of its contiguous block of
    neurons on its host cores (same NumPy random stream, so the banks are bit-identical to the serial build), the blocks are all-gathered on the
    device, and every rank returns the FULL population (match / MEI sharding indexes the same model on every rank).  BASELINE config 3: 1024
    neurons at 100x100 = 819 MB per GPU; ~5 s of host work on one process, 1/world of it here."""
    from .neuron_models.simulated_models import simulated_population
    from .neuron_models.utils.simulated_model import ArbitraryMultiNeuronModel, ArbitraryNeuronModel

    world, rank = _world()
    if world == 1:
        pop = simulated_population(n_neurons, img_res=list(img_res), seed=seed, num_filters=num_filters)
        return pop.to(device) if device is not None else pop
    lo, hi = shard_bounds(n_neurons, world, rank)
    local = simulated_population(n_neurons, img_res=list(img_res), seed=seed, num_filters=num_filters, only=range(lo, hi)) if hi > lo else None
    H, W = int(img_res[1]), int(img_res[0])
    block = local.packed_filters.reshape(hi - lo, num_filters, H, W) if local is not None else torch.empty(0, num_filters, H, W)
    if device is not None:
        block = block.to(device)
    packed = gather_ragged(block, n_neurons, dst=None)
    pop = ArbitraryMultiNeuronModel([ArbitraryNeuronModel.from_tensor(packed[i]) for i in range(n_neurons)], packed=packed)
    # the small index buffers (filter counts, neuron list) are created on the host: move them next to the bank (the bank itself and the per-neuron
    # views of it are already there: `.to` returns the same tensors)
    return pop.to(packed.device)


def match_sharded(im, target_neuron_idxs, **match_kwargs):
    """`InvarianceManifold.match` with the targets sharded over the ranks.  Call on every rank with the same arguments after `learn`
    ran on rank 0 (any rank may hold an untrained copy of the template: it is overwritten by rank 0's).
    Returns (images [D,N,C,H,W], activations [D,N]) as NumPy arrays on rank 0, (None, None) elsewhere.
    Afterwards every rank's `im` holds the coordinate transforms of ITS block only: `im.target_neuron_idxs` (and with it
    `get_images_from_matched_template`) is per rank; the full list is kept in `im.all_target_neuron_idxs`."""
    world, rank = _world()
    device = next(im.template.parameters()).device
    targets = list(target_neuron_idxs)
    D = len(targets)
    if world == 1:
        return im.match(targets, **match_kwargs)
    if D < world:  # decided from the arguments alone, so EVERY rank raises (a rank that raised alone would leave the others waiting in a collective)
        raise ValueError(f"match_sharded: {D} targets cannot be split over {world} ranks (every rank needs at least one target)")
    import os, sys, time
    timing = os.environ.get("LMR_TIMING") == "1"  # debug: per-phase wall clock of rank 0 on stderr (adds synchronisations)
    marks = []

    def mark(name):
        if timing:
            torch.cuda.synchronize(device)
            marks.append((name, time.perf_counter()))

    mark("start")
    tni = torch.tensor([int(getattr(im, "template_neuron_idx", -1))], dtype=torch.int64, device=device)
    dist.broadcast(tni, src=0)
    im.template_neuron_idx = int(tni.item())
    broadcast_template(im.template, src=0)
    broadcast_rng_state(src=0, device=device)
    lo, hi = shard_bounds(D, world, rank)
    fused = im._fused_ok()
    mark("broadcasts")
    snaps = match_kwargs.get("save_results_every_n_epochs") is not None
    hg = None
    if fused and not snaps:  # the result's shape is known: set up its host destination now, pages touched while the GPUs work
        C_, H_, W_ = list(im.meis_dict.values())[0]["mei"].shape
        hg = HostGather(D, (int(match_kwargs.get("grid_points_per_dim", 20)), C_, H_, W_), torch.float32)
    imgs, acts = im.match(targets[lo:hi], _total_targets=D, _stop_reduce=make_stop_reduce(device), _return_device=fused, **match_kwargs)
    im.all_target_neuron_idxs = targets
    mark("local match")
    if not fused:  # generic autograd path: results come back as host arrays
        imgs, acts = torch.from_numpy(np.ascontiguousarray(imgs)).to(device), torch.from_numpy(np.ascontiguousarray(acts)).to(device)
    if snaps:  # [snapshots, D_local, ...] -> the target axis first for the gather, and back
        imgs, acts = imgs.transpose(0, 1).contiguous(), acts.transpose(0, 1).contiguous()
    g_act = gather_ragged(acts.reshape(hi - lo, *acts.shape[1:]), D)
    mark("gather activations")
    block = imgs.reshape(hi - lo, *imgs.shape[1:])
    h_img = hg.finish(block) if hg is not None else gather_to_host(block, D)  # every rank copies its own block to the host (shared memory) when it can
    mark("images to host")
    if rank != 0:
        return None, None
    h_act = g_act.cpu().numpy()
    if snaps:
        h_img, h_act = np.swapaxes(h_img, 0, 1), np.swapaxes(h_act, 0, 1)
    out = h_img, h_act
    if timing:
        print("[match_sharded] " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.1f} ms" for a, b in zip(marks, marks[1:])), file=sys.__stderr__)
    return out


def get_mei_dict_sharded(model, input_shape, pixel_value_lower_bound, pixel_value_upper_bound, neuron_idxs=None, **kwargs):
    """`get_mei_dict` with the neurons sharded over the ranks; the full dict (keyed by enumeration index like the reference,
    laminr/utils/mei.py:114) is returned on EVERY rank (it is the input of `InvarianceManifold`)."""
    from .utils.mei import get_mei_dict

    world, rank = _world()
    if neuron_idxs is None:
        if not hasattr(model, "n_neurons"):
            raise ValueError("get_mei_dict_sharded: pass neuron_idxs for models without `n_neurons`")
        neuron_idxs = list(range(int(model.n_neurons)))
    neuron_idxs = list(neuron_idxs)
    if world == 1:
        return get_mei_dict(model, input_shape, pixel_value_lower_bound, pixel_value_upper_bound, neuron_idxs=neuron_idxs, **kwargs)
    lo, hi = shard_bounds(len(neuron_idxs), world, rank)
    # the reference draws one initial image per neuron, in order, from the CPU generator (mei.py:33): skip the draws of lower ranks
    items = list(model.parameters()) + list(model.buffers())
    broadcast_rng_state(src=0, device=items[0].device if items else None)
    # (one draw per neuron, like the unsharded call: the CPU generator's vectorised normal consumes its stream per CALL, so one big draw
    # is a different stream than `lo` single-image draws unless the image size is a multiple of 16)
    for _ in range(lo):
        torch.randn(1, *input_shape)
    # fewer neurons than ranks: the ranks without a neuron only keep the generator in step and take part in the gather
    local = get_mei_dict(model, input_shape, pixel_value_lower_bound, pixel_value_upper_bound, neuron_idxs=neuron_idxs[lo:hi], **kwargs) if hi > lo else {}
    if hi < len(neuron_idxs):  # ... and of the higher ranks: every rank leaves the generator where the unsharded call would (same jitter in `learn`)
        for _ in range(len(neuron_idxs) - hi):
            torch.randn(1, *input_shape)
    return _allgather_mei_entries(local, lo, len(neuron_idxs), items[0].device if items else torch.device("cpu"))


_ENTRY_KEYS = ("mei", "activation", "mask", "center_pos")


def _allgather_mei_entries(local, lo, n_total, device):
    """Every rank's `{k: {"mei": ndarray, "activation": float, "mask": ndarray, "center_pos": {..: float}}}` -> the full dict (keys lo + k) on every
    rank.  The arrays travel as three tensors (images, masks, float64 scalars) through one ragged all-gather each -- pickling the entries
    (`all_gather_object`) cost more than the whole sharded ascent for 1024 neurons at 100x100 (82 MB of arrays, profiles/r1k_config3.md).  The ranks
    first exchange a few bytes describing their entries; anything but uniform numpy entries falls back to the pickled gather, on all ranks alike."""
    world, rank = _world()
    vals = [local[k] for k in sorted(local)]

    def describe(v):
        if not (isinstance(v, dict) and tuple(v) == _ENTRY_KEYS and isinstance(v["mei"], np.ndarray) and isinstance(v["mask"], np.ndarray)
                and isinstance(v["center_pos"], dict) and isinstance(v["activation"], (float, np.floating))
                and all(isinstance(c, (float, np.floating)) for c in v["center_pos"].values())):
            return None
        return (v["mei"].dtype.str, v["mei"].shape, v["mask"].dtype.str, v["mask"].shape, tuple(v["center_pos"]))

    descr = {describe(v) for v in vals}
    metas = [None] * world
    dist.all_gather_object(metas, (len(vals), sorted(descr, key=repr) if None not in descr else None))
    kinds = {d for _, ds in metas if ds is not None for d in ds}
    if any(n > 0 and ds is None for n, ds in metas) or len(kinds) != 1:
        parts = [None] * world
        dist.all_gather_object(parts, {lo + k: v for k, v in local.items()})
        out = {}
        for part in parts:
            out.update(part)
        return dict(sorted(out.items()))
    mei_dt, mei_shape, mask_dt, mask_shape, ckeys = next(iter(kinds))
    dev = device if dist.get_backend() == "nccl" else torch.device("cpu")
    pack = lambda arrs, dt, shape: torch.from_numpy(np.stack(arrs).astype(dt, copy=False) if arrs else np.empty((0,) + tuple(shape), dt)).to(dev)
    mei = gather_ragged(pack([v["mei"] for v in vals], mei_dt, mei_shape), n_total, dst=None).cpu().numpy()
    mask = gather_ragged(pack([v["mask"] for v in vals], mask_dt, mask_shape), n_total, dst=None).cpu().numpy()
    scal = np.array([[v["activation"]] + [v["center_pos"][c] for c in ckeys] for v in vals], np.float64).reshape(len(vals), 1 + len(ckeys))
    scal = gather_ragged(torch.from_numpy(scal).to(dev), n_total, dst=None).cpu().numpy()
    return {k: {"mei": mei[k], "activation": float(scal[k, 0]), "mask": mask[k], "center_pos": {c: float(scal[k, 1 + j]) for j, c in enumerate(ckeys)}}
            for k in range(n_total)}
