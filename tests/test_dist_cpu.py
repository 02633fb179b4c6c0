"""Host-side logic of the multi-GPU sharding (laminr_b200/dist.py) on CPU: world_size-2 `gloo` process group.
The collectives carry plain tensors, so the plumbing (shard bounds, template broadcast, RNG agreement, global stop-rule reduce,
ragged gather) is exercised exactly as under NCCL; the CUDA compute is covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_targets, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from laminr_b200 import dist as D
        from laminr_b200.utils.training import ImprovementChecker, JitteringGridDatamodule

        # --- template broadcast: rank 0's parameters / buffers win
        torch.manual_seed(100 + rank)
        mod = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Tanh(), torch.nn.Linear(7, 1))
        mod.register_buffer("B", torch.randn(2, 4))
        D.broadcast_module_state(mod, src=0)
        # --- template broadcast of match_sharded: only what `learn` produced travels, so a rank that owns extra members (rank 0 after an unsharded
        #     match(): coordinate transform + RF-shift buffers) does not desynchronise the collectives
        tmpl = torch.nn.Module()
        tmpl.func = torch.nn.Sequential(torch.nn.Linear(3, 4), torch.nn.Tanh(), torch.nn.Linear(4, 1))
        for name, shape in (("B", (2, 5)), ("auxB", (2, 5)), ("coords", (1, 6, 2))):
            tmpl.register_buffer(name, torch.randn(*shape))
        if rank == 0:
            tmpl.coordinate_transform = torch.nn.Linear(2, 2)
            tmpl.register_buffer("coords_shift_to_center", torch.randn(1, 2))
        D.broadcast_template(tmpl, src=0)
        tmpl_state = [p.detach().numpy().copy() for p in tmpl.func.parameters()] + [tmpl.B.numpy().copy(), tmpl.auxB.numpy().copy(), tmpl.coords.numpy().copy()]
        # --- population built in parallel (every rank generates its block of neurons, the banks are all-gathered): the serial build, bit for bit
        from laminr_b200.neuron_models import simulated_population
        pop = D.simulated_population_sharded(5, (12, 10), seed=4, num_filters=6)
        assert torch.equal(pop.packed_filters, simulated_population(5, img_res=[12, 10], seed=4, num_filters=6).packed_filters)
        assert len(pop.neuron_models_list) == 5 and pop.neuron_models_list[3].filters.data_ptr() == pop.packed_filters[3].data_ptr()
        # --- RNG agreement: the jittered grids are drawn on the CPU generator
        D.broadcast_rng_state(src=0)
        dm = JitteringGridDatamodule(1, 20, 3)
        grids = torch.stack(list(dm.train_dataloader()))
        # --- global stop rule: per-rank partial (sum, min, max) -> identical global decision on every rank
        lo, hi = D.shard_bounds(n_targets, world, rank)
        rng = np.random.RandomState(7)
        curves = np.cumsum(rng.rand(40, n_targets) * np.clip(np.linspace(0.03, -0.03, 40), 0, None)[:, None], axis=0)  # saturating per-target activations
        reduce = D.make_stop_reduce()
        chk, stopped_at, means = ImprovementChecker(patience=5, ignore_diff_smaller_than=1e-3), None, []
        for epoch in range(40):
            a = curves[epoch, lo:hi]
            s, mn, mx = reduce(float(a.sum()), float(a.min()), float(a.max()))
            # the tensor form (the fused stepper's 3-float epoch statistics) must give the same decision inputs
            s2, mn2, mx2 = reduce(torch.tensor([a.sum(), a.min(), a.max()], dtype=torch.float64))
            assert (s2, mn2, mx2) == (s, mn, mx)
            # ... and so must the split form (collective queued now, result read an epoch later)
            assert reduce.end(reduce.begin(torch.tensor([a.sum(), a.min(), a.max()], dtype=torch.float64))) == (s, mn, mx)
            means.append(s / n_targets)
            assert abs(mn - curves[epoch].min()) < 1e-12 and abs(mx - curves[epoch].max()) < 1e-12
            if not chk.is_increasing(s / n_targets):
                stopped_at = epoch
                break
        # --- ragged gather of per-target results
        local = torch.arange(lo, hi, dtype=torch.float32)[:, None, None] * torch.ones(1, 3, 2)
        full = D.gather_ragged(local, n_targets, dst=0)
        # --- the same blocks collected on rank 0's HOST through one shared-memory segment that every rank fills itself
        host = D.gather_to_host(local, n_targets)
        assert (host is None) == (rank != 0)
        if rank == 0:
            assert isinstance(host, np.ndarray) and host.shape == (n_targets, 3, 2) and np.array_equal(host, full.numpy())
            host[0, 0, 0] = 5.0  # writable, and alive after the segment's name is gone
            assert not any(f.startswith("laminr_b200_") for f in os.listdir("/dev/shm"))
        # numpy payloads are pickled by value; torch tensors would travel as shared-memory handles that die with this process
        q.put((rank, [p.detach().numpy().copy() for p in mod.parameters()] + [mod.B.numpy().copy()] + tmpl_state, grids.numpy().copy(), stopped_at, means,
               None if full is None else full.numpy().copy(), (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_targets", [5, 8])
def test_sharded_match_plumbing_world2(n_targets):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_targets, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, state0, grids0, stop0, means0, full0, b0), (r1, state1, grids1, stop1, means1, full1, b1) = res
    for a, b in zip(state0, state1):
        assert np.array_equal(a, b)                    # broadcast template state
    assert np.array_equal(grids0, grids1)              # same jitter on every rank
    assert stop0 == stop1 and stop0 is not None        # same (global) stop epoch
    np.testing.assert_allclose(means0, means1, rtol=0, atol=0)
    # unsharded reference of the stop rule
    from laminr_b200.utils.training import ImprovementChecker
    rng = np.random.RandomState(7)
    curves = np.cumsum(rng.rand(40, n_targets) * np.clip(np.linspace(0.03, -0.03, 40), 0, None)[:, None], axis=0)
    chk, want = ImprovementChecker(patience=5, ignore_diff_smaller_than=1e-3), None
    for epoch in range(40):
        if not chk.is_increasing(float(curves[epoch].sum()) / n_targets):
            want = epoch
            break
    assert stop0 == want
    assert b0[1] == b1[0] and b0[0] == 0 and b1[1] == n_targets and (b0[1] - b0[0]) - (b1[1] - b1[0]) in (0, 1)
    assert full1 is None
    assert np.array_equal(full0[:, 0, 0], np.arange(n_targets, dtype=np.float32)) and full0.shape == (n_targets, 3, 2)


def test_shard_bounds_cover_everything():
    from laminr_b200.dist import shard_bounds
    for n in (1, 7, 8, 1023):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker_small(rank, world, port, n_neurons, arrays, q, shape=(1, 4, 4)):
    """Fewer items than ranks / ragged MEI shards: no rank may be left alone in a collective."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from laminr_b200 import dist as D
        import laminr_b200.utils.mei as mei_mod

        shape = list(shape)

        def fake_get_mei_dict(model, input_shape, lo_b, hi_b, neuron_idxs=None, **kw):
            # like the real one: one starting image per neuron, in order, from the CPU generator (laminr/utils/mei.py:33); keys = enumeration index
            x0 = torch.cat([torch.randn(1, *input_shape) for _ in neuron_idxs]) if len(neuron_idxs) else torch.empty(0, *input_shape)
            if arrays:  # entries shaped like the real ones: they travel as tensors (images, masks, float64 scalars)
                return {k: {"mei": x0[k].numpy().copy(), "activation": float(np.float32(i) / 3), "mask": x0[k, 0].numpy().copy() * 2,
                            "center_pos": {"y": i + 0.1, "x": i / 7}} for k, i in enumerate(neuron_idxs)}
            return {k: dict(idx=int(i), x0=x0[k].numpy().copy()) for k, i in enumerate(neuron_idxs)}  # anything else: pickled gather

        mei_mod.get_mei_dict = fake_get_mei_dict
        torch.manual_seed(50 + rank)  # the ranks start out of step: rank 0's generator must win
        out = D.get_mei_dict_sharded(torch.nn.Identity(), shape, -1.0, 1.0, neuron_idxs=list(range(10, 10 + n_neurons)))
        after = torch.rand(2).numpy().copy()

        class FakeManifold:
            template = torch.nn.Linear(2, 2)

        err = None
        try:
            D.match_sharded(FakeManifold(), [3] if world > 1 else [])
        except ValueError as e:
            err = str(e)
        if arrays:
            for k, v in out.items():
                i = 10 + k
                assert tuple(v) == ("mei", "activation", "mask", "center_pos") and v["mei"].dtype == np.float32 and v["mei"].shape == tuple(shape)
                assert type(v["activation"]) is float and v["activation"] == float(np.float32(i) / 3)
                assert v["center_pos"] == {"y": i + 0.1, "x": i / 7} and list(v["center_pos"]) == ["y", "x"]
                assert np.array_equal(v["mask"], v["mei"][0] * 2)
            out = {k: dict(idx=10 + k, x0=v["mei"]) for k, v in out.items()}
        q.put((rank, {k: (v["idx"], v["x0"]) for k, v in out.items()}, after, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_neurons,arrays,shape", [(1, False, (1, 4, 4)), (5, False, (1, 4, 4)), (1, True, (1, 4, 4)), (5, True, (1, 4, 4)),
                                                     # 15 / 196 pixels: NOT a multiple of 16, where the CPU generator's vectorised normal makes one
                                                     # [k, ...] draw differ from k single-image draws (the unsharded call draws per neuron)
                                                     (5, True, (1, 3, 5)), (5, True, (1, 14, 14))])
def test_sharded_mei_dict_and_too_few_targets_world2(n_neurons, arrays, shape):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_small, args=(r, world, port, n_neurons, arrays, q, shape)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the unsharded call: rank 0's seed
    torch.manual_seed(50)
    x0 = torch.cat([torch.randn(1, *shape) for _ in range(n_neurons)]).numpy()  # one draw per neuron (laminr/utils/mei.py:33 inside the loop of :98)
    want_after = torch.rand(2).numpy()
    for rank, out, after, err in res:
        assert list(out) == list(range(n_neurons))                                   # the full dict on EVERY rank, enumeration keys
        for k in range(n_neurons):
            assert out[k][0] == 10 + k and np.array_equal(out[k][1], x0[k])          # each neuron started from the image the unsharded call draws
        assert np.array_equal(after, want_after)                                     # ... and the generator is left where that call leaves it
        assert err is not None and "cannot be split" in err                          # one target over two ranks: every rank raises, nobody hangs
