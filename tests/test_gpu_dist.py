"""Sharded match == unsharded match per target (laminr_b200.dist.match_sharded), two ranks.  On a one-GPU box both ranks share cuda:0
and talk over gloo (NCCL refuses two ranks on one device); with >= 2 GPUs the same test runs over NCCL, one rank per GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _setup(res):
    import laminr_b200 as L
    from bench import synthetic_meis
    torch.manual_seed(0)
    np.random.seed(0)
    model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
    meis = synthetic_meis(res)
    im = L.InvarianceManifold(model, meis, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0, precision="bf16x3")
    im.template_neuron_idx = 0
    return im


def _worker(rank, world, port, backend, res, epochs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank if backend == "nccl" else 0)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from laminr_b200.dist import match_sharded
        im = _setup(res)
        if rank != 0:  # a different (untrained) template on the other ranks: it must be overwritten by rank 0's
            with torch.no_grad():
                im.template.flat_params().add_(0.05)
            torch.manual_seed(1234)
        imgs, acts = match_sharded(im, [1, 2], num_epochs=epochs)
        q.put((rank, imgs, acts, im.match_epochs_run))
    finally:
        dist.destroy_process_group()


def test_match_sharded_equals_unsharded():
    res, epochs = 24, 6
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, res, epochs, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    im = _setup(res)
    want_imgs, want_acts = im.match([1, 2], num_epochs=epochs)
    (_, imgs, acts, ep0), (_, imgs1, acts1, ep1) = out
    assert imgs1 is None and acts1 is None and ep0 == ep1 == im.match_epochs_run
    assert imgs.shape == want_imgs.shape == (2, 20, 1, res, res)
    np.testing.assert_allclose(acts, want_acts, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(imgs, want_imgs, rtol=0, atol=2e-5)


def _mei_worker(rank, world, port, backend, res, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank if backend == "nccl" else 0)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        import laminr_b200 as L
        from laminr_b200.dist import get_mei_dict_sharded
        model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
        torch.manual_seed(7 if rank == 0 else 99)  # rank 0's generator state is the one that counts
        meis = get_mei_dict_sharded(model, [1, res, res], -1.0, 1.0, neuron_idxs=[0, 1, 2], required_img_norm=1.0, num_iterations=50)
        q.put((rank, {k: (np.asarray(v["mei"]).copy(), float(v["activation"])) for k, v in meis.items()}, torch.get_rng_state().numpy().copy()))
    finally:
        dist.destroy_process_group()


def test_get_mei_dict_sharded_equals_unsharded():
    """Every rank gets the full dict, with the MEIs the unsharded call produces (the initial images are drawn per neuron, in order, from the CPU
    generator: laminr/utils/mei.py:33), and leaves the CPU generator where the unsharded call leaves it."""
    res = 24
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mei_worker, args=(r, world, port, backend, res, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    import laminr_b200 as L
    model = L.neuron_models.simulated("demo1", img_res=[res, res]).cuda()
    torch.manual_seed(7)
    want = L.get_mei_dict(model, [1, res, res], -1.0, 1.0, neuron_idxs=[0, 1, 2], required_img_norm=1.0, num_iterations=50)
    want_state = torch.get_rng_state().numpy()
    for rank, meis, state in out:
        assert sorted(meis) == [0, 1, 2]
        for k in range(3):
            np.testing.assert_allclose(meis[k][0], np.asarray(want[k]["mei"]), rtol=0, atol=1e-6)
            assert abs(meis[k][1] - float(want[k]["activation"])) < 1e-6
        assert np.array_equal(state, want_state)
