"""Fused step drivers: the learn / match hot steps as a fixed sequence of C-ABI launches on static buffers.

One `LearnStepper.step()` is the body of the reference's inner loop (laminr/inv_temp_learn_match.py:186-208: forward,
objective, backward, Adam) without autograd, without per-step allocation or host synchronisation, optionally replayed
from a CUDA graph.  `MatchStepper.step()` is the body of :351-366 for D targets, each neuron evaluated on its own
images only (the reference evaluates all neurons on all D*N images and keeps the diagonal, :551,:361).
"""
import os

import torch

from . import _lib
from . import ops


def _keep_activations(precision, workspace_bytes):
    """Keep the hidden activations' operand tiles between the INR forward and backward of a training step (LMR_PREC_KEEP_ACTIVATIONS)?  Yes on
    the tensor-core precisions when the enlarged forward workspace stays under LMR_KEEP_ACT_MAX_GIB (default 32 GiB of the 180 GB); the backward then
    loads the tiles instead of recomputing the forward.  LMR_KEEP_ACT=0 switches it off."""
    if os.environ.get("LMR_KEEP_ACT", "1") == "0" or ops.precision_code(precision) == _lib.PREC_FP32:
        return False
    return workspace_bytes <= float(os.environ.get("LMR_KEEP_ACT_MAX_GIB", "32")) * 2 ** 30


def match_block(precision, desc, N, P, D):
    """-> (targets per block, keep activations?) of the match training step: all D targets when their kept operand tiles fit the budget, else the
    largest block that does (the forward workspace grows linearly with the block), else -- fp32 path, LMR_KEEP_ACT=0, or not even one target fits --
    all D targets with the recompute backward.  LMR_MATCH_BLOCK forces a block size (testing).  Pure host arithmetic on the C ABI's workspace sizes."""
    import os

    block, keep = D, _keep_activations(precision, ops.forward_workspace_bytes(desc, N, P, D, True))
    if not keep and ops.precision_code(precision) != _lib.PREC_FP32 and os.environ.get("LMR_KEEP_ACT", "1") != "0":
        lo, hi = 0, D  # invariant: a block of `lo` targets fits (0 = nothing known to fit), a block of `hi` does not
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if _keep_activations(precision, ops.forward_workspace_bytes(desc, N, P, mid, True)):
                lo = mid
            else:
                hi = mid
        if lo >= 1:
            block, keep = lo, True
    if os.environ.get("LMR_MATCH_BLOCK"):
        block = max(1, min(D, int(os.environ["LMR_MATCH_BLOCK"])))
    return block, keep


def capture_graph(fn):
    """Stream-captures the launches of `fn` into a CUDA graph.  The raw capture API instead of the `torch.cuda.graph` context manager: that one
    synchronises the device, runs the Python garbage collector and empties the allocator cache around every capture (~15 ms, more than a whole
    small match() call); the launch sequences captured here work on static buffers and need none of it."""
    main = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    side.wait_stream(main)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        g.capture_begin()
        try:
            fn()
        finally:
            g.capture_end()
    main.wait_stream(side)
    return g


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


class _Constraint:
    """(mode, target, mean, pmin, pmax, eps) of a FixEnergyNormClip / StandardizeClip module."""

    def __init__(self, img_transf):
        for k in ("mode", "target", "mean", "pixel_min", "pixel_max", "eps"):
            if not hasattr(img_transf, k):
                raise TypeError("fused path needs a laminr_b200 FixEnergyNormClip / StandardizeClip module")
        self.mode, self.target, self.mean, self.eps = img_transf.mode, img_transf.target, img_transf.mean, img_transf.eps
        self.pmin, self.pmax = img_transf.pixel_min, img_transf.pixel_max


class _SimulatedResponse:
    """Response backend for `ArbitraryMultiNeuronModel` populations (lmr_simresp_*)."""
    launches_fwd, launches_fwd_bwd = 2, 3

    def __init__(self, model, D, N, P, dev, also=()):
        self.model = model
        self.rmax = torch.empty((D, N), dtype=torch.float32, device=dev)
        self.argmax = torch.empty((D, N), dtype=torch.int32, device=dev)
        # The partials' chunk size depends on the number of groups of a call (fewer groups -> smaller pixel chunks -> MORE partials per group), so
        # lmr_simresp_workspace_bytes is not monotonic in the group count: size the workspace for EVERY group count 1..D the backend may be called
        # with (a sweep's smaller tail batch, a block remainder).  Above 2 x 148 groups the chunk no longer changes, so 1..296 and D cover all.
        lib, Fmax = _lib.load(), model.packed_filters.shape[1]
        counts = set(range(1, min(int(D), 2 * 148) + 1)) | {int(D)} | {int(g) for g in also if g > 0}
        self.ws = _ws(max(lib.lmr_simresp_workspace_bytes(g, N, Fmax, P) for g in counts), dev)

    def forward(self, z, stride, N, D, Cc, P, nog, act):
        m = self.model
        ops.raw_simresp_fwd(z, stride, N, D, Cc, P, m.packed_filters, m.nfilt, nog, m.eps, out=(act, self.rmax, self.argmax), workspace=self.ws)

    def forward_backward(self, z, stride, N, D, Cc, P, nog, g_act, act, g_z, accumulate):
        self.forward(z, stride, N, D, Cc, P, nog, act)
        ops.raw_simresp_bwd(g_act, self.rmax, self.argmax, stride, N, D, Cc, P, self.model.packed_filters, nog, g_z, accumulate=accumulate)


class _ConvResponse:
    """Response backend for FiringRateEncoder(Stacked2dCore, FullGaussian2d): lmr_convcore_response evaluates the selected neuron's
    receptive-field cone, forward and image gradient in ONE launch (the upstream gradient of the loops is a constant per image)."""
    launches_fwd, launches_fwd_bwd = 1, 1

    def __init__(self, model, D, N, P, dev, also=()):
        self.cone = model.cone()
        if self.cone.HW != P:
            raise ValueError("response model and template resolution differ")

    def forward(self, z, stride, N, D, Cc, P, nog, act):
        ops.raw_convcore_response(self.cone, z, stride, N, D, nog, act=act)

    def forward_backward(self, z, stride, N, D, Cc, P, nog, g_act, act, g_z, accumulate):
        ops.raw_convcore_response(self.cone, z, stride, N, D, nog, g_act=g_act, act=act, g_img=g_z, accumulate=accumulate)


def response_backend(model, D, N, P, dev, also=()):
    from .neuron_models.utils.simulated_model import ArbitraryMultiNeuronModel
    if isinstance(model, ArbitraryMultiNeuronModel):
        return _SimulatedResponse(model, D, N, P, dev, also)
    if hasattr(model, "cone"):
        return _ConvResponse(model, D, N, P, dev, also)
    raise TypeError("fused path: unsupported response model %s" % type(model).__name__)


def fused_response_supported(model):
    from .neuron_models.utils.simulated_model import ArbitraryMultiNeuronModel
    return isinstance(model, ArbitraryMultiNeuronModel) or hasattr(model, "cone")


class LearnStepper:
    def __init__(self, template, img_transf, model, neuron_idx, rf_mask, mei_act, grid_reg, pixel_min, pixel_max, n_grid, lr=1e-3,
                 reg_scale=2.0, precision=None, use_graph=True):
        lib = _lib.load()
        import ctypes as C
        self.t, self.model, self.reg_mod = template, model, grid_reg
        self.con = _Constraint(img_transf)
        dev = next(template.parameters()).device
        self.dev, self.N, self.lr, self.precision = dev, int(n_grid), lr, precision
        self.C, (H, W) = template.out_channels, template.img_res
        self.P = H * W
        N, Cc, P = self.N, self.C, self.P
        self.pixel_min, self.pixel_max, self.mei_act = float(pixel_min), float(pixel_max), float(mei_act)
        self.flat = template.flat_params()
        self.coords = template.coords.reshape(-1, 2).contiguous()
        self.mask = torch.as_tensor(rf_mask, dtype=torch.float32).to(dev).reshape(-1).contiguous()
        self.nog = torch.tensor([int(neuron_idx)], dtype=torch.int32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.grid = torch.zeros(N, **f32)
        self.img_pre = torch.empty((N, Cc, H, W), **f32)
        self.z = torch.empty_like(self.img_pre)
        self.g_z = torch.empty_like(self.img_pre)
        self.g_x = torch.empty_like(self.img_pre)
        self.stats = torch.empty((N, 2), **f32)
        self.act = torch.empty((1, N), **f32)
        self.resp = response_backend(model, 1, N, P, dev)
        self.reg, self.K = torch.empty(1, **f32), torch.empty((N, N), **f32)
        self.g_act = torch.full((1, N), -1.0 / (N * self.mei_act), **f32)
        self.g_reg = torch.full((1,), -float(reg_scale), **f32)  # d loss / d reg = -reg_scale (host float, changes between epochs)
        self.g_flat = torch.zeros_like(self.flat)
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.step_count = torch.zeros(2, dtype=torch.int32, device=dev)  # [steps taken, ticket scratch]
        d = template._desc
        # the training forward keeps the operand tiles of the hidden activations for the backward (no recompute): 96 KB per 128-row tile
        self.keep = _keep_activations(precision, ops.forward_workspace_bytes(d, N, P, 1, True))
        self.ws_fwd = _ws(ops.forward_workspace_bytes(d, N, P, 1, self.keep), dev)
        self.ws_bwd = _ws(lib.lmr_inr_backward_workspace_bytes(C.byref(d), N, P, 1), dev)
        self.ws_con = _ws(lib.lmr_constrain_workspace_bytes(N, Cc * P), dev)
        self.ws_clr = _ws(lib.lmr_simclr_workspace_bytes(N, Cc, P), dev)
        self.use_graph, self.graph, self.eval_graph = use_graph, None, None
        self.host_graph, self.grid_map, self.loss_map, self.host_done = None, None, None, None  # host-fed steps (step() with a CPU row)
        self._host_pending, self.loss_np = False, None
        self._table_written = False
        self.side = torch.cuda.Stream(device=dev)
        # fused objective (one persistent launch for constraint + response + contrastive, forward and backward) for simulated neurons when it fits
        self.loss_buf = torch.zeros(1, **f32)
        self.ws_obj = None
        if isinstance(self.resp, _SimulatedResponse):
            self.bank = model.packed_filters[int(neuron_idx), :int(model.nfilt[int(neuron_idx)].item())].contiguous()
            self.ws_obj = ops.learn_objective_workspace(N, Cc, P, self.bank.shape[0], dev)
        # inr fwd (2) + [objective (1) | constrain fwd (2) + simclr fwd (2) + simclr bwd (1) + response + constrain bwd (2)] + inr bwd (4) + adam (1)
        # prep + forward + objective + backward tile kernel + fused tail (fp32 path: the four separate tail launches)
        self.kernels_per_step = (5 if ops.precision_code(precision) != _lib.PREC_FP32 else 8) if self.ws_obj is not None else 14 + self.resp.launches_fwd_bwd

    def set_reg_scale(self, reg_scale):
        self.g_reg.fill_(-float(reg_scale))

    def _table_valid(self):
        """Has a forward of this stepper already left the coordinate sin/cos table in `ws_fwd`?  The pixel grid, B and the (absent) coordinate shift
        never change during a learn loop, so every forward after the first one reuses the table (LMR_PREC_REUSE_COORD_TABLE).  The first call
        of a stepper always runs eagerly (graph capture comes after it), so a captured launch never carries the first-call flag value."""
        valid, self._table_written = self._table_written, True
        return valid and os.environ.get("LMR_REUSE_COORD_TABLE", "1") != "0"

    # ---- launches -------------------------------------------------------------------------------------------------
    def _forward(self):
        t, c, N, Cc, P = self.t, self.con, self.N, self.C, self.P
        ops.raw_inr_forward(t._desc, self.flat, t.B, t.auxB, self.grid, self.coords, None, None, out=self.img_pre.view(1, N, Cc, P),
                            precision=self.precision, workspace=self.ws_fwd, reuse_coord_table=self._table_valid())
        ops.raw_constrain_fwd(c.mode, self.img_pre, c.target, c.mean, c.pmin, c.pmax, c.eps, out=self.z, stats=self.stats, workspace=self.ws_con)
        self.resp.forward(self.z, 0, N, 1, Cc, P, self.nog, self.act)

    def _train(self, grid=None, loss=None):
        """grid / loss: where the step reads its latents / writes its loss (default: the stepper's device buffers; the host-fed step passes
        pinned host memory, ops.MappedHost)."""
        t, c, N, Cc, P = self.t, self.con, self.N, self.C, self.P
        r = self.reg_mod
        if self.ws_obj is not None:
            grid = self.grid if grid is None else grid
            ops.raw_inr_forward(t._desc, self.flat, t.B, t.auxB, grid, self.coords, None, None, out=self.img_pre.view(1, N, Cc, P),
                                precision=self.precision, workspace=self.ws_fwd, keep_activations=self.keep, reuse_coord_table=self._table_valid())
            ops.raw_learn_objective(c.mode, self.img_pre, c.target, c.mean, c.pmin, c.pmax, c.eps, self.bank, self.model.eps, self.mei_act, self.mask,
                                    r.flat_neighbor_mask_pos, r.flat_neighbor_mask_neg, r.temperature, r.eps, self.g_reg, self.z, self.act.view(-1),
                                    self.reg, self.loss_buf if loss is None else loss, self.g_x, self.ws_obj, stats=self.stats, rmax=self.resp.rmax.view(-1),
                                    argmax=self.resp.argmax.view(-1), K=self.K)
            # backward + Adam as one C-ABI call: one cooperative launch behind the tile kernel instead of four small ones
            ops.raw_inr_backward_adam(t._desc, self.flat, t.B, t.auxB, grid, self.coords, None, self.g_x.view(1, N, Cc, P), self.g_flat, self.m, self.v,
                                      self.step_count, lr=self.lr, precision=self.precision, workspace=self.ws_bwd, forward_workspace=self.ws_fwd,
                                      keep_activations=self.keep)
            return
        ops.raw_inr_forward(t._desc, self.flat, t.B, t.auxB, self.grid, self.coords, None, None, out=self.img_pre.view(1, N, Cc, P),
                            precision=self.precision, workspace=self.ws_fwd, keep_activations=self.keep, reuse_coord_table=self._table_valid())
        ops.raw_constrain_fwd(c.mode, self.img_pre, c.target, c.mean, c.pmin, c.pmax, c.eps, out=self.z, stats=self.stats, workspace=self.ws_con)
        # the response (+ its image gradient -> g_z) and the contrastive forward are independent: run them side by side (the response of a
        # 20-image batch occupies a fraction of the SMs), then add the contrastive gradient on top
        main = torch.cuda.current_stream()
        self.side.wait_stream(main)
        with torch.cuda.stream(self.side):
            self.resp.forward_backward(self.z, 0, N, 1, Cc, P, self.nog, self.g_act, self.act, self.g_z, False)
        ops.raw_simclr_fwd(self.z, self.mask, N, Cc, P, self.pixel_min, self.pixel_max, r.flat_neighbor_mask_pos, r.flat_neighbor_mask_neg,
                           r.temperature, r.eps, reg=self.reg, K=self.K, workspace=self.ws_clr)
        main.wait_stream(self.side)
        ops.raw_simclr_bwd(self.z, self.mask, self.K, self.g_reg, 1.0, N, Cc, P, self.pixel_min, self.pixel_max, self.g_z, accumulate=True)
        ops.raw_constrain_bwd(c.mode, self.img_pre, self.g_z, self.stats, c.target, c.mean, c.pmin, c.pmax, c.eps, out=self.g_x, workspace=self.ws_con)
        ops.raw_inr_backward(t._desc, self.flat, t.B, t.auxB, self.grid, self.coords, None, None, self.g_x.view(1, N, Cc, P), want_params=True,
                             precision=self.precision, workspace=self.ws_bwd, g_params=self.g_flat, forward_workspace=self.ws_fwd, keep_activations=self.keep)
        ops.raw_adam_step(self.flat, self.g_flat, self.m, self.v, self.step_count, lr=self.lr)

    def _capture(self, fn):
        return capture_graph(fn)

    def step(self, grid_row):
        """One optimisation step on the jittered latents `grid_row` ([N] tensor).  A device row is copied into the step's static input.  A HOST row
        (fused-objective path, graphs on) takes the host-fed step: see `_step_from_host`."""
        if not grid_row.is_cuda and self.use_graph and self.ws_obj is not None:
            return self._step_from_host(grid_row)
        self.grid.copy_(grid_row.reshape(-1), non_blocking=True)
        if not self.use_graph:
            self._train()
        elif self.graph is None:
            # the first step runs eagerly (it is also the warm-up that sets kernel attributes); capture afterwards
            self._train()
            self.graph = self._capture(self._train)  # stream capture records the launches without executing them
        else:
            self.graph.replay()

    def _step_from_host(self, grid_row):
        """The step fed from the host without a copy engine: the row is written into a pinned buffer that the preparation kernel reads directly
        (N loads over PCIe, hidden behind the kernel's other blocks) and the objective kernel writes the loss into pinned memory; one graph
        launch per step, `sync_loss()` waits for it.  (Measured against the alternatives on B200, host waiting for every step's loss: H2D copy +
        replay + D2H copy 176 us per step; the two copies as memcpy nodes inside the graph 183 us.)"""
        if self.grid_map is None:
            self.grid_map = torch.empty(self.N, dtype=torch.float32).pin_memory()
            self.loss_map = torch.zeros(1, dtype=torch.float32).pin_memory()
            self.grid_np, self.loss_np = self.grid_map.numpy(), self.loss_map.numpy()  # views: a NumPy write / read costs a tenth of a tensor op
            self.host_done = torch.cuda.Event()
            self._host_fn = lambda: self._train(ops.MappedHost(self.grid_map), ops.MappedHost(self.loss_map))
            self._host_pending = False
        if self._host_pending:
            self.host_done.synchronize()  # the previous host-fed step has read the buffer (and written its loss)
        self.grid_np[:] = grid_row.numpy().reshape(-1)
        if self.graph is None:
            self._host_fn()  # first step of the stepper: eager (see step); the device-fed graph is captured as usual
            self.graph = self._capture(self._train)
        elif self.host_graph is None:
            self.host_graph = self._capture(self._host_fn)
            self.host_graph.replay()
        else:
            self.host_graph.replay()
        self.host_done.record()
        self._host_pending = True

    def sync_loss(self):
        """Loss of the last host-fed step as a Python float (waits for that step)."""
        if self.loss_np is None:
            raise RuntimeError("sync_loss(): no host-fed step has run yet (call step() with a CPU tensor; device-fed steps report loss())")
        if self._host_pending:
            self.host_done.synchronize()
            self._host_pending = False
        return float(self.loss_np[0])

    def evaluate(self, grid_row):
        """No-grad forward: images and the template neuron's activations on `grid_row` (normalised by the MEI activation)."""
        self.grid.copy_(grid_row.reshape(-1), non_blocking=True)
        self._forward()
        return self.act[0] / self.mei_act

    def loss(self):
        """loss of the last training step (device scalar): -mean(act/a*) - reg_scale * reg  (inv_temp_learn_match.py:204)."""
        if self.ws_obj is not None:
            return self.loss_buf[0]  # written by the fused objective kernel itself
        return -(self.act[0] / self.mei_act).mean() + self.g_reg[0] * self.reg[0]

    def last_acts(self):
        return self.act[0] / self.mei_act


class MatchStepper:
    PIN_RING = 8  # pinned staging rows for steps fed with CPU tensors

    def __init__(self, template, img_transf, model, target_idxs, mei_acts, n_grid, lr=1e-3, total_targets=None, precision=None, use_graph=True):
        lib = _lib.load()
        import ctypes as C
        self.t, self.model = template, model
        self.con = _Constraint(img_transf)
        dev = next(template.parameters()).device
        self.dev, self.N, self.lr, self.precision = dev, int(n_grid), lr, precision
        self.C, (H, W) = template.out_channels, template.img_res
        self.P, self.D = H * W, len(target_idxs)
        N, D, Cc, P = self.N, self.D, self.C, self.P
        self.Dtot = int(total_targets) if total_targets is not None else D  # loss = -sum_local / D_total (sharded match)
        self.flat = template.flat_params()
        self.coords = template.coords.reshape(-1, 2).contiguous()
        self.nog = torch.as_tensor(list(target_idxs), dtype=torch.int32, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.mei_acts = torch.as_tensor(mei_acts, dtype=torch.float32).to(dev).reshape(D)
        self.grid = torch.zeros(N, **f32)
        self.img_pre = torch.empty((D, N, Cc, H, W), **f32)
        self.z = torch.empty_like(self.img_pre)
        self.g_z = torch.empty_like(self.img_pre)
        self.g_x = torch.empty_like(self.img_pre)
        self.stats = torch.empty((D * N, 2), **f32)
        self.act = torch.empty((D, N), **f32)
        self.g_act = (-1.0 / (N * self.Dtot * self.mei_acts))[:, None].expand(D, N).contiguous()
        self.aff_mod = template.coordinate_transform.transforms.Affine
        a = self.aff_mod
        self.params = [p for p in (a.angles, a.scalings, a.shears, a.translation)]
        self.trainable = [isinstance(p, torch.nn.Parameter) and p.requires_grad for p in self.params]
        self.grads = [torch.zeros_like(p) for p in self.params]
        self.m = [torch.zeros_like(p) for p in self.params]
        self.v = [torch.zeros_like(p) for p in self.params]
        self.steps = [torch.zeros(2, dtype=torch.int32, device=dev) for _ in self.params]
        d = template._desc
        # The training step walks the targets in blocks of `block` (all D when the kept activations of D targets fit the budget of
        # LMR_KEEP_ACT_MAX_GIB; e.g. BASELINE config 3 with 1024 targets on one GPU: 154 MB per target at 100x100 -> blocks of ~210): forward (keeping the
        # operand tiles), objective and backward of one block, then the next block on the same workspace; Adam once over all targets.  The no-grad
        # passes (evaluate, sweep) keep nothing and run all targets at once.
        self.block, self.keep = match_block(precision, d, N, P, D)
        Bk, rem = self.block, D % self.block
        self.resp = response_backend(model, D, N, P, dev, also=(Bk, rem))
        self.ws_fwd = _ws(max(ops.forward_workspace_bytes(d, N, P, D, False), ops.forward_workspace_bytes(d, N, P, Bk, self.keep)), dev)
        self.ws_bwd = _ws(lib.lmr_inr_backward_workspace_bytes(C.byref(d), N, P, Bk), dev)
        self.ws_con = _ws(lib.lmr_constrain_workspace_bytes(D * N, Cc * P), dev)
        self.use_graph = use_graph
        # (sum, min, max) over the local targets of acts_d = mean_n act[d,n] / a*_d, written by the last launch of every step into one of TWO slots
        # (epoch parity): the match loop queues epoch e + 1 before it reads epoch e's statistics (12 bytes through a pinned buffer, waited for on
        # an event), so the GPU never idles for the host's round trip; `backup` holds the state (parameters, Adam moments, step counters) from
        # before the running epoch, which is what `restore_state` brings back when the stop rule turns out to have fired one epoch earlier
        self.epoch2 = torch.zeros((2, 3), **f32)
        self.epoch = self.epoch2[0]
        self.epoch_host = torch.zeros((2, 3), dtype=torch.float32).pin_memory()
        self.stats_ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.backup = [[x.detach().clone() for x in grp] for grp in (self.params, self.m, self.v, self.steps)]
        self.graphs = {}
        self._pin = None

    def _affine(self):
        return self.t.affine_args()

    def _images(self, keep=False):
        t, c, N, D, Cc, P = self.t, self.con, self.N, self.D, self.C, self.P
        ops.raw_inr_forward(t._desc, self.flat, t.B, t.auxB, self.grid, self.coords, None, self._affine(), out=self.img_pre.view(D, N, Cc, P),
                            precision=self.precision, workspace=self.ws_fwd, keep_activations=keep)
        ops.raw_constrain_fwd(c.mode, self.img_pre.view(D * N, Cc, P), c.target, c.mean, c.pmin, c.pmax, c.eps, out=self.z.view(D * N, Cc, P),
                              stats=self.stats, workspace=self.ws_con)

    def _forward(self):
        N, D, Cc, P = self.N, self.D, self.C, self.P
        self._images()
        self.resp.forward(self.z, N * Cc * P, N, D, Cc, P, self.nog, self.act)

    def _affine_block(self, b0, b1):
        """Affine arguments of targets [b0, b1): views into the live parameters (a graph replay reads the updated values)."""
        D = self.D
        if b0 == 0 and b1 == D:
            return self._affine()
        a, const = self.aff_mod, self.t._affine_const()
        sf = const["shift_from"]
        return ops.AffineArgs(a.angles.detach()[b0:b1], a.scalings.detach()[b0:b1], a.shears.detach()[b0:b1], a.translation.detach().reshape(D, 2)[b0:b1],
                              tc=const["tc"], shift_to=const["shift_to"], shift_from=None if sf is None else sf.reshape(D, 2)[b0:b1], clamp=const["clamp"])

    def _save_state(self):
        for grp, bak in zip((self.params, self.m, self.v, self.steps), self.backup):
            for x, y in zip(grp, bak):
                y.copy_(x.detach(), non_blocking=True)

    def restore_state(self):
        """Back to the state saved by the last step(..., save=True): undoes the (speculatively queued) epoch that followed it."""
        with torch.no_grad():
            for grp, bak in zip((self.params, self.m, self.v, self.steps), self.backup):
                for x, y in zip(grp, bak):
                    x.detach().copy_(y, non_blocking=True)

    def _train(self, slot=0, save=False):
        t, c, N, D, Cc, P = self.t, self.con, self.N, self.D, self.C, self.P
        if save:
            self._save_state()
        for b0 in range(0, D, self.block):
            b1 = min(D, b0 + self.block)
            Db = b1 - b0
            aff = self._affine_block(b0, b1)
            img, z, g_z, g_x = (x[b0:b1] for x in (self.img_pre, self.z, self.g_z, self.g_x))
            ops.raw_inr_forward(t._desc, self.flat, t.B, t.auxB, self.grid, self.coords, None, aff, out=img.view(Db, N, Cc, P), precision=self.precision,
                                workspace=self.ws_fwd, keep_activations=self.keep)
            ops.raw_constrain_fwd(c.mode, img.view(Db * N, Cc, P), c.target, c.mean, c.pmin, c.pmax, c.eps, out=z.view(Db * N, Cc, P),
                                  stats=self.stats[b0 * N:b1 * N], workspace=self.ws_con)
            self.resp.forward_backward(z, N * Cc * P, N, Db, Cc, P, self.nog[b0:b1], self.g_act[b0:b1], self.act[b0:b1], g_z, False)
            ops.raw_constrain_bwd(c.mode, img.view(Db * N, Cc, P), g_z.view(Db * N, Cc, P), self.stats[b0 * N:b1 * N], c.target, c.mean, c.pmin, c.pmax,
                                  c.eps, out=g_x.view(Db * N, Cc, P), workspace=self.ws_con)
            ops.raw_inr_backward(t._desc, self.flat, t.B, t.auxB, self.grid, self.coords, None, aff, g_x.view(Db, N, Cc, P), want_params=False,
                                 want_affine=True, precision=self.precision, workspace=self.ws_bwd,
                                 g_affine=(self.grads[0][b0:b1], self.grads[1][b0:b1], self.grads[2][b0:b1], self.grads[3].view(D, 2)[b0:b1]),
                                 forward_workspace=self.ws_fwd, keep_activations=self.keep)
        groups = [(p.data, g, m, v, s) for p, g, m, v, s, tr in zip(self.params, self.grads, self.m, self.v, self.steps, self.trainable) if tr]
        if groups:  # torch.optim.Adam over the trainable scalars of all targets (:335,366): one launch for the (up to) four tensors
            ops.raw_adam_step_multi(groups, lr=self.lr)
        ops.raw_match_epoch_stats(self.act, self.mei_acts, self.epoch2[slot])

    def epoch_stats(self, slot=0):
        """(sum_d, min_d, max_d) of the last step's normalised activations acts_d (inv_temp_learn_match.py:362,368-371) as host floats."""
        self.queue_stats(slot)
        return self.wait_stats(slot)

    def queue_stats(self, slot):
        """Queues the copy of slot `slot` of the epoch statistics to pinned host memory behind the launches queued so far."""
        self.epoch_host[slot].copy_(self.epoch2[slot], non_blocking=True)
        self.stats_ready[slot].record()

    def wait_stats(self, slot):
        self.stats_ready[slot].synchronize()
        return tuple(self.epoch_host[slot].tolist())

    def step(self, grid_row, slot=0, save=False):
        """One Adam step of every target on the jittered latents `grid_row`.  slot: which of the two statistics slots the step writes;
        save: keep a copy of the state from before the step (first step of an epoch: what `restore_state` goes back to)."""
        if grid_row.is_cuda:
            self.grid.copy_(grid_row.reshape(-1), non_blocking=True)
        else:
            # A CPU row goes through a small ring of pinned rows: the H2D copy is then asynchronous.  (`row.to(device)` from pageable memory
            # returns only when the GPU has drained, so a loop built on it cannot queue an epoch ahead: 57 us of GPU idle time per epoch of
            # match([1, 2]), tools/match_hostprof.py.)  A ring slot is reused after PIN_RING steps; its event says the copy out of it is done.
            if self._pin is None:
                self._pin = torch.empty((self.PIN_RING, self.N), dtype=torch.float32).pin_memory()
                self._pin_np = self._pin.numpy()
                self._pin_ev = [None] * self.PIN_RING
                self._pin_i = 0
            k = self._pin_i
            self._pin_i = (k + 1) % self.PIN_RING
            if self._pin_ev[k] is not None:
                self._pin_ev[k].synchronize()
            else:
                self._pin_ev[k] = torch.cuda.Event()
            self._pin_np[k, :] = grid_row.numpy().reshape(-1)
            self.grid.copy_(self._pin[k], non_blocking=True)
            self._pin_ev[k].record()
        key = (int(slot), bool(save))
        if not self.use_graph:
            self._train(*key)
        elif key not in self.graphs:
            # the first step of a kind runs eagerly (it is also the warm-up that sets kernel attributes); captured afterwards
            self._train(*key)
            self.graphs[key] = capture_graph(lambda: self._train(*key))  # records the launches without executing them
        else:
            self.graphs[key].replay()

    def evaluate(self, grid_row):
        """No-grad forward on `grid_row`: act [D,N] of every target neuron on its own images."""
        self.grid.copy_(grid_row.reshape(-1), non_blocking=True)
        self._forward()
        return self.act

    def sweep(self, grid_row, angles, tx=None, ty=None, max_virtual=128):
        """Mean activation [K, D] of every target under K candidate initialisations (angle_k, optionally translation (tx_k, ty_k)) of its affine
        transform: the init sweep of `match` (inv_temp_learn_match.py:573-650), which the reference runs as K sequential no-grad forwards of all
        targets.  The candidate only enters through the affine matrix, so (candidate, target) pairs are laid out as VIRTUAL targets and evaluated
        `max_virtual` at a time by the same launches as `evaluate` (per-target arithmetic unchanged); with D > max_virtual / 8 this is the
        reference's loop, one candidate per pass, on the stepper's own buffers.  The table never leaves the device."""
        t, c, N, D, Cc, P, dev = self.t, self.con, self.N, self.D, self.C, self.P, self.dev
        lib = _lib.load()
        import ctypes as C
        f32 = dict(dtype=torch.float32, device=dev)
        ang = torch.as_tensor(angles, **f32).reshape(-1)
        K = ang.numel()
        txd = None if tx is None else torch.as_tensor(tx, **f32).reshape(-1)
        tyd = None if ty is None else torch.as_tensor(ty, **f32).reshape(-1)
        per = int(max_virtual) // D
        if per < 8:  # a pass over D >= max_virtual / 8 targets already fills the GPU for milliseconds: nothing to gain from extra buffers
            per = 1
        out = torch.empty((K, D), **f32)
        self.grid.copy_(grid_row.reshape(-1), non_blocking=True)
        a = self.aff_mod
        const = t._affine_const()
        scal, shear = a.scalings.detach(), a.shears.detach()
        trans0 = a.translation.detach().reshape(D, 2)
        sf = const["shift_from"]
        Dmax = min(per, K) * D
        own = Dmax == D  # one candidate per pass: the stepper's own buffers
        img = self.img_pre if own else torch.empty((Dmax, N, Cc, P), **f32)
        z = self.z if own else torch.empty_like(img)
        stats = self.stats if own else torch.empty((Dmax * N, 2), **f32)
        act = self.act if own else torch.empty((Dmax, N), **f32)
        ws_fwd = self.ws_fwd if own else _ws(ops.forward_workspace_bytes(t._desc, N, P, Dmax, False), dev)
        ws_con = self.ws_con if own else _ws(lib.lmr_constrain_workspace_bytes(Dmax * N, Cc * P), dev)
        resp = self.resp if own else response_backend(self.model, Dmax, N, P, dev)
        for k0 in range(0, K, per):
            nc = min(per, K - k0)
            Dv = nc * D
            trans = trans0.repeat(nc, 1)
            if txd is not None:
                trans[:, 0] = txd[k0:k0 + nc].repeat_interleave(D)
                trans[:, 1] = tyd[k0:k0 + nc].repeat_interleave(D)
            aff = ops.AffineArgs(ang[k0:k0 + nc].repeat_interleave(D), scal.repeat(nc, 1), shear.repeat(nc, 1), trans, tc=const["tc"],
                                 shift_to=const["shift_to"], shift_from=None if sf is None else sf.reshape(D, 2).repeat(nc, 1), clamp=const["clamp"])
            ops.raw_inr_forward(t._desc, self.flat, t.B, t.auxB, self.grid, self.coords, None, aff, out=img.view(-1, N, Cc, P)[:Dv],
                                precision=self.precision, workspace=ws_fwd)
            ops.raw_constrain_fwd(c.mode, img.view(-1, Cc, P)[:Dv * N], c.target, c.mean, c.pmin, c.pmax, c.eps, out=z.view(-1, Cc, P)[:Dv * N],
                                  stats=stats[:Dv * N], workspace=ws_con)
            resp.forward(z, N * Cc * P, N, Dv, Cc, P, self.nog.repeat(nc), act[:Dv])
            torch.mean(act[:Dv], dim=1, out=out[k0:k0 + nc].view(-1))
        return out

    def normalised_acts(self):
        """acts_d = mean_n act[d,n] / a*_d of the last step (inv_temp_learn_match.py:362)."""
        return self.act.mean(dim=1) / self.mei_acts
