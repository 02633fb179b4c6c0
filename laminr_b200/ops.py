"""PyTorch-facing operators over the C ABI (include/laminr_b200.h).

Two levels:
  * `raw_*` functions: thin stream-ordered wrappers (device pointers in, nothing allocated inside the library);
    these are what the fused learn/match/MEI drivers call directly and what CUDA-graph capture records.
  * `torch.autograd.Function`s (`inr_images`, `constrain`, `simresp`, `simclr_reg`): differentiable operators so
    the drop-in modules compose with stock autograd (e.g. a user-supplied nn.Module response model).

PyTorch is plumbing here (device memory, streams); every op fails loudly without a CUDA device / the library.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import Affine, InrDesc, check

_DEFAULT_PRECISION = _lib.PREC_FP32
_PREC_NAMES = {"fp32": _lib.PREC_FP32, "bf16x3_fast": _lib.PREC_BF16X3_FAST, "bf16x3": _lib.PREC_BF16X3}


def set_default_precision(name):
    """'fp32' (CUDA cores, bit-tight parity path) | 'bf16x3' | 'bf16x3_fast' (tcgen05 split-bf16 paths) for the INR hidden layers."""
    global _DEFAULT_PRECISION
    _DEFAULT_PRECISION = _PREC_NAMES[name]


def default_precision():
    return _DEFAULT_PRECISION


def precision_code(p):
    if p is None:
        return _DEFAULT_PRECISION
    return _PREC_NAMES[p] if isinstance(p, str) else int(p)


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "laminr_b200 operators run on CUDA (sm_100a) only; got a tensor on %s. There is no CPU fallback." % t.device
            )


class MappedHost:
    """A pinned (page-locked) host tensor handed to a kernel as a pointer argument: with unified addressing every cudaHostAlloc'ed block is mapped
    into the device's address space at the same address, so a kernel can read a few input scalars from it / write a result scalar into it without
    a copy engine in between (a 4..80-byte cudaMemcpyAsync costs more GPU idle time than the PCIe round trip of the load itself).  Only for such
    scalars: every access crosses PCIe."""
    is_cuda = True

    def __init__(self, t):
        if t.is_cuda or not t.is_pinned() or not t.is_contiguous() or t.dtype != torch.float32:
            raise ValueError("MappedHost: a contiguous pinned float32 host tensor is required")
        self.t = t

    def numel(self):
        return self.t.numel()

    def data_ptr(self):
        return self.t.data_ptr()


def _ptr(t):
    """Device pointer of `t` for the C ABI (every pointer argument of include/laminr_b200.h is device memory): a host tensor here would be an
    illegal address inside the kernel, so it is refused."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("laminr_b200 operators run on CUDA (sm_100a) only; got a tensor on %s. There is no CPU fallback." % t.device)
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t):
    if t.dtype != torch.float32:
        raise TypeError("laminr_b200 operators take float32 tensors")
    return t if t.is_contiguous() else t.contiguous()


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------------------------------
# results -> host (result getters, inv_temp_learn_match.py:394-422: `.cpu().data.numpy()`)
# ------------------------------------------------------------------------------------------------------------------
_HOST_RING = {}
HOST_CHUNK_BYTES = 8 << 20
_HOST_RING_LOCK = __import__("threading").Lock()


def _host_ring(dtype, step, nbuf):
    """The persistent ring of pinned staging buffers (+ its worker pool) for `to_host_numpy`.  Allocating it is a cudaHostAlloc (~1 ms per MB, once per
    process) that also stalls kernel launches from other threads while it runs, so the ring is kept small: nbuf x 8 MB chunks still move at PCIe speed."""
    import concurrent.futures as cf
    key = (dtype, step, nbuf)
    with _HOST_RING_LOCK:
        if key not in _HOST_RING:
            _HOST_RING[key] = ([torch.empty(step, dtype=dtype, pin_memory=True) for _ in range(nbuf)], cf.ThreadPoolExecutor(nbuf))
        return _HOST_RING[key]


def to_host_numpy(t, chunk_bytes=HOST_CHUNK_BYTES, nbuf=4, out=None):
    """`t.cpu().numpy()` for large device tensors (the [D,N,C,H,W] image blocks of match: 0.8 GB for 1023 targets at 100x100): the device-to-host
    copies run asynchronously through a small ring of persistent pinned buffers while worker threads move finished chunks into the destination
    array (a pageable `.cpu()` of that block takes 0.36 s on the B200 box, the first-touch host copy being the bottleneck; this takes about half).
    out: optional flat, writable numpy array of t.numel() elements to fill instead of a new one (e.g. a slice of a shared-memory segment)."""
    import numpy as np

    nbytes = t.numel() * t.element_size()
    if not t.is_cuda or nbytes < (chunk_bytes if out is not None else 16 * chunk_bytes):
        host = t.detach().cpu().numpy()
        if out is None:
            return host
        out[...] = host.reshape(-1)
        return out.reshape(tuple(t.shape))
    flat = t.detach().contiguous().reshape(-1)
    step = chunk_bytes // flat.element_size()
    bufs, pool = _host_ring(flat.dtype, step, nbuf)
    if out is None:
        out = np.empty(flat.numel(), dtype=bufs[0].numpy().dtype)
    side = torch.cuda.Stream(device=t.device)
    side.wait_stream(torch.cuda.current_stream(t.device))

    def drain(ev, b, m, o):
        ev.synchronize()
        out[o:o + m] = bufs[b][:m].numpy()

    pending = [None] * nbuf
    for i, o in enumerate(range(0, flat.numel(), step)):
        b, m = i % nbuf, min(step, flat.numel() - o)
        if pending[b] is not None:
            pending[b].result()  # the chunk that used this pinned buffer has reached the destination
        with torch.cuda.stream(side):
            bufs[b][:m].copy_(flat[o:o + m], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
        pending[b] = pool.submit(drain, ev, b, m, o)
    for f in pending:
        if f is not None:
            f.result()
    return out.reshape(tuple(t.shape))


class HostBuffer:
    """A host destination for a large result, made resident BEFORE the result exists: a background thread touches (zeroes) every page while the
    GPU still computes, so the final device-to-host copy runs at PCIe speed instead of at page-fault speed (first touch of 0.8 GB: 56-170 ms on
    the B200 box, as long as a whole 32-target match call).  `flat`: writable flat NumPy array (a new anonymous array, or a slice of a
    shared-memory segment handed in by laminr_b200.dist)."""

    def __init__(self, numel, dtype, flat=None, device=None):
        import ctypes
        import threading

        import numpy as np

        self.flat = np.empty(int(numel), dtype=dtype) if flat is None else flat
        addr, nbytes = self.flat.ctypes.data, self.flat.nbytes
        tdtype = torch.from_numpy(np.empty(0, dtype=self.flat.dtype)).dtype
        dev = torch.cuda.current_device() if device is None and torch.cuda.is_available() else device

        def touch():
            if dev is not None and torch.cuda.is_available():
                torch.cuda.set_device(dev)
                _host_ring(tdtype, HOST_CHUNK_BYTES // self.flat.dtype.itemsize, 4)  # the pinned staging ring of to_host_numpy, if this process has none yet
            step = 64 << 20
            for o in range(0, nbytes, step):  # (ctypes releases the GIL for the duration of each call)
                ctypes.memset(addr + o, 0, min(step, nbytes - o))

        self._thread = threading.Thread(target=touch, daemon=True)
        self._thread.start()

    def wait(self):
        self._thread.join()
        return self.flat


# ------------------------------------------------------------------------------------------------------------------
# INR
# ------------------------------------------------------------------------------------------------------------------
class AffineArgs:
    """Device tensors of the per-target affine transform (match stage); see lmr_affine in the header."""

    def __init__(self, angles, scalings, shears, translation, tc=None, shift_to=None, shift_from=None, clamp=True):
        self.angles, self.scalings, self.shears = _f32c(angles), _f32c(scalings), _f32c(shears)
        self.translation = _f32c(translation).reshape(-1, 2)
        self.tc = None if tc is None else _f32c(tc)
        self.shift_to = None if shift_to is None else _f32c(shift_to)
        self.shift_from = None if shift_from is None else _f32c(shift_from)
        self.clamp = bool(clamp)
        self.D = self.angles.numel()

    def struct(self):
        return Affine(
            self.angles.data_ptr(), self.scalings.data_ptr(), int(self.scalings.shape[-1]) if self.scalings.dim() > 1 else 1,
            self.shears.data_ptr(), self.translation.data_ptr(),
            None if self.tc is None else self.tc.data_ptr(), None if self.shift_to is None else self.shift_to.data_ptr(),
            None if self.shift_from is None else self.shift_from.data_ptr(), int(self.clamp),
        )


def make_desc(hidden, n_hidden, pe_dim, aux_pe_dim, channels):
    return InrDesc(int(hidden), int(n_hidden), int(pe_dim), int(aux_pe_dim), int(channels))


def forward_workspace_bytes(desc, N, P, D, keep_activations=False):
    lib = _lib.load()
    fn = lib.lmr_inr_forward_workspace_bytes_keep if keep_activations else lib.lmr_inr_forward_workspace_bytes
    return int(fn(C.byref(desc), N, P, D))


def raw_inr_forward(desc, params, B, auxB, grid, coords, coord_shift, affine, out=None, precision=None, workspace=None, keep_activations=False,
                    reuse_coord_table=False):
    """-> img [D, N, C, P] (float32).  params: flat parameter vector (see header); grid [N]; coords [P,2].
    keep_activations (tensor-core precisions): the forward keeps the operand tiles of the hidden activations in `workspace` (size:
    forward_workspace_bytes(..., keep_activations=True)) for a raw_inr_backward(..., forward_workspace=workspace, keep_activations=True).
    reuse_coord_table (tensor-core precisions, no affine): `workspace` still holds the coordinate sin/cos table of the previous call with the same
    B / coords / coord_shift (LMR_PREC_REUSE_COORD_TABLE); bit-identical results, no sin/cos evaluation."""
    lib = _lib.load()
    _require_cuda(params, B, auxB, grid, coords)
    N, P = grid.numel(), coords.shape[0]
    D = 1 if affine is None else affine.D
    if out is None:
        out = torch.empty((D, N, desc.channels, P), dtype=torch.float32, device=params.device)
    prec = precision_code(precision)
    keep = bool(keep_activations) and prec != _lib.PREC_FP32
    if workspace is None:
        workspace = _workspace(forward_workspace_bytes(desc, N, P, D, keep), params.device)
    aff = affine.struct() if affine is not None else None
    flags = (_lib.PREC_KEEP_ACTIVATIONS if keep else 0) | (_lib.PREC_REUSE_COORD_TABLE if reuse_coord_table and aff is None and prec != _lib.PREC_FP32 else 0)
    check(lib.lmr_inr_forward(
        C.byref(desc), _ptr(params), _ptr(B), _ptr(auxB), _ptr(grid), N, _ptr(coords), P, _ptr(coord_shift),
        C.byref(aff) if aff is not None else None, D, _ptr(out), _ptr(workspace), workspace.numel(), prec | flags, _stream(),
    ))
    return out


def raw_inr_backward(desc, params, B, auxB, grid, coords, coord_shift, affine, g_img, want_params=True, want_affine=False,
                     precision=None, workspace=None, g_params=None, g_affine=None, forward_workspace=None, keep_activations=False):
    """-> (g_params flat | None, (g_angles, g_scalings, g_shears, g_translation) | None).
    forward_workspace: workspace of the raw_inr_forward call on the same inputs (its per-step tables are reused: one launch less);
    keep_activations: that forward call kept the hidden activations' operand tiles there (no recompute in the backward)."""
    lib = _lib.load()
    _require_cuda(params, g_img)
    N, P = grid.numel(), coords.shape[0]
    D = 1 if affine is None else affine.D
    dev = params.device
    if want_params and g_params is None:
        g_params = torch.empty_like(params)
    if want_affine and g_affine is None:
        g_affine = (torch.empty_like(affine.angles), torch.empty_like(affine.scalings), torch.empty_like(affine.shears),
                    torch.empty_like(affine.translation))
    if workspace is None:
        workspace = _workspace(lib.lmr_inr_backward_workspace_bytes(C.byref(desc), N, P, D), dev)
    aff = affine.struct() if affine is not None else None
    ga = g_affine if want_affine else (None, None, None, None)
    common = (C.byref(desc), _ptr(params), _ptr(B), _ptr(auxB), _ptr(grid), N, _ptr(coords), P, _ptr(coord_shift),
              C.byref(aff) if aff is not None else None, D, _ptr(_f32c(g_img)), _ptr(g_params if want_params else None),
              _ptr(ga[0]), _ptr(ga[1]), _ptr(ga[2]), _ptr(ga[3]), _ptr(workspace), workspace.numel())
    if forward_workspace is not None:
        prec = precision_code(precision)
        keep = bool(keep_activations) and prec != _lib.PREC_FP32
        check(lib.lmr_inr_backward_after_forward(*common, _ptr(forward_workspace), forward_workspace.numel(), prec | (_lib.PREC_KEEP_ACTIVATIONS if keep else 0),
                                                 _stream()))
    else:
        check(lib.lmr_inr_backward(*common, precision_code(precision), _stream()))
    return (g_params if want_params else None), (g_affine if want_affine else None)


def raw_inr_backward_adam(desc, params, B, auxB, grid, coords, coord_shift, g_img, g_params, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8,
                          precision=None, workspace=None, forward_workspace=None, keep_activations=False):
    """The learn step's raw_inr_backward(..., forward_workspace=...) + raw_adam_step(params, g_params, m, v, step) as ONE C-ABI call
    (lmr_inr_backward_adam): on the tensor-core precisions the reductions behind the tile kernel and the Adam update share one launch.
    `params` is updated in place, `g_params` receives the gradient."""
    lib = _lib.load()
    _require_cuda(params, g_img, g_params, m, v, step)
    if step.numel() < 2 or step.dtype != torch.int32:
        raise ValueError("raw_inr_backward_adam: `step` must be an int32 tensor with 2 elements ([steps taken, zero-initialised scratch])")
    N, P = grid.numel(), coords.shape[0]
    if workspace is None:
        workspace = _workspace(lib.lmr_inr_backward_workspace_bytes(C.byref(desc), N, P, 1), params.device)
    prec = precision_code(precision)
    keep = bool(keep_activations) and prec != _lib.PREC_FP32
    check(lib.lmr_inr_backward_adam(C.byref(desc), _ptr(params), _ptr(B), _ptr(auxB), _ptr(grid), N, _ptr(coords), P, _ptr(coord_shift), _ptr(_f32c(g_img)),
                                    _ptr(g_params), _ptr(m), _ptr(v), _ptr(step), float(lr), float(b1), float(b2), float(eps), _ptr(workspace),
                                    workspace.numel(), _ptr(forward_workspace), 0 if forward_workspace is None else forward_workspace.numel(),
                                    prec | (_lib.PREC_KEEP_ACTIVATIONS if keep else 0), _stream()))
    return g_params


def inr_images(meta, flat_params, B, auxB, grid, coords, coord_shift, affine_params, param_views):
    """img[D,N,C,P] = INR(...), differentiable wrt the MLP parameters (learn) and the affine scalars (match): the registered operator
    torch.ops.laminr_b200.inr_images (laminr_b200/torch_ops.py), whose backward is torch.ops.laminr_b200.inr_backward."""
    from . import torch_ops  # noqa: F401  (registers the operators)
    desc, const = meta["desc"], meta["affine_const"]
    a = affine_params if affine_params is not None else (None, None, None, None)
    has_aff = a[0] is not None
    return torch.ops.laminr_b200.inr_images(
        flat_params, list(param_views), B, auxB, grid, coords, coord_shift, a[0], a[1], a[2], a[3],
        const.get("tc") if has_aff else None, const.get("shift_to") if has_aff else None, const.get("shift_from") if has_aff else None,
        bool(const.get("clamp", True)), int(desc.hidden), int(desc.n_hidden), int(desc.pe_dim), int(desc.aux_pe_dim), int(desc.channels),
        str(meta["precision"] or default_precision()))


# ------------------------------------------------------------------------------------------------------------------
# constraint projection
# ------------------------------------------------------------------------------------------------------------------
def raw_constrain_fwd(mode, x, target, mean, pmin, pmax, eps, out=None, stats=None, workspace=None):
    lib = _lib.load()
    _require_cuda(x)
    Bn, n = x.shape[0], x[0].numel()
    if out is None:
        out = torch.empty_like(x)
    if stats is None:
        stats = torch.empty((Bn, 2), dtype=torch.float32, device=x.device)
    if workspace is None:
        workspace = _workspace(lib.lmr_constrain_workspace_bytes(Bn, n), x.device)
    check(lib.lmr_constrain_fwd(mode, _ptr(x), Bn, n, float(target), float(mean), int(pmin is not None), float(pmin or 0.0),
                                int(pmax is not None), float(pmax or 0.0), float(eps), _ptr(out), _ptr(stats), _ptr(workspace),
                                workspace.numel(), _stream()))
    return out, stats


def raw_constrain_bwd(mode, x, g_z, stats, target, mean, pmin, pmax, eps, out=None, workspace=None):
    lib = _lib.load()
    Bn, n = x.shape[0], x[0].numel()
    if out is None:
        out = torch.empty_like(x)
    if workspace is None:
        workspace = _workspace(lib.lmr_constrain_workspace_bytes(Bn, n), x.device)
    check(lib.lmr_constrain_bwd(mode, _ptr(x), _ptr(g_z), _ptr(stats), Bn, n, float(target), float(mean), int(pmin is not None),
                                float(pmin or 0.0), int(pmax is not None), float(pmax or 0.0), float(eps), _ptr(out), _ptr(workspace),
                                workspace.numel(), _stream()))
    return out


def constrain(x, mode, target, mean, pmin, pmax, eps=1e-12):
    """Differentiable constraint projection: torch.ops.laminr_b200.constrain (backward: constrain_backward)."""
    from . import torch_ops  # noqa: F401
    z, _stats = torch.ops.laminr_b200.constrain(x, int(mode), float(target), float(mean), pmin is not None, float(pmin or 0.0), pmax is not None,
                                                float(pmax or 0.0), float(eps))
    return z


# ------------------------------------------------------------------------------------------------------------------
# simulated response
# ------------------------------------------------------------------------------------------------------------------
def raw_simresp_fwd(img, group_stride, NB, G, Cc, HW, filters, nfilt, neuron_of_group, eps, out=None, workspace=None):
    """-> act, rmax, argmax, each [G, NB]."""
    lib = _lib.load()
    _require_cuda(img, filters)
    Fmax = filters.shape[1]
    dev = img.device
    if out is None:
        out = (torch.empty((G, NB), dtype=torch.float32, device=dev), torch.empty((G, NB), dtype=torch.float32, device=dev),
               torch.empty((G, NB), dtype=torch.int32, device=dev))
    if workspace is None:
        workspace = _workspace(lib.lmr_simresp_workspace_bytes(G, NB, Fmax, HW), dev)
    check(lib.lmr_simresp_fwd(_ptr(img), int(group_stride), NB, G, Cc, HW, _ptr(filters), Fmax, _ptr(nfilt), _ptr(neuron_of_group),
                              float(eps), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(workspace), workspace.numel(), _stream()))
    return out


def raw_simresp_bwd(g_act, rmax, argmax, group_stride, NB, G, Cc, HW, filters, neuron_of_group, g_img, accumulate=False):
    lib = _lib.load()
    check(lib.lmr_simresp_bwd(_ptr(g_act), _ptr(rmax), _ptr(argmax), int(group_stride), NB, G, Cc, HW, _ptr(filters), filters.shape[1],
                              _ptr(neuron_of_group), _ptr(g_img), int(accumulate), _stream()))
    return g_img


def simresp(x, filters, nfilt, neuron_of_group, eps=1e-6):
    """x [B,C,H,W] -> act [G,B] for the neurons in `neuron_of_group` (every neuron sees every image), differentiable wrt x:
    torch.ops.laminr_b200.simresp (backward: simresp_backward)."""
    from . import torch_ops  # noqa: F401
    act, _rmax, _argmax = torch.ops.laminr_b200.simresp(x, filters, nfilt, neuron_of_group, float(eps))
    return act


# ------------------------------------------------------------------------------------------------------------------
# stacked-conv core + Gaussian readout response (receptive-field cone of the selected neuron)
# ------------------------------------------------------------------------------------------------------------------
def _pack_cone(w):
    """[c_in][k][c_out][k] -> [c_in][k][c_out / CB][k][CB], CB = 4 / 2 / 1 by the divisibility of c_out (csrc/convcone.cu pick_cb): the k * CB
    weights a thread needs for one (input channel, kernel row) are contiguous."""
    cin, k, cout, _ = w.shape
    cb = 4 if cout % 4 == 0 else 2 if cout % 2 == 0 else 1
    return w.reshape(cin, k, cout // cb, cb, k).permute(0, 1, 2, 4, 3).contiguous()


class ConvCone:
    """Device-side weight pack + descriptor of a FiringRateEncoder(Stacked2dCore, FullGaussian2d) in eval mode, in the layout
    lmr_convcore_response reads (include/laminr_b200.h).  Built once from the module's parameters (the response model is frozen)."""

    def __init__(self, layers, mu, features, bias, height, width, xshift=0.0, yshift=0.0, offset=0.0, align_corners=True):
        """layers: list of dict(weight [Co,Ci,k,k], conv_bias|None, bn=(running_mean, running_var, gamma|None, beta|None, eps)|None, nonlin: bool)."""
        _require_cuda(mu, features)
        if not 1 <= len(layers) <= _lib.CONV_MAX_LAYERS:
            raise ValueError(f"ConvCone: 1..{_lib.CONV_MAX_LAYERS} layers supported")
        d, w = _lib.ConvcoreDesc(), _lib.ConvcoreWeights()
        self.keep = []
        hidden = layers[0]["weight"].shape[0]
        d.n_layers, d.c_in, d.hidden = len(layers), layers[0]["weight"].shape[1], hidden
        for l, L in enumerate(layers):
            W = L["weight"].detach().to(torch.float32)
            co, ci, k, k2 = W.shape
            if co != hidden or k != k2 or k % 2 == 0 or (l > 0 and ci != hidden):
                raise ValueError("ConvCone: every layer must be a square odd-kernel convolution with `hidden` output channels")
            d.kernel[l], d.nonlin[l] = k, int(bool(L["nonlin"]))
            wf = _pack_cone(W.permute(1, 2, 0, 3))                # [ci][ky][co][kx]
            wb = _pack_cone(W.flip(2, 3).permute(0, 2, 1, 3))     # [co][ky][ci][kx] of the flipped kernel
            # eval-mode batch norm + conv bias folded into v*scale + shift (float64 on the way, like the oracle)
            scale, shift = torch.ones(co, dtype=torch.float64, device=W.device), torch.zeros(co, dtype=torch.float64, device=W.device)
            if L.get("conv_bias") is not None:
                shift = shift + L["conv_bias"].detach().double()
            if L.get("bn") is not None:
                rm, rv, gamma, beta, eps = L["bn"]
                scale = 1.0 / torch.sqrt(rv.detach().double() + eps)
                if gamma is not None:
                    scale = scale * gamma.detach().double()
                shift = (shift - rm.detach().double()) * scale
                if beta is not None:
                    shift = shift + beta.detach().double()
            scale, shift = scale.float().contiguous(), shift.float().contiguous()
            self.keep += [wf, wb, scale, shift]
            w.w_fwd[l], w.w_bwd[l], w.scale[l], w.shift[l] = wf.data_ptr(), wb.data_ptr(), scale.data_ptr(), shift.data_ptr()
        self.mu = mu.detach().to(torch.float32).reshape(-1, 2).contiguous()
        self.n_neurons = self.mu.shape[0]
        self.features = features.detach().to(torch.float32).reshape(hidden, self.n_neurons).t().contiguous()  # [n, hidden]
        self.bias = None if bias is None else bias.detach().to(torch.float32).contiguous()
        w.mu, w.features, w.bias = self.mu.data_ptr(), self.features.data_ptr(), (self.bias.data_ptr() if self.bias is not None else None)
        d.elu_xshift, d.elu_yshift, d.elu_offset = float(xshift), float(yshift), float(offset)
        d.height, d.width, d.align_corners = int(height), int(width), int(bool(align_corners))
        self.desc, self.weights, self.device = d, w, self.mu.device
        self.c_in, self.HW = d.c_in, int(height) * int(width)
        self.cone_edge = 2 + sum(int(d.kernel[l]) - 1 for l in range(d.n_layers))


def raw_convcore_response(cone, img, group_stride, NB, G, neuron_of_group, g_act=None, act=None, g_img=None, accumulate=False):
    """act [G,NB] (and, if g_img is given, g_img (+)= d sum(g_act*act)/d img) of neuron_of_group[g] on the NB images of group g."""
    lib = _lib.load()
    _require_cuda(img, neuron_of_group, g_act, g_img)
    if act is None:
        act = torch.empty((G, NB), dtype=torch.float32, device=img.device)
    check(lib.lmr_convcore_response(C.byref(cone.desc), C.byref(cone.weights), _ptr(img), int(group_stride), NB, G, _ptr(neuron_of_group), _ptr(g_act),
                                    _ptr(act), _ptr(g_img), int(accumulate), _stream()))
    return act


class _ConvResp(torch.autograd.Function):
    """x [B,C,H,W] -> act [G,B] of the neurons in `neuron_of_group` (every listed neuron sees every image)."""

    @staticmethod
    def forward(ctx, x, cone, neuron_of_group):
        x = _f32c(x)
        ctx.cone, ctx.nog = cone, neuron_of_group
        ctx.save_for_backward(x)
        return raw_convcore_response(cone, x, 0, x.shape[0], neuron_of_group.numel(), neuron_of_group)

    @staticmethod
    def backward(ctx, g_act):
        (x,) = ctx.saved_tensors
        cone, nog = ctx.cone, ctx.nog
        g_act = _f32c(g_act)
        g_img = torch.zeros_like(x)
        scratch = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        live = torch.nonzero(g_act.abs().sum(dim=1) != 0).flatten().tolist() if nog.numel() > 1 else [0]  # usually ONE neuron carries gradient
        for gi in live:  # one launch per neuron with upstream gradient: each CTA owns its image, so the groups accumulate launch after launch
            raw_convcore_response(cone, x, 0, x.shape[0], 1, nog[gi:gi + 1], g_act=g_act[gi].contiguous(), act=scratch, g_img=g_img, accumulate=True)
        return g_img, None, None


def convresp(x, cone, neuron_of_group):
    return _ConvResp.apply(x, cone, neuron_of_group)


def fold_affine(c_out, conv_bias=None, bn=None, device=None):
    """Eval-mode batch norm (running_mean, running_var, gamma|None, beta|None, eps) + convolution bias folded into v*scale + shift (float64 on
    the way, like the oracle).  Returns (None, None) when there is neither."""
    if conv_bias is None and bn is None:
        return None, None
    scale, shift = torch.ones(c_out, dtype=torch.float64, device=device), torch.zeros(c_out, dtype=torch.float64, device=device)
    if conv_bias is not None:
        shift = shift + conv_bias.detach().double()
    if bn is not None:
        rm, rv, gamma, beta, eps = bn
        scale = 1.0 / torch.sqrt(rv.detach().double() + eps)
        if gamma is not None:
            scale = scale * gamma.detach().double()
        shift = (shift - rm.detach().double()) * scale
        if beta is not None:
            shift = shift + beta.detach().double()
    return scale.float().contiguous(), shift.float().contiguous()


def _pack_dense(W):
    """[co, ci, k, k] -> the [ci][co / CO][k][kp][CO] layout of LMR_FF_DENSE (include/laminr_b200.h)."""
    co, ci, k, _ = W.shape
    CO, kp = (4 if co % 4 == 0 else 1), (k + 7) // 8 * 8
    Wp = torch.zeros(co, ci, k, kp, dtype=torch.float32, device=W.device)
    Wp[..., :k] = W
    return Wp.reshape(co // CO, CO, ci, k, kp).permute(2, 0, 3, 4, 1).contiguous()


class FullFrameNet:
    """Device-side weight pack + lmr_ff_net descriptor of a frozen, eval-mode conv response model, in the layouts lmr_ff_forward / lmr_ff_backward
    read (include/laminr_b200.h).  `layers`: list of dicts
        dict(kind="dense"|"depthwise"|"pointwise", weight=conv.weight, pad=, dilation=, conv_bias=|None, bn=(mean, var, gamma, beta, eps)|None,
             elu=(xshift, yshift)|None)
        dict(kind="se", w1=[r,c], b1=[r], w2=[c,r], b2=[c])
    readout: grid [n,2] (or any shape with n*2 elements), features [n_scales*C, n] (scale-major rows, the reference's layout), bias [n]|None."""

    def __init__(self, layers, in_shape, grid, features, bias, n_scales=1, pool_kern=1, offset=0.0, align_corners=True, need_grad=True):
        _require_cuda(grid, features)
        if not 1 <= len(layers) <= _lib.FF_MAX_LAYERS:
            raise ValueError(f"FullFrameNet: 1..{_lib.FF_MAX_LAYERS} layers supported")
        net = _lib.FFNet()
        self.keep = []
        dev = grid.device
        hold = lambda t: (self.keep.append(t), t.data_ptr())[1]
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        net.n_layers = len(layers)
        net.c_in, net.height, net.width = (int(v) for v in in_shape)
        c = net.c_in
        for l, L in enumerate(layers):
            d = net.layers[l]
            kind = L["kind"]
            d.c_in, d.dilation, d.k = c, 1, 1
            if kind == "se":
                w1, b1, w2, b2 = f32(L["w1"]), f32(L["b1"]), f32(L["w2"]), f32(L["b2"])
                if w1.shape[1] != c or tuple(w2.shape) != (c, w1.shape[0]):
                    raise ValueError(f"FullFrameNet layer {l}: squeeze-excitation weights do not match {c} channels")
                d.kind, d.c_out, d.se_hidden = _lib.FF_SE, c, w1.shape[0]
                d.se_w1, d.se_b1, d.se_w2, d.se_b2 = hold(w1), hold(b1), hold(w2), hold(b2)
                continue
            W = f32(L["weight"])
            co, cig, k, k2 = W.shape
            if k != k2:
                raise ValueError(f"FullFrameNet layer {l}: square kernels only")
            d.c_out, d.k, d.pad, d.dilation = co, k, int(L.get("pad", 0)), int(L.get("dilation", 1))
            if kind == "dense":
                if cig != c:
                    raise ValueError(f"FullFrameNet layer {l}: {cig} input channels, {c} expected")
                d.kind = _lib.FF_DENSE
                d.w_fwd = hold(_pack_dense(W))
                if need_grad:
                    d.w_bwd = hold(_pack_dense(W.flip(2, 3).transpose(0, 1).contiguous()))
            elif kind == "depthwise":
                if cig != 1 or co != c:
                    raise ValueError(f"FullFrameNet layer {l}: depthwise convolution needs groups == channels")
                d.kind = _lib.FF_DEPTHWISE
                d.w_fwd = hold(W.reshape(co, k, k).contiguous())
                if need_grad:
                    d.w_bwd = hold(W.flip(2, 3).reshape(co, k, k).contiguous())
            elif kind == "pointwise":
                if cig != c or k != 1:
                    raise ValueError(f"FullFrameNet layer {l}: pointwise convolution needs a 1 x 1 kernel over {c} channels")
                d.kind = _lib.FF_POINTWISE
                d.w_fwd = hold(W.reshape(co, cig).t().contiguous())
                if need_grad:
                    d.w_bwd = hold(W.reshape(co, cig).contiguous())
            else:
                raise ValueError(f"FullFrameNet layer {l}: unknown kind {kind!r}")
            scale, shift = fold_affine(co, L.get("conv_bias"), L.get("bn"), device=dev)
            if scale is not None:
                d.affine, d.scale, d.shift = 1, hold(scale), hold(shift)
            if L.get("elu") is not None:
                d.elu, d.elu_xshift, d.elu_yshift = 1, float(L["elu"][0]), float(L["elu"][1])
            c = co
        self.grid = f32(grid).reshape(-1, 2).clamp(-1, 1).contiguous()
        n = self.grid.shape[0]
        feats = f32(features).reshape(-1, n)
        if feats.shape[0] != n_scales * c:
            raise ValueError(f"FullFrameNet: features have {feats.shape[0]} rows, {n_scales} scales x {c} channels expected")
        self.features = feats.t().contiguous()
        self.bias = None if bias is None else f32(bias)
        net.n_neurons, net.n_scales, net.pool_kern, net.align_corners = n, int(n_scales), int(pool_kern), int(bool(align_corners))
        net.elu_offset = float(offset)
        net.grid, net.features, net.bias = self.grid.data_ptr(), self.features.data_ptr(), (self.bias.data_ptr() if self.bias is not None else None)
        self.net, self.device, self.n_neurons, self.need_grad = net, dev, n, need_grad
        self.in_shape = (net.c_in, net.height, net.width)
        self._ws = {}

    def workspace(self, B):
        ws = self._ws.get(B)
        if ws is None:
            nbytes = _lib.load().lmr_ff_workspace_bytes(C.byref(self.net), int(B))
            if nbytes == 0:
                raise _lib.LaminrB200Error("laminr_b200: " + _lib.load().lmr_last_error().decode())
            self._ws = {B: _workspace(nbytes, self.device)}  # one batch size at a time
            ws = self._ws[B]
        return ws


def raw_ff_forward(net, img, act=None):
    """act [B, n_neurons] of every neuron on every image; the layer outputs stay in net.workspace(B) for raw_ff_backward."""
    lib = _lib.load()
    _require_cuda(img)
    if tuple(img.shape[1:]) != net.in_shape:
        raise ValueError(f"expected images [B,{net.in_shape[0]},{net.in_shape[1]},{net.in_shape[2]}], got {tuple(img.shape)}")
    B = img.shape[0]
    ws = net.workspace(B)
    if act is None:
        act = torch.empty((B, net.n_neurons), dtype=torch.float32, device=img.device)
    check(lib.lmr_ff_forward(C.byref(net.net), _ptr(img), B, _ptr(act), _ptr(ws), ws.numel(), _stream()))
    return act


def raw_ff_backward(net, g_act, g_img):
    """g_img [B, C, H, W] = d sum(g_act * act) / d img of the forward that last used net.workspace(B)."""
    lib = _lib.load()
    _require_cuda(g_act, g_img)
    B = g_act.shape[0]
    ws = net.workspace(B)
    check(lib.lmr_ff_backward(C.byref(net.net), _ptr(g_act), B, _ptr(g_img), _ptr(ws), ws.numel(), _stream()))
    return g_img


class _FullFrameResp(torch.autograd.Function):
    """x [B,C,H,W] -> act [B, n_neurons]; gradient wrt x only (the response model is frozen)."""

    @staticmethod
    def forward(ctx, x, net):
        x = _f32c(x)
        ctx.net, ctx.shape = net, x.shape
        net._fwd_serial = getattr(net, "_fwd_serial", 0) + 1
        ctx.serial = net._fwd_serial
        ctx.save_for_backward(x)
        return raw_ff_forward(net, x)

    @staticmethod
    def backward(ctx, g_act):
        net = ctx.net
        (x,) = ctx.saved_tensors
        if not net.need_grad:
            raise RuntimeError("laminr_b200: this FullFrameNet was packed without gradient weights")
        if ctx.serial != net._fwd_serial:  # another forward overwrote the kept layer outputs: run this one again
            raw_ff_forward(net, x)
            net._fwd_serial += 1
        g_img = torch.empty(ctx.shape, dtype=torch.float32, device=g_act.device)
        raw_ff_backward(net, _f32c(g_act), g_img)
        return g_img, None


def fullframe_response(x, net):
    return _FullFrameResp.apply(x, net)


def raw_ascent_update(mode, x, g, step_size, target, mean, pmin, pmax, eps=1e-9, workspace=None):
    """x <- clip(renorm(x + step_size*g)) in place; x, g: [B, ...]."""
    lib = _lib.load()
    _require_cuda(x, g)
    Bn = x.shape[0]
    n = x.numel() // Bn
    if workspace is None:
        workspace = _workspace(lib.lmr_constrain_workspace_bytes(Bn, n), x.device)
    check(lib.lmr_ascent_update(int(mode), _ptr(x), _ptr(g), float(step_size), Bn, n, float(target), float(mean), int(pmin is not None),
                                float(pmin if pmin is not None else 0.0), int(pmax is not None), float(pmax if pmax is not None else 0.0), float(eps),
                                _ptr(workspace), workspace.numel(), _stream()))
    return x


# ------------------------------------------------------------------------------------------------------------------
# masked contrastive term
# ------------------------------------------------------------------------------------------------------------------
def raw_simclr_fwd(img, mask, N, Cc, HW, pmin, pmax, pos, neg, T, eps, reg=None, K=None, workspace=None):
    lib = _lib.load()
    _require_cuda(img, mask, pos, neg)
    dev = img.device
    if reg is None:
        reg = torch.empty((1,), dtype=torch.float32, device=dev)
    if K is None:
        K = torch.empty((N, N), dtype=torch.float32, device=dev)
    if workspace is None:
        workspace = _workspace(lib.lmr_simclr_workspace_bytes(N, Cc, HW), dev)
    check(lib.lmr_simclr_fwd(_ptr(img), _ptr(mask), N, Cc, HW, float(pmin), float(pmax), _ptr(pos), _ptr(neg), float(T), float(eps),
                             _ptr(reg), _ptr(K), _ptr(workspace), workspace.numel(), _stream()))
    return reg, K


def raw_simclr_bwd(img, mask, K, g_reg, scale, N, Cc, HW, pmin, pmax, g_img, accumulate=False):
    lib = _lib.load()
    check(lib.lmr_simclr_bwd(_ptr(img), _ptr(mask), _ptr(K), _ptr(g_reg), float(scale), N, Cc, HW, float(pmin), float(pmax), _ptr(g_img),
                             int(accumulate), _stream()))
    return g_img


def simclr_reg(img, mask, pmin, pmax, pos, neg, T, eps=1e-6):
    """apply_mask + SimCLROnGrid.reg_term fused, differentiable wrt img.  img [N,C,H,W]; mask [H,W]: torch.ops.laminr_b200.simclr_reg
    (backward: simclr_reg_backward)."""
    from . import torch_ops  # noqa: F401
    reg, _K = torch.ops.laminr_b200.simclr_reg(img, _f32c(mask), float(pmin), float(pmax), pos, neg, float(T), float(eps))
    return reg.reshape(())


# ------------------------------------------------------------------------------------------------------------------
# fused learn-stage objective (forward + backward in one launch)
# ------------------------------------------------------------------------------------------------------------------
def learn_objective_workspace(N, Cc, HW, F, device):
    """Zero-initialised workspace of lmr_learn_objective, or None when the configuration does not fit the fused kernel."""
    nbytes = _lib.load().lmr_learn_objective_workspace_bytes(int(N), int(Cc), int(HW), int(F))
    return None if nbytes == 0 else torch.zeros(int(nbytes), dtype=torch.uint8, device=device)


def raw_learn_objective(mode, x, target, mean, pmin, pmax, eps_con, bank, act_eps, mei_act, mask, pos, neg, T, eps_clr, g_reg, z, act, reg, loss, g_x,
                        workspace, stats=None, rmax=None, argmax=None, K=None):
    """x [N,C,H,W] -> z, act [N], reg [1], loss [1], g_x = d loss / d x.  bank [F,HW]: filters of the template neuron."""
    lib = _lib.load()
    _require_cuda(x, bank, mask, pos, neg, g_reg)
    N, Cc = x.shape[0], x.shape[1]
    HW = x[0, 0].numel()
    check(lib.lmr_learn_objective(int(mode), _ptr(x), N, Cc, HW, float(target), float(mean), int(pmin is not None), float(pmin or 0.0),
                                  int(pmax is not None), float(pmax or 0.0), float(eps_con), _ptr(bank), int(bank.shape[0]), float(act_eps),
                                  float(mei_act), _ptr(mask), _ptr(pos), _ptr(neg), float(T), float(eps_clr), _ptr(g_reg), _ptr(z), _ptr(stats),
                                  _ptr(act), _ptr(rmax), _ptr(argmax), _ptr(reg), _ptr(K), _ptr(loss), _ptr(g_x), _ptr(workspace), workspace.numel(),
                                  _stream()))


# ------------------------------------------------------------------------------------------------------------------
# Adam / MEI
# ------------------------------------------------------------------------------------------------------------------
def raw_adam_step(p, g, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """step: int32[2] device tensor {steps taken, scratch (zero)}; incremented by the call."""
    lib = _lib.load()
    _require_cuda(p, g, m, v, step)
    if step.numel() < 2 or step.dtype != torch.int32:
        raise ValueError("raw_adam_step: `step` must be an int32 tensor with 2 elements ([steps taken, zero-initialised scratch])")
    check(lib.lmr_adam_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), float(lr), float(b1), float(b2), float(eps), _ptr(step), _stream()))


class _AdamGroup(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("step", C.c_void_p), ("n", C.c_int)]


def raw_adam_step_multi(groups, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """One launch for up to 8 parameter tensors: `groups` = [(p, g, m, v, step), ...] with the arguments of raw_adam_step (lmr_adam_step_multi:
    same element arithmetic, every group its own moments and step word)."""
    lib = _lib.load()
    if not 1 <= len(groups) <= 8:
        raise ValueError("raw_adam_step_multi: 1..8 groups")
    arr = (_AdamGroup * len(groups))()
    for k, (p, g, m, v, step) in enumerate(groups):
        _require_cuda(p, g, m, v, step)
        if step.numel() < 2 or step.dtype != torch.int32:
            raise ValueError("raw_adam_step_multi: `step` must be an int32 tensor with 2 elements ([steps taken, zero-initialised scratch])")
        arr[k] = _AdamGroup(_ptr(p), _ptr(g), _ptr(m), _ptr(v), _ptr(step), p.numel())
    check(lib.lmr_adam_step_multi(arr, len(groups), float(lr), float(b1), float(b2), float(eps), _stream()))


def raw_match_epoch_stats(act, mei_act, out3):
    """act [D,N], mei_act [D] -> out3 = (sum_d, min_d, max_d) of acts_d = mean_n act[d,n] / mei_act[d] (inv_temp_learn_match.py:361-371)."""
    lib = _lib.load()
    _require_cuda(act, mei_act, out3)
    D, N = act.shape
    check(lib.lmr_match_epoch_stats(_ptr(act), _ptr(mei_act), int(D), int(N), _ptr(out3), _stream()))


def raw_mei_ascent(x, filters, nfilt, neuron_of_group, step_size, iters, mode, target, mean, pmin, pmax, act_eps=1e-6, apply_post_first=True):
    """x [G,C,H,W] in/out -> fevals [G, iters+1]."""
    lib = _lib.load()
    _require_cuda(x, filters)
    G, Cc = x.shape[0], x.shape[1]
    HW = x[0, 0].numel()
    fevals = torch.empty((G, iters + 1), dtype=torch.float32, device=x.device)
    check(lib.lmr_mei_ascent(_ptr(x), G, Cc, HW, _ptr(filters), filters.shape[1], _ptr(nfilt), _ptr(neuron_of_group), float(step_size), int(iters),
                             int(mode), float(target), float(mean), float(pmin), float(pmax), float(act_eps), int(apply_post_first),
                             _ptr(fevals), _stream()))
    return fevals


# --------------------------------------------------------------------------------------------------------------------
# RF masks (create_mask, laminr/utils/mask.py:25-64) on the device
# --------------------------------------------------------------------------------------------------------------------
def gaussian_blur_weights(sigma, truncate=4.0):
    """scipy.ndimage.gaussian_filter1d's kernel (float64): radius = int(truncate*sigma + 0.5), normalised exp(-0.5 (k/sigma)^2).
    Returned outermost tap first, centre last (the layout lmr_create_mask takes)."""
    import numpy as np

    radius = int(truncate * float(sigma) + 0.5)
    k = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * k.astype(np.float64) ** 2)
    phi = phi / phi.sum()
    return phi[: radius + 1].copy(), radius


def raw_create_mask(mei, zscore_thresh, closing_iters, gaussian_sigma, mask_out=None, stats=None, workspace=None):
    """mei: [G,H,W] float32 (channel-averaged MEIs).  Returns (mask [G,H,W] float32, stats int64 [G,4] =
    hull pixel count, row-index sum, column-index sum, empty flag)."""
    lib = _lib.load()
    _require_cuda(mei)
    mei = _f32c(mei)
    G, H, W = mei.shape
    w, radius = gaussian_blur_weights(gaussian_sigma)
    w_dev = torch.from_numpy(w).to(mei.device)
    mask_out = torch.empty_like(mei) if mask_out is None else mask_out
    stats = torch.empty((G, 4), dtype=torch.int32, device=mei.device) if stats is None else stats
    ws = _workspace(lib.lmr_create_mask_workspace_bytes(G, H, W), mei.device) if workspace is None else workspace
    check(lib.lmr_create_mask(_ptr(mei), G, H, W, float(zscore_thresh), int(closing_iters), _ptr(w_dev), radius, _ptr(mask_out), _ptr(stats), _ptr(ws),
                              ws.numel(), _stream()))
    return mask_out, stats


# --------------------------------------------------------------------------------------------------------------------
# Gaussian blur (featurevis.ops.GaussianBlur, featurevis/ops.py:378-423)
# --------------------------------------------------------------------------------------------------------------------
PAD_MODES = {"constant": 0, "reflect": 1, "replicate": 2}


def gaussian_window(sigma, truncate=4):
    """The window featurevis builds with scipy.signal.gaussian(2 h + 1, std=sigma): exp(-0.5 (k / sigma)^2), h = max(round(sigma * truncate), 1)."""
    import numpy as np

    h = max(int(round(sigma * truncate)), 1)
    k = np.arange(2 * h + 1, dtype=np.float64) - h
    return np.exp(-0.5 * (k / float(sigma)) ** 2), h


def raw_gaussian_blur(x, sigma_y, sigma_x, truncate=4, pad_mode="reflect", out=None):
    """x: [..., H, W] float32 on CUDA; every leading dimension is a plane.  Returns the blurred tensor (same shape)."""
    if pad_mode not in PAD_MODES:
        raise ValueError(f"pad_mode {pad_mode!r}: valid values are 'constant', 'reflect' and 'replicate'")
    lib = _lib.load()
    _require_cuda(x)
    x = _f32c(x)
    H, W = x.shape[-2:]
    B = x.numel() // (H * W)
    wy, hy = gaussian_window(sigma_y, truncate)
    wx, hx = gaussian_window(sigma_x, truncate)
    wy_t, wx_t = torch.as_tensor(wy, dtype=torch.float32), torch.as_tensor(wx, dtype=torch.float32)
    inv_norm = 1.0 / float(wy_t.sum() * wx_t.sum())  # float32 sums, like featurevis (ops.py:420)
    out = torch.empty_like(x) if out is None else out
    ws = torch.empty(x.numel(), dtype=torch.float32, device=x.device)
    wy_d, wx_d = wy_t.to(x.device), wx_t.to(x.device)  # named: the device copies must outlive the pointer extraction
    check(lib.lmr_gaussian_blur(_ptr(x), B, H, W, _ptr(wy_d), hy, _ptr(wx_d), hx, PAD_MODES[pad_mode], inv_norm, _ptr(out), _ptr(ws), ws.numel() * 4, _stream()))
    return out
