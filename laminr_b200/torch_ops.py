"""PyTorch custom operators over the C ABI: `torch.ops.laminr_b200.*` (torch.library).

The differentiable building blocks of the generic path -- INR images, constraint projection, simulated response, masked contrastive
term -- are registered with the dispatcher (CUDA kernels only: every implementation is a call into include/laminr_b200.h through
laminr_b200.ops.raw_*; fake kernels give shapes for tracing; `register_autograd` wires each forward to its backward operator).  The functions
of laminr_b200.ops that the drop-in modules call (`inr_images`, `constrain`, `simresp`, `simclr_reg`) are thin wrappers around these operators.
The fused steppers (laminr_b200/fused.py) call the same C ABI directly on static buffers inside CUDA graphs and do not go through the dispatcher.

    img           = torch.ops.laminr_b200.inr_forward(flat, B, auxB, grid, coords, coord_shift, angles, scalings, shears, translation,
                                                      tc, shift_to, shift_from, clamp, hidden, n_hidden, pe_dim, aux_pe_dim, channels, precision)
    g_flat, g_ang, g_scal, g_shear, g_trans = torch.ops.laminr_b200.inr_backward(..., g_img, want_params, want_affine)
    z, stats      = torch.ops.laminr_b200.constrain(x, mode, target, mean, has_min, pmin, has_max, pmax, eps)
    act, rmax, am = torch.ops.laminr_b200.simresp(x, filters, nfilt, neuron_of_group, eps)
    reg, K        = torch.ops.laminr_b200.simclr_reg(img, mask, has_bounds, pmin, pmax, pos, neg, T, eps)
"""
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor
from torch.library import custom_op, register_autograd

from . import ops as _o

_NS = "laminr_b200"


def _affine(angles, scalings, shears, translation, tc, shift_to, shift_from, clamp):
    if angles is None:
        return None
    return _o.AffineArgs(angles, scalings, shears, translation, tc=tc, shift_to=shift_to, shift_from=shift_from, clamp=clamp)


# ---------------------------------------------------------------------------------------------------------------------------------------
# INR generator
# ---------------------------------------------------------------------------------------------------------------------------------------
@custom_op(f"{_NS}::inr_forward", mutates_args=(), device_types="cuda")
def inr_forward(flat: Tensor, B: Tensor, auxB: Tensor, grid: Tensor, coords: Tensor, coord_shift: Optional[Tensor], angles: Optional[Tensor],
                scalings: Optional[Tensor], shears: Optional[Tensor], translation: Optional[Tensor], tc: Optional[Tensor], shift_to: Optional[Tensor],
                shift_from: Optional[Tensor], clamp: bool, hidden: int, n_hidden: int, pe_dim: int, aux_pe_dim: int, channels: int,
                precision: str) -> Tensor:
    desc = _o.make_desc(hidden, n_hidden, pe_dim, aux_pe_dim, channels)
    aff = _affine(angles, scalings, shears, translation, tc, shift_to, shift_from, clamp)
    return _o.raw_inr_forward(desc, flat, B, auxB, _o._f32c(grid.reshape(-1)), coords, coord_shift, aff, precision=precision or None)


@inr_forward.register_fake
def _(flat, B, auxB, grid, coords, coord_shift, angles, scalings, shears, translation, tc, shift_to, shift_from, clamp, hidden, n_hidden, pe_dim,
      aux_pe_dim, channels, precision):
    D = 1 if angles is None else angles.numel()
    return flat.new_empty((D, grid.numel(), channels, coords.shape[0]))


@custom_op(f"{_NS}::inr_backward", mutates_args=(), device_types="cuda")
def inr_backward(flat: Tensor, B: Tensor, auxB: Tensor, grid: Tensor, coords: Tensor, coord_shift: Optional[Tensor], angles: Optional[Tensor],
                 scalings: Optional[Tensor], shears: Optional[Tensor], translation: Optional[Tensor], tc: Optional[Tensor], shift_to: Optional[Tensor],
                 shift_from: Optional[Tensor], clamp: bool, hidden: int, n_hidden: int, pe_dim: int, aux_pe_dim: int, channels: int, precision: str,
                 g_img: Tensor, want_params: bool, want_affine: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """-> (g_flat, g_angles, g_scalings, g_shears, g_translation); the ones not asked for are empty tensors."""
    desc = _o.make_desc(hidden, n_hidden, pe_dim, aux_pe_dim, channels)
    aff = _affine(angles, scalings, shears, translation, tc, shift_to, shift_from, clamp)
    g_flat, g_aff = _o.raw_inr_backward(desc, flat, B, auxB, _o._f32c(grid.reshape(-1)), coords, coord_shift, aff, g_img, want_params=want_params,
                                        want_affine=want_affine and aff is not None, precision=precision or None)
    e = lambda: flat.new_empty(0)
    ga = g_aff if g_aff is not None else (e(), e(), e(), e())
    return (g_flat if g_flat is not None else e()), ga[0], ga[1], ga[2], ga[3]


@inr_backward.register_fake
def _(flat, B, auxB, grid, coords, coord_shift, angles, scalings, shears, translation, tc, shift_to, shift_from, clamp, hidden, n_hidden, pe_dim,
      aux_pe_dim, channels, precision, g_img, want_params, want_affine):
    e = lambda: flat.new_empty(0)
    aff = want_affine and angles is not None
    return (torch.empty_like(flat) if want_params else e(), torch.empty_like(angles) if aff else e(), torch.empty_like(scalings) if aff else e(),
            torch.empty_like(shears) if aff else e(), torch.empty_like(translation) if aff else e())


@custom_op(f"{_NS}::inr_images", mutates_args=(), device_types="cuda")
def inr_images_op(flat: Tensor, param_views: Sequence[Tensor], B: Tensor, auxB: Tensor, grid: Tensor, coords: Tensor, coord_shift: Optional[Tensor],
                  angles: Optional[Tensor], scalings: Optional[Tensor], shears: Optional[Tensor], translation: Optional[Tensor], tc: Optional[Tensor],
                  shift_to: Optional[Tensor], shift_from: Optional[Tensor], clamp: bool, hidden: int, n_hidden: int, pe_dim: int, aux_pe_dim: int,
                  channels: int, precision: str) -> Tensor:
    """inr_forward whose gradient reaches the network's parameters: `param_views` are the nn.Parameters of INRTemplates.func (views into `flat`,
    in flat order); the forward reads `flat`, the backward hands every view its slice of the flat gradient."""
    desc = _o.make_desc(hidden, n_hidden, pe_dim, aux_pe_dim, channels)
    det = lambda t: None if t is None else t.detach()
    aff = _affine(det(angles), det(scalings), det(shears), det(translation), tc, shift_to, shift_from, clamp)
    return _o.raw_inr_forward(desc, flat, B, auxB, _o._f32c(grid.detach().reshape(-1)), coords, coord_shift, aff, precision=precision or None)


@inr_images_op.register_fake
def _(flat, param_views, B, auxB, grid, coords, coord_shift, angles, scalings, shears, translation, tc, shift_to, shift_from, clamp, hidden, n_hidden,
      pe_dim, aux_pe_dim, channels, precision):
    D = 1 if angles is None else angles.numel()
    return flat.new_empty((D, grid.numel(), channels, coords.shape[0]))


def _inr_setup(ctx, inputs, output):
    (flat, param_views, B, auxB, grid, coords, coord_shift, angles, scalings, shears, translation, tc, shift_to, shift_from, clamp, hidden, n_hidden,
     pe_dim, aux_pe_dim, channels, precision) = inputs
    ctx.save_for_backward(flat, B, auxB, grid, coords, coord_shift, angles, scalings, shears, translation, tc, shift_to, shift_from)
    ctx.cfg = (clamp, hidden, n_hidden, pe_dim, aux_pe_dim, channels, precision)
    # every view is a strided window of `flat` (contiguous for 50-wide layers, the corner of a zero-padded block for narrower ones): its
    # gradient is the same window of the flat gradient
    base = flat.storage_offset()
    ctx.view_geom = [(tuple(v.shape), tuple(v.stride()), v.storage_offset() - base) if v.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
                     else None for v in param_views]
    ctx.view_shapes = [tuple(v.shape) for v in param_views]
    ctx.want_params = any(v.requires_grad for v in param_views)
    ctx.want_affine = angles is not None and any(t is not None and t.requires_grad for t in (angles, scalings, shears, translation))
    ctx.translation_shape = None if translation is None else tuple(translation.shape)


def _inr_backward(ctx, g_img):
    flat, B, auxB, grid, coords, coord_shift, angles, scalings, shears, translation, tc, shift_to, shift_from = ctx.saved_tensors
    clamp, hidden, n_hidden, pe_dim, aux_pe_dim, channels, precision = ctx.cfg
    out = [None] * 21
    if not (ctx.want_params or ctx.want_affine):
        return tuple(out)
    det = lambda t: None if t is None else t.detach()
    g_flat, g_ang, g_scal, g_shear, g_trans = torch.ops.laminr_b200.inr_backward(
        flat.detach(), B, auxB, grid.detach(), coords, coord_shift, det(angles), det(scalings), det(shears), det(translation), tc, shift_to, shift_from, clamp,
        hidden, n_hidden, pe_dim, aux_pe_dim, channels, precision, g_img.contiguous(), ctx.want_params, ctx.want_affine)
    if ctx.want_params:
        grads, off = [], 0
        for shp, geom in zip(ctx.view_shapes, ctx.view_geom):
            n = 1
            for s_ in shp:
                n *= s_
            grads.append(g_flat.as_strided(*geom) if geom is not None else g_flat[off:off + n].view(shp))
            off += n
        out[1] = grads
    if ctx.want_affine:
        out[7], out[8], out[9], out[10] = g_ang, g_scal, g_shear, g_trans.reshape(ctx.translation_shape)
    return tuple(out)


register_autograd(f"{_NS}::inr_images", _inr_backward, setup_context=_inr_setup)


# ---------------------------------------------------------------------------------------------------------------------------------------
# constraint projection
# ---------------------------------------------------------------------------------------------------------------------------------------
@custom_op(f"{_NS}::constrain", mutates_args=(), device_types="cuda")
def constrain_op(x: Tensor, mode: int, target: float, mean: float, has_min: bool, pmin: float, has_max: bool, pmax: float, eps: float) -> Tuple[Tensor, Tensor]:
    z, stats = _o.raw_constrain_fwd(mode, _o._f32c(x), target, mean, pmin if has_min else None, pmax if has_max else None, eps)
    return z, stats


@constrain_op.register_fake
def _(x, mode, target, mean, has_min, pmin, has_max, pmax, eps):
    return torch.empty_like(x), x.new_empty((x.shape[0], 2))


@custom_op(f"{_NS}::constrain_backward", mutates_args=(), device_types="cuda")
def constrain_backward_op(x: Tensor, g_z: Tensor, stats: Tensor, mode: int, target: float, mean: float, has_min: bool, pmin: float, has_max: bool, pmax: float,
                          eps: float) -> Tensor:
    return _o.raw_constrain_bwd(mode, _o._f32c(x), _o._f32c(g_z), stats, target, mean, pmin if has_min else None, pmax if has_max else None, eps)


@constrain_backward_op.register_fake
def _(x, g_z, stats, mode, target, mean, has_min, pmin, has_max, pmax, eps):
    return torch.empty_like(x)


def _constrain_setup(ctx, inputs, output):
    x, mode, target, mean, has_min, pmin, has_max, pmax, eps = inputs
    ctx.save_for_backward(x, output[1])
    ctx.cfg = (mode, target, mean, has_min, pmin, has_max, pmax, eps)


def _constrain_backward(ctx, g_z, g_stats):
    x, stats = ctx.saved_tensors
    return (torch.ops.laminr_b200.constrain_backward(x.detach(), g_z.contiguous(), stats, *ctx.cfg),) + (None,) * 8


register_autograd(f"{_NS}::constrain", _constrain_backward, setup_context=_constrain_setup)


# ---------------------------------------------------------------------------------------------------------------------------------------
# simulated response
# ---------------------------------------------------------------------------------------------------------------------------------------
@custom_op(f"{_NS}::simresp", mutates_args=(), device_types="cuda")
def simresp_op(x: Tensor, filters: Tensor, nfilt: Tensor, neuron_of_group: Tensor, eps: float) -> Tuple[Tensor, Tensor, Tensor]:
    """x [B,C,H,W] -> (act, rmax, argmax), each [G,B], for the neurons in `neuron_of_group` (every neuron sees every image)."""
    x = _o._f32c(x)
    return _o.raw_simresp_fwd(x, 0, x.shape[0], neuron_of_group.numel(), x.shape[1], x.shape[2] * x.shape[3], filters, nfilt, neuron_of_group, eps)


@simresp_op.register_fake
def _(x, filters, nfilt, neuron_of_group, eps):
    shape = (neuron_of_group.numel(), x.shape[0])
    return x.new_empty(shape), x.new_empty(shape), x.new_empty(shape, dtype=torch.int32)


@custom_op(f"{_NS}::simresp_backward", mutates_args=(), device_types="cuda")
def simresp_backward_op(g_act: Tensor, rmax: Tensor, argmax: Tensor, filters: Tensor, neuron_of_group: Tensor, batch: int, channels: int, height: int,
                        width: int) -> Tensor:
    g_img = torch.empty((batch, channels, height, width), dtype=torch.float32, device=g_act.device)
    _o.raw_simresp_bwd(_o._f32c(g_act), rmax, argmax, 0, batch, neuron_of_group.numel(), channels, height * width, filters, neuron_of_group, g_img)
    return g_img


@simresp_backward_op.register_fake
def _(g_act, rmax, argmax, filters, neuron_of_group, batch, channels, height, width):
    return g_act.new_empty((batch, channels, height, width))


def _simresp_setup(ctx, inputs, output):
    x, filters, nfilt, neuron_of_group, eps = inputs
    ctx.save_for_backward(output[1], output[2], filters, neuron_of_group)
    ctx.shape = tuple(x.shape)


def _simresp_backward(ctx, g_act, g_rmax, g_argmax):
    rmax, argmax, filters, neuron_of_group = ctx.saved_tensors
    return torch.ops.laminr_b200.simresp_backward(g_act.contiguous(), rmax, argmax, filters, neuron_of_group, *ctx.shape), None, None, None, None


register_autograd(f"{_NS}::simresp", _simresp_backward, setup_context=_simresp_setup)


# ---------------------------------------------------------------------------------------------------------------------------------------
# masked contrastive term
# ---------------------------------------------------------------------------------------------------------------------------------------
@custom_op(f"{_NS}::simclr_reg", mutates_args=(), device_types="cuda")
def simclr_reg_op(img: Tensor, mask: Tensor, pmin: float, pmax: float, pos: Tensor, neg: Tensor, T: float, eps: float) -> Tuple[Tensor, Tensor]:
    """img [N,C,H,W], mask [H,W] -> (reg [1], K [N,N] the backward's coefficient matrix)."""
    img = _o._f32c(img)
    N, Cc, HW = img.shape[0], img.shape[1], img.shape[2] * img.shape[3]
    return _o.raw_simclr_fwd(img, _o._f32c(mask), N, Cc, HW, pmin, pmax, pos, neg, T, eps)


@simclr_reg_op.register_fake
def _(img, mask, pmin, pmax, pos, neg, T, eps):
    return img.new_empty(1), img.new_empty((img.shape[0], img.shape[0]))


@custom_op(f"{_NS}::simclr_reg_backward", mutates_args=(), device_types="cuda")
def simclr_reg_backward_op(img: Tensor, mask: Tensor, K: Tensor, g_reg: Tensor, pmin: float, pmax: float) -> Tensor:
    img = _o._f32c(img)
    N, Cc, HW = img.shape[0], img.shape[1], img.shape[2] * img.shape[3]
    g_img = torch.empty_like(img)
    _o.raw_simclr_bwd(img, _o._f32c(mask), K, _o._f32c(g_reg.reshape(1)), 1.0, N, Cc, HW, pmin, pmax, g_img)
    return g_img


@simclr_reg_backward_op.register_fake
def _(img, mask, K, g_reg, pmin, pmax):
    return torch.empty_like(img)


def _simclr_setup(ctx, inputs, output):
    img, mask, pmin, pmax, pos, neg, T, eps = inputs
    ctx.save_for_backward(img, mask, output[1])
    ctx.cfg = (pmin, pmax)


def _simclr_backward(ctx, g_reg, g_K):
    img, mask, K = ctx.saved_tensors
    return (torch.ops.laminr_b200.simclr_reg_backward(img.detach(), mask, K, g_reg.contiguous(), *ctx.cfg),) + (None,) * 7


register_autograd(f"{_NS}::simclr_reg", _simclr_backward, setup_context=_simclr_setup)

OPERATORS = ("inr_forward", "inr_backward", "inr_images", "constrain", "constrain_backward", "simresp", "simresp_backward", "simclr_reg",
             "simclr_reg_backward")
