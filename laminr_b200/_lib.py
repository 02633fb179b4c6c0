"""ctypes binding of the C ABI in include/laminr_b200.h.  There is no fallback: if the library is missing the import fails."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblaminr_b200.so")

PREC_FP32, PREC_BF16X3_FAST, PREC_BF16X3 = 0, 1, 2
PREC_REUSE_COORD_TABLE = 0x200  # flag: the forward reuses the coordinate sin/cos table its previous call left in the workspace (learn stage)
PREC_KEEP_ACTIVATIONS = 0x100  # flag: the forward keeps the hidden activations' operand tiles for the backward (include/laminr_b200.h)
MODE_NORM, MODE_STD = 0, 1


class InrDesc(C.Structure):
    _fields_ = [("hidden", C.c_int), ("n_hidden", C.c_int), ("pe_dim", C.c_int), ("aux_pe_dim", C.c_int), ("channels", C.c_int)]


class Affine(C.Structure):
    _fields_ = [("angles", C.c_void_p), ("scalings", C.c_void_p), ("scal_cols", C.c_int), ("shears", C.c_void_p), ("translation", C.c_void_p),
                ("tc", C.c_void_p), ("shift_to", C.c_void_p), ("shift_from", C.c_void_p), ("clamp", C.c_int)]


CONV_MAX_LAYERS = 6


class ConvcoreDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("c_in", C.c_int), ("hidden", C.c_int), ("kernel", C.c_int * CONV_MAX_LAYERS), ("nonlin", C.c_int * CONV_MAX_LAYERS),
                ("elu_xshift", C.c_float), ("elu_yshift", C.c_float), ("height", C.c_int), ("width", C.c_int), ("align_corners", C.c_int),
                ("elu_offset", C.c_float)]


class ConvcoreWeights(C.Structure):
    _fields_ = [("w_fwd", C.c_void_p * CONV_MAX_LAYERS), ("w_bwd", C.c_void_p * CONV_MAX_LAYERS), ("scale", C.c_void_p * CONV_MAX_LAYERS),
                ("shift", C.c_void_p * CONV_MAX_LAYERS), ("mu", C.c_void_p), ("features", C.c_void_p), ("bias", C.c_void_p)]


FF_MAX_LAYERS = 16
FF_DENSE, FF_DEPTHWISE, FF_POINTWISE, FF_SE = 0, 1, 2, 3


class FFLayer(C.Structure):
    _fields_ = [("kind", C.c_int), ("c_in", C.c_int), ("c_out", C.c_int), ("k", C.c_int), ("dilation", C.c_int), ("pad", C.c_int), ("affine", C.c_int),
                ("elu", C.c_int), ("elu_xshift", C.c_float), ("elu_yshift", C.c_float), ("se_hidden", C.c_int), ("w_fwd", C.c_void_p), ("w_bwd", C.c_void_p),
                ("scale", C.c_void_p), ("shift", C.c_void_p), ("se_w1", C.c_void_p), ("se_b1", C.c_void_p), ("se_w2", C.c_void_p), ("se_b2", C.c_void_p)]


class FFNet(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("c_in", C.c_int), ("height", C.c_int), ("width", C.c_int), ("n_neurons", C.c_int), ("n_scales", C.c_int),
                ("pool_kern", C.c_int), ("align_corners", C.c_int), ("elu_offset", C.c_float), ("grid", C.c_void_p), ("features", C.c_void_p),
                ("bias", C.c_void_p), ("layers", FFLayer * FF_MAX_LAYERS)]


_vp, _i, _f, _sz, _ll = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong

# name -> (restype, argtypes): one entry per symbol declared in include/laminr_b200.h
SIGNATURES = {
    "lmr_abi_version": (_i, []),
    "lmr_last_error": (C.c_char_p, []),
    "lmr_check_device": (_i, []),
    "lmr_inr_params_count": (_sz, [C.POINTER(InrDesc)]),
    "lmr_inr_forward_workspace_bytes": (_sz, [C.POINTER(InrDesc), _i, _i, _i]),
    "lmr_inr_backward_workspace_bytes": (_sz, [C.POINTER(InrDesc), _i, _i, _i]),
    "lmr_inr_forward_workspace_bytes_keep": (_sz, [C.POINTER(InrDesc), _i, _i, _i]),
    "lmr_inr_forward": (_i, [C.POINTER(InrDesc), _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, C.POINTER(Affine), _i, _vp, _vp, _sz, _i, _vp]),
    "lmr_inr_backward": (_i, [C.POINTER(InrDesc), _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, C.POINTER(Affine), _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "lmr_inr_backward_after_forward": (_i, [C.POINTER(InrDesc), _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, C.POINTER(Affine), _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                                            _vp, _sz, _i, _vp]),
    "lmr_inr_backward_adam": (_i, [C.POINTER(InrDesc), _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _vp, _sz, _vp, _sz, _i, _vp]),
    "lmr_constrain_workspace_bytes": (_sz, [_i, _i]),
    "lmr_constrain_fwd": (_i, [_i, _vp, _i, _i, _f, _f, _i, _f, _i, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "lmr_constrain_bwd": (_i, [_i, _vp, _vp, _vp, _i, _i, _f, _f, _i, _f, _i, _f, _f, _vp, _vp, _sz, _vp]),
    "lmr_ascent_update": (_i, [_i, _vp, _vp, _f, _i, _i, _f, _f, _i, _f, _i, _f, _f, _vp, _sz, _vp]),
    "lmr_convcore_response": (_i, [C.POINTER(ConvcoreDesc), C.POINTER(ConvcoreWeights), _vp, _ll, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "lmr_ff_workspace_bytes": (_sz, [C.POINTER(FFNet), _i]),
    "lmr_ff_forward": (_i, [C.POINTER(FFNet), _vp, _i, _vp, _vp, _sz, _vp]),
    "lmr_ff_backward": (_i, [C.POINTER(FFNet), _vp, _i, _vp, _vp, _sz, _vp]),
    "lmr_simresp_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "lmr_simresp_fwd": (_i, [_vp, _ll, _i, _i, _i, _i, _vp, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "lmr_simresp_bwd": (_i, [_vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp, _i, _vp, _vp, _i, _vp]),
    "lmr_simclr_workspace_bytes": (_sz, [_i, _i, _i]),
    "lmr_simclr_fwd": (_i, [_vp, _vp, _i, _i, _i, _f, _f, _vp, _vp, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "lmr_simclr_bwd": (_i, [_vp, _vp, _vp, _vp, _f, _i, _i, _i, _f, _f, _vp, _i, _vp]),
    "lmr_learn_objective_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "lmr_learn_objective": (_i, [_i, _vp, _i, _i, _i, _f, _f, _i, _f, _i, _f, _f, _vp, _i, _f, _f, _vp, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                 _vp, _vp, _vp, _vp, _sz, _vp]),
    "lmr_adam_step": (_i, [_vp, _vp, _vp, _vp, _i, _f, _f, _f, _f, _vp, _vp]),
    "lmr_adam_step_multi": (_i, [_vp, _i, _f, _f, _f, _f, _vp]),
    "lmr_match_epoch_stats": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "lmr_create_mask_workspace_bytes": (_sz, [_i, _i, _i]),
    "lmr_create_mask": (_i, [_vp, _i, _i, _i, _f, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "lmr_gaussian_blur": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _i, _f, _vp, _vp, _sz, _vp]),
    "lmr_mei_ascent": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _f, _i, _i, _f, _f, _f, _f, _f, _i, _vp, _vp]),
}

_lib = None


def load():
    """Loads liblaminr_b200.so (built by laminr_b200/csrc/build.py).  Raises if it is missing: no CPU / eager fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m laminr_b200.csrc.build` (nvcc, sm_100a). "
                "laminr_b200 has no CPU or eager-PyTorch fallback for the fused path."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class LaminrB200Error(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise LaminrB200Error(f"laminr_b200 C-ABI call failed ({rc}): {load().lmr_last_error().decode()}")
