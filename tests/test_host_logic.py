"""CPU: host-side logic of the product (no kernels are launched): library exports, RNG-parity of constructors and the
jittered-grid stream, filter banks, control logic, error behaviour."""
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest
import torch

from tests._util import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from laminr_b200.csrc import build
    build.build()
    import laminr_b200
    return laminr_b200


def test_library_exports_every_header_symbol(L):
    from laminr_b200 import _lib
    header = open(os.path.join(ROOT, "include", "laminr_b200.h")).read()
    declared = set(re.findall(r"\b(lmr_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 19
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(_lib.SIGNATURES) == declared
    assert lib.lmr_abi_version() == 1


def test_no_cpu_fallback(L):
    """ops refuse CPU tensors instead of silently computing something else."""
    from laminr_b200 import ops
    x = torch.zeros(2, 1, 4, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.raw_constrain_fwd(0, x, 1.0, 0.0, -1.0, 1.0, 1e-12)
    m = L.neuron_models.simulated("demo1", img_res=[8, 8])
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 1, 8, 8))


def test_demo1_banks_bit_identical_to_reference(L):
    d = golden("banks")
    for res in (100, 200):
        m = L.neuron_models.simulated("demo1", img_res=[res, res])
        for i, sub in enumerate(m.neuron_models_list):
            f = sub.filters.numpy()
            assert tuple(d[f"shape_{res}_{i}"]) == f.shape
            assert hashlib.sha256(f.tobytes()).digest() == d[f"sha_{res}_{i}"].tobytes(), (res, i)
    assert m.packed_filters.shape == (3, 60, 200 * 200) and m.nfilt.tolist() == [20, 20, 60]
    # ragged banks are padded by repeating a filter
    assert torch.equal(m.packed_filters[0, 20:], m.packed_filters[0, :1].expand(40, -1))
    pop = L.neuron_models.simulated_population(4, img_res=[100, 100], seed=0)
    got = np.stack([s.filters.numpy() for s in pop.neuron_models_list])
    assert hashlib.sha256(got.tobytes()).digest() == d["pop4_sha"].tobytes()
    with pytest.raises(ValueError):
        L.neuron_models.simulated("nope")


def _meis(res):
    return {i: dict(mei=np.zeros((1, res, res), np.float32), activation=1.0, mask=np.ones((res, res), np.float32), center_pos=dict(x=res / 2, y=res / 2))
            for i in range(3)}


def test_constructor_consumes_rng_like_the_reference(L):
    """torch.manual_seed(0); InvarianceManifold(...) must give the reference's initial network (inr.py:49-51 draw order)."""
    d = golden("learn_traj_20")
    m = L.neuron_models.simulated("demo1", img_res=[20, 20])
    torch.manual_seed(0)
    im = L.InvarianceManifold(m, _meis(20), device="cpu", required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    sd = im.template.state_dict()
    for l in range(5):
        np.testing.assert_array_equal(sd[f"func.layer{l}.weight"].numpy(), d[f"init_W{l}"])
        np.testing.assert_array_equal(sd[f"func.layer{l}.bias"].numpy(), d[f"init_b{l}"])
    np.testing.assert_array_equal(sd["B"].numpy(), d["init_B"])
    np.testing.assert_array_equal(sd["auxB"].numpy(), d["init_auxB"])
    np.testing.assert_array_equal(sd["coords"].numpy().reshape(-1, 2), d["init_coords"])
    assert sum(p.numel() for p in im.template.parameters()) == 17801 and im.template.in_dim == 201
    # the parameters are views of one flat vector, in func.parameters() order
    flat = im.template.flat_params()
    assert flat.numel() == 17801
    np.testing.assert_array_equal(flat[:50 * 201].numpy().reshape(50, 201), d["init_W0"])
    flat[0] = 123.0
    assert im.template.func.layer0.weight[0, 0].item() == 123.0
    with pytest.raises(ValueError):
        L.InvarianceManifold(m, _meis(20), device="cpu", pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)
    with pytest.raises(ValueError):
        L.InvarianceManifold(m, _meis(20), device="cpu", required_img_norm=1.0, required_img_std=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)


def test_jittered_grid_stream_matches_reference(L):
    from laminr_b200.utils.training import JitteringGridDatamodule
    d = golden("training_utils")
    for n in (20, 7, 32):
        np.testing.assert_array_equal(JitteringGridDatamodule(1, n, 3).grid[:, 0].numpy(), d[f"grid_{n}"])
    torch.manual_seed(4)
    dm = JitteringGridDatamodule(1, 20, 5)
    batches = torch.stack(list(dm.train_dataloader()))
    np.testing.assert_array_equal(batches[:, :, 0].numpy(), d["jitter_batches"])
    t = golden("learn_traj_20")
    torch.manual_seed(123)
    dm = JitteringGridDatamodule(1, 20, 30)
    np.testing.assert_array_equal(torch.stack(list(dm.train_dataloader()))[:, :, 0].numpy(), t["grids"])


def test_iterating_consumes_the_dataloader_seed_draw(L):
    """DataLoader.__iter__ draws one int64 from the CPU generator; our iterable must leave the generator in the same state."""
    from torch.utils.data import DataLoader
    from laminr_b200.utils.training import JitteringGridDatamodule
    torch.manual_seed(7)
    dm = JitteringGridDatamodule(1, 20, 4)
    for _ in dm.train_dataloader():
        pass
    ours = torch.rand(3)
    torch.manual_seed(7)
    jitter = torch.rand([4, 1]) * dm.sigma - 0.5 * dm.sigma
    grids = dm.base_grid + jitter.repeat_interleave(20, dim=0)
    for _ in DataLoader(grids, batch_size=20):
        pass
    assert torch.equal(ours, torch.rand(3))


def test_control_logic_traces(L):
    from laminr_b200.utils.training import ImprovementChecker, ParamReduceOnPlateau, check_activity_requirements
    d = golden("training_utils")
    sch = ParamReduceOnPlateau(factor=0.8, patience=5, threshold=0.005, mode="max")
    val, vals = 2.0, []
    for s in d["plateau_seq"]:
        val = sch.step(torch.tensor(s), val)
        vals.append(val)
    np.testing.assert_array_equal(np.array(vals), d["plateau_vals"])
    chk = ImprovementChecker(patience=4, ignore_diff_smaller_than=1e-3)
    np.testing.assert_array_equal(np.array([chk.is_increasing(float(s)) for s in d["plateau_seq"]]), d["improve_flags"])
    req = dict(avg=0.99, std=1.0, necessary_min=0.98)
    np.testing.assert_array_equal(np.array([check_activity_requirements(torch.tensor(a), req) for a in d["req_acts"]]), d["req_pass"])


def test_neighbour_masks(L):
    from laminr_b200.contrastive_loss import SimCLROnGrid
    d = golden("simclr")
    for tag, n, periodic, nb in (("p20", 20, True, 0.1), ("np7", 7, False, 0.3), ("p10", 10, True, 0.1)):
        reg = SimCLROnGrid(1, n, nb, 0.3, periodic)
        np.testing.assert_array_equal(reg.flat_neighbor_mask_pos.numpy(), d[f"{tag}_pos"])
        np.testing.assert_array_equal(reg.flat_neighbor_mask_neg.numpy(), d[f"{tag}_neg"])


def test_affine_parameter_names_and_init(L):
    from laminr_b200.inr import INRTemplates
    from torch import nn
    t = INRTemplates((8, 8), widths=[50] * 4, periodic_invariance=True, positional_encoding_dim=50, positional_encoding_projection_scale=10,
                     aux_positional_encoding_dim=50, aux_positional_encoding_projection_scale=0.1, final_nonlinearity=nn.Tanh, bias=True, weights_scale=0.1)
    t.initialize_coordinate_transform(3, True, False, 0.1, True, True, False, True)
    keys = set(t.state_dict())
    for k in ("angles", "scalings", "shears", "translation"):
        assert f"coordinate_transform.transforms.Affine.{k}" in keys
    aff = t.coordinate_transform.transforms.Affine
    assert aff.translation.shape == (3, 1, 2) and torch.all(aff.scalings == 1) and torch.all(aff.angles == 0)
    assert torch.allclose(aff.As, torch.eye(2).expand(3, 2, 2))
    t.register_coords_shifts(torch.tensor([5.0, 3.0]), torch.tensor([[2.0, 6.0], [4.0, 4.0], [1.0, 1.0]]))
    assert {"coords_shift_to_center", "template_center_in_coords_space", "coords_shift_from_center"} <= set(t.state_dict())
    with pytest.raises(NotImplementedError):
        t.initialize_coordinate_transform(3, False, False, 0.1, True, True, False, True)  # RealNVP warp: out of scope


def test_mei_argument_errors(L):
    m = L.neuron_models.simulated("demo1", img_res=[8, 8])
    with pytest.raises(ValueError, match="Either std or norm"):
        L.get_mei_dict(m, [1, 8, 8], -1.0, 1.0, neuron_idxs=[0])
    with pytest.raises(ValueError, match="Only one"):
        L.get_mei_dict(m, [1, 8, 8], -1.0, 1.0, neuron_idxs=[0], required_img_std=1.0, required_img_norm=1.0)


@pytest.mark.gpu
def test_create_mask_centroid(L):
    from laminr_b200.utils.mask import create_mask
    yy, xx = np.meshgrid(np.arange(40), np.arange(40), indexing="ij")
    mei = (np.exp(-0.5 * (((xx - 25) / 4.0) ** 2 + ((yy - 12) / 4.0) ** 2)) * np.cos(0.9 * (xx - 25))).astype(np.float32)  # Gabor-like, zero-mean
    mask, c = create_mask(torch.from_numpy(mei), zscore_thresh=0.3, return_mask_centroid=True)
    assert mask.shape == (40, 40) and abs(c["x"] - 25.5) < 1.0 and abs(c["y"] - 12.5) < 1.0
    assert float(mask.max()) <= 1.0 + 1e-6 and float(mask[12, 25]) > 0.95 and float(mask[35, 3]) < 1e-3


def test_match_block_selection(L, monkeypatch):
    """Block size of the match training step (laminr_b200.fused.match_block): pure host arithmetic on the C ABI's workspace sizes."""
    from laminr_b200 import ops
    from laminr_b200.fused import match_block
    d = ops.make_desc(50, 4, 50, 50, 1)
    N, P = 20, 100 * 100
    for k in ("LMR_KEEP_ACT", "LMR_KEEP_ACT_MAX_GIB", "LMR_MATCH_BLOCK"):
        monkeypatch.delenv(k, raising=False)
    # the kept record of a 128-row tile (6 pixels x 20 latents): 3 layers x 32 KB of operand tiles + 2 KB of output values y
    # (include/laminr_b200.h, LMR_PREC_KEEP_ACTIVATIONS)
    per_target = ops.forward_workspace_bytes(d, N, P, 1, True) - ops.forward_workspace_bytes(d, N, P, 1, False)
    assert per_target == -(-P // 6) * (3 * 32768 + 2048)
    assert match_block("bf16x3", d, N, P, 32) == (32, True)          # 4.9 GiB of tiles: one pass
    blk, keep = match_block("bf16x3", d, N, P, 1024)                  # BASELINE config 3 on one GPU: 157 GiB of tiles > 32 GiB budget
    assert keep and 190 <= blk <= 210
    assert ops.forward_workspace_bytes(d, N, P, blk, True) <= 32 * 2 ** 30 < ops.forward_workspace_bytes(d, N, P, blk + 1, True)
    assert match_block("fp32", d, N, P, 1024) == (1024, False)        # CUDA-core path keeps nothing
    monkeypatch.setenv("LMR_KEEP_ACT_MAX_GIB", "0.1")                 # not even one target fits: recompute backward, one pass
    assert match_block("bf16x3", d, N, P, 8) == (8, False)
    monkeypatch.setenv("LMR_KEEP_ACT_MAX_GIB", "2")
    assert match_block("bf16x3", d, N, P, 32) == (12, True)
    monkeypatch.setenv("LMR_KEEP_ACT", "0")
    assert match_block("bf16x3", d, N, P, 32) == (32, False)
    monkeypatch.setenv("LMR_MATCH_BLOCK", "5")
    assert match_block("bf16x3", d, N, P, 32)[0] == 5


def test_simresp_workspace_follows_the_group_count(L):
    """The response partials' pixels-per-CTA depend on the number of groups of a call (about two CTAs per SM for few groups), so the workspace is
    NOT monotonic in the group count: callers size it for every group count they use (fused._SimulatedResponse)."""
    from laminr_b200 import _lib
    lib = _lib.load()
    NB, F, HW = 20, 20, 100 * 100
    chunks = {g: lib.lmr_simresp_workspace_bytes(g, NB, F, HW) // (g * NB * F * 4) for g in range(1, 41)}
    assert set(chunks.values()) == {20, 40, 79}                       # 512 / 256 / 128 pixels per CTA at 100 x 100
    assert chunks[1] == 79 and chunks[8] == 40 and chunks[32] == 20
    assert lib.lmr_simresp_workspace_bytes(7, NB, F, HW) > lib.lmr_simresp_workspace_bytes(8, NB, F, HW)


def _header_prototypes():
    """name -> (return type, [parameter type strings]) parsed from include/laminr_b200.h (comments stripped)."""
    src = open(os.path.join(ROOT, "include", "laminr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(lmr_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, params = m.group(1).replace("LMR_API", "").strip(), m.group(2), " ".join(m.group(3).split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        out[name] = (ret, plist)
    return out


def _ctype_of(decl):
    """C parameter declaration -> the ctypes type the binding must use."""
    from laminr_b200 import _lib
    if "*" in decl:
        for struct, ct in (("lmr_inr_desc", _lib.InrDesc), ("lmr_affine", _lib.Affine), ("lmr_convcore_desc", _lib.ConvcoreDesc),
                           ("lmr_convcore_weights", _lib.ConvcoreWeights), ("lmr_ff_net", _lib.FFNet)):
            if re.search(rf"\b{struct}\b", decl):
                return ctypes.POINTER(ct)
        return ctypes.c_char_p if re.match(r"const char\s*\*$", decl) else ctypes.c_void_p
    base = re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*$", "", decl).strip() or decl  # drop the parameter name
    base = base.replace("const ", "").strip()
    return {"int": ctypes.c_int, "float": ctypes.c_float, "size_t": ctypes.c_size_t, "long long": ctypes.c_longlong}[base]


def test_ctypes_signatures_match_the_header_prototypes(L):
    """Every prototype of include/laminr_b200.h against the ctypes binding: same parameter count and, position by position, the same C type class
    (an int / float / size_t / pointer mix-up would corrupt the call frame silently)."""
    from laminr_b200 import _lib
    protos = _header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    for name, (ret, params) in protos.items():
        res, args = _lib.SIGNATURES[name]
        assert len(args) == len(params), (name, len(args), len(params))
        for i, (decl, got) in enumerate(zip(params, args)):
            assert _ctype_of(decl) is got, (name, i, decl, got)
        assert _ctype_of(ret + " x" if "*" not in ret else ret) is res, (name, ret, res)


def test_host_only_abi_functions(L):
    """The entry points that never touch the device: parameter count, workspace sizes (monotone in the extents; the kept-activation forward
    needs room for three 32 KB operand tiles per 128 rows on top), and argument errors reported through lmr_last_error."""
    from laminr_b200 import _lib
    lib = _lib.load()
    d1, d3 = _lib.InrDesc(50, 4, 50, 50, 1), _lib.InrDesc(50, 4, 50, 50, 3)
    assert lib.lmr_inr_params_count(ctypes.byref(d1)) == 17801
    assert lib.lmr_inr_params_count(ctypes.byref(d3)) == 17801 + 2 * 51
    for fn in (lib.lmr_inr_forward_workspace_bytes, lib.lmr_inr_backward_workspace_bytes, lib.lmr_inr_forward_workspace_bytes_keep):
        base = fn(ctypes.byref(d1), 20, 10000, 1)
        assert base > 0
        assert fn(ctypes.byref(d1), 20, 40000, 1) >= base and fn(ctypes.byref(d1), 40, 10000, 1) >= base
        # D == 1 additionally holds the sin/cos(coords.B) table of the layer-0 weight gradient (learn); from D = 2 on the size grows with D
        assert fn(ctypes.byref(d1), 20, 10000, 4) >= fn(ctypes.byref(d1), 20, 10000, 2) > 0
    plain, keep = lib.lmr_inr_forward_workspace_bytes(ctypes.byref(d1), 20, 10000, 1), lib.lmr_inr_forward_workspace_bytes_keep(ctypes.byref(d1), 20, 10000, 1)
    rows = 20 * 10000
    assert keep - plain >= (rows // 128) * 3 * 32768
    assert lib.lmr_constrain_workspace_bytes(20, 1) > 0 and lib.lmr_simclr_workspace_bytes(20, 1, 10000) > 0
    assert lib.lmr_learn_objective_workspace_bytes(20, 1, 10000, 20) > 0 and lib.lmr_create_mask_workspace_bytes(3, 100, 100) > 0
    # an unknown precision is rejected before anything is launched
    rc = lib.lmr_inr_forward(ctypes.byref(d1), None, None, None, None, 20, None, 100, None, None, 1, None, None, 0, 77, None)
    assert rc != 0 and b"precision" in lib.lmr_last_error()
    with pytest.raises(_lib.LaminrB200Error, match="precision"):
        _lib.check(rc)


def test_conv_model_constructors_consume_rng_like_the_reference(L):
    """`torch.manual_seed(0)` + `Stacked2dCore(...)`, `FullGaussian2d(...)` of the host mirror give the state_dict (names, shapes, values) and leave
    the CPU generator where the reference's vendored neuralpredictors constructors do (BASELINE config 4 and a small 2-channel variant;
    tests/golden/conv_init.npz, oracle/make_golden_conv_init.py)."""
    from oracle.make_golden_conv_init import digest
    from laminr_b200.neuron_models.conv_models import FullGaussian2d, Stacked2dCore
    want = golden("conv_init")
    got = digest(Stacked2dCore, FullGaussian2d)
    assert sorted(got) == sorted(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k


def test_state_dicts_are_the_references(L):
    """Checkpoint compatibility: the template's state_dict (INR, coordinate transform of two targets, RF shift buffers) and the demo1 response model's
    have the reference's keys in the reference's order, shapes, dtypes and -- from the same seed -- values (tests/golden/state_dicts_20.npz,
    oracle/make_golden_state_dicts.py)."""
    from oracle.make_golden_state_dicts import digests
    want = golden("state_dicts_20")
    got = digests(L, device="cpu")
    for k in want:
        assert np.array_equal(got[k], want[k]), (k, got[k] if got[k].ndim == 1 else None)


def test_unsupported_inr_shapes_fail_at_construction(L):
    """Shapes the kernels are not built for are refused when the template is constructed (NotImplementedError naming the reason), not at the first
    launch: a hidden width above 50, fewer than 2 / more than 8 hidden layers, more than 4 output channels (csrc/inr_common.cuh: make_net)."""
    from laminr_b200.inr import INRTemplates
    from torch import nn
    kw = dict(periodic_invariance=True, positional_encoding_dim=50, positional_encoding_projection_scale=10, aux_positional_encoding_dim=50,
              aux_positional_encoding_projection_scale=0.1, final_nonlinearity=nn.Tanh, bias=True, weights_scale=0.1)
    INRTemplates((8, 8), widths=[50] * 2, **kw), INRTemplates((8, 8), widths=[50] * 8, out_channels=4, **kw)   # the supported corners construct
    for bad, why in ((dict(widths=[64] * 4), "hidden widths"), (dict(widths=[50]), "1 hidden layers"), (dict(widths=[50] * 9), "9 hidden layers"),
                     (dict(widths=[50] * 4, out_channels=5), "5 output channels"), (dict(widths=[50, 51, 50, 50]), "hidden widths")):
        with pytest.raises(NotImplementedError, match=why):
            INRTemplates((8, 8), **bad, **kw)
    m = L.neuron_models.simulated("demo1", img_res=[8, 8])
    with pytest.raises(NotImplementedError, match="hidden widths"):
        L.InvarianceManifold(m, _meis(8), device="cpu", widths=[64] * 4, required_img_norm=1.0, pixel_value_lower_bound=-1.0, pixel_value_upper_bound=1.0)


def test_narrow_and_non_uniform_widths_are_zero_padded_views(L):
    """`widths=` below the kernels' 50 (uniform or not; reference inr.py:156-174): the parameters keep the reference's shapes, names and seeded
    values (same CPU draws, in the same order: nn.Linear constructions, B, auxB, then normal_ per layer, inr.py:49-51,207-211) and are strided
    views of zero-padded 50-wide blocks of the flat vector the kernels read; state_dict round-trips."""
    from laminr_b200.inr import INRTemplates
    from torch import nn
    widths, C, pe = [32, 20, 50], 2, 50
    kw = dict(out_channels=C, periodic_invariance=True, positional_encoding_dim=pe, positional_encoding_projection_scale=10, aux_positional_encoding_dim=pe,
              aux_positional_encoding_projection_scale=0.1, final_nonlinearity=nn.Tanh, bias=True, weights_scale=0.1)
    torch.manual_seed(3)
    t = INRTemplates((8, 8), widths=widths, **kw)
    # the reference's draw order restated with plain torch
    torch.manual_seed(3)
    dims = [1 + 4 * pe] + widths + [C]
    lin = [nn.Linear(dims[i], dims[i + 1], bias=True) for i in range(len(dims) - 1)]
    B, auxB = torch.randn(2, pe) * 10, torch.randn(2, pe) * 0.1
    for m_ in lin:
        m_.weight.data.normal_(0.0, 0.1), m_.bias.data.fill_(0)
    assert torch.equal(t.B, B) and torch.equal(t.auxB, auxB)
    for i, m_ in enumerate(lin):
        assert torch.equal(getattr(t.func, f"layer{i}").weight, m_.weight) and tuple(getattr(t.func, f"layer{i}").bias.shape) == (dims[i + 1],)
    flat = t.flat_params()
    assert flat.numel() == 50 * 201 + 50 + 2 * (2500 + 50) + C * 50 + C and t._desc.hidden == 50 and t._desc.n_hidden == 3
    assert sum(p.numel() for p in t.func.parameters()) == sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))
    blk = flat[50 * 201 + 50: 50 * 201 + 50 + 2500].view(50, 50)            # layer 1: [20, 32] in the corner of a zero [50, 50] block
    assert torch.equal(blk[:20, :32], lin[1].weight) and float(blk[20:].abs().sum()) == 0 and float(blk[:, 32:].abs().sum()) == 0
    assert float(flat.abs().sum()) == pytest.approx(float(sum(m_.weight.abs().sum() for m_ in lin)), rel=1e-5)
    flat[50 * 201 + 50] = 7.0                                                 # the parameters ARE views of the flat vector
    assert t.func.layer1.weight[0, 0].item() == 7.0 and t.flat_params().data_ptr() == flat.data_ptr()
    torch.manual_seed(4)
    t2 = INRTemplates((8, 8), widths=widths, **kw)
    t2.flat_params()
    t2.load_state_dict(t.state_dict())
    assert torch.equal(t2.flat_params(), flat) and [k for k in t.state_dict()] == [k for k in t2.state_dict()]


def test_torch_library_operators_are_registered(L):
    """The C-ABI entry points of the generic path are PyTorch custom operators (torch.ops.laminr_b200.*) with CUDA-only kernels: a CPU tensor
    finds no kernel instead of a fallback; the fake kernels give the output shapes without running anything."""
    from laminr_b200 import torch_ops
    for name in torch_ops.OPERATORS:
        assert hasattr(torch.ops.laminr_b200, name), name
    schema = str(torch.ops.laminr_b200.inr_images.default._schema)
    assert "Tensor[] param_views" in schema and "Tensor? angles" in schema and schema.endswith("-> Tensor")
    with pytest.raises(NotImplementedError):
        torch.ops.laminr_b200.constrain(torch.zeros(2, 1, 4, 4), 0, 1.0, 0.0, True, -1.0, True, 1.0, 1e-12)
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        x = torch.empty(3, 1, 5, 6, device="cuda")
        z, stats = torch.ops.laminr_b200.constrain(x, 0, 1.0, 0.0, True, -1.0, True, 1.0, 1e-12)
        assert z.shape == x.shape and stats.shape == (3, 2)
        act, rmax, am = torch.ops.laminr_b200.simresp(x, torch.empty(4, 7, 30, device="cuda"), torch.empty(4, dtype=torch.int32, device="cuda"),
                                                      torch.empty(2, dtype=torch.int32, device="cuda"), 1e-6)
        assert act.shape == (2, 3) and am.dtype == torch.int32
        flat = torch.empty(17801, device="cuda")
        img = torch.ops.laminr_b200.inr_forward(flat, torch.empty(2, 50, device="cuda"), torch.empty(2, 50, device="cuda"), torch.empty(20, device="cuda"),
                                                torch.empty(100, 2, device="cuda"), None, torch.empty(3, device="cuda"), torch.empty(3, 2, device="cuda"),
                                                torch.empty(3, 2, device="cuda"), torch.empty(3, 1, 2, device="cuda"), None, None, None, True, 50, 4, 50, 50, 1,
                                                "fp32")
        assert img.shape == (3, 20, 1, 100)


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """sizeof / field offsets of every struct passed by pointer through the C ABI, as gcc lays out include/laminr_b200.h, against the ctypes mirrors."""
    import subprocess
    from laminr_b200 import _lib
    pairs = [("lmr_inr_desc", _lib.InrDesc), ("lmr_affine", _lib.Affine), ("lmr_convcore_desc", _lib.ConvcoreDesc), ("lmr_convcore_weights", _lib.ConvcoreWeights),
             ("lmr_ff_layer", _lib.FFLayer), ("lmr_ff_net", _lib.FFNet)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "laminr_b200.h"', "int main(void) {"]
    for cname, ct in pairs:
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    for (cname, ct), line in zip(pairs, got):
        nums = [int(v) for v in line.split()[1:]]
        assert nums[0] == ctypes.sizeof(ct), (cname, nums[0], ctypes.sizeof(ct))
        assert nums[1:] == [getattr(ct, f).offset for f, _ in ct._fields_], cname


def test_package_modules_have_no_undefined_global_names():
    """The GPU-only code paths (steppers, sharded drivers) cannot run in the CPU tier: at least every name they load must be bound somewhere in its
    module (an `os.environ` without `import os` would otherwise surface on the GPU box only)."""
    import ast
    import builtins
    import pathlib

    root = pathlib.Path(__file__).resolve().parent.parent
    bad = []
    for f in list((root / "laminr_b200").rglob("*.py")) + [root / "bench.py", root / "__graft_entry__.py"]:
        tree = ast.parse(f.read_text())
        bound = set(dir(builtins)) | {"__file__"}
        for n in ast.walk(tree):
            if isinstance(n, (ast.Import, ast.ImportFrom)):
                bound.update((a.asname or a.name).split(".")[0] for a in n.names)
            elif isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef, ast.ClassDef)):
                bound.add(n.name)
            elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
                bound.add(n.id)
            elif isinstance(n, ast.arg):
                bound.add(n.arg)
            elif isinstance(n, ast.ExceptHandler) and n.name:
                bound.add(n.name)
        bad += [f"{f.relative_to(root)}:{n.lineno} {n.id}" for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load) and n.id not in bound]
    assert not bad, bad


def test_affine_transform_forward_maps_coordinates_like_the_reference_formula():
    """AffineTransform.forward / CoordinateTransform.forward (coordinate_transforms.py:299-309,351-354): out[d, p, a] = sum_b M_d[a, b] x[d, p, b] + t_d[a]
    with M_d = R(angle_d) diag(s_d) [[1, sh_d0], [sh_d1, 1]] -- against an explicit float64 evaluation, uniform and per-axis scale."""
    import math

    from laminr_b200.coordinate_transforms import CoordinateTransform

    g = torch.Generator().manual_seed(0)
    for uniform in (False, True):
        ct = CoordinateTransform(3, only_affine=True, uniform_scale=uniform)
        aff = ct.transforms.Affine
        with torch.no_grad():
            aff.angles.copy_(torch.rand(3, generator=g) * 2 - 1)
            aff.scalings.copy_(torch.rand(aff.scalings.shape, generator=g) + 0.5)
            aff.shears.copy_(torch.rand(3, 2, generator=g) * 0.4 - 0.2)
            aff.translation.copy_(torch.rand(3, 1, 2, generator=g) * 0.2 - 0.1)
        x = torch.rand(3, 7, 2, generator=g) * 2 - 1
        got = ct(x).detach().numpy()
        for d in range(3):
            th = float(aff.angles[d])
            s0 = float(aff.scalings[d, 0])
            s1 = s0 if uniform else float(aff.scalings[d, 1])
            R = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
            M = R @ np.diag([s0, s1]) @ np.array([[1.0, float(aff.shears[d, 0])], [float(aff.shears[d, 1]), 1.0]])
            want = x[d].double().numpy() @ M.T + aff.translation[d].detach().double().numpy()
            np.testing.assert_allclose(got[d], want, rtol=0, atol=2e-6)
    (ct(x).sum()).backward()  # differentiable in the 7 scalars per target
    assert aff.angles.grad is not None and aff.translation.grad is not None


def test_cone_weight_pack_follows_the_header_layout():
    """ops._pack_cone: w_fwd[ci][ky][co / CB][kx][co % CB] = weight[co][ci][ky][kx] and w_bwd[co][ky][ci / CB][kx][ci % CB] =
    weight[co][ci][k-1-ky][k-1-kx] with CB = 4 / 2 / 1 by the divisibility of the blocked channel count (include/laminr_b200.h)."""
    import numpy as np
    import torch
    from laminr_b200 import ops
    rs = np.random.RandomState(0)
    for co, ci, k in ((8, 3, 5), (6, 4, 3), (5, 2, 7), (32, 1, 9)):
        W = torch.from_numpy(rs.randn(co, ci, k, k).astype(np.float32))
        wf = ops._pack_cone(W.permute(1, 2, 0, 3)).numpy()
        wb = ops._pack_cone(W.flip(2, 3).permute(0, 2, 1, 3)).numpy()
        cbf = 4 if co % 4 == 0 else 2 if co % 2 == 0 else 1
        cbb = 4 if ci % 4 == 0 else 2 if ci % 2 == 0 else 1
        assert wf.shape == (ci, k, co // cbf, k, cbf) and wb.shape == (co, k, ci // cbb, k, cbb)
        Wn = W.numpy()
        for o in range(co):
            for i in range(ci):
                for ky in range(k):
                    for kx in range(k):
                        assert wf[i, ky, o // cbf, kx, o % cbf] == Wn[o, i, ky, kx]
                        assert wb[o, ky, i // cbb, kx, i % cbb] == Wn[o, i, k - 1 - ky, k - 1 - kx]
