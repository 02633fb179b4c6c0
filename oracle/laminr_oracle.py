"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the LAMINR hot path, used as the parity oracle.

This module restates, with explicit analytic backward passes (no autograd), the algorithm of the reference
sinzlab/laminr v0.0.7 for the invariance-manifold optimisation loop.  Every function cites the reference
file:line it follows (paths relative to the reference's `src/`).  It is pinned against golden vectors produced by
running the unmodified reference (see `oracle/make_golden.py`, fixtures under `tests/golden/`), so parity is
*pinned by the reference itself*: the reference has no tests / golden vectors of its own.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module,
and only as the checker.  The product (laminr_b200) never imports it and has no CPU fallback.

All functions take a `dtype` through their inputs: feed float32 arrays for a same-precision restatement, float64
for a high-precision check.
"""
import numpy as np

# ----------------------------------------------------------------------------------------------------------------
# latent grid / control logic  (laminr/utils/training.py)
# ----------------------------------------------------------------------------------------------------------------


def latent_grid(grid_points_per_dim, dtype=np.float32):
    """Centres of N equal bins on [0, 2pi)  -- training.py:27-34 (torch.linspace in float32, then + sigma/2)."""
    n = grid_points_per_dim
    sigma = 2 * np.pi / n
    # torch.linspace(0, 2pi, n+1) in float32: start + step*i for the first half, end - step*(n-i) for the second
    step = np.float32((np.float32(2 * np.pi) - np.float32(0)) / np.float32(n))
    idx = np.arange(n + 1)
    lo = np.float32(0) + step * idx.astype(np.float32)
    hi = np.float32(2 * np.pi) - step * (n - idx).astype(np.float32)
    lin = np.where(idx < (n + 1) // 2, lo, hi).astype(np.float32)[:-1]
    # torch adds a python double to a float32 tensor: the scalar is rounded to float32 first
    return (lin + np.float32(sigma / 2)).astype(dtype)


def jittered_grids(grid, jitter_u01, sigma):
    """training.py:42-53: one uniform jitter per step, `rand*sigma - 0.5*sigma`, added to every grid point."""
    jitter = (jitter_u01.astype(np.float32) * np.float32(sigma) - np.float32(0.5 * sigma)).astype(np.float32)
    return (grid[None, :].astype(np.float32) + jitter[:, None]).astype(grid.dtype)


class ParamReduceOnPlateau:
    """training.py:56-81."""

    def __init__(self, factor=0.8, patience=10, mode="min", threshold=0.0):
        self.factor, self.patience, self.mode = factor, patience, mode
        self.threshold = -threshold if mode == "min" else threshold
        self.best_metric = np.float32(np.inf if mode == "min" else -np.inf)
        self.num_epochs_no_improvement = 0

    def step(self, metric_to_monitor, value_to_reduce):
        metric = np.float32(metric_to_monitor)
        bound = np.float32(self.best_metric + np.float32(self.threshold))
        better = metric < bound if self.mode == "min" else metric > bound
        if better:
            self.best_metric = metric
            self.num_epochs_no_improvement = 0
        else:
            self.num_epochs_no_improvement += 1
        if self.num_epochs_no_improvement > self.patience:
            value_to_reduce = value_to_reduce * self.factor
            self.num_epochs_no_improvement = 0
        return value_to_reduce


def check_activity_requirements(acts, requirements):
    """training.py:7-15 (Tensor.std is unbiased)."""
    acts = np.asarray(acts)
    return bool(
        acts.mean() > requirements["avg"]
        and acts.std(ddof=1) < requirements["std"]
        and acts.min() > requirements["necessary_min"]
    )


class ImprovementChecker:
    """training.py:84-102."""

    def __init__(self, patience=5, ignore_diff_smaller_than=1e-3):
        self.current_value = np.nan
        self.iter = 0
        self.has_not_improved_counter = 0
        self.patience = patience
        self.ignore_diff_smaller_than = ignore_diff_smaller_than

    def is_increasing(self, value):
        if self.iter == 0:
            self.current_value = value
            self.iter += 1
        elif value <= self.current_value + self.ignore_diff_smaller_than:
            self.has_not_improved_counter += 1
        else:
            self.has_not_improved_counter = 0
            self.current_value = value
        return self.has_not_improved_counter != self.patience


# ----------------------------------------------------------------------------------------------------------------
# INR generator  (laminr/inr.py, laminr/coordinate_transforms.py)
# ----------------------------------------------------------------------------------------------------------------


def pixel_coords(img_res, dtype=np.float32):
    """inr.py:194-198: meshgrid(linspace(-1,1,H), linspace(-1,1,W), 'ij') stacked -> [P, 2] in (row=y, col=x)."""

    def tl(n):  # torch.linspace(-1, 1, n) in float32 (symmetric two-sided evaluation)
        if n == 1:
            return np.array([-1.0], np.float32)
        step = np.float32(np.float32(2.0) / np.float32(n - 1))
        i = np.arange(n)
        lo = np.float32(-1) + step * i.astype(np.float32)
        hi = np.float32(1) - step * (n - 1 - i).astype(np.float32)
        return np.where(i < n // 2, lo, hi).astype(np.float32)

    yy, xx = np.meshgrid(tl(img_res[0]), tl(img_res[1]), indexing="ij")
    return np.stack([yy, xx], -1).reshape(-1, 2).astype(dtype)


def latent_encode(grid, auxB):
    """inr.py:371-386 (periodic: [sin g, cos g]) then :401-409 ([sin(a.auxB), cos(a.auxB)]) -> [N, 2*apd]."""
    g = np.asarray(grid).reshape(-1, 1)
    a = np.concatenate([np.sin(g), np.cos(g)], 1)
    phi = a @ auxB
    return np.concatenate([np.sin(phi), np.cos(phi)], 1), a, phi


def affine_matrices(angles, scalings, shears):
    """coordinate_transforms.py:276-296: M = R(theta) . diag(s) . [[1, sh0], [sh1, 1]]   -> [D,2,2]."""
    c, s = np.cos(angles), np.sin(angles)
    R = np.stack([np.stack([c, -s], -1), np.stack([s, c], -1)], -2)
    S = np.zeros(R.shape, R.dtype)
    S[:, 0, 0], S[:, 1, 1] = scalings[:, 0], scalings[:, 1]
    Hs = np.stack(
        [np.stack([np.ones_like(shears[:, 0]), shears[:, 0]], -1), np.stack([shears[:, 1], np.ones_like(shears[:, 1])], -1)],
        -2,
    )
    return np.einsum("nab,nbc,ncd->nad", R, S, Hs), (R, S, Hs)


def coords_shifts(template_rf_xy, target_rf_xy, img_res, dtype=np.float32):
    """inr.py:260-296.  rf centres are (x, y) in pixels; everything is flipped to (y, x)."""
    res = np.asarray(img_res, dtype)
    t_yx = np.asarray(template_rf_xy, dtype)[::-1]
    shift_to = (t_yx - res / 2) / (res / 2)  # coords_shift_to_center
    tc = t_yx / res * 2 - 1  # template_center_in_coords_space
    targ_yx = np.asarray(target_rf_xy, dtype)[:, ::-1]
    shift_from = -1 * ((targ_yx - res / 2) / (res / 2))  # coords_shift_from_center
    return shift_to.astype(dtype), tc.astype(dtype), shift_from.astype(dtype)


def match_coords(coords, M, translation, tc, shift_to, shift_from, clamp=True):
    """inr.py:337-369 + coordinate_transforms.py:298-309.

    c1 = M (c - tc) + t + tc (differentiable); c2 = M (shift_from + shift_to) + t (no_grad);
    out = clamp(c1 + c2, -1, 1) with straight-through gradient (`.data.clamp_`).   -> [D,P,2]
    """
    rel = coords - tc  # [P,2]
    c1 = np.einsum("nab,cb->nca", M, rel) + translation[:, None, :] + tc
    adj = shift_from + shift_to  # [D,2]
    c2 = np.einsum("nab,nb->na", M, adj) + translation
    out = c1 + c2[:, None, :]
    if clamp:
        out = np.clip(out, -1, 1)
    return out.astype(coords.dtype)


def inr_forward(weights, biases, B, auxB, grid, coords_dp, img_res, keep=False):
    """inr.py:53-100.  weights[l]: [out,in] (torch Linear layout); tanh after every layer incl. the last.

    coords_dp: [D,P,2].  Layer 0 is evaluated in its separable form
        z0 = W0[:,1:1+2a] alpha_n + W0[:,1+2a:] beta_dp + b0        (template_input == 0, inr.py:200-205)
    which equals the reference's dense product over the cartesian input [template, aux_pe, xy_pe] (inr.py:436-452)
    up to float rounding.  Returns img_pre [D,N,C,H,W].
    """
    alpha, a, aphi = latent_encode(grid, auxB)
    na = alpha.shape[1]
    W0 = weights[0]
    phi = coords_dp @ B  # [D,P,pe]
    beta = np.concatenate([np.sin(phi), np.cos(phi)], -1)  # [D,P,2pe]
    A = alpha @ W0[:, 1 : 1 + na].T + biases[0]  # [N,h]
    Cp = beta @ W0[:, 1 + na :].T  # [D,P,h]
    z = A[None, :, None, :] + Cp[:, None, :, :]  # [D,N,P,h]
    hs = [np.tanh(z)]
    for W, b in zip(weights[1:], biases[1:]):
        hs.append(np.tanh(hs[-1] @ W.T + b))
    y = hs[-1]  # [D,N,P,C]
    D, N, P, C = y.shape
    img = np.transpose(y, (0, 1, 3, 2)).reshape(D, N, C, img_res[0], img_res[1])
    if keep:
        return img, dict(alpha=alpha, beta=beta, phi=phi, hs=hs, na=na)
    return img


def inr_backward(weights, B, cache, g_img, want_weights=True, want_coords=False):
    """Analytic backward of inr_forward.  g_img: [D,N,C,H,W].

    Returns (dW list, db list) and/or dL/dcoords [D,P,2].
    """
    hs, alpha, beta, phi, na = cache["hs"], cache["alpha"], cache["beta"], cache["phi"], cache["na"]
    D, N, C = g_img.shape[:3]
    g = np.transpose(g_img.reshape(D, N, C, -1), (0, 1, 3, 2))  # [D,N,P,C]
    L = len(weights)
    dW, db = [None] * L, [None] * L
    delta = g * (1 - hs[-1] ** 2)
    for l in range(L - 1, 0, -1):
        if want_weights:
            dW[l] = np.einsum("dnpo,dnpi->oi", delta, hs[l - 1])
            db[l] = delta.sum((0, 1, 2))
        delta = (delta @ weights[l]) * (1 - hs[l - 1] ** 2)
    # layer 0, separable
    out = {}
    if want_weights:
        W0 = weights[0]
        dW0 = np.zeros_like(W0)
        dA = delta.sum((0, 2))  # [N,h]
        dW0[:, 1 : 1 + na] = dA.T @ alpha
        dW0[:, 1 + na :] = np.einsum("dpo,dpk->ok", delta.sum(1), beta)
        dW[0], db[0] = dW0, dA.sum(0)
        out["dW"], out["db"] = dW, db
    if want_coords:
        G0 = delta.sum(1)  # [D,P,h]
        dbeta = G0 @ weights[0][:, 1 + na :]  # [D,P,2pe]
        pe = B.shape[1]
        dphi = dbeta[..., :pe] * np.cos(phi) - dbeta[..., pe:] * np.sin(phi)
        out["dcoords"] = dphi @ B.T
    return out


def affine_backward(G, coords, tc, angles, scalings, shears):
    """Gradient of the 7 affine scalars per target given G = dL/dcoords [D,P,2]  (SURVEY Appendix A.2).

    The clamp is straight-through and c2 is a constant (torch.no_grad), so only c1 = M (c - tc) + t + tc carries
    gradient (inr.py:345-369).
    """
    rel = coords - tc
    dT = G.sum(1)  # [D,2]
    dM = np.einsum("dpa,pb->dab", G, rel)
    _, (R, S, Hs) = affine_matrices(angles, scalings, shears)
    c, s = np.cos(angles), np.sin(angles)
    dR = np.stack([np.stack([-s, -c], -1), np.stack([c, -s], -1)], -2)
    d_angles = np.einsum("dab,dab->d", dM, np.einsum("nab,nbc,ncd->nad", dR, S, Hs))
    d_scal = np.zeros_like(scalings)
    for k in range(2):
        E = np.zeros((2, 2), angles.dtype)
        E[k, k] = 1
        d_scal[:, k] = np.einsum("dab,dab->d", dM, np.einsum("nab,bc,ncd->nad", R, E, Hs))
    Q = np.einsum("nba,nbc->nac", np.einsum("nab,nbc->nac", R, S), dM)
    d_shear = np.stack([Q[:, 0, 1], Q[:, 1, 0]], -1)
    return d_angles, d_scal, d_shear, dT


# ----------------------------------------------------------------------------------------------------------------
# image-constraint projection  (laminr/utils/img_transforms.py; featurevis/ops.py for the MEI twins)
# ----------------------------------------------------------------------------------------------------------------


def constrain_norm_fwd(x, norm, mean, pmin, pmax, eps=1e-12):
    """FixEnergyNorm (img_transforms.py:26-32) + Clip (:63-65).  x: [B, ...]."""
    dt = x.dtype
    u = (x - dt.type(mean)).reshape(len(x), -1)
    s = np.sqrt((u * u).sum(-1, keepdims=True))
    y = (u / (s + dt.type(eps)) * dt.type(norm) + dt.type(mean)).reshape(x.shape)
    z = y if pmin is None and pmax is None else np.clip(y, pmin, pmax)
    return z.astype(dt), (u, s, y)


def constrain_norm_bwd(g_z, cache, norm, pmin, pmax, eps=1e-12):
    """SURVEY Appendix A.4.  clamp passes gradient where lo <= y <= hi (inclusive)."""
    u, s, y = cache
    gy = g_z.copy()
    if pmin is not None:
        gy = gy * (y >= pmin)
    if pmax is not None:
        gy = gy * (y <= pmax)
    gyf = gy.reshape(len(gy), -1)
    dot = (gyf * u).sum(-1, keepdims=True)
    gx = norm / (s + eps) * gyf - norm * dot / ((s + eps) ** 2 * s) * u
    return gx.reshape(g_z.shape).astype(g_z.dtype)


def constrain_std_fwd(x, std, mean, pmin, pmax, eps=1e-12):
    """Normalize (img_transforms.py:46-51, unbiased std over dims [1,2,3]) + Clip."""
    dt = x.dtype
    xf = x.reshape(len(x), -1)
    mu = xf.mean(-1, keepdims=True)
    sd = xf.std(-1, ddof=1, keepdims=True)
    y = (dt.type(std) * (xf - mu) / (sd + dt.type(eps)) + dt.type(mean)).reshape(x.shape)
    z = y if pmin is None and pmax is None else np.clip(y, pmin, pmax)
    return z.astype(dt), (xf - mu, sd, y)


def constrain_std_bwd(g_z, cache, std, pmin, pmax, eps=1e-12):
    xc, sd, y = cache
    n = xc.shape[1]
    gy = g_z.copy()
    if pmin is not None:
        gy = gy * (y >= pmin)
    if pmax is not None:
        gy = gy * (y <= pmax)
    gyf = gy.reshape(len(gy), -1)
    k = std / (sd + eps)
    dot = (gyf * xc).sum(-1, keepdims=True)
    gx = k * (gyf - gyf.mean(-1, keepdims=True)) - std * dot / ((sd + eps) ** 2 * (n - 1) * sd) * xc
    return gx.reshape(g_z.shape).astype(g_z.dtype)


def change_norm(x, norm, mean):
    """featurevis/ops.py:475-481 (eps 1e-9)."""
    dt = x.dtype
    u = (x - dt.type(mean)).reshape(len(x), -1)
    s = np.sqrt((u * u).sum(-1, keepdims=True))
    return (u / (s + dt.type(1e-9)) * dt.type(norm) + dt.type(mean)).reshape(x.shape).astype(dt)


def change_stats(x, std, mean):
    """featurevis/ops.py:452-459 (unbiased std, eps 1e-9)."""
    dt = x.dtype
    xf = x.reshape(len(x), -1)
    sd = xf.std(-1, ddof=1)
    mu = xf.mean(-1)
    shape = (len(x),) + (1,) * (x.ndim - 1)
    return ((x - mu.reshape(shape)) * (dt.type(std) / (sd + dt.type(1e-9))).reshape(shape) + dt.type(mean)).astype(dt)


# ----------------------------------------------------------------------------------------------------------------
# simulated response model  (laminr/neuron_models/utils/simulated_model.py:8-29)
# ----------------------------------------------------------------------------------------------------------------


def elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def simresp_fwd(x, filters, eps=1e-6):
    """ArbitraryNeuronModel.forward: r_f = sum_{c,h,w} x w_f; m = max_f; act = (elu(m)+1)/2 + eps.

    x: [B,C,H,W], filters: [F,H,W] -> act [B], (r [B,F], argmax [B])
    """
    r = np.einsum("bchw,fhw->bf", x, filters)
    am = r.argmax(1)
    m = r[np.arange(len(r)), am]
    act = (elu(m) + 1) / 2 + x.dtype.type(eps)
    return act.astype(x.dtype), (r, am)


def simresp_bwd(g_act, cache, x_shape, filters):
    """d act / d x[c] = 1/2 * (1 if m > 0 else e^m) * w_{f*}  (SURVEY Appendix A.5)."""
    r, am = cache
    m = r[np.arange(len(r)), am]
    de = np.where(m > 0, 1.0, np.exp(np.minimum(m, 0))).astype(filters.dtype)
    coef = g_act * 0.5 * de
    g = coef[:, None, None, None] * filters[am][:, None, :, :]
    return np.broadcast_to(g, x_shape).astype(filters.dtype).copy()


def multi_resp_fwd(x, banks):
    """ArbitraryMultiNeuronModel.forward (simulated_model.py:25-29): stack over neurons -> [B, n]."""
    return np.stack([simresp_fwd(x, f)[0] for f in banks], 1)


# ----------------------------------------------------------------------------------------------------------------
# masked contrastive objective  (laminr/utils/mask.py:7-22, laminr/contrastive_loss.py)
# ----------------------------------------------------------------------------------------------------------------


def pos_neg_masks(n, neighbor_size, periodic, dtype=np.float32):
    """contrastive_loss.py:9-34 (num_invariances == 1) and :115-137.  -> pos [n,n], neg [n,n] as 0/1."""
    ns = int(neighbor_size * n)
    full = -np.ones((n, n), dtype)
    for x in range(n):
        m = full[x]
        if x < ns:
            m[: x + ns + 1] = 1
            if periodic:
                m[x - ns :] = 1
        elif x + ns >= n:
            m[x - ns :] = 1
            if periodic:
                m[: ns - n + x + 1] = 1
        else:
            m[x - ns : x + ns + 1] = 1
        m[x] = 0
    return (full > 0).astype(dtype), (full < 0).astype(dtype)


def apply_mask(x, mask, pmin, pmax):
    """mask.py:7-22: rescale to [-1,1], multiply by the mask, rescale back (same op order)."""
    dt = x.dtype
    in_mean, gain = (pmin + pmax) / 2, (1 - (-1)) / (pmax - pmin)
    r = (x - dt.type(in_mean)) * dt.type(gain) + dt.type(0.0)
    r = r * mask.astype(dt)  # mask [H,W] broadcast over [N,C,H,W]
    out_gain = (pmax - pmin) / (1 - (-1))
    return ((r - dt.type(0.0)) * dt.type(out_gain) + dt.type((pmin + pmax) / 2)).astype(dt)


def simclr_reg_fwd(xm, pos, neg, T, eps=1e-6):
    """contrastive_loss.py:139-154 with cosine_similarity (img_similarity_metrics.py:4-10).  xm: [N, ...]."""
    dt = xm.dtype
    X = xm.reshape(len(xm), -1)
    rho = np.sqrt((X * X).sum(1, keepdims=True))
    Xh = X / rho
    cos = Xh @ Xh.T
    S = np.exp(cos / dt.type(T))
    ps, npos = (S * pos).sum(-1), pos.sum(-1)
    ngs, nneg = (S * neg).sum(-1), neg.sum(-1)
    num = np.where(ps == 0, 0.0, ps / npos).astype(dt) + dt.type(eps)
    den = np.where(ngs == 0, 0.0, ngs / nneg).astype(dt) + dt.type(eps)
    reg = np.log(num / den).mean() * dt.type(T) / 2
    return dt.type(reg), (X, rho, Xh, cos, S, ps, ngs, npos, nneg, num, den)


def simclr_reg_bwd(cache, pos, neg, T):
    """d reg / d xm  (SURVEY Appendix A.6)."""
    X, rho, Xh, cos, S, ps, ngs, npos, nneg, num, den = cache
    N = len(X)
    A = (T / (2 * N)) * (
        pos * np.where(ps == 0, 0.0, 1.0 / (npos * num))[:, None] - neg * np.where(ngs == 0, 0.0, 1.0 / (nneg * den))[:, None]
    )
    Gam = A * S / T
    Lam = Gam + Gam.T
    gXh = Lam @ Xh
    gX = (gXh - (gXh * Xh).sum(1, keepdims=True) * Xh) / rho
    return gX.astype(X.dtype)


# ----------------------------------------------------------------------------------------------------------------
# step-level objectives  (laminr/inv_temp_learn_match.py)
# ----------------------------------------------------------------------------------------------------------------


def learn_step(weights, biases, B, auxB, grid, coords, img_res, filters, rf_mask, mei_act, reg_scale, norm, pmin, pmax, pos, neg, T, response=None,
               std=None):
    """One `learn` hot step up to (and excluding) the optimiser: inv_temp_learn_match.py:187-207.

    `response(z, g_act) -> (act[N], g_z)` replaces the simulated filter bank when given (e.g. the conv-core oracle,
    oracle/convcore_oracle.py:response_and_grad).  `std` (with norm=None): StandardizeClip instead of FixEnergyNormClip
    (inv_temp_learn_match.py:52-73).  Returns dict(loss, acts (normalised), sim, img_pre, img_post, dW, db).
    """
    con_fwd, con_bwd, target = (constrain_std_fwd, constrain_std_bwd, std) if std is not None else (constrain_norm_fwd, constrain_norm_bwd, norm)
    dt = weights[0].dtype
    mean = (pmin + pmax) / 2
    img_pre, cache = inr_forward(weights, biases, B, auxB, grid, coords[None], img_res, keep=True)
    x = img_pre[0]  # [N,C,H,W]
    z, ccache = con_fwd(x, target, mean, pmin, pmax)
    N = len(x)
    g_act = np.full(N, -1.0 / (N * mei_act), dt)
    if response is not None:
        act, g_z = response(z, g_act)
        act, g_z = act.astype(dt), g_z.astype(dt)
    else:
        act, rcache = simresp_fwd(z, filters)
        g_z = simresp_bwd(g_act, rcache, z.shape, filters)
    acts = act / dt.type(mei_act)
    xm = apply_mask(z, rf_mask, pmin, pmax)
    sim, scache = simclr_reg_fwd(xm, pos, neg, T)
    loss = -acts.mean() - dt.type(reg_scale) * sim
    # backward
    g_xm = simclr_reg_bwd(scache, pos, neg, T).reshape(z.shape) * dt.type(-reg_scale)
    g_z = g_z + g_xm * rf_mask.astype(dt)  # d apply_mask / dx = gain_in * mask * gain_out = mask
    g_x = con_bwd(g_z, ccache, target, pmin, pmax)
    grads = inr_backward(weights, B, cache, g_x[None], want_weights=True)
    return dict(loss=loss, acts=acts, sim=sim, img_pre=x, img_post=z, dW=grads["dW"], db=grads["db"], g_img_pre=g_x)


def match_step(weights, biases, B, auxB, grid, coords, img_res, banks, mei_acts, affine, shifts, norm, pmin, pmax, clamp=True, responses=None,
               std=None):
    """One `match` hot step: inv_temp_learn_match.py:352-365.  banks: list of [F,H,W] for the D targets.

    affine = dict(angles[D], scalings[D,2], shears[D,2], translation[D,2]); shifts = (shift_to, tc, shift_from).
    Only neuron d's response to its own images enters the loss (the diagonal, :361).  `std` (with norm=None): StandardizeClip.
    """
    con_fwd, con_bwd, target = (constrain_std_fwd, constrain_std_bwd, std) if std is not None else (constrain_norm_fwd, constrain_norm_bwd, norm)
    dt = weights[0].dtype
    shift_to, tc, shift_from = shifts
    M, _ = affine_matrices(affine["angles"], affine["scalings"], affine["shears"])
    cdp = match_coords(coords, M, affine["translation"], tc, shift_to, shift_from, clamp)
    img_pre, cache = inr_forward(weights, biases, B, auxB, grid, cdp, img_res, keep=True)
    D, N = img_pre.shape[:2]
    mean = (pmin + pmax) / 2
    flat = img_pre.reshape((D * N,) + img_pre.shape[2:])
    z, ccache = con_fwd(flat, target, mean, pmin, pmax)
    zs = z.reshape(img_pre.shape)
    acts_dn = np.zeros((D, N), dt)
    g_z = np.zeros_like(zs)
    for d in range(D):
        g_act = np.full(N, -1.0 / (N * D * mei_acts[d]), dt)
        if responses is not None:  # responses[d](z, g_act) -> (act[N], g_z): any other response model (conv-core oracle)
            acts_dn[d], g_z[d] = responses[d](zs[d], g_act)
            continue
        a, rc = simresp_fwd(zs[d], banks[d])
        acts_dn[d] = a
        g_z[d] = simresp_bwd(g_act, rc, zs[d].shape, banks[d])
    acts = acts_dn.mean(1) / mei_acts.astype(dt)
    loss = -acts.mean()
    g_x = con_bwd(g_z.reshape(z.shape), ccache, target, pmin, pmax).reshape(img_pre.shape)
    G = inr_backward(weights, B, cache, g_x, want_weights=False, want_coords=True)["dcoords"]
    d_ang, d_scal, d_shear, d_t = affine_backward(G, coords, tc, affine["angles"], affine["scalings"], affine["shears"])
    return dict(
        loss=loss, acts=acts, acts_dn=acts_dn, img_pre=img_pre, img_post=zs, coords=cdp,
        d_angles=d_ang, d_scalings=d_scal, d_shears=d_shear, d_translation=d_t,
    )


def adam_step(p, g, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (non-amsgrad, no weight decay), single-tensor formulation, in the array dtype."""
    dt = p.dtype
    m = m + (g - m) * dt.type(1 - b1)  # exp_avg.lerp_(grad, 1-beta1)
    v = v * dt.type(b2) + dt.type(1 - b2) * g * g
    bc1 = 1 - b1**step
    bc2 = 1 - b2**step
    denom = np.sqrt(v) / dt.type(np.sqrt(bc2)) + dt.type(eps)
    p = p - dt.type(lr / bc1) * (m / denom)
    return p.astype(dt), m.astype(dt), v.astype(dt)


# ----------------------------------------------------------------------------------------------------------------
# Gaussian blur, the optional gradient preconditioner (featurevis/ops.py:378-423)
# ----------------------------------------------------------------------------------------------------------------
def gaussian_blur(x, sigma, truncate=4, pad_mode="reflect"):
    """featurevis.ops.GaussianBlur.__call__ without decay (ops.py:398-423).  x: [B,C,H,W].  The window is scipy.signal.gaussian(2h+1, std)
    = exp(-0.5 (k/std)^2) (ops.py:409-412), cast to x's dtype; F.pad's modes map to numpy's: 'reflect' -> 'reflect' (no edge repeat),
    'replicate' -> 'edge', 'constant' -> zeros (ops.py:417); depthwise y then x correlation (:418-421); division by the product of the
    window sums (:422)."""
    sig = sigma if isinstance(sigma, tuple) else (sigma,) * 2
    x = np.asarray(x)
    hy, hx = max(int(round(sig[0] * truncate)), 1), max(int(round(sig[1] * truncate)), 1)
    wy = np.exp(-0.5 * ((np.arange(2 * hy + 1) - hy) / sig[0]) ** 2).astype(x.dtype)
    wx = np.exp(-0.5 * ((np.arange(2 * hx + 1) - hx) / sig[1]) ** 2).astype(x.dtype)
    mode = {"reflect": "reflect", "replicate": "edge", "constant": "constant"}[pad_mode]
    xp = np.pad(x, ((0, 0), (0, 0), (hy, hy), (hx, hx)), mode=mode)
    H, W = x.shape[-2:]
    by = np.zeros(xp.shape[:2] + (H, xp.shape[3]), x.dtype)
    for k in range(2 * hy + 1):
        by += wy[k] * xp[:, :, k:k + H, :]
    bx = np.zeros(x.shape, x.dtype)
    for k in range(2 * hx + 1):
        bx += wx[k] * by[:, :, :, k:k + W]
    return bx / (wy.sum() * wx.sum())


# ----------------------------------------------------------------------------------------------------------------
# MEI gradient ascent  (featurevis/core.py:54-136 with SGD; laminr/utils/mei.py:9-71)
# ----------------------------------------------------------------------------------------------------------------


def mei_ascent(x0, filters, step_size, num_iterations, pmin, pmax, norm=None, std=None):
    """x <- post(x + lr * d act/dx), post = ChangeNorm|ChangeStats -> ClipRange; SGD without momentum.

    x0: [1,C,H,W] *before* the initial post_update (mei.py:33-42 applies it once).  Returns (x, fevals[iters+1]).
    """
    mean = (pmin + pmax) / 2
    dt = x0.dtype

    def post(x):
        x = change_norm(x, norm, mean) if norm is not None else change_stats(x, std, mean)
        return np.clip(x, pmin, pmax).astype(dt)

    x = post(x0)
    fevals = []
    for _ in range(num_iterations):
        act, cache = simresp_fwd(x, filters)
        fevals.append(float(act[0]))
        g = simresp_bwd(np.full(1, -1.0, dt), cache, x.shape, filters)  # grad of (-f)
        x = post((x - dt.type(step_size) * g).astype(dt))
    fevals.append(float(simresp_fwd(x, filters)[0][0]))
    return x, fevals


# ----------------------------------------------------------------------------------------------------------------
# simulated Gabor banks  (laminr/neuron_models/utils/simulated_model.py:52-115, simulated_models.py:74-139,243-274,380-394)
# ----------------------------------------------------------------------------------------------------------------


def _transform_pattern(xx, yy, mat, centre):
    px, py = centre
    xy = np.stack((xx - px, yy - py)).reshape(2, -1)
    xr, yr = (mat @ xy).reshape(2, *xx.shape)
    return xr + px, yr + py


def gabor_filter(theta, width_x, width_y, frequency, phase, center, grid_size, transformation_matrix=None):
    """simulated_model.py:71-115 (float64; 'xy' meshgrid so filter[i,j] has x = col j, y = row i)."""
    x = np.linspace(-1, 1, grid_size[0])
    y = np.linspace(-1, 1, grid_size[1])
    xx, yy = np.meshgrid(x, y)
    dx, dy = center
    xx, yy = xx - dx, yy - dy
    px, py = center
    rot = np.array([[np.cos(theta), np.sin(theta)], [-np.sin(theta), np.cos(theta)]])
    xx, yy = _transform_pattern(xx, yy, rot, (px - dx, py - dy))
    if transformation_matrix is not None:
        c = rot @ (0 - dx, 0 - dy)
        xx, yy = _transform_pattern(xx, yy, transformation_matrix, c)
    envelope = np.exp(-0.5 * (xx**2 / width_x**2 + yy**2 / width_y**2))
    return envelope * np.cos(2 * np.pi * frequency * xx + phase)


def phase_invariant_bank(loc, img_res, transformation_matrix=None, num_filters=32):
    """simulated_models.py:243-274."""
    out = []
    for phase in np.linspace(0, 2 * np.pi, num_filters, endpoint=False):
        gb = gabor_filter(0, 0.14, 0.14, 2.3, phase, loc, img_res, transformation_matrix)
        out.append((gb / np.linalg.norm(gb)).astype(np.float32))
    return np.stack(out)


def neuron2_bank(loc, img_res, transformation_matrix=None, num_filters=30):
    """simulated_models.py:74-139."""
    sx, sy, sf, lo, hi = 0.11, 0.15, 3, 0.9, 1.5
    locs = np.array([[-0.09, 0.01], [-0.01, -0.05], [0.11, 0.04]]) + loc
    out = []
    for x in np.linspace(0, 2 * np.pi, num_filters):
        sc = [(np.sin(x + k * np.pi / 3) + 1) / 2 * (hi - lo) + lo for k in range(3)]
        th = [-np.pi / 8, 0, np.pi / 8]
        gbs = [
            gabor_filter(th[k], sx * sc[k], sy * sc[k], sf / sc[k], 0, locs[k], img_res, transformation_matrix)
            for k in range(3)
        ]
        gb = (gbs[0] + gbs[1] + gbs[2]) / 3
        out.append((gb / np.linalg.norm(gb)).astype(np.float32))
    return np.stack(out)


def demo1_banks(img_res):
    """simulated_models.py:380-391."""
    t45 = np.array([[0.70710678, -0.70710678], [0.70710678, 0.70710678]])
    return [
        phase_invariant_bank([0.2, 0.2], img_res, num_filters=20),
        phase_invariant_bank([-0.2, -0.2], img_res, transformation_matrix=t45, num_filters=20),
        neuron2_bank([-0.2, 0.2], img_res, num_filters=60),
    ]
