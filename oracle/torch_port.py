"""TEST / BASELINE INFRASTRUCTURE ONLY -- torch-CPU restatement of the reference's *op sequence* for one learn / match step.

Purpose: the CPU baseline leg of bench.py on the GPU box, where /root/reference does not exist.  Unlike the numpy
oracle (which evaluates layer 0 in its separable form), this port executes what the reference executes: the dense
[rows, 201] cartesian-product input (inr.py:436-452), five nn.Linear+tanh layers (inr.py:156-174), FixEnergyNorm+Clip
(img_transforms.py:26-32,63-65), the simulated response (simulated_model.py:14-17), apply_mask (mask.py:17-22),
SimCLROnGrid.reg_term (contrastive_loss.py:139-154), autograd backward and torch.optim.Adam (inv_temp_learn_match.py:204-208),
so its wall time is representative of the reference's CPU path.  It is pinned against the same golden vectors as the oracle.
Never imported by the product.
"""
import numpy as np
import torch


class LearnPort:
    def __init__(self, W, b, B, auxB, coords, bank, mask, mei_act, pos, neg, norm=1.0, pmin=-1.0, pmax=1.0, T=0.3, lr=1e-3, res=None,
                 anomaly_mode=False, response=None):
        t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)
        self.W = [t(w).clone().requires_grad_() for w in W]
        self.b = [t(x).clone().requires_grad_() for x in b]
        self.B, self.auxB, self.coords, self.bank, self.mask = t(B), t(auxB), t(coords), (t(bank) if bank is not None else None), t(mask)
        self.response = response  # optional: z [N,C,H,W] -> act [N] (e.g. ConvResponsePort), replaces the simulated filter bank
        self.pos, self.neg = t(pos), t(neg)
        self.mei_act, self.norm, self.pmin, self.pmax, self.T = mei_act, norm, pmin, pmax, T
        self.res = res
        self.opt = torch.optim.Adam(self.W + self.b, lr=lr)
        self.anomaly_mode = anomaly_mode  # the reference leaves torch.autograd.set_detect_anomaly(True) on (:206)

    def images(self, grid):
        aux = grid.reshape(-1, 1)
        a = torch.cat([torch.sin(aux), torch.cos(aux)], 1)
        aux_pe = torch.cat([torch.sin(a @ self.auxB), torch.cos(a @ self.auxB)], -1)
        xy = self.coords[None]
        xy_pe = torch.cat([torch.sin(xy @ self.B), torch.cos(xy @ self.B)], -1).reshape(-1, 2 * self.B.shape[1])
        N, P = aux_pe.shape[0], xy_pe.shape[0]
        tmpl = torch.zeros(1, 1)
        first = torch.cat([tmpl[:, None, :].repeat(1, N, 1), aux_pe[None].repeat(1, 1, 1)], -1).reshape(N, -1)
        inputs = torch.cat([first[:, None, :].repeat(1, P, 1), xy_pe[None].repeat(N, 1, 1)], -1).reshape(N * P, -1)  # [R, 201] materialised
        h = inputs
        for W, b in zip(self.W, self.b):
            h = torch.tanh(torch.nn.functional.linear(h, W, b))
        C = self.W[-1].shape[0]
        return h.reshape(1, N, 1, *self.res, C).permute(0, 2, 1, 5, 3, 4).reshape(N, C, *self.res)

    def loss(self, grid, reg_scale=2.0):
        x = self.images(torch.as_tensor(np.asarray(grid), dtype=torch.float32))
        mean = (self.pmin + self.pmax) / 2
        u = (x - mean).view(len(x), -1)
        s = torch.sqrt(torch.sum(u**2, dim=-1, keepdim=True))
        z = torch.clamp((u / (s + 1e-12) * self.norm + mean).reshape(x.shape), self.pmin, self.pmax)
        if self.response is not None:
            act = self.response(z)
        else:
            r = torch.einsum("bchw,fhw->bf", z, self.bank)
            act = (torch.nn.functional.elu(r.max(dim=1)[0]) + 1) / 2 + 1e-6
        acts = act / self.mei_act
        gain = 2 / (self.pmax - self.pmin)
        xm = ((z - mean) * gain * self.mask) / gain + mean
        X = xm.flatten(1, -1)
        Xn = X / torch.norm(X, p=2, dim=1, keepdim=True)
        S = torch.exp(Xn @ Xn.T / self.T)
        ps, ngs = (S * self.pos).sum(-1), (S * self.neg).sum(-1)
        num = torch.where(ps == 0, 0.0, ps / self.pos.sum(-1)) + 1e-6
        den = torch.where(ngs == 0, 0.0, ngs / self.neg.sum(-1)) + 1e-6
        sim = torch.log(num / den).mean() * self.T / 2
        return -acts.mean() - reg_scale * sim, acts, sim

    def step(self, grid, reg_scale=2.0):
        loss, acts, sim = self.loss(grid, reg_scale)
        self.opt.zero_grad()
        if self.anomaly_mode:
            torch.autograd.set_detect_anomaly(True)
        loss.backward()
        self.opt.step()
        return float(loss.detach())


class ConvResponsePort:
    """SingleNeuronModel(FiringRateEncoder(Stacked2dCore, FullGaussian2d).eval(), idx) as the reference executes it: every layer on the FULL
    frame (conv2d -> eval batch_norm -> elu, conv2d.py:202-253), all neurons read out with grid_sample (gaussian.py:532-581),
    elu(x + offset) + 1 (firing_rate.py:36-52), then one column selected (general.py:5-12).  `arrays`: the flat export of
    oracle.convcore_oracle.model_from_arrays / make_golden.export_convcore_arrays."""

    def __init__(self, arrays, neuron, prefix=""):
        t = lambda k: torch.as_tensor(np.asarray(arrays[prefix + k]), dtype=torch.float32)
        self.L = int(arrays[prefix + "n_layers"])
        self.layers = []
        for l in range(self.L):
            bn = (t(f"bn{l}_mean"), t(f"bn{l}_var"), t(f"bn{l}_gamma"), t(f"bn{l}_beta"), float(arrays[prefix + f"bn{l}_eps"])) if prefix + f"bn{l}_mean" in arrays else None
            self.layers.append((t(f"w{l}"), t(f"b{l}") if prefix + f"b{l}" in arrays else None, bn, bool(arrays[prefix + f"nonlin{l}"])))
        self.xs, self.ys, self.offset = float(arrays[prefix + "xs"]), float(arrays[prefix + "ys"]), float(arrays[prefix + "offset"])
        self.mu, self.features = t("mu").clamp(-1, 1), t("features")
        self.bias = t("bias") if prefix + "bias" in arrays else None
        self.align_corners, self.neuron = bool(arrays[prefix + "align_corners"]), int(neuron)

    def __call__(self, x):
        F = torch.nn.functional
        h = x
        for w, b, bn, nonlin in self.layers:
            h = F.conv2d(h, w, b, padding=w.shape[-1] // 2)
            if bn is not None:
                h = F.batch_norm(h, bn[0], bn[1], bn[2], bn[3], training=False, eps=bn[4])
            if nonlin:
                h = F.elu(h - self.xs) + self.ys
        n = self.mu.shape[0]
        grid = self.mu.reshape(1, n, 1, 2).expand(len(x), n, 1, 2)
        y = F.grid_sample(h, grid, align_corners=self.align_corners)
        y = (y.squeeze(-1) * self.features[None]).sum(1)
        if self.bias is not None:
            y = y + self.bias
        return (F.elu(y + self.offset) + 1)[:, self.neuron].squeeze()
