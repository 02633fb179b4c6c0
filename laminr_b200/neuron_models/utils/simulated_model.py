"""Simulated Gabor-bank response models.

Mirror of the reference interface laminr/neuron_models/utils/simulated_model.py:8-29,71-115,171-204:
`ArbitraryNeuronModel(filters)`, `ArbitraryMultiNeuronModel(list)`, `generate_gabor_filter`,
`random_transformation_matrices`.  Filter *construction* is a one-off NumPy float64 computation on the host in the
same operation order as the reference (bit-identical float32 banks are a precondition for parity and are tested
against the reference's sha256).  The *response* (filter-bank dot products, max, elu) runs in the fused kernels
lmr_simresp_fwd/bwd on a packed `[n, Fmax, H*W]` bank.
"""
import numpy as np
import torch
from torch import nn

from ... import ops


def _transform_pattern(xx, yy, mat, centre):
    px, py = centre
    shifted = np.stack((xx - px, yy - py)).reshape(2, -1)
    xr, yr = (mat @ shifted).reshape(2, *xx.shape)
    return xr + px, yr + py


def gabor_geometry(theta, width_x, width_y, center, grid_size, transformation_matrix=None):
    """The phase- and frequency-independent part of `generate_gabor_filter`: transformed x coordinates and the Gaussian envelope.  A bank whose
    filters differ only in phase / frequency (phase_invariant_neuron_generator: 20..32 phases per neuron) computes this once; the per-filter values
    are the same floating-point operations on the same operands, so the banks stay bit-identical to the reference's."""
    xx, yy = np.meshgrid(np.linspace(-1, 1, grid_size[0]), np.linspace(-1, 1, grid_size[1]))
    dx, dy = center
    xx, yy = xx - dx, yy - dy  # translate
    rot = np.array([[np.cos(theta), np.sin(theta)], [-np.sin(theta), np.cos(theta)]])
    xx, yy = _transform_pattern(xx, yy, rot, (center[0] - dx, center[1] - dy))  # local rotation about the (moved) centre
    if transformation_matrix is not None:  # global transform about the image centre
        xx, yy = _transform_pattern(xx, yy, transformation_matrix, rot @ (0 - dx, 0 - dy))
    envelope = np.exp(-0.5 * (xx**2 / width_x**2 + yy**2 / width_y**2))
    return xx, envelope


def generate_gabor_filter(theta, width_x, width_y, frequency, phase, center, grid_size, transformation_matrix=None):
    """Gaussian envelope x cosine carrier on a [-1,1]^2 grid ('xy' meshgrid: filter[i, j] has x = column j, y = row i)."""
    xx, envelope = gabor_geometry(theta, width_x, width_y, center, grid_size, transformation_matrix)
    return envelope * np.cos(2 * np.pi * frequency * xx + phase)


def random_transformation_matrices(n):
    """rotation . scaling . shearing with the reference's draw order from the NumPy global RNG (simulated_model.py:188-204)."""
    out = []
    for _ in range(n):
        theta = np.random.uniform(0, 2 * np.pi)
        s_x = np.random.uniform(0.5, 1.3)
        s_y = s_x + np.random.uniform(-0.3, 0.3)
        s_h = np.random.uniform(-0.15, 0.15)
        s_v = np.random.uniform(-0.15, 0.15)
        rot = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])
        out.append(np.dot(rot, np.dot(np.array([[s_x, 0], [0, s_y]]), np.array([[1, s_h], [s_v, 1]]))))
    return out


class ArbitraryNeuronModel(nn.Module):
    """act = (elu(max_f <x, w_f>) + 1)/2 + eps for one neuron (simulated_model.py:8-17)."""

    def __init__(self, filters, eps=1e-6):
        super().__init__()
        self.register_buffer("filters", torch.from_numpy(np.ascontiguousarray(filters.astype(np.float32))))
        self.eps = eps
        self._nf = None

    @classmethod
    def from_tensor(cls, filters, eps=1e-6):
        """A neuron whose `filters` buffer IS the given [F,H,W] tensor (no copy; e.g. a row of a packed population bank on the device)."""
        m = cls.__new__(cls)
        nn.Module.__init__(m)
        m.register_buffer("filters", filters)
        m.eps, m._nf = eps, None
        return m

    def forward(self, x):
        F, H, W = self.filters.shape
        if self._nf is None or self._nf.device != x.device:
            self._nf = (torch.tensor([F], dtype=torch.int32, device=x.device), torch.zeros(1, dtype=torch.int32, device=x.device))
        return ops.simresp(x, self.filters.reshape(1, F, H * W), self._nf[0], self._nf[1], self.eps)[0]


class ArbitraryMultiNeuronModel(nn.Module):
    """Population of simulated neurons -> [B, n] (simulated_model.py:20-29).  The per-neuron banks are packed into one
    `[n, Fmax, H*W]` buffer (ragged banks padded by repeating their first filter: the max is unaffected)."""

    def __init__(self, neuron_models_list, packed=None):
        """`packed` (optional): a `[n, F, H, W]` tensor that already holds the banks (equal filter counts) and that the per-neuron `filters` buffers
        are views of -- large populations are then built without a second copy of the bank (819 MB for 1024 neurons at 100x100)."""
        super().__init__()
        self.neuron_models_list = nn.ModuleList(neuron_models_list)
        banks = [m.filters for m in neuron_models_list]
        self.eps = neuron_models_list[0].eps
        if any(m.eps != self.eps for m in neuron_models_list):  # the kernels take ONE epsilon for the population
            raise ValueError("ArbitraryMultiNeuronModel: every neuron must use the same eps")
        fmax = max(b.shape[0] for b in banks)
        self.img_res = tuple(banks[0].shape[1:])
        if packed is None:
            packed = torch.stack([torch.cat([b, b[:1].expand(fmax - b.shape[0], -1, -1)]) if b.shape[0] < fmax else b for b in banks])
        elif tuple(packed.shape) != (len(banks), fmax) + self.img_res or any(b.shape[0] != fmax for b in banks):
            raise ValueError("ArbitraryMultiNeuronModel: `packed` must be [n, F, H, W] with every neuron holding F filters")
        self.register_buffer("packed_filters", packed.reshape(len(banks), fmax, -1).contiguous(), persistent=False)
        self.register_buffer("nfilt", torch.tensor([b.shape[0] for b in banks], dtype=torch.int32), persistent=False)
        self.register_buffer("all_neurons", torch.arange(len(banks), dtype=torch.int32), persistent=False)
        # `packed_filters` is a derived, non-persistent copy of the per-neuron `filters` buffers (the state_dict keys of the reference): rebuild it
        # whenever a state_dict is loaded, or the kernels would keep reading the old banks
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module.repack())

    def repack(self):
        """Rebuilds the packed bank the kernels read from the per-neuron `filters` buffers (after they were modified in place)."""
        fmax = self.packed_filters.shape[1]
        with torch.no_grad():
            for i, m in enumerate(self.neuron_models_list):
                b = m.filters.reshape(m.filters.shape[0], -1).to(self.packed_filters.device)
                if b.data_ptr() == self.packed_filters[i].data_ptr():
                    continue  # the buffer is a view of the packed bank: already current
                self.packed_filters[i, :b.shape[0]].copy_(b)
                if b.shape[0] < fmax:
                    self.packed_filters[i, b.shape[0]:].copy_(b[:1].expand(fmax - b.shape[0], -1))

    @property
    def n_neurons(self):
        return self.packed_filters.shape[0]

    def forward_neurons(self, x, neuron_idxs):
        """Responses of the listed neurons only -> [B, len(neuron_idxs)]."""
        idx = torch.as_tensor(list(neuron_idxs), dtype=torch.int32, device=x.device)
        return ops.simresp(x, self.packed_filters, self.nfilt, idx, self.eps).t()

    def forward(self, x):
        return ops.simresp(x, self.packed_filters, self.nfilt, self.all_neurons, self.eps).t()
