"""SimCLR-on-grid contrastive objective.

Mirror of the reference interface laminr/contrastive_loss.py:9-154 (`SimCLROnGrid`, `GetPosNegMask`,
`get_mask_of_point`) for num_invariances == 1 (what `learn` uses): same constructor, same buffers
(`flat_neighbor_mask`, `flat_neighbor_mask_pos`, `flat_neighbor_mask_neg`).  `reg_term` runs the fused kernels
(lmr_simclr_fwd/bwd); `masked_reg_term` additionally fuses `apply_mask` (laminr/utils/mask.py:17-22).
"""
import numpy as np
import torch
from torch import nn

from . import ops


def get_mask_of_point(point, grid_points_per_dim, num_invariances, neighbor_size, with_periodic_invariances, with_round_neighbor=False):
    """+1 for the `ns` neighbours on either side of `point` (wrapping if periodic), 0 at the point, -1 elsewhere."""
    if num_invariances != 1:
        raise NotImplementedError("laminr_b200: 2-D invariance grids are outside the fused path (the pipeline uses num_invariances=1)")
    n = grid_points_per_dim
    ns = int(neighbor_size * n)
    x = int(point[0])
    mask = -torch.ones(n)
    for off in range(-ns, ns + 1):
        j = x + off
        if with_periodic_invariances:
            mask[j % n] = 1
        elif 0 <= j < n:
            mask[j] = 1
    mask[x] = 0
    return mask


def GetPosNegMask(num_invariances, grid_points_per_dim, neighbor_size, with_periodic_invariances, with_round_neighbor):
    return np.stack([
        get_mask_of_point((x,), grid_points_per_dim, num_invariances, neighbor_size, with_periodic_invariances, with_round_neighbor).numpy()
        for x in range(grid_points_per_dim)
    ])


class SimCLROnGrid(nn.Module):
    def __init__(self, num_invariances, grid_points_per_dim, neighbor_size, temperature, with_periodic_invariances, with_round_neighbor=False, **args):
        super().__init__()
        self.num_invariances, self.points_per_dim = num_invariances, grid_points_per_dim
        self.neighbor_size, self.temperature, self.with_round_neighbor = neighbor_size, temperature, with_round_neighbor
        self.eps = 1e-6
        self.neighbor_mask = torch.tensor(GetPosNegMask(num_invariances, grid_points_per_dim, neighbor_size, with_periodic_invariances,
                                                        with_round_neighbor).astype(np.float32))
        n = grid_points_per_dim**num_invariances
        self.register_buffer("flat_neighbor_mask", self.neighbor_mask.reshape(n, n))
        self.register_buffer("flat_neighbor_mask_pos", torch.where(self.flat_neighbor_mask > 0, self.flat_neighbor_mask, 0.0))
        self.register_buffer("flat_neighbor_mask_neg", torch.where(self.flat_neighbor_mask < 0, self.flat_neighbor_mask, 0.0) * -1)

    def masked_reg_term(self, images, mask, pixel_min, pixel_max):
        """reg_term(apply_mask(images, mask, pixel_min, pixel_max)) in one fused op (inv_temp_learn_match.py:196-203)."""
        if images.dim() != 4:
            images = images.reshape(images.shape[0], 1, 1, -1)
        return ops.simclr_reg(images, mask.to(images.device, torch.float32).reshape(-1), pixel_min, pixel_max, self.flat_neighbor_mask_pos,
                              self.flat_neighbor_mask_neg, self.temperature, self.eps)

    def reg_term(self, images):
        """contrastive_loss.py:139-154 on already-masked images [N, ...]."""
        x = images.reshape(images.shape[0], 1, 1, -1)
        ones = torch.ones(x.shape[-1], dtype=torch.float32, device=x.device)
        # identity rescale: (x - 0) * 1 * 1 * 1 + 0
        return ops.simclr_reg(x, ones, -1.0, 1.0, self.flat_neighbor_mask_pos, self.flat_neighbor_mask_neg, self.temperature, self.eps)
