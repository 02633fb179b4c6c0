"""Boundary glue: mirror of laminr/utils/general.py:5-20."""
import torch
from torch import nn


class SingleNeuronModel(nn.Module):
    """model(x)[:, idx].squeeze().  For the packed simulated population only neuron `idx` is evaluated
    (the reference evaluates every neuron and discards all but one, general.py:11-12)."""

    def __init__(self, model, idx):
        super().__init__()
        self.model = model
        self.idx = idx

    def forward(self, x):
        single = getattr(self.model, "forward_neurons", None)
        if single is not None:
            return single(x, [self.idx])[:, 0].squeeze()
        return self.model(x)[:, self.idx].squeeze()


def infer_device(module):
    """Device of the module's first parameter, else of its first buffer, else the CPU (general.py:15-20)."""
    import itertools

    first = next(itertools.chain(module.parameters(), module.buffers()), None)
    return torch.device("cpu") if first is None else first.device
