"""Host-side control logic of the optimisation loops.  Mirror of laminr/utils/training.py:7-102 (semantics preserved
exactly: these decide when `learn` / `match` stop).  All randomness is drawn from the CPU generator in the same
order as the reference, including the draw `DataLoader.__iter__` makes for its base seed."""
import numpy as np
import torch


def check_activity_requirements(acts, requirements):
    return bool(acts.mean() > requirements["avg"] and acts.std() < requirements["std"] and acts.min() > requirements["necessary_min"])


class _GridBatches:
    """Iterable over the jittered grids of one epoch, `[N, num_invariances]` per step (DataLoader(grids, batch_size=N) in the reference)."""

    def __init__(self, grids, batch):
        self.grids, self.batch = grids, batch

    def __len__(self):
        return self.grids.shape[0] // self.batch

    def __iter__(self):
        # torch's DataLoader draws one int64 from the global CPU generator per iterator (its worker base seed);
        # replay it so that the jitter stream of later epochs stays aligned with the reference
        torch.empty((), dtype=torch.int64).random_()
        for i in range(len(self)):
            yield self.grids[i * self.batch:(i + 1) * self.batch]


class JitteringGridDatamodule:
    def __init__(self, num_invariances, grid_points_per_dim, steps_per_epoch):
        if num_invariances != 1:
            raise NotImplementedError("laminr_b200: num_invariances == 1 on the fused path")
        self.num_invariances, self.grid_points_per_dim, self.steps_per_epoch = num_invariances, grid_points_per_dim, steps_per_epoch
        self.sigma = 2 * np.pi / grid_points_per_dim
        grid = torch.linspace(0, 2 * np.pi, grid_points_per_dim + 1)[:-1] + self.sigma / 2
        self.grid = grid.reshape(-1, 1)
        self.grid0 = torch.linspace(0, 2 * np.pi, grid_points_per_dim + 1)[:-1].reshape(-1, 1)
        self.base_grid = self.grid.repeat(steps_per_epoch, 1)

    def train_dataloader(self):
        jitter = torch.rand([self.steps_per_epoch, self.num_invariances]) * self.sigma - 0.5 * self.sigma
        grids = self.base_grid + jitter.repeat_interleave(self.grid_points_per_dim**self.num_invariances, dim=0)
        return _GridBatches(grids, self.grid_points_per_dim**self.num_invariances)


class ParamReduceOnPlateau:
    def __init__(self, factor=0.8, patience=10, mode="min", threshold=0):
        self.factor, self.patience, self.mode, self.threshold = factor, patience, mode, threshold
        if mode == "min":
            self.op, self.threshold, self.best_metric = torch.lt, -threshold, torch.tensor(float("inf"))
        if mode == "max":
            self.op, self.best_metric = torch.gt, torch.tensor(float("-inf"))
        self.num_epochs_no_improvement = 0

    def step(self, metric_to_monitor, value_to_reduce):
        metric_to_monitor = torch.as_tensor(metric_to_monitor).detach().cpu()
        if self.op(metric_to_monitor, self.best_metric + self.threshold):
            self.best_metric = metric_to_monitor
            self.num_epochs_no_improvement = 0
        else:
            self.num_epochs_no_improvement += 1
        if self.num_epochs_no_improvement > self.patience:
            value_to_reduce = value_to_reduce * self.factor
            self.num_epochs_no_improvement = 0
        return value_to_reduce


class ImprovementChecker:
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
        return False if self.has_not_improved_counter == self.patience else True
