"""SymbolicRegression — fitness = -error on a fixed dataset
(reference: src/evogp/problem/symbolic_regression.py:9-96)."""
from typing import Callable, Optional

import torch
from torch import Tensor

from .. import _native
from ..tree import Forest
from .base import BaseProblem

_MODES = ("torch", "hybrid parallel", "data parallel", "tree parallel", "auto")


class SymbolicRegression(BaseProblem):
    def __init__(self, datapoints: Optional[Tensor] = None, labels: Optional[Tensor] = None,
                 func: Optional[Callable] = None, num_inputs: Optional[int] = None, num_data: Optional[int] = 100,
                 lower_bounds=-1, upper_bounds=1, execute_mode: str = "auto"):
        assert execute_mode in _MODES, f"execute_mode should be one of {list(_MODES)}, but got {execute_mode}"
        self.execute_mode = execute_mode
        if datapoints is not None and labels is not None:
            self.datapoints, self.labels = datapoints, labels
            return
        assert func is not None and num_inputs is not None, \
            "func and num_inputs, must be provided when datapoints and labels are not provided"
        self.datapoints, self.labels = self.generate_data(func, num_inputs, num_data, lower_bounds, upper_bounds)

    def generate_data(self, func, num_inputs, num_data, lower_bounds, upper_bounds):
        dev = _native.device()

        def bound(b):
            if isinstance(b, (int, float)):
                return torch.full((num_inputs,), float(b), device=dev)
            return torch.as_tensor(b, dtype=torch.float32, device=dev)

        lo, hi = bound(lower_bounds)[None, :], bound(upper_bounds)[None, :]
        inputs = torch.rand(num_data, num_inputs, device=dev) * (hi - lo) + lo
        return inputs, torch.vmap(func)(inputs)

    def evaluate(self, forest: Forest, use_MSE: bool = True):
        if self.execute_mode == "torch":
            pred = forest.batch_forward(self.datapoints)   # [P, N, O]
            err = pred - self.labels[None, :, :]
            err = err**2 if use_MSE else err.abs()
            return -torch.mean(err, dim=(1, 2))            # note: divides by N*O (reference :80)
        return -forest.SR_fitness(self.datapoints, self.labels, use_MSE, self.execute_mode)

    @property
    def problem_dim(self):
        return self.datapoints.shape[1]

    @property
    def solution_dim(self):
        return self.labels.shape[1]
