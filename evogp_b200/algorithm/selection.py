"""Selection operators.  DefaultSelection: truncation with elitism
(reference: src/evogp/algorithm/selection/default.py:15-71)."""
from typing import Optional

import torch

from ..tree import Forest


class BaseSelection:
    def __call__(self, forest: Forest, fitness: torch.Tensor):
        raise NotImplementedError


class DefaultSelection(BaseSelection):
    """Keep the best `elite` individuals unchanged and let the best `survival_rate` fraction breed."""

    def __init__(self, survival_rate: float = 0.3, elite_cnt: Optional[int] = None, elite_rate: Optional[float] = None):
        assert 0 <= survival_rate <= 1, "survival_rate should be in [0, 1]"
        assert elite_cnt is None or elite_rate is None, "elite_cnt and elite_rate should not be set at the same time"
        self.survival_rate = survival_rate
        self.elite_cnt = elite_cnt
        self.elite_rate = elite_rate

    def counts(self, pop_size):
        survive = int(pop_size * self.survival_rate)
        if self.elite_cnt is not None:
            elite = self.elite_cnt
        elif self.elite_rate is not None:
            elite = int(pop_size * self.elite_rate)
        else:
            elite = 0
        return elite, survive

    def __call__(self, forest: Forest, fitness: torch.Tensor):
        elite, survive = self.counts(forest.pop_size)
        order = torch.sort(fitness, descending=True).indices
        return order[:elite].to(torch.int32), order[:survive].to(torch.int32)
