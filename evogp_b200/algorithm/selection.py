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


def _counts(pop_size, rate, cnt):
    return cnt if cnt is not None else int(pop_size * rate)


def _elite(fitness, elite_cnt):
    if elite_cnt == 0:
        return torch.empty(0, dtype=torch.int64, device=fitness.device)
    return torch.sort(fitness, descending=True).indices[:elite_cnt]


class TournamentSelection(BaseSelection):
    """Survivors are the winners of `survivor_cnt` tournaments of `tournament_size` contenders; the nth best contender
    wins with probability best_probability * (1 - best_probability)^n (reference:
    src/evogp/algorithm/selection/tournament.py:11-133).  The reference materialises a [k_times, P] matrix of ones for
    torch.multinomial under vmap and argsorts every tournament under vmap; here one kernel runs all tournaments
    (`tree_tournament_select`, counter-based Philox draws keyed by two words taken from the torch CUDA generator; without
    replacement = slices of a keyed pseudo-random permutation per round).  Same distribution, not the same random stream."""

    def __init__(self, tournament_size: int, best_probability: float = 1, replace: bool = True, survivor_rate: float = 0.5,
                 elite_rate: float = 0, survivor_cnt: Optional[int] = None, elite_cnt: Optional[int] = None):
        assert 0 <= survivor_rate <= 1, "survival_rate should be in [0, 1]"
        assert 0 <= elite_rate <= 1, "elite_rate should be in [0, 1]"
        assert 0 < best_probability <= 1, "best_probability should be in (0, 1]"
        self.t_size, self.best_p, self.replace = tournament_size, best_probability, replace
        self.survivor_rate, self.survivor_cnt = survivor_rate, survivor_cnt
        self.elite_rate, self.elite_cnt = elite_rate, elite_cnt

    def __call__(self, forest: Forest, fitness: torch.Tensor, keys: Optional[torch.Tensor] = None):
        survivor_cnt = _counts(forest.pop_size, self.survivor_rate, self.survivor_cnt)
        elite_cnt = _counts(forest.pop_size, self.elite_rate, self.elite_cnt)
        fitness = fitness.to(torch.float32).contiguous()
        if keys is None:
            keys = torch.randint(low=0, high=1000000, size=(2,), dtype=torch.uint32, device=fitness.device)
        survivors = torch.ops.evogp_cuda.tree_tournament_select(fitness, self.t_size, float(self.best_p), bool(self.replace),
                                                                max(survivor_cnt, 1), keys)[:survivor_cnt]
        return _elite(fitness, elite_cnt), survivors


class TruncationSelection(BaseSelection):
    """Uniform draws (with replacement) among the best `survivor_rate` fraction (reference:
    src/evogp/algorithm/selection/truncation.py:8-81)."""

    def __init__(self, survivor_rate: float = 0.5, elite_rate: float = 0, survivor_cnt: Optional[int] = None,
                 elite_cnt: Optional[int] = None):
        assert 0 <= survivor_rate <= 1, "survival_rate should be in [0, 1]"
        assert 0 <= elite_rate <= 1, "elite_rate should be in [0, 1]"
        self.survivor_rate, self.survivor_cnt = survivor_rate, survivor_cnt
        self.elite_rate, self.elite_cnt = elite_rate, elite_cnt

    def __call__(self, forest: Forest, fitness: torch.Tensor):
        survivor_cnt = _counts(forest.pop_size, self.survivor_rate, self.survivor_cnt)
        elite_cnt = _counts(forest.pop_size, self.elite_rate, self.elite_cnt)
        order = torch.sort(fitness, descending=True).indices
        selectable = max(int(forest.pop_size * self.survivor_rate), 1)
        picks = torch.randint(0, selectable, (survivor_cnt,), device=fitness.device)
        return order[:elite_cnt], order[picks]
