"""GeneticProgramming — one generation = select, cross over, mutate, re-assemble
(reference: src/evogp/algorithm/genetic_programming.py:8-120)."""
import torch
from torch import Tensor

from ..tree import Forest
from .crossover import BaseCrossover
from .mutation import BaseMutation
from .selection import BaseSelection


class ParetoFront:
    """Best fitness seen for every tree size, and the tree that achieved it."""

    def __init__(self, size, forest_descriptor):
        self.solution = Forest.zero_generate(size, *forest_descriptor)
        self.fitness = torch.full((size,), float("-inf"), dtype=torch.float32, device=self.solution.batch_node_value.device)

    def update(self, fitness: Tensor, forest: Forest):
        """Segmented arg-max of fitness by tree size, then a masked row replace."""
        L = forest.max_tree_len
        raw = forest.batch_subtree_size[:, 0].long()
        sizes = raw.clamp(0, L - 1)
        # slots are indexed by size 0..L-1 (a full-width tree has no slot, as in the reference)
        fit = torch.where(torch.isnan(fitness) | (raw < 0) | (raw >= L), torch.full_like(fitness, float("-inf")), fitness)
        best = torch.full((L,), float("-inf"), dtype=fit.dtype, device=fit.device)
        best.scatter_reduce_(0, sizes, fit, reduce="amax", include_self=True)
        # first individual reaching the per-size best
        hit = fit == best[sizes]
        idx = torch.where(hit, torch.arange(fit.shape[0], device=fit.device), fit.shape[0])
        first = torch.full((L,), fit.shape[0], dtype=torch.long, device=fit.device)
        first.scatter_reduce_(0, sizes, idx, reduce="amin", include_self=True)
        better = (best > self.fitness) & (first < fit.shape[0])
        if not bool(better.any()):
            return
        src = first.clamp(max=fit.shape[0] - 1)
        self.fitness = torch.where(better, best, self.fitness)
        sol = self.solution
        for name in ("batch_node_value", "batch_node_type", "batch_subtree_size"):
            setattr(sol, name, torch.where(better.unsqueeze(1), getattr(forest, name)[src], getattr(sol, name)))

    def __str__(self):
        return "\n".join(f"size: {i}, fitness: {float(self.fitness[i]):.2e}, solution: {self.solution[i]}"
                         for i in range(len(self.fitness)))

    def __repr__(self):
        return repr(self.fitness) + repr(self.solution)


class GeneticProgramming:
    def __init__(self, initial_forest: Forest, crossover: BaseCrossover, mutation: BaseMutation,
                 selection: BaseSelection, enable_pareto_front: bool = False):
        self.forest = initial_forest
        self.pop_size = initial_forest.pop_size
        self.crossover = crossover
        self.mutation = mutation
        self.selection = selection
        self.enable_pareto_front = enable_pareto_front
        if enable_pareto_front:
            f = self.forest
            self.pareto_front = ParetoFront(f.max_tree_len, (f.max_tree_len, f.input_len, f.output_len))

    def step(self, fitness: Tensor) -> Forest:
        assert self.forest is not None, "forest is not initialized"
        assert fitness.shape == (self.forest.pop_size,), \
            f"fitness shape should be ({self.forest.pop_size}, ), but got {fitness.shape}"
        if self.enable_pareto_front:
            self.pareto_front.update(fitness, self.forest)
        elite_indices, survivor_indices = self.selection(self.forest, fitness)
        children = self.crossover(forest=self.forest, survivor_indices=survivor_indices,
                                  target_cnt=self.pop_size - elite_indices.shape[0], fitness=fitness)
        children = self.mutation(children)
        self.forest = self.forest[elite_indices] + children
        return self.forest
