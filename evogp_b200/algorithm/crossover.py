"""Crossover operators.  DefaultCrossover: random recipient/donor pairs among the survivors,
random subtree positions (reference: src/evogp/algorithm/crossover/default.py:13-66)."""
import torch

from ..tree import Forest


class BaseCrossover:
    def __call__(self, forest: Forest, survivor_indices: torch.Tensor, target_cnt: int, fitness: torch.Tensor):
        raise NotImplementedError


class DefaultCrossover(BaseCrossover):
    def __call__(self, forest: Forest, survivor_indices: torch.Tensor, target_cnt: int, fitness: torch.Tensor):
        parents = forest[survivor_indices]
        dev = parents.batch_node_value.device
        # draw order and ranges follow the reference (default.py:40-61) so seeded runs agree
        pair = torch.randint(low=0, high=len(parents), size=(2, target_cnt), dtype=torch.int32, device=dev)
        raw = torch.randint(low=0, high=torch.iinfo(torch.int32).max, size=(2, target_cnt), dtype=torch.int32, device=dev)
        sizes = parents.batch_subtree_size[:, 0]
        left, right = pair[0], pair[1]
        left_pos = raw[0] % sizes[left]
        right_pos = raw[1] % sizes[right]
        return parents.crossover(left, right, left_pos, right_pos)
