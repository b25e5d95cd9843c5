"""Mutation operators.  DefaultMutation: with probability `mutation_rate` replace a random
subtree by a freshly generated one (reference: src/evogp/algorithm/mutation/default.py:19-75)."""
import torch

from ..tree import Forest, MAX_STACK, GenerateDescriptor


class BaseMutation:
    def __call__(self, forest: Forest):
        raise NotImplementedError


class DefaultMutation(BaseMutation):
    def __init__(self, mutation_rate: float, descriptor: GenerateDescriptor):
        self.mutation_rate = mutation_rate
        self.descriptor = descriptor

    def __call__(self, forest: Forest):
        dev = forest.batch_node_value.device
        # the reference draws this mask on the CPU generator (default.py:43); keep that stream
        chosen = torch.rand(forest.pop_size) < self.mutation_rate
        count = int(chosen.sum())
        if count == 0:
            return forest
        rows = chosen.nonzero(as_tuple=True)[0].to(dev)
        mutants = forest[rows]
        donors = Forest.random_generate(pop_size=count, descriptor=self.descriptor)
        raw = torch.randint(low=0, high=MAX_STACK, size=(count,), dtype=torch.int32, device=dev)
        positions = raw % mutants.batch_subtree_size[:, 0]
        forest[rows] = mutants.mutate(positions, donors)
        return forest
