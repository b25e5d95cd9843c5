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
        # the reference draws these uniforms on the CPU generator (default.py:43); keep that stream, but do the
        # comparison on the GPU: the CPU elementwise kernels are multi-threaded and their wake-up costs up to
        # ~10 ms per call on a many-core host (measured on the B200 box), the H2D of P floats costs ~30 us
        chosen = torch.rand(forest.pop_size).to(dev, non_blocking=True) < self.mutation_rate
        rows = chosen.nonzero(as_tuple=True)[0]
        count = rows.shape[0]
        if count == 0:
            return forest
        mutants = forest[rows]
        donors = Forest.random_generate(pop_size=count, descriptor=self.descriptor)
        raw = torch.randint(low=0, high=MAX_STACK, size=(count,), dtype=torch.int32, device=dev)
        positions = raw % mutants.batch_subtree_size[:, 0]
        forest[rows] = mutants.mutate(positions, donors)
        return forest
