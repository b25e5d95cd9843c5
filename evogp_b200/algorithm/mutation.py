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


def vmap_subtree(forest: Forest, pos: torch.Tensor) -> Forest:
    """Row n -> the subtree of tree n rooted at pos[n], moved to the front (reference:
    src/evogp/algorithm/mutation/mutation_utils.py:6-48, built there from gather / where over three tensors; here one
    kernel, `tree_extract_subtree`)."""
    dev = forest.batch_node_value.device
    pos = pos.to(dev).reshape(-1).to(torch.int32).contiguous()
    v, t, s = torch.ops.evogp_cuda.tree_extract_subtree(forest.pop_size, forest.max_tree_len, forest.batch_node_value.contiguous(),
                                                        forest.batch_node_type.contiguous(), forest.batch_subtree_size.contiguous(), pos)
    return Forest(forest.input_len, forest.output_len, v, t, s)


def _randint_below(high: torch.Tensor) -> torch.Tensor:
    """floor(U[0,1) * high) per element (reference tree/utils.py:306-310 with low = 0)."""
    return (torch.rand(high.shape, device=high.device) * high).to(torch.int64)


class HoistMutation(BaseMutation):
    """With probability `mutation_rate` a random subtree is replaced by one of its own subtrees, which can only shrink
    the tree (reference: src/evogp/algorithm/mutation/hoist.py:10-77)."""

    def __init__(self, mutation_rate: float):
        self.mutation_rate = mutation_rate

    def __call__(self, forest: Forest):
        dev = forest.batch_node_value.device
        chosen = torch.rand(forest.pop_size).to(dev, non_blocking=True) < self.mutation_rate   # CPU generator, as the reference
        rows = chosen.nonzero(as_tuple=True)[0]
        if rows.shape[0] == 0:
            return forest
        sub = forest[rows]
        at = _randint_below(sub.batch_subtree_size[:, 0])                       # the subtree to shrink ...
        inner = _randint_below(sub.batch_subtree_size.gather(1, at[:, None])[:, 0])   # ... and the part of it that survives
        forest[rows] = sub.mutate(at.to(torch.int32), vmap_subtree(sub, at + inner))
        return forest


class DeleteMutation(BaseMutation):
    """With probability `mutation_rate` a random function node is replaced by one of its children (reference:
    src/evogp/algorithm/mutation/delete.py:10-107)."""

    def __init__(self, mutation_rate: float, max_mutatable_size: int = None):
        self.mutation_rate = mutation_rate
        self.max_mutatable_size = max_mutatable_size

    def __call__(self, forest: Forest):
        dev = forest.batch_node_value.device
        chosen = (torch.rand(forest.pop_size, device=dev) < self.mutation_rate) & (forest.batch_subtree_size[:, 0] > 1)
        rows = chosen.nonzero(as_tuple=True)[0]
        if rows.shape[0] == 0:
            return forest
        sub = forest[rows]
        size = sub.batch_subtree_size
        # a uniformly random non-leaf position: arg-max of random keys over the eligible slots
        keys = torch.rand(size.shape, device=dev)
        eligible = (torch.arange(size.shape[1], device=dev)[None, :] < size[:, :1]) & (size != 1)
        if self.max_mutatable_size:
            eligible &= size <= self.max_mutatable_size
        at = torch.argmax(torch.where(eligible, keys, torch.zeros_like(keys)), 1)
        arity = (sub.batch_node_type.gather(1, at[:, None])[:, 0].long() & 0x7F) - 1
        nth = 1 + _randint_below(arity.clamp(min=1))                           # which child takes the father's place
        c1 = at + 1
        c2 = c1 + size.gather(1, c1.clamp(max=size.shape[1] - 1)[:, None])[:, 0].long()
        c3 = c2 + size.gather(1, c2.clamp(max=size.shape[1] - 1)[:, None])[:, 0].long()
        child = torch.where(nth == 3, c3, torch.where(nth == 2, c2, c1))
        forest[rows] = sub.mutate(at.to(torch.int32), vmap_subtree(sub, child))
        return forest
