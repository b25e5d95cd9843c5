"""Multi-GPU evaluation: one process per GPU (torchrun), population REPLICATED on every
rank, fitness evaluation SHARDED by rows, one all-gather of fitness scalars per generation.

Why replicate: parents of the next generation come from the global top 30 %, so sharding the
genetic operators would need an all-gather of P*L*8 bytes per generation; recomputing them on
every rank from the same seeds costs less and keeps the collective to 4 bytes per individual.
Populations stay bit-identical across ranks because every genetic kernel is a deterministic
integer move and every rank consumes identical RNG streams (see `seed_all`).
"""
import torch
import torch.distributed as dist


def shard_bounds(pop_size: int, world_size: int, rank: int):
    """Rows [lo, hi) evaluated by `rank`; shards are ceil(P/W) rows, the last one ragged."""
    per = (pop_size + world_size - 1) // world_size
    lo = min(rank * per, pop_size)
    return lo, min(lo + per, pop_size), per


def all_gather_fitness(local: torch.Tensor, pop_size: int, group=None) -> torch.Tensor:
    """local: this rank's fitness slice (length hi-lo) -> full fitness [pop_size] on every rank.
    Exactly one collective (all_gather_into_tensor; NCCL on GPUs, gloo in the CPU tests)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi, per = shard_bounds(pop_size, world, rank)
    assert local.shape == (hi - lo,), f"rank {rank}: expected {(hi - lo,)} fitness values, got {tuple(local.shape)}"
    if hi - lo < per:
        pad = torch.zeros(per - (hi - lo), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad])
    full = torch.empty(world * per, dtype=local.dtype, device=local.device)
    if dist.get_backend(group) == "gloo":
        chunks = list(full.chunk(world))
        dist.all_gather(chunks, local.contiguous(), group=group)
    else:
        dist.all_gather_into_tensor(full, local.contiguous(), group=group)
    return full[:pop_size]


def seed_all(seed: int):
    """Identical CPU + CUDA generator state on every rank (the GP loop draws from both)."""
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


class ShardedSymbolicRegression:
    """Wraps a SymbolicRegression problem: evaluate() runs SR_fitness on this rank's row shard
    and all-gathers.  Drop-in for `problem` in StandardPipeline / GeneticProgramming loops."""

    def __init__(self, problem, group=None):
        self.problem = problem
        self.group = group

    def evaluate(self, forest, use_MSE: bool = True):
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        lo, hi, _ = shard_bounds(forest.pop_size, world, rank)
        if hi > lo:
            local = self.problem.evaluate(forest[lo:hi], use_MSE)
        else:
            local = torch.empty(0, dtype=torch.float32, device=forest.batch_node_value.device)
        return all_gather_fitness(local, forest.pop_size, self.group)

    @property
    def problem_dim(self):
        return self.problem.problem_dim

    @property
    def solution_dim(self):
        return self.problem.solution_dim
