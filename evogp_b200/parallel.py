"""Multi-GPU evaluation: one process per GPU (torchrun), population REPLICATED on every
rank, fitness evaluation SHARDED by rows, one all-gather of fitness scalars per generation.

Why replicate: parents of the next generation come from the global top 30 %, so sharding the
genetic operators would need an all-gather of P*L*8 bytes per generation; recomputing them on
every rank from the same seeds costs less and keeps the collective to 4 bytes per individual.
Populations stay bit-identical across ranks because every genetic kernel is a deterministic
integer move and every rank consumes identical RNG streams (see `seed_all`).

The exchange itself comes in two forms: `all_gather_fitness` (one NCCL all-gather; gloo in the CPU
tests) and `FitnessExchange` (NVLink boxes): every rank's fitness slice is written straight into
each rank's full-population buffer through peer-mapped symmetric memory - by a small kernel of its
own after the evaluation (`evogp_push_fitness`, the default) or by the evaluation kernel itself
(`evogp_SR_fitness_scatter`, EVOGP_EXCHANGE=fused) - and a ~7 us inter-GPU barrier follows it.
"""
import ctypes
import os

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


class FitnessExchange:
    """Fitness all-gather over peer-mapped memory: own kernels, no NCCL collective on the data path.

    Two full-population fitness buffers in torch symmetric memory (peer-mapped over NVLink / NVSwitch), used
    alternately: the kernel of generation g stores into buffer g % 2 of every rank while a slower rank may still be
    reading generation g - 1 from the other one; the barrier after each kernel keeps ranks at most one generation
    apart, so two buffers are enough.  `available` is False (reason in `why`) when symmetric memory cannot be set up
    (gloo / CPU, no peer access); callers then use `all_gather_fitness`."""

    def __init__(self, pop_size: int, device, group=None):
        self.pop_size, self.group = pop_size, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.lo, self.hi, _ = shard_bounds(pop_size, self.world, self.rank)
        self.available, self.why, self._turn = False, "", 0
        self._bufs, self._handles = [], []
        if self.world > 32:
            self.why = "more than 32 ranks"
            return
        # EVOGP_EXCHANGE = push (default): plain evaluation kernel, then one small kernel that copies this rank's slice into
        # every rank's buffer over peer memory; fused: the evaluation kernel itself pushes finished chunks
        # (evogp_SR_fitness_scatter); nccl: one all-gather.  Measured at 2 x B200, 500000 trees per rank: the fused
        # form costs +0.21 ms per step (fence + counter per tree, and the protocol's code next to the replay loop in the
        # instruction cache), the push kernel ~0.01 ms.
        self.mode = os.environ.get("EVOGP_EXCHANGE", "push")
        if os.environ.get("EVOGP_FUSED_EXCHANGE", "1") == "0" or self.mode == "nccl":
            self.why = "disabled by EVOGP_EXCHANGE=nccl"
            return
        try:
            if dist.get_backend(group) != "nccl":
                raise RuntimeError("backend is not nccl")
            import torch.distributed._symmetric_memory as symm

            for _ in range(2):
                buf = symm.empty(pop_size, dtype=torch.float32, device=device)
                self._handles.append(symm.rendezvous(buf, group if group is not None else dist.group.WORLD))
                self._bufs.append(buf)
                buf.zero_()
            self._handles[0].barrier(channel=0)
            self.available = True
        except Exception as e:   # no symmetric memory here: fall back to the NCCL all-gather
            self.why = f"{type(e).__name__}: {e}"
            self._bufs, self._handles = [], []

    def sr_fitness(self, shard, datapoints, labels, use_MSE: bool = True):
        """shard: this rank's rows [lo, hi) as a Forest.  Returns the full-population fitness [pop_size]."""
        from . import _native

        assert shard.pop_size == self.hi - self.lo, "shard does not match this rank's row range"
        if not self.available:
            return all_gather_fitness(shard.SR_fitness(datapoints, labels, use_MSE), self.pop_size, self.group)
        turn = self._turn & 1
        self._turn += 1
        buf, hdl = self._bufs[turn], self._handles[turn]
        if shard.pop_size > 0:
            abi = _native.abi()
            dev = shard.batch_node_value.device
            P, L = shard.batch_node_value.shape
            N, V = datapoints.shape
            O = labels.shape[1] if labels.dim() > 1 else 1
            ws_bytes = abi.evogp_eval_workspace_bytes(P, L)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            local = torch.empty(P, dtype=torch.float32, device=dev)
            vp = lambda t: ctypes.c_void_p(t.data_ptr())
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            if self.mode != "fused":
                rc = abi.evogp_SR_fitness(P, N, L, V, O, 1 if use_MSE else 0, vp(shard.batch_node_value), vp(shard.batch_node_type),
                                          vp(shard.batch_subtree_size), vp(datapoints.contiguous()), vp(labels.contiguous()),
                                          vp(local), 4, vp(ws), ctypes.c_size_t(ws_bytes), stream)
                _native.check(rc, "evogp_SR_fitness")
                rc = abi.evogp_push_fitness(vp(local), P, ctypes.c_void_p(hdl.buffer_ptrs_dev), self.world, self.lo, stream)
                _native.check(rc, "evogp_push_fitness")
                hdl.barrier(channel=0)
                return buf
            rc = abi.evogp_SR_fitness_scatter(P, N, L, V, O, 1 if use_MSE else 0, vp(shard.batch_node_value),
                                              vp(shard.batch_node_type), vp(shard.batch_subtree_size),
                                              vp(datapoints.contiguous()), vp(labels.contiguous()), vp(local),
                                              ctypes.c_void_p(hdl.buffer_ptrs_dev), self.world, self.lo, vp(ws),
                                              ctypes.c_size_t(ws_bytes), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            _native.check(rc, "evogp_SR_fitness_scatter")
        hdl.barrier(channel=0)     # every rank's stores have landed; also orders this generation against the next
        return buf


class ShardedSymbolicRegression:
    """Wraps a SymbolicRegression problem: evaluate() runs SR_fitness on this rank's row shard and exchanges the
    fitness - fused into the kernel over peer memory when symmetric memory is available (`FitnessExchange`), else
    one NCCL / gloo all-gather.  Drop-in for `problem` in StandardPipeline / GeneticProgramming loops."""

    def __init__(self, problem, group=None, fused_exchange: bool = True):
        self.problem = problem
        self.group = group
        self.fused_exchange = fused_exchange
        self._exchange = None

    def evaluate(self, forest, use_MSE: bool = True):
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        lo, hi, _ = shard_bounds(forest.pop_size, world, rank)
        dev = forest.batch_node_value.device
        fused = (self.fused_exchange and dev.type == "cuda" and getattr(self.problem, "execute_mode", "") != "torch"
                 and hasattr(self.problem, "datapoints"))
        if fused:
            if self._exchange is None or self._exchange.pop_size != forest.pop_size:
                self._exchange = FitnessExchange(forest.pop_size, dev, self.group)
            if self._exchange.available:
                return -self._exchange.sr_fitness(forest[lo:hi], self.problem.datapoints, self.problem.labels, use_MSE)
        if hi > lo:
            local = self.problem.evaluate(forest[lo:hi], use_MSE)
        else:
            local = torch.empty(0, dtype=torch.float32, device=dev)
        return all_gather_fitness(local, forest.pop_size, self.group)

    @property
    def problem_dim(self):
        return self.problem.problem_dim

    @property
    def solution_dim(self):
        return self.problem.solution_dim
