#!/usr/bin/env python
"""Multi-GPU check, spawned by tests/test_multi_gpu.py (-m gpu, >= 2 GPUs) or by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Verifies, over NCCL: (1) sharded SR fitness + one all-gather == the full evaluation, bit for bit;
(2) populations stay bit-identical on every rank across generations without exchanging trees."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from evogp_b200.algorithm import DefaultCrossover, DefaultMutation, DefaultSelection, GeneticProgramming
    from evogp_b200.parallel import ShardedSymbolicRegression, seed_all
    from evogp_b200.problem import SymbolicRegression
    from evogp_b200.tree import Forest, GenerateDescriptor

    rank, world = dist.get_rank(), dist.get_world_size()
    seed_all(0)
    X = torch.rand(1024, 10, device="cuda") * 2 - 1
    y = (X[:, :1] ** 2 + X[:, 1:2] * X[:, 2:3]).contiguous()
    desc = GenerateDescriptor(max_tree_len=64, input_len=10, output_len=1, using_funcs=["+", "-", "*", "/"],
                              max_layer_cnt=6, const_samples=[-1, 0, 1])
    forest = Forest.random_generate(100003, desc)            # odd size: ragged last shard
    base = SymbolicRegression(datapoints=X, labels=y)
    sharded = ShardedSymbolicRegression(base)
    full = base.evaluate(forest)
    got = sharded.evaluate(forest)
    assert torch.equal(torch.nan_to_num(full, nan=-1.0), torch.nan_to_num(got, nan=-1.0)), "sharded != full"
    fused = sharded._exchange is not None and sharded._exchange.available
    why = "" if fused else (sharded._exchange.why if sharded._exchange is not None else "not attempted")
    # the NCCL all-gather form must give the same answer
    got2 = ShardedSymbolicRegression(base, fused_exchange=False).evaluate(forest)
    assert torch.equal(torch.nan_to_num(full, nan=-1.0), torch.nan_to_num(got2, nan=-1.0)), "sharded (all-gather) != full"
    for _ in range(5):     # alternating buffers, back to back
        again = sharded.evaluate(forest)
        assert torch.equal(torch.nan_to_num(full, nan=-1.0), torch.nan_to_num(again, nan=-1.0)), "repeat != full"
    algo = GeneticProgramming(forest, DefaultCrossover(), DefaultMutation(0.2, desc.update(max_layer_cnt=3)),
                              DefaultSelection(survival_rate=0.3, elite_rate=0.01))
    for gen in range(3):
        fit = sharded.evaluate(algo.forest)
        fit = torch.where(torch.isnan(fit), torch.full_like(fit, float("-inf")), fit)
        algo.step(fit)
        f = algo.forest
        lens = f.batch_subtree_size[:, 0].long()
        cols = torch.arange(f.max_tree_len, device="cuda")[None, :]
        valid = cols < lens[:, None]
        digest = torch.stack([(f.batch_node_value.view(torch.int32).long() * valid).sum(),
                              (f.batch_node_type.long() * valid * (cols + 1)).sum(),
                              (f.batch_subtree_size.long() * valid * (cols + 3)).sum()])
        gathered = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        assert all(torch.equal(g, gathered[0]) for g in gathered), f"generation {gen}: populations diverged"
    # the fused generation step (one kernel, Philox draws keyed by the generation's keys): replicated populations stay
    # identical on every rank, evaluation sharded with the fitness exchange fused into the kernel
    from evogp_b200.algorithm import FusedGeneticProgramming
    seed_all(1)
    fused_algo = FusedGeneticProgramming(Forest.random_generate(60001, desc), desc.update(max_layer_cnt=3), 0.2, 0.3, elite_rate=0.01)
    sharded2 = ShardedSymbolicRegression(base)
    for gen in range(3):
        fused_algo.step(sharded2.evaluate(fused_algo.forest))
        f = fused_algo.forest
        digest = torch.stack([f.batch_node_value.view(torch.int32).long().sum(), f.batch_node_type.long().sum(),
                              f.batch_subtree_size.long().sum()])
        gathered = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(gathered, digest)
        assert all(torch.equal(g, gathered[0]) for g in gathered), f"fused generation {gen}: populations diverged"
    dist.barrier()
    if rank == 0:
        print(f"multi-gpu check ok on {world} ranks: sharded fitness bit-identical, populations identical for 3 generations; "
              f"fitness exchange over peer-mapped memory: {fused} ({getattr(sharded._exchange, 'mode', '-')}) {why}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
