#!/usr/bin/env python
"""Runs each HBM-bound genetic kernel a few times at BASELINE config-5 sizes (for ncu launch lists / captures)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from evogp_b200.tree import Forest, GenerateDescriptor
dev = torch.device("cuda", 0)
c = bench.CONFIG5
d = GenerateDescriptor(**bench.descriptor_args(c)); dm = d.update(max_layer_cnt=3)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for r in range(reps):
    f = Forest.generate_with_keys(100000, GenerateDescriptor(**bench.descriptor_args(bench.WORKLOADS[2])), bench.keys_for(r, dev))
    donors = Forest.generate_with_keys(99000, dm, bench.keys_for(r, dev))
pop = Forest.generate_with_keys(c["pop"], d, bench.keys_for(6, dev))
surv = bench.contiguous_forest(Forest, pop, 0, 150000)
g = torch.Generator(device=dev).manual_seed(3)
n_new = 495000
li = torch.randint(0, 150000, (n_new,), dtype=torch.int32, device=dev, generator=g); ri = torch.randint(0, 150000, (n_new,), dtype=torch.int32, device=dev, generator=g)
sizes = surv.batch_subtree_size[:, 0].int()
lp = torch.randint(0, 2**31 - 1, (n_new,), dtype=torch.int32, device=dev, generator=g) % sizes[li.long()]
rp = torch.randint(0, 2**31 - 1, (n_new,), dtype=torch.int32, device=dev, generator=g) % sizes[ri.long()]
for r in range(reps):
    child = surv.crossover(li, ri, lp, rp)
mut = bench.contiguous_forest(Forest, child, 0, 99000)
pos = torch.randint(0, 1024, (99000,), dtype=torch.int32, device=dev, generator=g) % mut.batch_subtree_size[:, 0].int()
for r in range(reps):
    res = mut.mutate(pos, donors)
order = torch.sort(torch.rand(c["pop"], device=dev, generator=g), descending=True, stable=True).indices
for r in range(reps):
    out = torch.ops.evogp_cuda.tree_next_generation(c["pop"], 64, pop.batch_node_value, pop.batch_node_type, pop.batch_subtree_size, order,
                                                    5000, 150000, 0.2, 10, 1, dm.out_prob, dm.const_prob, dm.depth2leaf_probs, dm.roulette_funcs, dm.const_samples, bench.keys_for(8, dev))
torch.cuda.synchronize()
print("ok")
