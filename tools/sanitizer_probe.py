#!/usr/bin/env python
"""Small invocations of every kernel, for compute-sanitizer (racecheck / memcheck / synccheck) runs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from evogp_b200.tree import Forest, GenerateDescriptor
from evogp_b200.algorithm import FusedGeneticProgramming, TournamentSelection, HoistMutation
from evogp_b200.problem import Classification
dev = torch.device("cuda", 0)
torch.manual_seed(0)
w = bench.WORKLOADS[2]
d = GenerateDescriptor(**bench.descriptor_args(w))
f = Forest.generate_with_keys(3000, d, bench.keys_for(0, dev))                     # generate_fast_kernel
X, y = bench.dataset(w, dev)
fit = f.SR_fitness(X, y)                                                            # lower_fast + replay<16>
fit8 = f.SR_fitness(X[:200].contiguous(), y[:200].contiguous())                     # replay<8, tmem>
algo = FusedGeneticProgramming(f, d.update(max_layer_cnt=3), 0.3, 0.3, elite_rate=0.01)
nxt = algo.step(-torch.nan_to_num(fit, nan=1e30))                                   # nextgen_batch_kernel
dall = GenerateDescriptor(max_tree_len=64, input_len=3, output_len=1, using_funcs=["+", "-", "*", "/", "if", "sin", "pow"], max_layer_cnt=4, const_samples=[-1, 0, 1])
fa = Forest.generate_with_keys(2000, dall, bench.keys_for(1, dev))
fa.SR_fitness(X, y)                                                                 # fallback rows of lower_fast, slow-path opcodes
dm = GenerateDescriptor(max_tree_len=64, input_len=3, output_len=3, using_funcs=["+", "-", "*", "/"], max_layer_cnt=5, const_samples=[-1, 0, 1])
fm = Forest.generate_with_keys(2000, dm, bench.keys_for(2, dev))                    # generate_kernel<multi>
lab = torch.randint(0, 3, (1024,), device=dev).float()
Classification(datapoints=X, labels=lab).evaluate(fm)                               # multi PTX loop + accuracy epilogue
fm.batch_forward(X[:64].contiguous())
sel = TournamentSelection(5, 0.8, replace=False, survivor_rate=0.5)(f, -torch.nan_to_num(fit, nan=1e30))
HoistMutation(0.5)(f)                                                               # extract_subtree + mutate
li = torch.randint(0, 3000, (4000,), dtype=torch.int32, device=dev)
lp = torch.randint(0, 1 << 30, (4000,), dtype=torch.int32, device=dev) % f.batch_subtree_size[li.long(), 0].int()
f.crossover(li, li.flip(0), lp, lp.flip(0) % f.batch_subtree_size[li.flip(0).long(), 0].int())
Forest.generate_with_keys(1000, d, bench.keys_for(3, dev), rng="philox")
torch.cuda.synchronize()
print("sanitizer probe ok", float(torch.nan_to_num(fit, nan=0.0).clamp(max=1e6).mean()))
