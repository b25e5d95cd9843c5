#!/usr/bin/env python
"""Times evogp_generate and evogp_next_generation at the sizes of bench.py's hbm_kernels report (developer tool; run once
per EVOGP_GENERATE_BALANCED / EVOGP_NEXTGEN_PIPELINE setting - the switches are read when the library loads)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from evogp_b200.tree import Forest, GenerateDescriptor
dev = torch.device("cuda", 0)
d2 = GenerateDescriptor(**bench.descriptor_args(bench.WORKLOADS[2]))
d5 = GenerateDescriptor(**bench.descriptor_args(bench.CONFIG5))
dm = d5.update(max_layer_cnt=3)
out = {"EVOGP_GENERATE_BALANCED": os.environ.get("EVOGP_GENERATE_BALANCED", "default")}
for name, pop, d in (("pop100000", 100000, d2), ("donors99000", 99000, dm), ("pop500000", 500000, d5), ("pop1000000", 1000000, d5)):
    for r in range(3):
        f = Forest.generate_with_keys(pop, d, bench.keys_for(r, dev))
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for r in range(10):
        f = Forest.generate_with_keys(pop, d, bench.keys_for(r, dev))
        ev[r + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
    out[name] = {"us_median": ts[5] * 1e3, "us_min": ts[0] * 1e3, "digest": int(f.batch_subtree_size[:, 0].long().sum()),
                 "GBps_written": pop * 64 * 8 / (ts[5] * 1e-3) / 1e9}
c = bench.CONFIG5
pop = Forest.generate_with_keys(c["pop"], d5, bench.keys_for(6, dev))
g = torch.Generator(device=dev).manual_seed(3)
order = torch.sort(torch.rand(c["pop"], device=dev, generator=g), descending=True, stable=True).indices
def nextgen():
    return torch.ops.evogp_cuda.tree_next_generation(c["pop"], 64, pop.batch_node_value, pop.batch_node_type, pop.batch_subtree_size, order,
                                                     5000, 150000, 0.2, 10, 1, dm.out_prob, dm.const_prob, dm.depth2leaf_probs, dm.roulette_funcs,
                                                     dm.const_samples, bench.keys_for(8, dev))
for r in range(3):
    res = nextgen()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
ev[0].record()
for r in range(10):
    res = nextgen()
    ev[r + 1].record()
torch.cuda.synchronize()
ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
out["EVOGP_NEXTGEN_PIPELINE"] = os.environ.get("EVOGP_NEXTGEN_PIPELINE", "default")
out["nextgen_pop500000"] = {"us_median": ts[5] * 1e3, "us_min": ts[0] * 1e3, "digest": int(res[2][:, 0].long().sum())}
print(json.dumps(out))
