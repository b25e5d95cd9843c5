"""Probe: does torch symmetric memory (peer-mapped buffers over NVLink) work on this box?  Run under torchrun."""
import os
import time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 20
t = symm.empty(n, dtype=torch.float32, device=torch.device("cuda", local))
h = symm.rendezvous(t, dist.group.WORLD)
print(rank, "rendezvous ok; multicast:", h.has_multicast_support, "ptrs", [hex(p) for p in h.buffer_ptrs][:4], flush=True)
t.zero_()
h.barrier(channel=0)
shard = n // world
for p in range(world):
    peer = h.get_buffer(p, (n,), torch.float32)
    peer[rank * shard:(rank + 1) * shard] = float(rank + 1)
h.barrier(channel=0)
torch.cuda.synchronize()
want = torch.cat([torch.full((shard,), float(r + 1)) for r in range(world)])
print(rank, "peer writes visible:", bool(torch.equal(t.cpu(), want)), flush=True)
# timing: symm barrier vs NCCL all_gather of the same payload
x = torch.ones(shard, device="cuda"); out = torch.empty(n, device="cuda")
for name, fn in (("symm barrier", lambda: h.barrier(channel=0)), ("nccl all_gather", lambda: dist.all_gather_into_tensor(out, x))):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): fn()
    e1.record(); torch.cuda.synchronize()
    print(rank, name, "%.1f us" % (e0.elapsed_time(e1) * 1000 / 200), flush=True)
dist.destroy_process_group()
