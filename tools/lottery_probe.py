#!/usr/bin/env python
"""Does the speed of the replay loop depend on where the module lands in device memory?  Loads N private copies of the
kernel library into ONE process (each copy registers its own CUDA module and is loaded at its own address) and times the
same evaluation through each (developer tool; see profiles/README.md "placement")."""
import ctypes, json, os, shutil, sys, tempfile
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from evogp_b200 import _native
from evogp_b200.tree import Forest, GenerateDescriptor

n_copies = int(sys.argv[1]) if len(sys.argv) > 1 else 8
src = os.environ.get("EVOGP_B200_LIB") or os.path.join(ROOT, "evogp_b200", "lib", "libevogp_b200.so")
dev = torch.device("cuda", 0)
w = bench.WORKLOADS[2]
X, y = bench.dataset(w, dev)
d = GenerateDescriptor(**bench.descriptor_args(w))
P = w["pop"]
pops = [Forest.generate_with_keys(P, d, bench.keys_for(r, dev)) for r in range(2)]
vp = lambda t: ctypes.c_void_p(t.data_ptr())
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
tmp = tempfile.mkdtemp()
out = []
pad = []
for c in range(n_copies):
    path = os.path.join(tmp, f"libcopy{c}.so")
    shutil.copy(src, path)
    L = ctypes.CDLL(path)
    L.evogp_eval_workspace_bytes.restype = ctypes.c_size_t
    L.evogp_eval_workspace_bytes.argtypes = [ctypes.c_uint, ctypes.c_uint]
    ws_bytes = L.evogp_eval_workspace_bytes(P, w["L"])
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    fit = torch.empty(P, dtype=torch.float32, device=dev)
    L.evogp_SR_fitness.restype = ctypes.c_int
    L.evogp_eval_set_timing_events.restype = None
    L.evogp_eval_set_timing_events.argtypes = [ctypes.c_void_p, ctypes.c_void_p]

    def run(f):
        rc = L.evogp_SR_fitness(ctypes.c_uint(P), ctypes.c_uint(w["N"]), ctypes.c_uint(w["L"]), ctypes.c_uint(w["V"]), ctypes.c_uint(w["O"]), 1,
                                vp(f.batch_node_value), vp(f.batch_node_type), vp(f.batch_subtree_size), vp(X), vp(y), vp(fit), ctypes.c_int(4),
                                vp(ws), ctypes.c_size_t(ws_bytes), stream)
        assert rc == 0
    for i in range(3):
        run(pops[i % 2])
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    for a, b in evs:
        a.record(); b.record()
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(evs):
        L.evogp_eval_set_timing_events(ctypes.c_void_p(a.cuda_event), ctypes.c_void_p(b.cuda_event))
        run(pops[i % 2])
    torch.cuda.synchronize()
    L.evogp_eval_set_timing_events(None, None)
    out.append(round(float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e3, 1))
    pad.append(torch.empty((c + 1) * 1234567, dtype=torch.uint8, device=dev))     # perturb the allocator between loads
print(json.dumps({"lib": src[-40:], "replay_us_per_copy": out, "digest": float(torch.nan_to_num(fit, nan=0.0, posinf=0.0, neginf=0.0).clamp(max=1e6).double().sum())}))
shutil.rmtree(tmp, ignore_errors=True)
