#!/usr/bin/env python
"""Device-side timing of the evaluation step at a BASELINE config, lowering and replay apart (developer tool).
    python tools/time_eval.py [config 2|3] [reps]   (EVOGP_B200_LIB=<variant .so> selects another build;
    TIME_FUNCS="if,+,pow,..." / TIME_LAYERS=n replace the config's function set / generation depth)"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from evogp_b200 import _native  # noqa: E402


def main():
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    from evogp_b200.tree import Forest, GenerateDescriptor
    w = bench.WORKLOADS[cfg]
    dev = torch.device("cuda", 0)
    X, y = bench.dataset(w, dev)
    da = bench.descriptor_args(w)
    if os.environ.get("TIME_FUNCS"):
        da["using_funcs"] = os.environ["TIME_FUNCS"].split(",")
    if os.environ.get("TIME_LAYERS"):
        da["max_layer_cnt"] = int(os.environ["TIME_LAYERS"])
    d = GenerateDescriptor(**da)
    P = w["pop"] if cfg == 2 else 125000
    pops = [Forest.generate_with_keys(P, d, bench.keys_for(r, dev)) for r in range(4)]
    abi = _native.abi()
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    ws_bytes = abi.evogp_eval_workspace_bytes(P, w["L"])
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    out = torch.empty(P, dtype=torch.float32, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def fitness(f):   # straight through the C ABI of the library under test (EVOGP_B200_LIB), not the torch op
        rc = abi.evogp_SR_fitness(P, w["N"], w["L"], w["V"], w["O"], 1, vp(f.batch_node_value), vp(f.batch_node_type),
                                  vp(f.batch_subtree_size), vp(X), vp(y), vp(out), 4, vp(ws), ctypes.c_size_t(ws_bytes), stream)
        _native.check(rc, "evogp_SR_fitness")
        return out
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in kev:
        a.record(); b.record()
    for i in range(5):
        fitness(pops[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        abi.evogp_eval_set_timing_events(ctypes.c_void_p(kev[i][0].cuda_event), ctypes.c_void_p(kev[i][1].cuda_event))
        fit = fitness(pops[i % 4])
    e1.record()
    torch.cuda.synchronize()
    abi.evogp_eval_set_timing_events(None, None)
    step = e0.elapsed_time(e1) / reps
    rep = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(0)
        for i in range(20):
            fitness(pops[i % 4])
        clock = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)     # while the queue is still busy
        torch.cuda.synchronize()
    except Exception:
        clock = None
    f = torch.nan_to_num(fit, nan=0.0, posinf=0.0, neginf=0.0).clamp(max=1e6)
    print(json.dumps({"lib": os.environ.get("EVOGP_B200_LIB", "default"), "config": cfg, "pop": P, "step_us": step * 1e3, "replay_us": rep * 1e3,
                      "lower_us": (step - rep) * 1e3, "sm_mhz": clock, "tree_evals_per_s": P * w["N"] / (step * 1e-3),
                      "fitness_digest": float(f.double().sum()),
                      "mean_len": float(pops[0].batch_subtree_size[:, 0].float().mean()), "funcs": os.environ.get("TIME_FUNCS", "config"), "env": {k: v for k, v in os.environ.items() if k.startswith("EVOGP_")}}))


if __name__ == "__main__":
    main()
