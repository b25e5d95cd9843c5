#!/usr/bin/env python
"""bench.py — tree-evals/s of the batched fitness evaluation (BASELINE.json metric).

A step = one SR-fitness pass (lower_kernel + replay_kernel) over this rank's population shard
against the whole dataset, plus — for N > 1 — the single all-gather of fitness scalars.
Workload at N = 1: BASELINE.json configs[1] (synthetic SR, 3 inputs, pop 100000, max_tree_len 64,
1024 datapoints).  For N > 1 every rank holds a shard of that size (weak scaling; the population
is N x 100000).  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(workload="configs[1]: synthetic SR, 3 inputs, pop 100000/GPU, max_tree_len 64, 1024 datapoints, funcs + - * /",
           pop_per_gpu=100000, max_tree_len=64, datapoints=1024, inputs=3, outputs=1,
           funcs=["+", "-", "*", "/"], max_layer_cnt=6, const_samples=[-1.0, 0.0, 1.0], rotating_populations=4)
METRIC = "tree-evals/sec (pop x datapoints)"


def algorithmic_bytes(P, L, N, V, O):
    # node_value + node_type at fixed width, one subtree_size per tree, dataset once, fitness out
    return P * (6 * L + 2) + 4 * N * (V + O) + 4 * P


def target_fn(X):
    # fixed closed-form target (shape of the reference's sr_test.py:17-19)
    return (X[:, 0:1] ** 4 / (X[:, 0:1] ** 4 + 1) + X[:, 1:2] ** 4 / (X[:, 1:2] ** 4 + 1)) + 0.0 * X[:, 2:3]


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.active, self.ready = False, threading.Event()   # NVML is initialised before the timed region; samples only inside it
        self.force = False    # one extra sample right after a timed region too short to catch one

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            self.ready.set()
            while not self.stop_flag:
                if not (self.active or self.force):
                    time.sleep(0.0005)
                    continue
                self.force = False
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
                time.sleep(0.002)
        except Exception as e:   # NVML missing: report that rather than fail the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")
            self.ready.set()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_reference_leg(steps, warmup, target_seconds=20.0):
    """The reference has no CPU implementation of this path (torch_wrapper.cu:301-307 registers CUDA only):
    the CPU arm is the oracle's restatement, all host threads, on a bounded sample of the same workload."""
    import oracle

    L, N, V = CFG["max_tree_len"], CFG["datapoints"], CFG["inputs"]
    threads = oracle.max_threads()
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, (N, V)).astype(np.float32)
    y = target_fn(X).astype(np.float32)
    inner = CFG["max_layer_cnt"] - 1
    d2l = np.array([0.2] * inner + [1.0] * (10 - inner), np.float32)
    p = np.zeros(29, np.float32); p[1:5] = 0.25
    roul = np.cumsum(p, dtype=np.float32)
    consts = np.array(CFG["const_samples"], np.float32)

    def forest(n, key):
        return oracle.generate(n, L, V, 1, 0.5, 0.5, np.array([key, 1], np.uint32), d2l, roul, consts, nthreads=threads)

    probe = forest(2048, 0)
    t0 = time.perf_counter(); oracle.sr_fitness(*probe, X, y, nthreads=threads); dt = time.perf_counter() - t0
    # bounded sample: the whole --steps K --warmup W run takes about target_seconds whatever K is (>= 512 trees per step)
    per_step = target_seconds / max(steps + warmup, 1)
    sample = int(min(CFG["pop_per_gpu"], max(512, 2048 * per_step / dt)))
    pops = [forest(sample, k) for k in range(2)]
    for i in range(warmup):
        oracle.sr_fitness(*pops[i % 2], X, y, nthreads=threads)
    t0 = time.perf_counter()
    for i in range(steps):
        oracle.sr_fitness(*pops[i % 2], X, y, nthreads=threads)
    total = time.perf_counter() - t0
    value = sample * N * steps / total
    return value, total / steps * 1e3, threads, f"{sample} of {CFG['pop_per_gpu']} trees x {N} datapoints per step (same generator, same dataset)"


class StdoutGuard:
    """Keeps stdout to the ONE JSON line: while active, file descriptor 1 points at stderr, so banners that native
    libraries write to stdout (NCCL prints its version there) cannot precede the result; emit() writes to the real one."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, threads, sample = cpu_reference_leg(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "tree-evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": dict(CFG),
            "cpu_baseline": {"value": value, "unit": "tree-evals/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "tree-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args, out):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from evogp_b200 import _native
    from evogp_b200.tree import Forest, GenerateDescriptor
    from evogp_b200.parallel import all_gather_fitness, shard_bounds

    dev = torch.device("cuda", local)
    P, L, N, V, O = CFG["pop_per_gpu"], CFG["max_tree_len"], CFG["datapoints"], CFG["inputs"], CFG["outputs"]
    P_total = P * world
    torch.manual_seed(0)
    X = (torch.rand(N, V, device=dev) * 2 - 1).contiguous()
    y = target_fn(X).contiguous()
    desc = GenerateDescriptor(max_tree_len=L, input_len=V, output_len=O, using_funcs=CFG["funcs"],
                              max_layer_cnt=CFG["max_layer_cnt"], const_samples=CFG["const_samples"])
    # rotating populations: this rank's shard of R different populations, so no step finds its inputs in L2
    R = CFG["rotating_populations"]
    lo, hi, _ = shard_bounds(P_total, world, rank)
    pops = []
    for r in range(R):
        keys = torch.tensor([1000 + r, 7], dtype=torch.uint32, device=dev)
        full = Forest.generate_with_keys(P_total, desc, keys)   # replicated population; this rank evaluates [lo, hi)
        pops.append(full[lo:hi])
        pops[-1] = Forest(V, O, pops[-1].batch_node_value.contiguous(), pops[-1].batch_node_type.contiguous(),
                          pops[-1].batch_subtree_size.contiguous())
        del full
    torch.cuda.synchronize()
    mean_len = float(torch.stack([p.batch_subtree_size[:, 0].float().mean() for p in pops]).mean())

    # N > 1: the fitness all-gather is fused into the evaluation kernel (stores into every rank's buffer through
    # peer-mapped symmetric memory, then one inter-GPU barrier); NCCL all-gather when symmetric memory is unavailable
    exch = None
    if world > 1:
        from evogp_b200.parallel import FitnessExchange
        exch = FitnessExchange(P_total, dev)
    exchange_kind = ("none (1 GPU)" if world == 1 else
                     ("fused into the evaluation kernel over peer-mapped memory + barrier" if exch.available
                      else "NCCL all_gather (symmetric memory unavailable: %s)" % exch.why))

    def step(i):
        if exch is not None:
            return exch.sr_fitness(pops[i % R], X, y)
        return pops[i % R].SR_fitness(X, y)

    abi = _native.abi()
    sampler = ClockSampler(local); sampler.start()
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    sampler.ready.wait(timeout=10)
    if world > 1:
        dist.barrier()
    launches0 = _native.launch_count()
    # Device-side timing, no host synchronisation inside the timed region: one event pair around the K steps gives the
    # total, one pair per step (recorded inside the C ABI around the replay launch) the kernel's own duration.
    import ctypes
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in kev:
        a.record(); b.record()     # materialise the handles
    ev_t0 = torch.cuda.Event(enable_timing=True); ev_t1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    sampler.active = True
    ev_t0.record()
    for i in range(args.steps):
        abi.evogp_eval_set_timing_events(ctypes.c_void_p(kev[i][0].cuda_event), ctypes.c_void_p(kev[i][1].cuda_event))
        step(i)
    ev_t1.record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    sampler.active = False
    if not sampler.samples:     # very short runs: read the clocks immediately after the last step
        sampler.force = True
        time.sleep(0.02)
    abi.evogp_eval_set_timing_events(None, None)
    kern_ms = [a.elapsed_time(b) for a, b in kev]
    launches = _native.launch_count() - launches0
    sampler.stop_flag = True; sampler.join(timeout=2)
    step_ms = [ev_t0.elapsed_time(ev_t1) / args.steps] * args.steps
    total_ms = torch.tensor([ev_t0.elapsed_time(ev_t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    total_s = float(total_ms) / 1e3
    value = P_total * N * args.steps / total_s

    # ---- e2e: host buffers through the C ABI (H2D of the forest + dataset, D2H of fitness, every step) ----
    hv = [p.batch_node_value.cpu().pin_memory() for p in pops]
    ht = [p.batch_node_type.cpu().pin_memory() for p in pops]
    hs = [p.batch_subtree_size.cpu().pin_memory() for p in pops]
    hX, hy = X.cpu().pin_memory(), y.cpu().pin_memory()
    hfit = torch.empty(hi - lo, dtype=torch.float32).pin_memory()
    vp = lambda t: ctypes.c_void_p(t.data_ptr())

    def e2e_step(i):
        r = i % R
        rc = abi.evogp_SR_fitness_host(hi - lo, N, L, V, O, 1, vp(hv[r]), vp(ht[r]), vp(hs[r]), vp(hX), vp(hy), vp(hfit), local)
        _native.check(rc, "evogp_SR_fitness_host")

    for i in range(max(args.warmup, 1)):
        e2e_step(i)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = P_total * N * args.steps / float(e2e_t)
    h2d = (hi - lo) * (L * 6 + 2) + N * (V + O) * 4      # value + type rows, one length per tree, dataset
    d2h = (hi - lo) * 4
    fit_host_check = float(np.nanmean(hfit.numpy()))

    line = None
    if rank == 0:
        peak, peak_src = measured_peak()
        kms = float(np.mean(kern_ms))
        ach = algorithmic_bytes(hi - lo, L, N, V, O) / (kms * 1e-3) / 1e9
        clocks = sampler.summary()
        line = {"metric": METRIC, "value": value, "unit": "tree-evals/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": float(np.mean(step_ms)) if world == 1 else total_s * 1e3 / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": dict(CFG, population_total=P_total, mean_tree_len=round(mean_len, 2),
                               l2="inputs rotate over %d populations (%.0f MB + %.0f MB programs) > 126 MB L2"
                                  % (R, R * (hi - lo) * L * 8 / 1e6, (hi - lo) * L * 8 / 1e6),
                               parallelism="population replicated, eval sharded x%d, fitness exchange: %s" % (world, exchange_kind)),
                "e2e": {"value": e2e_value, "unit": "tree-evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "path": "evogp_SR_fitness_host (C ABI, pinned host buffers, chunked copy/compute overlap)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "replay_kernel<16,false,false,true>", "achieved": ach, "peak": peak,
                             "unit": "GB/s", "frac": ach / peak,
                             "traffic": 53.6e6,   # dram read+write per launch, profiles/r1_final_ncu.txt
                             "traffic_source": "ncu --set full, profiles/r1_final_ncu.txt (same workload shard)",
                             "algorithmic_bytes": algorithmic_bytes(hi - lo, L, N, V, O), "peak_source": peak_src,
                             "kernel_ms": kms, "kernel_share_of_step": kms / float(np.mean(step_ms)),
                             "issue_slots_busy_pct": 80.8,   # smsp__issue_active, same capture: the resource this kernel is bound by
                             "note": "interpreter kernel: bound by instruction issue (81 % of issue slots busy, shared-memory pipe 62 %), not HBM (DESIGN.md 3.2)"},
                "clocks": clocks, "wall_s_timed_region": t_wall, "fitness_mean_check": fit_host_check}
        if world == 1 and not args.no_cpu:
            v, ms, threads, sample = cpu_reference_leg(3, 1, target_seconds=10.0)
            line["cpu_baseline"] = {"value": v, "unit": "tree-evals/s", "cores": threads, "kind": "port", "sample": sample}
            if not args.no_ref_gpu:
                try:
                    import oracle
                    if oracle.ref_gpu_available():
                        ref = oracle.ref_gpu()
                        for r in range(3):   # warm-up: the reference kernel's local-memory frames are sized on first use
                            p0 = pops[r % R]
                            ref.sr_fitness(p0.batch_node_value, p0.batch_node_type, p0.batch_subtree_size, X, y)
                        torch.cuda.synchronize()
                        best = float("inf")
                        for r in range(5):   # best of five single calls: the most favourable reading for the reference
                            pr = pops[(r + 1) % R]
                            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                            e0.record()
                            ref.sr_fitness(pr.batch_node_value, pr.batch_node_type, pr.batch_subtree_size, X, y)
                            e1.record(); torch.cuda.synchronize()
                            best = min(best, e0.elapsed_time(e1))
                        line["reference_cuda_same_gpu"] = {"value": (hi - lo) * N / (best * 1e-3), "unit": "tree-evals/s",
                                                           "what": "reference forward.cu SR_fitness(kernel_type=4) compiled unmodified for sm_100a (oracle/_ref); best of 5 calls"}
                except Exception as e:
                    line["reference_cuda_same_gpu"] = {"unavailable": repr(e)}
        out.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip timing the reference's CUDA kernels")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        with StdoutGuard() as out:
            run_ours(args, out)


if __name__ == "__main__":
    main()
