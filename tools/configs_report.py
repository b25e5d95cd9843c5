#!/usr/bin/env python
"""Times every BASELINE.json config on ONE B200 (configs 3 and 5 also as the per-GPU shard of the 8-GPU layout)
and prints a JSON report (kept as profiles/r2_configs.json).  CUDA-event timed, median of several repetitions."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from evogp_b200.algorithm import (DefaultCrossover, DefaultMutation, DefaultSelection, FusedGeneticProgramming,  # noqa: E402
                                  GeneticProgramming, GraphedGeneration)
from evogp_b200.problem import Classification, SymbolicRegression  # noqa: E402
from evogp_b200.tree import Forest, GenerateDescriptor  # noqa: E402


def ev_time(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def ref_time(fn):
    """Reference kernels: best of five warmed-up calls (their first launches size local-memory frames and are erratic)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def sr_data(N, V):
    X = torch.rand(N, V, device="cuda") * 2 - 1
    y = (X[:, :1] ** 4 / (X[:, :1] ** 4 + 1) + X[:, 1:2] ** 4 / (X[:, 1:2] ** 4 + 1) + X[:, 2:].sum(1, keepdim=True) * 0.1).contiguous()
    return X, y


def main():
    torch.manual_seed(0)
    rep = {}
    ref = oracle.ref_gpu() if oracle.ref_gpu_available() else None
    funcs = ["+", "-", "*", "/"]
    # ---- config 1: XOR-3d, pop 5000, L 32 (example/basic.py plumbing) ----
    X = torch.tensor([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)], dtype=torch.float32, device="cuda")
    y = (X.sum(1, keepdim=True) % 2).contiguous()
    d = GenerateDescriptor(max_tree_len=32, input_len=3, output_len=1, using_funcs=funcs, max_layer_cnt=4, const_samples=[-1, 0, 1])
    f = Forest.random_generate(5000, d)
    prob = SymbolicRegression(datapoints=X, labels=y)
    algo = GeneticProgramming(f, DefaultCrossover(), DefaultMutation(0.2, d.update(max_layer_cnt=3)), DefaultSelection(0.3, elite_rate=0.01))
    ms = ev_time(lambda: algo.step(prob.evaluate(algo.forest)), reps=20, warm=5)
    hv = [a.cpu().numpy() for a in (f.batch_node_value, f.batch_node_type, f.batch_subtree_size)]
    t0 = time.perf_counter(); oracle.sr_fitness(*hv, X.cpu().numpy(), y.cpu().numpy(), nthreads=oracle.max_threads()); cpu_ms = (time.perf_counter() - t0) * 1e3
    rep["config1_xor3d_pop5000_L32"] = {"gpu_ms_per_generation": ms, "cpu_oracle_fitness_ms": cpu_ms,
                                        "gpu_fitness_ms": ev_time(lambda: prob.evaluate(f), reps=20)}
    # ---- config 2 ----
    X, y = sr_data(1024, 3)
    d = GenerateDescriptor(max_tree_len=64, input_len=3, output_len=1, using_funcs=funcs, max_layer_cnt=6, const_samples=[-1, 0, 1])
    f = Forest.random_generate(100000, d)
    ms = ev_time(lambda: f.SR_fitness(X, y))
    rep["config2_pop1e5_L64_N1024_V3"] = {"ms": ms, "tree_evals_per_s": 1e5 * 1024 / ms * 1e3}
    if ref:
        rms = ref_time(lambda: ref.sr_fitness(f.batch_node_value, f.batch_node_type, f.batch_subtree_size, X, y))
        rep["config2_pop1e5_L64_N1024_V3"]["reference_cuda_ms"] = rms
    # ---- SURVEY.md 8(d) variants of config 2: transcendental function set; post-evolution population ----
    dt = GenerateDescriptor(max_tree_len=64, input_len=3, output_len=1, using_funcs=["+", "-", "*", "/", "sin", "cos", "tan"],
                            max_layer_cnt=6, const_samples=[-1, 0, 1])
    ft = Forest.random_generate(100000, dt)
    ms_t = ev_time(lambda: ft.SR_fitness(X, y))
    var = {"transcendental_funcs_ms": ms_t, "transcendental_tree_evals_per_s": 1e5 * 1024 / ms_t * 1e3,
           "transcendental_mean_tree_len": float(ft.batch_subtree_size[:, 0].float().mean())}
    if ref:
        var["transcendental_reference_cuda_ms"] = ref_time(lambda: ref.sr_fitness(ft.batch_node_value, ft.batch_node_type, ft.batch_subtree_size, X, y))
    # every function of the reference (if / pow / sinh / cosh leave the PTX fast path for the generic interpreter)
    ALL = ["if", "+", "-", "*", "/", "loose_div", "pow", "loose_pow", "max", "min", "<", ">", "<=", ">=", "sin", "cos", "tan", "sinh", "cosh",
           "tanh", "log", "loose_log", "exp", "inv", "loose_inv", "neg", "abs", "sqrt", "loose_sqrt"]
    da = GenerateDescriptor(max_tree_len=64, input_len=3, output_len=1, using_funcs=ALL, max_layer_cnt=4, const_samples=[-1, 0, 1])
    fa = Forest.random_generate(100000, da)
    ms_a = ev_time(lambda: fa.SR_fitness(X, y))
    var.update({"all_29_funcs_ms": ms_a, "all_29_funcs_tree_evals_per_s": 1e5 * 1024 / ms_a * 1e3,
                "all_29_funcs_mean_tree_len": float(fa.batch_subtree_size[:, 0].float().mean())})
    if ref:
        var["all_29_funcs_reference_cuda_ms"] = ref_time(lambda: ref.sr_fitness(fa.batch_node_value, fa.batch_node_type, fa.batch_subtree_size, X, y))
    del fa
    prob2 = SymbolicRegression(datapoints=X, labels=y)
    algo2 = GeneticProgramming(f, DefaultCrossover(), DefaultMutation(0.2, d.update(max_layer_cnt=3)), DefaultSelection(0.3, elite_rate=0.01))
    for _ in range(20):
        algo2.step(torch.nan_to_num(prob2.evaluate(algo2.forest), nan=float("-inf")))
    fe = algo2.forest
    ms_e = ev_time(lambda: fe.SR_fitness(X, y))
    var.update({"after_20_generations_ms": ms_e, "after_20_generations_tree_evals_per_s": 1e5 * 1024 / ms_e * 1e3,
                "after_20_generations_mean_tree_len": float(fe.batch_subtree_size[:, 0].float().mean())})
    if ref:
        var["after_20_generations_reference_cuda_ms"] = ref_time(lambda: ref.sr_fitness(fe.batch_node_value, fe.batch_node_type, fe.batch_subtree_size, X, y))
    rep["config2_variants"] = var
    del ft, fe, algo2
    # ---- config 3: pop 1e6, V 10 ----
    X, y = sr_data(1024, 10)
    d = GenerateDescriptor(max_tree_len=64, input_len=10, output_len=1, using_funcs=funcs, max_layer_cnt=6, const_samples=[-1, 0, 1])
    f = Forest.random_generate(1000000, d)
    ms_full = ev_time(lambda: f.SR_fitness(X, y), reps=5)
    shard = f[0:125000]
    shard = Forest(10, 1, shard.batch_node_value.contiguous(), shard.batch_node_type.contiguous(), shard.batch_subtree_size.contiguous())
    ms_shard = ev_time(lambda: shard.SR_fitness(X, y))
    rep["config3_pop1e6_L64_N1024_V10"] = {"one_gpu_full_population_ms": ms_full, "tree_evals_per_s_one_gpu": 1e6 * 1024 / ms_full * 1e3,
                                           "per_gpu_shard_of_8_ms": ms_shard}
    if ref:
        rms = ref_time(lambda: ref.sr_fitness(shard.batch_node_value, shard.batch_node_type, shard.batch_subtree_size, X, y))
        rep["config3_pop1e6_L64_N1024_V10"]["reference_cuda_shard_ms"] = rms
    del f, shard
    # ---- config 4: multi-output classification shape ----
    rng = np.random.default_rng(0)
    Xc = torch.from_numpy(rng.normal(size=(4096, 13)).astype(np.float32)).cuda()
    labels = torch.from_numpy(rng.integers(0, 3, 4096).astype(np.float32)).cuda()
    d = GenerateDescriptor(max_tree_len=128, input_len=13, output_len=3, using_funcs=funcs, max_layer_cnt=7, const_samples=[-1, 0, 1], out_prob=0.5)
    f = Forest.random_generate(200000, d)
    onehot = torch.nn.functional.one_hot(labels.long(), 3).float().contiguous()
    ms_fit = ev_time(lambda: f.SR_fitness(Xc, onehot), reps=5)
    cls = Classification(datapoints=Xc, labels=labels, multi_output=True)
    ms_cls = ev_time(lambda: cls.evaluate(f), reps=5, warm=2)
    ms_unf = ev_time(lambda: cls.evaluate_unfused(f), reps=2, warm=1)
    rep["config4_pop2e5_L128_N4096_V13_O3"] = {"sr_fitness_onehot_ms": ms_fit, "tree_evals_per_s": 2e5 * 4096 / ms_fit * 1e3,
                                              "classification_accuracy_fused_ms": ms_cls, "classification_tree_evals_per_s": 2e5 * 4096 / ms_cls * 1e3,
                                              "classification_accuracy_unfused_ms (batch_forward 9.8 GB + softmax/argmax in torch)": ms_unf,
                                              "datapoint_tiles": "16 x 4096 floats > staging area: tiled launches"}
    del f
    torch.cuda.empty_cache()
    # ---- config 5: full GP loop, pop 500000 (replicated on every GPU in the 8-GPU layout) ----
    X, y = sr_data(1024, 10)
    d = GenerateDescriptor(max_tree_len=64, input_len=10, output_len=1, using_funcs=funcs, max_layer_cnt=6, const_samples=[-1, 0, 1])
    prob = SymbolicRegression(datapoints=X, labels=y)
    out = {}
    for name in ("unfused", "fused", "fused_cuda_graph"):
        torch.manual_seed(1)
        f = Forest.random_generate(500000, d)
        if name == "unfused":
            algo = GeneticProgramming(f, DefaultCrossover(), DefaultMutation(0.2, d.update(max_layer_cnt=3)), DefaultSelection(0.3, elite_rate=0.01))
            one = lambda: algo.step(torch.nan_to_num(prob.evaluate(algo.forest), nan=float("-inf")))
        else:
            algo = FusedGeneticProgramming(f, d.update(max_layer_cnt=3), 0.2, 0.3, elite_rate=0.01)
            if name == "fused":
                one = lambda: algo.step(prob.evaluate(algo.forest))
            else:
                gen = GraphedGeneration(algo, prob)
                one = gen.replay
        for _ in range(3):
            one()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        G = 30
        for _ in range(G):
            one()
        torch.cuda.synchronize()
        out[name + "_ms_per_generation"] = (time.perf_counter() - t0) / G * 1e3
        out[name + "_mean_tree_len_after"] = float(algo.forest.batch_subtree_size[:, 0].float().mean())
        del f, algo
        torch.cuda.empty_cache()
    out["note"] = "single GPU evaluating all 500000 trees; in the 8-GPU layout each rank evaluates 62500 and runs the same step"
    rep["config5_gp_loop_pop5e5_L64_N1024_V10"] = out
    print(json.dumps(rep, indent=1))


if __name__ == "__main__":
    main()
