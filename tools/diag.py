#!/usr/bin/env python
"""Developer diagnostics (GPU): per-operator timings with CUDA events, achieved GB/s of the genetic
operators against their algorithmic bytes, and a per-phase breakdown of one GP generation."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from evogp_b200.tree import Forest, GenerateDescriptor  # noqa: E402
from evogp_b200.algorithm import DefaultCrossover, DefaultMutation, DefaultSelection, GeneticProgramming  # noqa: E402
from evogp_b200.problem import SymbolicRegression  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    V = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    L, N = 64, 1024
    torch.manual_seed(0)
    desc = GenerateDescriptor(max_tree_len=L, input_len=V, output_len=1, using_funcs=["+", "-", "*", "/"], max_layer_cnt=6,
                              const_samples=[-1, 0, 1])
    f = Forest.random_generate(P, desc)
    X = torch.rand(N, V, device="cuda") * 2 - 1
    y = (X[:, :1] ** 2).contiguous()
    lens = f.batch_subtree_size[:, 0].float()
    print(f"P={P} V={V} L={L} N={N} mean_len={lens.mean():.2f}")
    t_full = timeit(lambda: f.SR_fitness(X, y))
    t_n1 = timeit(lambda: f.SR_fitness(X[:1], y[:1]))
    print(f"SR_fitness N={N}: {t_full:.1f} us   N=1 (lowering + tiny replay): {t_n1:.1f} us")
    keys = torch.tensor([1, 2], dtype=torch.uint32, device="cuda")
    t_gen = timeit(lambda: Forest.generate_with_keys(P, desc, keys))
    gen_bytes = P * L * 8
    print(f"generate: {t_gen:.1f} us  ({gen_bytes / t_gen / 1e3:.0f} GB/s written, full rows)")
    n_new = int(P * 0.99)
    surv = f[: int(P * 0.3)]
    li = torch.randint(0, len(surv), (n_new,), dtype=torch.int32, device="cuda")
    ri = torch.randint(0, len(surv), (n_new,), dtype=torch.int32, device="cuda")
    lp = torch.randint(0, 2**31 - 1, (n_new,), dtype=torch.int32, device="cuda") % surv.batch_subtree_size[li.long(), 0]
    rp = torch.randint(0, 2**31 - 1, (n_new,), dtype=torch.int32, device="cuda") % surv.batch_subtree_size[ri.long(), 0]
    t_cx = timeit(lambda: surv.crossover(li, ri, lp, rp))
    child = surv.crossover(li, ri, lp, rp)
    clen = child.batch_subtree_size[:, 0].float().sum().item()
    alg = 8 * (2 * clen) + 16 * n_new          # read spans ~= child length, write child (valid prefix)
    print(f"crossover {len(surv)} -> {n_new}: {t_cx:.1f} us  algorithmic {alg / t_cx / 1e3:.0f} GB/s, moved (full rows written) {(8 * clen + n_new * L * 8) / t_cx / 1e3:.0f} GB/s")
    donors = Forest.random_generate(P, desc.update(max_layer_cnt=3))
    pos = torch.randint(0, 1024, (P,), dtype=torch.int32, device="cuda") % f.batch_subtree_size[:, 0]
    t_mu = timeit(lambda: f.mutate(pos, donors))
    print(f"mutate {P}: {t_mu:.1f} us  moved {(2 * P * L * 8) / t_mu / 1e3:.0f} GB/s")
    # one GP generation, phase by phase (wall clock with syncs)
    algo = GeneticProgramming(f, DefaultCrossover(), DefaultMutation(0.2, desc.update(max_layer_cnt=3)),
                              DefaultSelection(survival_rate=0.3, elite_rate=0.01))
    prob = SymbolicRegression(datapoints=X, labels=y)
    for _ in range(3):
        algo.step(prob.evaluate(algo.forest))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    G = 10
    t_eval = 0.0
    for _ in range(G):
        a = time.perf_counter()
        fit = prob.evaluate(algo.forest)
        torch.cuda.synchronize()
        t_eval += time.perf_counter() - a
        algo.step(fit)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) / G
    print(f"GP generation: {tot * 1e3:.3f} ms total, evaluate {t_eval / G * 1e3:.3f} ms, select+crossover+mutate+glue {(tot - t_eval / G) * 1e3:.3f} ms; "
          f"mean_len now {algo.forest.batch_subtree_size[:, 0].float().mean():.1f}")


if __name__ == "__main__" and (len(sys.argv) <= 3):
    main()


def gp_phase_breakdown(P=100000, V=3):
    """Wall-clock (synchronised) time of every phase of GeneticProgramming.step, reference-equivalent flow."""
    from evogp_b200.tree import MAX_STACK
    L, N = 64, 1024
    torch.manual_seed(0)
    desc = GenerateDescriptor(max_tree_len=L, input_len=V, output_len=1, using_funcs=["+", "-", "*", "/"], max_layer_cnt=6,
                              const_samples=[-1, 0, 1])
    mdesc = desc.update(max_layer_cnt=3)
    forest = Forest.random_generate(P, desc)
    fitness = torch.rand(P, device="cuda")
    T = {}

    def lap(name, fn):
        torch.cuda.synchronize(); a = time.perf_counter()
        r = fn()
        torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + (time.perf_counter() - a) * 1e3
        return r

    for it in range(6):
        if it == 1:
            T.clear()
        order = lap("sort", lambda: torch.sort(fitness, descending=True).indices)
        surv_idx = order[: int(P * 0.3)].to(torch.int32); elite_idx = order[: int(P * 0.01)].to(torch.int32)
        parents = lap("gather survivors", lambda: forest[surv_idx])
        n_new = P - len(elite_idx)
        pair = lap("randint x2", lambda: (torch.randint(0, len(parents), (2, n_new), dtype=torch.int32, device="cuda"),
                                          torch.randint(0, 2**31 - 1, (2, n_new), dtype=torch.int32, device="cuda")))
        sizes = parents.batch_subtree_size[:, 0]
        lp, rp = lap("positions", lambda: (pair[1][0] % sizes[pair[0][0]], pair[1][1] % sizes[pair[0][1]]))
        kids = lap("crossover kernel", lambda: parents.crossover(pair[0][0], pair[0][1], lp, rp))
        chosen = lap("cpu mask", lambda: torch.rand(P - len(elite_idx)) < 0.2)
        cnt = lap("mask sum", lambda: int(chosen.sum()))
        rows = lap("nonzero + h2d", lambda: chosen.nonzero(as_tuple=True)[0].to("cuda"))
        mutants = lap("gather mutants", lambda: kids[rows])
        donors = lap("generate donors", lambda: Forest.random_generate(cnt, mdesc))
        pos = lap("mut positions", lambda: torch.randint(0, MAX_STACK, (cnt,), dtype=torch.int32, device="cuda") % mutants.batch_subtree_size[:, 0])
        mut = lap("mutate kernel", lambda: mutants.mutate(pos, donors))
        lap("scatter back", lambda: kids.__setitem__(rows, mut))
        forest = lap("cat elites", lambda: forest[elite_idx] + kids)
    tot = sum(T.values()) / 5
    print(f"GP step phases at P={P} (ms, mean of 5): total {tot:.3f}")
    for k, v in T.items():
        print(f"   {k:18s} {v / 5:.3f}")


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "phases":
    gp_phase_breakdown(int(sys.argv[1]), int(sys.argv[2]))


def fused_loop(P=100000, V=3, G=20):
    """Generations per second of the reference-equivalent loop vs the fused step (same algorithm, different RNG stream)."""
    from evogp_b200.algorithm import FusedGeneticProgramming
    L, N = 64, 1024
    torch.manual_seed(0)
    desc = GenerateDescriptor(max_tree_len=L, input_len=V, output_len=1, using_funcs=["+", "-", "*", "/"], max_layer_cnt=6,
                              const_samples=[-1, 0, 1])
    X = torch.rand(N, V, device="cuda") * 2 - 1
    y = (X[:, :1] ** 2 + X[:, 1:2]).contiguous()
    prob = SymbolicRegression(datapoints=X, labels=y)
    for name in ("unfused", "fused"):
        f = Forest.random_generate(P, desc)
        if name == "fused":
            algo = FusedGeneticProgramming(f, desc.update(max_layer_cnt=3), 0.2, 0.3, elite_rate=0.01)
        else:
            algo = GeneticProgramming(f, DefaultCrossover(), DefaultMutation(0.2, desc.update(max_layer_cnt=3)),
                                      DefaultSelection(survival_rate=0.3, elite_rate=0.01))
        for _ in range(3):
            algo.step(prob.evaluate(algo.forest))
        torch.cuda.synchronize()
        t0 = time.perf_counter(); te = 0.0
        for _ in range(G):
            a = time.perf_counter(); fit = prob.evaluate(algo.forest); torch.cuda.synchronize(); te += time.perf_counter() - a
            algo.step(fit)
        torch.cuda.synchronize()
        tot = (time.perf_counter() - t0) / G
        best = float(torch.nan_to_num(prob.evaluate(algo.forest), nan=float("-inf")).max())
        print(f"{name:8s} P={P}: {tot * 1e3:.3f} ms/generation (evaluate {te / G * 1e3:.3f} ms, step {(tot - te / G) * 1e3:.3f} ms), "
              f"mean_len {algo.forest.batch_subtree_size[:, 0].float().mean():.1f}, best fitness {best:.4g}")


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "fused":
    fused_loop(int(sys.argv[1]), int(sys.argv[2]))
