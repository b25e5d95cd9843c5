"""CPU suite, part 3: the C-ABI library and the host-side logic (no compute on a GPU-less box)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ARITH_FUNCS, ROOT, make_data, make_forest

HEADER = os.path.join(ROOT, "include", "evogp_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(evogp_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native):
    lib = C.CDLL(native.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/evogp_b200.h but not exported"
    assert sorted(native.ABI_SYMBOLS) == names
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (evogp_\w+)", out))
    assert exported == set(names), f"undeclared exports: {exported - set(names)}"


def test_operator_library_registers_reference_schemas(native):
    native.load_ops()
    ns = torch.ops.evogp_cuda
    # schemas of src/evogp/cuda/torch_wrapper.cu:294-298, argument for argument
    want = {
        "tree_generate": ["pop_size", "gp_len", "var_len", "out_len", "const_samples_len", "out_prob", "const_prob", "keys",
                          "depth2leaf_probs", "roulette_funcs", "const_samples"],
        "tree_mutate": ["pop_size", "gp_len", "value_ori", "type_ori", "subtree_size_ori", "mutateIndices", "value_new",
                        "type_new", "subtree_size_new"],
        "tree_crossover": ["pop_size_ori", "pop_size_new", "gp_len", "value_ori", "type_ori", "subtree_size_ori", "left_idx",
                           "right_idx", "left_node_idx", "right_node_idx"],
        "tree_evaluate": ["pop_size", "gp_len", "var_len", "out_len", "value", "node_type", "subtree_size", "variables"],
        "tree_SR_fitness": ["pop_size", "data_points", "gp_len", "var_len", "out_len", "useMSE", "value", "node_type",
                            "subtree_size", "variables", "labels", "kernel_type"],
    }
    for op, args in want.items():
        schema = getattr(ns, op).default._schema
        assert [a.name for a in schema.arguments] == args


def test_reference_front_end_import_path():
    """The reference reaches its native code through `import evogp.evogp_cuda` (src/evogp/tree/__init__.py:2) and the
    `evogp.tree / algorithm / problem / pipeline` modules; the shim provides every one of them."""
    import importlib

    mod = importlib.import_module("evogp.evogp_cuda")
    assert mod is not None and hasattr(torch.ops.evogp_cuda, "tree_SR_fitness")
    from evogp.tree import Forest, GenerateDescriptor          # noqa: F401
    from evogp.algorithm import GeneticProgramming, DefaultSelection, DefaultMutation, DefaultCrossover   # noqa: F401
    from evogp.problem import SymbolicRegression               # noqa: F401
    from evogp.pipeline import StandardPipeline                # noqa: F401


def test_scalar_argument_errors_need_no_gpu(native):
    L = native.abi()
    assert L.evogp_crossover(0, 1, 64, *([None] * 11)) == 1 and b"pop_size_ori" in L.evogp_last_error()
    assert L.evogp_mutate(4, 0, *([None] * 11)) == 1 and b"gp_len" in L.evogp_last_error()
    assert L.evogp_generate(4, 2000, 1, 1, 1, 0.5, 0.5, *([None] * 8)) == 1
    assert L.evogp_generate(4, 64, 1, 1, 1, 1.5, 0.5, *([None] * 8)) == 1 and b"out_prob" in L.evogp_last_error()
    assert L.evogp_SR_fitness(4, 5, 8, 2, 1, 1, None, None, None, None, None, None, 4, None, 0, None) == 3
    assert L.evogp_eval_workspace_bytes(1000, 64) >= 256      # scheduler words only: programs live in shared memory


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_a_gpu(native):
    L = native.abi()
    assert L.evogp_crossover(4, 4, 8, *([None] * 11)) == 2          # EVOGP_ERR_CUDA, not a silent fallback
    assert b"no CPU path" in L.evogp_last_error()
    native.load_ops()
    with pytest.raises(NotImplementedError):                          # only the CUDA dispatch key is registered
        torch.ops.evogp_cuda.tree_evaluate(1, 4, 1, 1, torch.zeros(1, 4), torch.zeros(1, 4, dtype=torch.int16),
                                           torch.ones(1, 4, dtype=torch.int16), torch.zeros(1, 1))


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "evogp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text, f


def test_shard_bounds_partition_the_population():
    from evogp_b200.parallel import shard_bounds

    for P in (1, 7, 100000, 1000003):
        for W in (1, 2, 3, 8):
            covered = []
            for r in range(W):
                lo, hi, per = shard_bounds(P, W, r)
                assert 0 <= lo <= hi <= P and hi - lo <= per
                covered.extend([lo, hi])
            assert covered[0] == 0 and covered[-1] == P
            assert all(covered[i] == covered[i + 1] for i in range(1, len(covered) - 1, 2))


_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import oracle
from conftest import ARITH_FUNCS, make_data, make_forest
from evogp_b200.parallel import all_gather_fitness, shard_bounds
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
P = 1001                                   # odd: the last shard is ragged
v, t, s = make_forest(oracle, P, 32, 3, 1, ARITH_FUNCS, 5, keys=(3, 3))
X, y = make_data(40, 3)
lo, hi, _ = shard_bounds(P, 2, rank)
local = torch.from_numpy(oracle.sr_fitness(v[lo:hi], t[lo:hi], s[lo:hi], X, y))
full = all_gather_fitness(local, P)
want = torch.from_numpy(oracle.sr_fitness(v, t, s, X, y))
assert full.shape == (P,)
assert torch.equal(torch.nan_to_num(full, nan=-1.0), torch.nan_to_num(want, nan=-1.0)), "gathered fitness differs"
# the fused exchange needs NCCL + peer memory: on gloo it must decline (and say why), never half-initialise
from evogp_b200.parallel import FitnessExchange
ex = FitnessExchange(P, torch.device("cpu"))
assert not ex.available and "nccl" in ex.why and (ex.lo, ex.hi) == (lo, hi)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_sharded_fitness_all_gather_world_size_2(orc, tmp_path):
    """N > 1 path on CPU: two gloo ranks evaluate their row shards (with the oracle standing in for the
    kernel) and exchange exactly one all-gather of fitness scalars."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=29500 + os.getpid() % 2000))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-2000:]}"
        assert f"rank {r} ok" in o


def test_selection_counts_follow_the_reference():
    # selection/default.py:56-69: int(P * rate), prefixes of the descending sort
    from evogp_b200.algorithm.selection import DefaultSelection

    sel = DefaultSelection(survival_rate=0.3, elite_rate=0.01)
    assert sel.counts(5000) == (50, 1500) and sel.counts(99) == (0, 29)
    assert DefaultSelection(0.5, elite_cnt=7).counts(10) == (7, 5)
    with pytest.raises(AssertionError):
        DefaultSelection(0.3, elite_cnt=1, elite_rate=0.1)
