"""CPU suite, part 2: the product's lowering pass (evogp_b200/csrc/lower.cuh), compiled for the host
by tests/host_lower_harness.cu, replayed by a scalar interpreter, against the oracle's direct stack
evaluation.  Bit-exact: reordering siblings never changes an operator's operands."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ALL_FUNCS, ARITH_FUNCS, EXACT_FUNCS, ROOT, make_data, make_forest

SRC = os.path.join(ROOT, "tests", "host_lower_harness.cu")
OUT = os.path.join(ROOT, "tests", "_build", "libhost_lower.so")
DEPS = [SRC] + [os.path.join(ROOT, "evogp_b200", "csrc", f) for f in ("lower.cuh", "program.cuh", "common.cuh")]


@pytest.fixture(scope="module")
def harness(orc):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        odir = os.path.join(ROOT, "oracle")
        subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-arch=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", OUT, SRC,
                               "-L", odir, "-loracle", "-Xlinker", "-rpath", "-Xlinker", odir])
    orc.lib()
    return C.CDLL(OUT)


def run(harness, v, t, s, X, O, use_sizes=1):
    v = np.ascontiguousarray(v, np.float32); t = np.ascontiguousarray(t, np.int16); s = np.ascontiguousarray(s, np.int16)
    X = np.ascontiguousarray(X, np.float32)
    P, L = v.shape
    N, V = X.shape
    out = np.zeros((P, N, O), np.float32)
    need = np.zeros(P, np.int32); ninstr = np.zeros(P, np.int32); maxsp = np.zeros(P, np.int32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = harness.harness_batch_forward(P, N, L, V, O, vp(v), vp(t), vp(s), vp(X), vp(out), vp(need), vp(ninstr), vp(maxsp), use_sizes)
    assert rc == 0, f"harness failed rc={rc}"
    return out, need, ninstr, maxsp


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same(a, b):
    """bit-equal, except that any NaN equals any NaN (payloads are not part of the contract)"""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(bits(a)[~na], bits(b)[~nb])


@pytest.mark.parametrize("funcs,L,layers,V", [(ARITH_FUNCS, 64, 6, 3), (ALL_FUNCS, 64, 4, 5), (EXACT_FUNCS, 121, 5, 2),
                                               (ARITH_FUNCS + ["if"], 40, 4, 3), (["neg", "abs"], 16, 9, 1)])
def test_lowered_program_equals_stack_machine(harness, orc, funcs, L, layers, V):
    v, t, s = make_forest(orc, 3000, L, V, 1, funcs, layers, keys=(11, 12), consts=(-1.0, 0.5, 2.0), leaf_prob=0.15)
    X, _ = make_data(33, V, seed=7)
    got, need, ninstr, maxsp = run(harness, v, t, s, X, 1)
    want = orc.batch_forward(v, t, s, X, 1)
    assert same(got, want)
    lens = s[:, 0]
    assert (ninstr <= lens).all() and (ninstr >= 1).all()
    assert (need >= 0).all() and (maxsp <= need).all()
    bound = harness.harness_depth_bound(L)
    assert need.max() <= bound
    # split mode (what the K = 16 replay kernel is fed): LOAD + acc-form instead of the fresh-value forms
    got2, need2, ninstr2, maxsp2 = run(harness, v, t, s, X, 1, use_sizes=3)
    assert same(got2, want)
    assert (ninstr2 <= lens).all() and (ninstr2 >= ninstr).all()
    assert (need2 == need).all() and (maxsp2 <= need2).all()
    # ... with slots >= 1 marked deep (the harness checks that exactly those accesses use the deep opcodes)
    got3, need3, ninstr3, _ = run(harness, v, t, s, X, 1, use_sizes=7)
    assert same(got3, want) and (need3 == need).all() and (ninstr3 == ninstr2).all()
    got4, need4, ninstr4, _ = run(harness, v, t, s, X, 1, use_sizes=5)   # default emission, deep pops / LOADs marked
    assert same(got4, want) and (need4 == need).all() and (ninstr4 == ninstr).all()


def test_multi_output_programs(harness, orc):
    v, t, s = make_forest(orc, 3000, 64, 4, 3, ALL_FUNCS, 4, keys=(21, 22), consts=(-1.0, 0.5, 2.0), out_prob=0.6)
    X, _ = make_data(20, 4, seed=8)
    got, need, _, _ = run(harness, v, t, s, X, 3)
    want = orc.batch_forward(v, t, s, X, 3)
    assert same(got, want)      # bit-exact: out nodes are emitted in the reference's processing order
    assert (need == 0).all()
    v2, t2, s2 = make_forest(orc, 2000, 64, 4, 2, EXACT_FUNCS, 4, keys=(23, 24), out_prob=0.9)
    got2, *_ = run(harness, v2, t2, s2, X, 2)
    assert same(got2, orc.batch_forward(v2, t2, s2, X, 2))
    # out nodes nested in out nodes, constants everywhere (the two-constant and IF3 encodings)
    v3, t3, s3 = make_forest(orc, 2000, 40, 2, 4, ["if", "+", "*", "neg"], 4, keys=(25, 26), out_prob=1.0, const_prob=0.9)
    got3, *_ = run(harness, v3, t3, s3, X[:, :2], 4)
    assert same(got3, orc.batch_forward(v3, t3, s3, X[:, :2], 4))


def test_depth_bound_is_tight_enough(harness):
    # M(d) table of program.cuh: fewest nodes needing depth d
    table = {3: 1, 4: 2, 8: 2, 9: 3, 13: 4, 26: 4, 27: 5, 40: 6, 64: 6, 81: 7, 121: 8, 128: 8, 243: 9, 364: 10, 729: 11, 1024: 11}
    for length, d in table.items():
        assert harness.harness_depth_bound(length) == d


def _nested(kind, depth):
    """Worst-case shapes for the operand stack: complete binary trees of unary-wrapped leaves and
    ternary trees of ternaries."""
    nodes = []

    def rec(d):
        if kind == "binary":
            if d == 0:
                nodes.append((2, 25.0)); nodes.append((0, 0.0))       # neg(x0): smallest non-leaf
            else:
                nodes.append((3, 1.0)); rec(d - 1); rec(d - 1)
        else:
            if d == 0:
                nodes.append((4, 0.0)); nodes.extend([(0, 0.0), (1, 1.0), (0, 1.0)])
            else:
                nodes.append((4, 0.0)); rec(d - 1); rec(d - 1); rec(d - 1)
    rec(depth)
    return nodes


@pytest.mark.parametrize("kind,depth,expect_need", [("binary", 1, 1), ("binary", 4, 4), ("ternary", 0, 2), ("ternary", 1, 4), ("ternary", 2, 6)])
def test_stack_need_of_worst_case_shapes(harness, orc, kind, depth, expect_need):
    nodes = _nested(kind, depth)
    L = max(8, len(nodes) + (len(nodes) & 1))
    t = np.zeros((1, L), np.int16); v = np.zeros((1, L), np.float32); s = np.zeros((1, L), np.int16)
    for i, (ty, val) in enumerate(nodes):
        t[0, i], v[0, i] = ty, val
    sizes = [0] * len(nodes)
    for i in range(len(nodes) - 1, -1, -1):
        ar = 0 if t[0, i] <= 1 else t[0, i] - 1
        sz, c = 1, i + 1
        for _ in range(ar):
            sz += sizes[c]; c += sizes[c]
        sizes[i] = sz
    s[0, :len(nodes)] = sizes
    orc.check_forest(v, t, s, input_len=2)
    X, _ = make_data(9, 2, seed=1)
    got, need, ninstr, maxsp = run(harness, v, t, s, X, 1)
    assert same(got, orc.batch_forward(v, t, s, X, 1))
    assert need[0] == expect_need and maxsp[0] == expect_need
    assert need[0] <= harness.harness_depth_bound(len(nodes))
    # the same worst cases through the other emission modes: split, deep slots (>= 1) marked, both
    for flags in (3, 5, 7):
        got_f, need_f, _, maxsp_f = run(harness, v, t, s, X, 1, use_sizes=flags)
        assert same(got_f, got) and need_f[0] == expect_need and maxsp_f[0] == expect_need


def test_malformed_rows_lower_to_nan(harness, orc):
    v, t, s = make_forest(orc, 8, 16, 2, 1, ARITH_FUNCS, 3, keys=(1, 1))
    s = s.copy(); t = t.copy()
    s[0, 0] = 0; s[1, 0] = 17; s[2, 0] = -3
    t[3, : s[3, 0]] = 3
    s[4, 0] = max(1, s[4, 0] - 1) if s[4, 0] > 1 else s[4, 0]    # truncated prefix (does not close) unless single leaf
    X, _ = make_data(3, 2)
    got, need, *_ = run(harness, v, t, s, X, 1)
    assert np.isnan(got[:4]).all() and (need[:4] == -1).all()


def test_sizes_are_verified_not_trusted(harness, orc):
    """The reference's evaluator reads only subtree_size[0] (forward.cu:283): rows whose interior sizes are
    stale, or that come without a size row at all, must lower to the same program."""
    v, t, s = make_forest(orc, 1500, 64, 3, 1, ARITH_FUNCS + ["if", "sin"], 4, keys=(31, 32), consts=(-1.0, 0.5))
    X, _ = make_data(17, 3, seed=5)
    want = orc.batch_forward(v, t, s, X, 1)
    got_nosize, need0, *_ = run(harness, v, t, s, X, 1, use_sizes=0)
    assert same(got_nosize, want)
    rng = np.random.default_rng(0)
    stale = s.copy()
    rows = rng.integers(0, 1500, 600)
    cols = (rng.integers(1, 64, 600) % np.maximum(s[rows, 0], 2)).clip(1)
    stale[rows, cols] = rng.integers(0, 70, 600).astype(np.int16)        # corrupt interior entries, keep size[:,0]
    got_stale, need1, *_ = run(harness, v, t, stale, X, 1)
    assert same(got_stale, want)
    got_ref, need2, *_ = run(harness, v, t, s, X, 1)
    assert np.array_equal(need0, need2) and np.array_equal(need1, need2)


@pytest.mark.parametrize("funcs,const_prob", [(ARITH_FUNCS, 0.5), (ALL_FUNCS, 0.8), (ARITH_FUNCS + ["neg", "sin"], 1.0)])
def test_constant_folding_changes_the_program_not_the_values(harness, orc, funcs, const_prob):
    """A function of constant leaves is folded at lowering time with the interpreter's own operator: fewer
    instructions, the same bits (also with all-constant trees, which fold level by level only once)."""
    layers = 4 if "if" in funcs else 6
    v, t, s = make_forest(orc, 3000, 64, 3, 1, funcs, layers, keys=(41, 42), consts=(-1.0, 0.0, 0.5, 2.0), const_prob=const_prob)
    X, _ = make_data(21, 3, seed=11)
    want = orc.batch_forward(v, t, s, X, 1)
    got_f, need_f, ninstr_f, _ = run(harness, v, t, s, X, 1, use_sizes=1)
    got_n, need_n, ninstr_n, _ = run(harness, v, t, s, X, 1, use_sizes=1 | 8)
    assert same(got_f, want) and same(got_n, want)
    assert (ninstr_f <= ninstr_n).all() and ninstr_f.sum() < 0.95 * ninstr_n.sum()
    assert (need_f <= need_n).all()
    # the other emission modes see the folded row too
    for flags in (3, 5, 7):
        got_m, *_ = run(harness, v, t, s, X, 1, use_sizes=flags)
        assert same(got_m, want)
