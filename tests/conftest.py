import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def native():
    """Native libraries, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from evogp_b200 import build

    if not (os.path.exists(build.LIB_SO) and os.path.exists(build.OPS_SO)):
        build.build_all()
    from evogp_b200 import _native

    return _native


@pytest.fixture(autouse=True)
def _default_replay_width(request):
    """The evaluation-kernel width is process-wide state the front-end sets from the function set of the last generated
    forest (evogp_eval_set_replay_width): every GPU test starts from the automatic choice."""
    if request.node.get_closest_marker("gpu") is not None and not os.environ.get("EVOGP_REPLAY_K"):
        from evogp_b200 import _native

        _native.set_replay_width(0)
    yield


# ---------------------------------------------------------------------------
# shared builders of seeded test inputs (numpy, CPU)
# ---------------------------------------------------------------------------
FUNC_NAMES = ["if", "+", "-", "*", "/", "loose_div", "pow", "loose_pow", "max", "min", "<", ">", "<=", ">=",
              "sin", "cos", "tan", "sinh", "cosh", "tanh", "log", "loose_log", "exp", "inv", "loose_inv", "neg", "abs",
              "sqrt", "loose_sqrt"]


def roulette(names):
    p = np.zeros(29, np.float32)
    for n in names:
        p[FUNC_NAMES.index(n)] = 1.0
    p /= p.sum()
    return np.cumsum(p, dtype=np.float32)


def depth2leaf(max_layer_cnt, leaf_prob=0.2):
    inner = max_layer_cnt - 1
    return np.array([leaf_prob] * inner + [1.0] * (10 - inner), np.float32)


EXACT_FUNCS = ["+", "-", "*", "max", "min", "<", ">", "<=", ">=", "neg", "abs", "if"]          # IEEE-exact on both sides
ARITH_FUNCS = ["+", "-", "*", "/"]
ALL_FUNCS = FUNC_NAMES


def make_forest(orc, pop, L, V, O=1, funcs=ARITH_FUNCS, layers=5, keys=(42, 0), consts=(-1.0, 0.0, 1.0), out_prob=0.5,
                const_prob=0.5, leaf_prob=0.2):
    """Random forest produced by the CPU oracle's generator (bit-identical to the reference's)."""
    return orc.generate(pop, L, V, O, out_prob, const_prob, np.array(keys, np.uint32), depth2leaf(layers, leaf_prob),
                        roulette(funcs), np.array(consts, np.float32))


def make_data(N, V, O=1, seed=0):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, size=(N, V)).astype(np.float32)
    y = (X[:, :1] ** 2 + (X[:, 1:2] if V > 1 else 0)).astype(np.float32)
    if O > 1:
        y = np.concatenate([y + o for o in range(O)], axis=1).astype(np.float32)
    return X, y


def prefix_equal(a, b, lens):
    """Compare [P, L] arrays on their valid prefixes only."""
    L = a.shape[1]
    m = np.arange(L)[None, :] < np.asarray(lens)[:, None]
    if a.dtype.kind == "f":
        return np.array_equal(a.view(np.uint32)[m], b.view(np.uint32)[m])
    return np.array_equal(a[m], b[m])
