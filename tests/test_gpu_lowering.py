"""-m gpu: the register-resident lowering pass (csrc/lower_fast.cuh) must emit, word for word, the programs of the
generic pass (csrc/lower.cuh — the one the host harness and the oracle pin), for every row; rows outside its class
(ternary nodes, stale sizes, bad lengths) must come out of its fallback identical too."""
import ctypes as C

import numpy as np
import pytest
import torch

import gpu_util as G
from conftest import ALL_FUNCS, ARITH_FUNCS, EXACT_FUNCS, make_forest

pytestmark = pytest.mark.gpu
C_END = 0


def lower(native, v, t, s, V, O, fast, deep_from=0):
    P, L = v.shape
    Lp = (L + 2) & ~1
    ws, n = G._ws(native, P, L)
    prog = torch.zeros((P, Lp), dtype=torch.int64, device=v.device)
    rc = native.abi().evogp_debug_lower(P, L, V, O, G._p(v), G._p(t), G._p(s), int(fast), deep_from, G._p(ws), C.c_size_t(n),
                                        G._p(prog), G._stream())
    native.check(rc, "evogp_debug_lower")
    torch.cuda.synchronize()
    return prog.cpu().numpy().view(np.uint64)


def programs_equal(a, b):
    """Compare up to and including the first C_END of every row (slots after it are never read)."""
    code = (a & np.uint64(0x1FF)).astype(np.int64)
    is_end = code == C_END
    first_end = np.where(is_end.any(1), is_end.argmax(1), a.shape[1] - 1)
    live = np.arange(a.shape[1])[None, :] <= first_end[:, None]
    return np.array_equal(a[live], b[live]), int((a[live] != b[live]).sum())


@pytest.mark.parametrize("funcs,L,layers,V,deep", [(ARITH_FUNCS, 64, 6, 3, 0), (ARITH_FUNCS, 64, 6, 10, 4), (ARITH_FUNCS, 32, 5, 3, 0),
                                                    (ARITH_FUNCS + ["sin", "cos", "neg"], 64, 6, 4, 4), (ALL_FUNCS, 64, 4, 5, 0),
                                                    (EXACT_FUNCS, 48, 4, 2, 1), (["neg", "abs"], 16, 9, 1, 0), (ARITH_FUNCS, 7, 3, 2, 0),
                                                    (ARITH_FUNCS, 63, 6, 3, 2), (ARITH_FUNCS, 33, 5, 3, 0)])
def test_fast_pass_emits_the_generic_programs(native, orc, funcs, L, layers, V, deep):
    v, t, s = make_forest(orc, 20000, L, V, 1, funcs, layers, keys=(31, 7), consts=(-1.0, 0.5, 2.0), leaf_prob=0.15)
    dv, dt, ds = G.to_dev(v, t, s)
    slow = lower(native, dv, dt, ds, V, 1, fast=False, deep_from=deep)
    fast = lower(native, dv, dt, ds, V, 1, fast=True, deep_from=deep)
    ok, ndiff = programs_equal(slow, fast)
    assert ok, f"{ndiff} program words differ"
    assert (slow[:, 0] & np.uint64(0x1FF) != 3).all()        # no C_NAN programs in a well-formed population


def test_fast_pass_falls_back_on_rows_outside_its_class(native, orc):
    L, V = 64, 3
    v, t, s = make_forest(orc, 4000, L, V, 1, ARITH_FUNCS, 6, keys=(5, 9))
    rng = np.random.default_rng(1)
    s = s.copy(); t = t.copy(); v = v.copy()
    lens = s[:, 0].astype(np.int64)
    rows = np.arange(4000)
    # stale interior sizes (the reference's evaluator never reads them, forward.cu:283)
    stale = rows[(rows % 4 == 0) & (lens > 3)]
    s[stale, 1] = 1
    # impossible lengths and truncated prefixes -> C_NAN programs
    s[rows % 97 == 1, 0] = 0
    s[rows % 97 == 2, 0] = L + 5
    trunc = rows[(rows % 97 == 3) & (lens > 4)]
    s[trunc, 0] = (lens[trunc] - 2).astype(np.int16)
    # unknown type codes count as ternary in single-output mode (forward.cu:91-94 does not mask)
    weird = rows[(rows % 97 == 4) & (lens > 2)]
    t[weird, 1] = 0x83
    dv, dt, ds = G.to_dev(v, t, s)
    slow = lower(native, dv, dt, ds, V, 1, fast=False)
    fast = lower(native, dv, dt, ds, V, 1, fast=True)
    ok, ndiff = programs_equal(slow, fast)
    assert ok, f"{ndiff} program words differ"
    assert ((slow[:, 0] & np.uint64(0x1FF)) == 3).sum() > 50   # the malformed rows did lower to C_NAN


def test_fitness_is_identical_with_either_pass(native, orc, monkeypatch):
    """End to end through the C ABI: EVOGP_LOWER_FAST is read at load time, so this compares against the oracle-pinned
    generic programs by evaluating them: same programs -> same bits."""
    v, t, s = make_forest(orc, 5000, 64, 3, 1, ARITH_FUNCS, 6, keys=(77, 1))
    from conftest import make_data
    X, y = make_data(1024, 3, seed=4)
    dv, dt, ds, dX, dy = G.to_dev(v, t, s, X, y)
    got = G.abi_sr_fitness(native, dv, dt, ds, dX, dy)
    # the host-buffer path uploads one length per tree, which keeps it on the generic pass
    fit = np.zeros(5000, np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = native.abi().evogp_SR_fitness_host(5000, 1024, 64, 3, 1, 1, vp(v), vp(t), vp(s), vp(X), vp(y), vp(fit), torch.cuda.current_device())
    native.check(rc, "evogp_SR_fitness_host")
    assert G.same_bits(got, fit)
