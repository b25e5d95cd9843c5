"""-m gpu: this library's own operators (no counterpart in the reference's kernel.h), bit-exact against their CPU
restatements in oracle/evogp_oracle.c: the Philox mode of generate (BASELINE.json north_star), the fused generation
step, subtree extraction and tournament selection (f-3)."""
import ctypes as C

import numpy as np
import pytest
import torch

import gpu_util as G
from conftest import ALL_FUNCS, ARITH_FUNCS, depth2leaf, make_forest, roulette

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("pop,L,V,O,funcs,layers,keys", [(20000, 64, 3, 1, ARITH_FUNCS, 6, (1000, 7)), (3000, 128, 13, 3, ALL_FUNCS, 4, (5, 5)),
                                                       (999, 33, 2, 2, ARITH_FUNCS + ["sin"], 5, (0, 0)), (64, 1024, 5, 1, ARITH_FUNCS, 9, (3, 1))])
def test_generate_philox_bit_exact(native, orc, pop, L, V, O, funcs, layers, keys):
    d2l, roul, consts = depth2leaf(layers), roulette(funcs), np.array([-1.0, 0.0, 1.0], np.float32)
    want = orc.generate_philox(pop, L, V, O, 0.5, 0.5, np.array(keys, np.uint32), d2l, roul, consts, nthreads=8)
    k, a, r, c = G.to_dev(np.array(keys, np.uint32), d2l, roul, consts)
    v = torch.empty((pop, L), dtype=torch.float32, device=G.dev()); t = torch.empty((pop, L), dtype=torch.int16, device=G.dev())
    s = torch.empty((pop, L), dtype=torch.int16, device=G.dev())
    rc = native.abi().evogp_generate_philox(pop, L, V, O, 3, 0.5, 0.5, _p(k), _p(a), _p(r), _p(c), _p(v), _p(t), _p(s), G._stream())
    native.check(rc, "evogp_generate_philox")
    torch.cuda.synchronize()
    for g, w in zip((v, t, s), want):
        assert G.same_bits(g, w)
    lens = orc.check_forest(*want, input_len=V, output_len=O)
    # the same growth rules as the taus88 mode: the two populations are different trees of the same distribution
    tv, tt, ts = orc.generate(pop, L, V, O, 0.5, 0.5, np.array(keys, np.uint32), d2l, roul, consts)
    assert not np.array_equal(tt, want[1])
    if pop >= 3000:
        assert abs(lens.mean() - ts[:, 0].mean()) < 0.08 * ts[:, 0].mean()


@pytest.mark.parametrize("pop,L", [(5000, 64), (700, 33), (64, 1024)])
def test_extract_subtree_bit_exact(native, orc, pop, L):
    v, t, s = make_forest(orc, pop, L, 3, 1, ARITH_FUNCS, 5 if L < 1024 else 9, keys=(8, 8), leaf_prob=0.1)
    rng = np.random.default_rng(2)
    pos = (rng.integers(0, 1 << 30, pop) % s[:, 0]).astype(np.int32)
    pos[:5] = [-1, L, L + 7, 0, int(s[4, 0]) - 1]
    want = orc.extract_subtree(v, t, s, pos)
    dv, dt, ds, dp = G.to_dev(v, t, s, pos)
    ov, ot, os_ = torch.empty_like(dv), torch.empty_like(dt), torch.empty_like(ds)
    native.check(native.abi().evogp_extract_subtree(pop, L, _p(dv), _p(dt), _p(ds), _p(dp), _p(ov), _p(ot), _p(os_), G._stream()), "extract")
    torch.cuda.synchronize()
    for g, w in zip((ov, ot, os_), want):
        assert G.same_bits(g, w)
    # the reference's torch formulation (mutation_utils.py:6-48) on the in-range positions
    ok = (pos >= 0) & (pos < L)
    start = torch.from_numpy(pos[ok]).long().to(G.dev())[:, None]
    length = ds[torch.from_numpy(ok).to(G.dev())].gather(1, start).long()
    idx = (torch.arange(L, device=G.dev())[None, :] + start).clamp(max=L - 1)
    for got, src in ((ov, dv), (ot, dt), (os_, ds)):
        ref = torch.where(idx < start + length, src[torch.from_numpy(ok).to(G.dev())].gather(1, idx), torch.zeros((), dtype=src.dtype, device=G.dev()))
        assert torch.equal(got[torch.from_numpy(ok).to(G.dev())], ref)
    sub = np.ascontiguousarray(want[2][ok][5:])
    assert (sub[:, 0] >= 1).all()


@pytest.mark.parametrize("P,T,best_p,replace,count", [(10000, 5, 1.0, True, 7000), (10000, 7, 0.8, True, 10000), (10000, 4, 1.0, False, 6000),
                                                       (1001, 3, 0.6, False, 2000), (50, 50, 0.9, True, 40), (300, 1, 1.0, False, 300)])
def test_tournament_select_bit_exact_and_distribution(native, orc, P, T, best_p, replace, count):
    rng = np.random.default_rng(P + T)
    fit = rng.normal(size=P).astype(np.float32)
    fit[rng.integers(0, P, P // 50)] = np.nan
    fit[rng.integers(0, P, P // 50)] = fit[0]                       # ties
    keys = np.array([123, 456], np.uint32)
    want = orc.tournament(fit, T, best_p, replace, count, keys)
    df, dk = G.to_dev(fit, keys)
    out = torch.empty(count, dtype=torch.int32, device=G.dev())
    rc = native.abi().evogp_tournament_select(P, _p(df), T, best_p, int(replace), count, _p(dk), _p(out), G._stream())
    native.check(rc, "evogp_tournament_select")
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.array_equal(got, want)
    assert (got >= 0).all() and (got < P).all()
    if not replace:      # every round of P // T tournaments draws from one permutation: with best_p = 1 and T = 1 winners never repeat
        per_round = P // T
        if T == 1:
            assert len(set(got[:per_round].tolist())) == min(per_round, count)
    if P >= 10000 and best_p == 1.0:
        # selection pressure of a size-T tournament: P(winner among the best fraction q) = 1 - (1 - q)^T
        rank = np.argsort(np.argsort(-np.nan_to_num(fit, nan=-np.inf), kind="stable"), kind="stable")
        for q in (0.1, 0.5):
            emp = (rank[got] < q * P).mean()
            assert abs(emp - (1 - (1 - q) ** T)) < 0.03


@pytest.mark.parametrize("P,L,V,O,funcs,rate", [(20000, 64, 3, 1, ARITH_FUNCS, 0.2), (3000, 128, 5, 3, ARITH_FUNCS + ["sin", "if"], 0.5),
                                                 (2000, 32, 2, 1, ARITH_FUNCS, 1.0), (1500, 33, 4, 1, ARITH_FUNCS, 0.0)])
def test_next_generation_bit_exact(native, orc, P, L, V, O, funcs, rate):
    """The fused generation step against its restatement built from the reference-style splice walk (f-1)."""
    layers = 4 if "if" in funcs else 5
    v, t, s = make_forest(orc, P, L, V, O, funcs, layers, keys=(12, 21), leaf_prob=0.15)
    rng = np.random.default_rng(4)
    order = rng.permutation(P).astype(np.int64)
    elite, survivors = P // 100, int(P * 0.3)
    d2l, roul, consts = depth2leaf(3), roulette(funcs), np.array([-1.0, 0.0, 1.0, 0.5], np.float32)
    keys = np.array([777, 31337], np.uint32)
    want = orc.next_generation(v, t, s, order, elite, survivors, rate, V, O, 0.5, 0.5, d2l, roul, consts, keys, nthreads=8)
    dv, dt, ds, do, da, dr, dc, dk = G.to_dev(v, t, s, order, d2l, roul, consts, keys)
    ov, ot, os_ = torch.empty_like(dv), torch.empty_like(dt), torch.empty_like(ds)
    rc = native.abi().evogp_next_generation(P, L, _p(dv), _p(dt), _p(ds), _p(do), elite, survivors, rate, V, O, consts.shape[0], 0.5, 0.5,
                                            _p(da), _p(dr), _p(dc), _p(dk), _p(ov), _p(ot), _p(os_), G._stream())
    native.check(rc, "evogp_next_generation")
    torch.cuda.synchronize()
    for g, w, name in zip((ov, ot, os_), want, ("value", "type", "size")):
        assert G.same_bits(g, w), f"node_{name} differs in {(g.cpu().numpy() != w).any(1).sum()} rows"
    lens = orc.check_forest(*want, input_len=V, output_len=O)
    assert (lens <= L).all()
    assert np.array_equal(want[1][:elite], t[order[:elite]])          # elites are verbatim copies
    changed = (want[1][elite:] != t[order[rng.integers(0, survivors, P - elite)]]).any(1).mean()
    assert changed > 0.5                                               # children are not copies of random survivors
