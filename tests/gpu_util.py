"""Helpers for the -m gpu tests: call the C ABI (include/evogp_b200.h) through ctypes with torch
device pointers, and move oracle/numpy data to the GPU."""
import ctypes as C

import numpy as np
import torch


def dev():
    return torch.device("cuda", torch.cuda.current_device())


def to_dev(*arrs):
    out = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev()) for a in arrs)
    return out if len(out) > 1 else out[0]


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ws(native, P, L):
    n = native.abi().evogp_eval_workspace_bytes(P, L)
    return torch.empty(n, dtype=torch.uint8, device=dev()), n


def abi_sr_fitness(native, v, t, s, X, y, use_mse=True):
    P, L = v.shape
    N, V = X.shape
    O = y.shape[1]
    fit = torch.empty(P, dtype=torch.float32, device=v.device)
    ws, n = _ws(native, P, L)
    rc = native.abi().evogp_SR_fitness(P, N, L, V, O, int(use_mse), _p(v), _p(t), _p(s), _p(X), _p(y), _p(fit), 4,
                                       _p(ws), n, _stream())
    native.check(rc, "evogp_SR_fitness")
    return fit


def abi_evaluate(native, v, t, s, X, O):
    P, L = v.shape
    V = X.shape[1]
    res = torch.empty((P, O), dtype=torch.float32, device=v.device)
    ws, n = _ws(native, P, L)
    native.check(native.abi().evogp_evaluate(P, L, V, O, _p(v), _p(t), _p(s), _p(X), _p(res), _p(ws), n, _stream()),
                 "evogp_evaluate")
    return res


def abi_batch_forward(native, v, t, s, X, O):
    P, L = v.shape
    N, V = X.shape
    res = torch.empty((P, N, O), dtype=torch.float32, device=v.device)
    ws, n = _ws(native, P, L)
    native.check(native.abi().evogp_batch_forward(P, N, L, V, O, _p(v), _p(t), _p(s), _p(X), _p(res), _p(ws), n,
                                                  _stream()), "evogp_batch_forward")
    return res


def abi_generate(native, pop, L, V, O, out_prob, const_prob, keys, d2l, roul, consts):
    v = torch.empty((pop, L), dtype=torch.float32, device=dev())
    t = torch.empty((pop, L), dtype=torch.int16, device=dev())
    s = torch.empty((pop, L), dtype=torch.int16, device=dev())
    rc = native.abi().evogp_generate(pop, L, V, O, consts.shape[0], out_prob, const_prob, _p(keys), _p(d2l), _p(roul),
                                     _p(consts), _p(v), _p(t), _p(s), _stream())
    native.check(rc, "evogp_generate")
    return v, t, s


def abi_crossover(native, v, t, s, li, ri, lp, rp):
    P, L = v.shape
    Pn = li.shape[0]
    ov = torch.empty((Pn, L), dtype=torch.float32, device=v.device)
    ot = torch.empty((Pn, L), dtype=torch.int16, device=v.device)
    os_ = torch.empty((Pn, L), dtype=torch.int16, device=v.device)
    rc = native.abi().evogp_crossover(P, Pn, L, _p(v), _p(t), _p(s), _p(li), _p(ri), _p(lp), _p(rp), _p(ov), _p(ot),
                                      _p(os_), _stream())
    native.check(rc, "evogp_crossover")
    return ov, ot, os_


def abi_mutate(native, v, t, s, pos, nv, nt, ns):
    P, L = v.shape
    ov, ot, os_ = torch.empty_like(v), torch.empty_like(t), torch.empty_like(s)
    rc = native.abi().evogp_mutate(P, L, _p(v), _p(t), _p(s), _p(pos), _p(nv), _p(nt), _p(ns), _p(ov), _p(ot),
                                   _p(os_), _stream())
    native.check(rc, "evogp_mutate")
    return ov, ot, os_


def same_bits(a, b):
    a, b = a.cpu().numpy(), b if isinstance(b, np.ndarray) else b.cpu().numpy()
    if a.dtype.kind == "f":
        return np.array_equal(a.view(np.uint32), b.view(np.uint32))
    return np.array_equal(a, b)


def assert_close_fitness(got, want, rtol=1e-5, atol=0.0, what=""):
    """NaN == NaN, +-inf == +-inf, otherwise relative tolerance (north star: 1e-5 on fp32 fitness)."""
    got = got.cpu().numpy() if hasattr(got, "cpu") else np.asarray(got)
    want = want.cpu().numpy() if hasattr(want, "cpu") else np.asarray(want)
    assert got.shape == want.shape
    gn, wn = np.isnan(got), np.isnan(want)
    assert np.array_equal(gn, wn), f"{what}: NaN pattern differs at {np.nonzero(gn != wn)[0][:10]}"
    gi, wi = np.isinf(got), np.isinf(want)
    assert np.array_equal(gi, wi) and np.array_equal(got[gi], want[wi]), f"{what}: inf pattern differs"
    m = ~(gn | gi)
    err = np.abs(got[m].astype(np.float64) - want[m]) - atol
    tol = rtol * np.abs(want[m].astype(np.float64))
    bad = err > tol
    assert not bad.any(), (f"{what}: {bad.sum()} of {m.sum()} beyond rtol={rtol}; worst rel "
                           f"{(err[bad] / np.maximum(np.abs(want[m][bad]), 1e-30)).max():.3e}")
