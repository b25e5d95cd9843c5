"""ctypes front-end of the CPU oracle (numpy in / numpy out) and of the
reference's own CUDA kernels (``oracle/_ref``).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "evogp_oracle.c")
_SO = os.path.join(_HERE, "liboracle.so")
_REF_DIR = os.path.join(_HERE, "_ref")
_REF_SO = os.path.join(_REF_DIR, "libevogp_ref.so")
_REFERENCE_ROOT = "/root/reference"

_lib = None
_ref = None

f32p = C.POINTER(C.c_float)
i16p = C.POINTER(C.c_int16)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)


def build(force=False):
    """gcc the C restatement into oracle/liboracle.so."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O2", "-march=native", "-fopenmp", "-fno-fast-math", "-ffp-contract=off", "-shared", "-fPIC",
           "-o", _SO, _SRC, "-lm"]
    subprocess.check_call(cmd)
    return _SO


def build_ref(force=False):
    """Compile the reference's own three .cu files, where they lie, with plain
    nvcc (no torch, no reference build system) -> oracle/_ref/libevogp_ref.so.
    Only possible where /root/reference exists (this container)."""
    if os.path.exists(_REF_SO) and not force:
        return _REF_SO
    if not os.path.isdir(_REFERENCE_ROOT):
        return None
    subprocess.check_call(["bash", os.path.join(_HERE, "build_ref.sh")])
    return _REF_SO


def build_reference_package(force=False):
    """Installs the reference's own package, unmodified, into the git-ignored baseline/_ref (what `bench.py --impl
    reference` and tests/test_dropin.py run): pip install from a copy of /root/reference (the tree is read-only and the
    build writes into it), offline, no dependency resolution.  ~4 minutes (its CUDA extension); skipped when already
    there or where /root/reference does not exist (the GPU box uses the prebuilt files)."""
    import shutil
    import sys
    import tempfile

    target = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")
    if os.path.isdir(os.path.join(target, "evogp")) and not force:
        return target
    if not os.path.isdir(_REFERENCE_ROOT):
        return None
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "refsrc")
        shutil.copytree(_REFERENCE_ROOT, src)
        env = dict(os.environ, TORCH_CUDA_ARCH_LIST="10.0a")
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                               "--find-links", "/opt/wheelhouse", "--target", target, src], env=env,
                              stdout=subprocess.DEVNULL)
    return target


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oracle_hash.restype = C.c_uint32
        L.oracle_hash.argtypes = [C.c_uint32] * 3
        L.oracle_taus88_nth.restype = C.c_uint32
        L.oracle_taus88_nth.argtypes = [C.c_uint32, C.c_int]
        L.oracle_taus88_draws.argtypes = [C.c_uint32, C.c_int, u32p]
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def max_threads():
    return int(lib().oracle_max_threads())


def hash32(n, k1, k2):
    return int(lib().oracle_hash(n, k1, k2))


def taus88_nth(seed, n):
    return int(lib().oracle_taus88_nth(seed, n))


def taus88_draws(seed, n):
    out = np.zeros(n, np.uint32)
    lib().oracle_taus88_draws(seed, n, _p(out, u32p))
    return out


def evaluate(value, ntype, size, variables, out_len, nthreads=1):
    value, ntype, size = _c(value, np.float32), _c(ntype, np.int16), _c(size, np.int16)
    variables = _c(variables, np.float32)
    P, L = value.shape
    V = variables.shape[1]
    res = np.zeros((P, out_len), np.float32)
    lib().oracle_evaluate(C.c_uint(P), C.c_uint(L), C.c_uint(V), C.c_uint(out_len), _p(value, f32p), _p(ntype, i16p),
                          _p(size, i16p), _p(variables, f32p), _p(res, f32p), C.c_int(nthreads))
    return res


def batch_forward(value, ntype, size, variables, out_len, nthreads=1):
    value, ntype, size = _c(value, np.float32), _c(ntype, np.int16), _c(size, np.int16)
    variables = _c(variables, np.float32)
    P, L = value.shape
    N, V = variables.shape
    res = np.zeros((P, N, out_len), np.float32)
    lib().oracle_batch_forward(C.c_uint(P), C.c_uint(N), C.c_uint(L), C.c_uint(V), C.c_uint(out_len),
                               _p(value, f32p), _p(ntype, i16p), _p(size, i16p), _p(variables, f32p), _p(res, f32p),
                               C.c_int(nthreads))
    return res


def sr_fitness(value, ntype, size, variables, labels, use_mse=True, nthreads=1):
    value, ntype, size = _c(value, np.float32), _c(ntype, np.int16), _c(size, np.int16)
    variables, labels = _c(variables, np.float32), _c(labels, np.float32)
    P, L = value.shape
    N, V = variables.shape
    O = labels.shape[1]
    fit = np.zeros(P, np.float32)
    lib().oracle_sr_fitness(C.c_uint(P), C.c_uint(N), C.c_uint(L), C.c_uint(V), C.c_uint(O), C.c_int(int(use_mse)),
                            _p(value, f32p), _p(ntype, i16p), _p(size, i16p), _p(variables, f32p), _p(labels, f32p),
                            _p(fit, f32p), C.c_int(nthreads))
    return fit


def generate(pop, gp_len, var_len, out_len, out_prob, const_prob, keys, depth2leaf, roulette, const_samples,
             nthreads=1):
    keys = _c(keys, np.uint32)
    depth2leaf, roulette = _c(depth2leaf, np.float32), _c(roulette, np.float32)
    const_samples = _c(const_samples, np.float32)
    assert depth2leaf.shape == (10,) and roulette.shape == (29,)
    v = np.zeros((pop, gp_len), np.float32)
    t = np.zeros((pop, gp_len), np.int16)
    s = np.zeros((pop, gp_len), np.int16)
    lib().oracle_generate(C.c_uint(pop), C.c_uint(gp_len), C.c_uint(var_len), C.c_uint(out_len),
                          C.c_uint(const_samples.shape[0]), C.c_float(out_prob), C.c_float(const_prob),
                          _p(keys, u32p), _p(depth2leaf, f32p), _p(roulette, f32p), _p(const_samples, f32p),
                          _p(v, f32p), _p(t, i16p), _p(s, i16p), C.c_int(nthreads))
    return v, t, s


def generate_philox(pop, gp_len, var_len, out_len, out_prob, const_prob, keys, depth2leaf, roulette, const_samples, nthreads=1):
    """evogp_generate_philox restated: same growth, Philox4x32-10 counter-based draws."""
    keys = _c(keys, np.uint32)
    depth2leaf, roulette = _c(depth2leaf, np.float32), _c(roulette, np.float32)
    const_samples = _c(const_samples, np.float32)
    v = np.zeros((pop, gp_len), np.float32)
    t = np.zeros((pop, gp_len), np.int16)
    s = np.zeros((pop, gp_len), np.int16)
    lib().oracle_generate_philox(C.c_uint(pop), C.c_uint(gp_len), C.c_uint(var_len), C.c_uint(out_len),
                                 C.c_uint(const_samples.shape[0]), C.c_float(out_prob), C.c_float(const_prob),
                                 _p(keys, u32p), _p(depth2leaf, f32p), _p(roulette, f32p), _p(const_samples, f32p),
                                 _p(v, f32p), _p(t, i16p), _p(s, i16p), C.c_int(nthreads))
    return v, t, s


def philox(c0, c1, k0, k1):
    out = np.zeros(4, np.uint32)
    lib().oracle_philox(C.c_uint32(c0), C.c_uint32(c1), C.c_uint32(k0), C.c_uint32(k1), _p(out, u32p))
    return out


def extract_subtree(value, ntype, size, pos):
    value, ntype, size = _c(value, np.float32), _c(ntype, np.int16), _c(size, np.int16)
    pos = _c(pos, np.int32)
    P, L = value.shape
    v, t, s = np.zeros_like(value), np.zeros_like(ntype), np.zeros_like(size)
    lib().oracle_extract_subtree(C.c_int(P), C.c_int(L), _p(value, f32p), _p(ntype, i16p), _p(size, i16p), _p(pos, i32p),
                                 _p(v, f32p), _p(t, i16p), _p(s, i16p))
    return v, t, s


def tournament(fitness, t_size, best_p, replace, count, keys):
    fitness, keys = _c(fitness, np.float32), _c(keys, np.uint32)
    out = np.zeros(count, np.int32)
    lib().oracle_tournament(C.c_int(fitness.shape[0]), _p(fitness, f32p), C.c_int(t_size), C.c_float(best_p), C.c_int(int(replace)),
                            C.c_int(count), _p(keys, u32p), _p(out, i32p))
    return out


def feistel_perm(n, round_, keys):
    keys = _c(keys, np.uint32)
    out = np.zeros(n, np.uint32)
    lib().oracle_feistel_perm(C.c_uint32(n), C.c_uint32(round_), _p(keys, u32p), _p(out, u32p))
    return out


def next_generation(value, ntype, size, order, elite, survivors, rate, var_len, out_len, out_prob, const_prob, depth2leaf,
                    roulette, const_samples, keys, nthreads=1):
    """evogp_next_generation restated (elitism + crossover + mutation of a whole generation, Philox draws)."""
    value, ntype, size = _c(value, np.float32), _c(ntype, np.int16), _c(size, np.int16)
    order = _c(order, np.int64)
    depth2leaf, roulette, const_samples = _c(depth2leaf, np.float32), _c(roulette, np.float32), _c(const_samples, np.float32)
    keys = _c(keys, np.uint32)
    P, L = value.shape
    v, t, s = np.zeros_like(value), np.zeros_like(ntype), np.zeros_like(size)
    lib().oracle_next_generation(C.c_int(P), C.c_int(L), _p(value, f32p), _p(ntype, i16p), _p(size, i16p),
                                 order.ctypes.data_as(C.POINTER(C.c_longlong)), C.c_int(elite), C.c_int(survivors), C.c_float(rate),
                                 C.c_uint(var_len), C.c_uint(out_len), C.c_uint(const_samples.shape[0]), C.c_float(out_prob),
                                 C.c_float(const_prob), _p(depth2leaf, f32p), _p(roulette, f32p), _p(const_samples, f32p),
                                 _p(keys, u32p), _p(v, f32p), _p(t, i16p), _p(s, i16p), C.c_int(nthreads))
    return v, t, s


def crossover(value, ntype, size, left_idx, right_idx, left_node, right_node, nthreads=1):
    value, ntype, size = _c(value, np.float32), _c(ntype, np.int16), _c(size, np.int16)
    li, ri, ln, rn = (_c(a, np.int32) for a in (left_idx, right_idx, left_node, right_node))
    P, L = value.shape
    Pn = li.shape[0]
    v = np.zeros((Pn, L), np.float32)
    t = np.zeros((Pn, L), np.int16)
    s = np.zeros((Pn, L), np.int16)
    lib().oracle_crossover(C.c_int(P), C.c_int(Pn), C.c_int(L), _p(value, f32p), _p(ntype, i16p), _p(size, i16p),
                           _p(li, i32p), _p(ri, i32p), _p(ln, i32p), _p(rn, i32p), _p(v, f32p), _p(t, i16p),
                           _p(s, i16p), C.c_int(nthreads))
    return v, t, s


def mutate(value, ntype, size, mut_idx, nvalue, ntype_new, nsize, nthreads=1):
    value, ntype, size = _c(value, np.float32), _c(ntype, np.int16), _c(size, np.int16)
    nvalue, ntype_new, nsize = _c(nvalue, np.float32), _c(ntype_new, np.int16), _c(nsize, np.int16)
    mi = _c(mut_idx, np.int32)
    P, L = value.shape
    v = np.zeros((P, L), np.float32)
    t = np.zeros((P, L), np.int16)
    s = np.zeros((P, L), np.int16)
    lib().oracle_mutate(C.c_int(P), C.c_int(L), _p(value, f32p), _p(ntype, i16p), _p(size, i16p), _p(mi, i32p),
                        _p(nvalue, f32p), _p(ntype_new, i16p), _p(nsize, i16p), _p(v, f32p), _p(t, i16p),
                        _p(s, i16p), C.c_int(nthreads))
    return v, t, s


def check_forest(value, ntype, size, input_len=None, output_len=1):
    """Structural invariants of a packed forest (the checks the reference keeps
    private in tree/tree.py:361-411): every row's prefix closes under arity,
    size[0] is the real length and every size[i] equals 1 + sizes of children.
    Returns the per-row length array; raises AssertionError on the first violation."""
    value, ntype, size = np.asarray(value), np.asarray(ntype), np.asarray(size)
    P, L = ntype.shape
    lens = size[:, 0].astype(np.int64)
    assert (lens >= 1).all() and (lens <= L).all(), "size[:,0] out of range"
    base = ntype.astype(np.int64) & 0x7F
    assert (base <= 4).all(), "unknown node type"
    arity = np.where(base <= 1, 0, base - 1)
    cols = np.arange(L)[None, :]
    valid = cols < lens[:, None]
    # reverse scan: size[i] = 1 + sum of children sizes, children at i+1, i+1+size[i+1], ...
    calc = np.zeros((P, L + 1), np.int64)
    rows = np.arange(P)
    for i in range(L - 1, -1, -1):
        a = np.where(valid[:, i], arity[:, i], 0)
        tot = np.ones(P, np.int64)
        nxt = np.full(P, i + 1, np.int64)
        for k in range(3):
            use = a > k
            idx = np.minimum(nxt, L)
            sz = np.where(use, calc[rows, idx], 0)
            assert (~use | ((nxt < lens) & (sz > 0))).all(), f"node {i}: child {k} runs past the tree"
            tot += sz
            nxt = nxt + sz
        calc[:, i] = np.where(valid[:, i], tot, 0)
    got = np.where(valid, size.astype(np.int64), 0)
    assert (calc[:, :L] == got).all(), "subtree_size inconsistent with prefix arities"
    if input_len is not None:
        is_var = valid & (base == 0)
        vv = value[is_var]
        assert ((vv >= 0) & (vv < input_len) & (vv == np.floor(vv))).all(), "variable index out of range"
    return lens


# ---------------------------------------------------------------------------
# the reference's own CUDA kernels (GPU box only)
# ---------------------------------------------------------------------------

def ref_gpu_available():
    return os.path.exists(_REF_SO)


class _RefGPU:
    """Raw-device-pointer bindings of the reference's kernel.h entry points
    (src/evogp/cuda/kernel.h:23-97), C++-mangled, launched on the legacy
    default stream exactly as the reference does.  Arguments are torch CUDA
    tensors; callers synchronize."""

    def __init__(self, path):
        self.dll = C.CDLL(path)
        g = self.dll
        self._sr = getattr(g, "_Z10SR_fitnessjjjjjbPKfPKsS2_S0_S0_Pfj")
        self._ev = getattr(g, "_Z8evaluatejjjjPKfPKsS2_S0_Pf")
        self._gen = getattr(g, "_Z8generatejjjjjffPKjPKfS2_S2_PfPsS4_")
        self._mut = getattr(g, "_Z6mutateiiPKfPKsS2_PKiS0_S2_S2_PfPsS6_")
        self._cx = getattr(g, "_Z9crossoveriiiPKfPKsS2_PKiS4_S4_S4_PfPsS6_")
        for f in (self._sr, self._ev, self._gen, self._mut, self._cx):
            f.restype = None

    @staticmethod
    def _dp(t):
        return C.c_void_p(t.data_ptr())

    def sr_fitness(self, value, ntype, size, variables, labels, use_mse=True, kernel_type=4):
        import torch
        P, L = value.shape
        N, V = variables.shape
        O = labels.shape[1]
        fit = torch.empty(P, dtype=torch.float32, device=value.device)
        self._sr(C.c_uint(P), C.c_uint(N), C.c_uint(L), C.c_uint(V), C.c_uint(O), C.c_bool(use_mse), self._dp(value),
                 self._dp(ntype), self._dp(size), self._dp(variables), self._dp(labels), self._dp(fit),
                 C.c_uint(kernel_type))
        return fit

    def evaluate(self, value, ntype, size, variables, out_len):
        import torch
        P, L = value.shape
        V = variables.shape[1]
        res = torch.empty((P, out_len), dtype=torch.float32, device=value.device)
        self._ev(C.c_uint(P), C.c_uint(L), C.c_uint(V), C.c_uint(out_len), self._dp(value), self._dp(ntype),
                 self._dp(size), self._dp(variables), self._dp(res))
        return res

    def generate(self, pop, gp_len, var_len, out_len, out_prob, const_prob, keys, depth2leaf, roulette, const_samples):
        import torch
        dev = keys.device
        # zeros (not empty): the reference only writes valid prefixes
        v = torch.zeros((pop, gp_len), dtype=torch.float32, device=dev)
        t = torch.zeros((pop, gp_len), dtype=torch.int16, device=dev)
        s = torch.zeros((pop, gp_len), dtype=torch.int16, device=dev)
        self._gen(C.c_uint(pop), C.c_uint(gp_len), C.c_uint(var_len), C.c_uint(out_len),
                  C.c_uint(const_samples.shape[0]), C.c_float(out_prob), C.c_float(const_prob), self._dp(keys),
                  self._dp(depth2leaf), self._dp(roulette), self._dp(const_samples), self._dp(v), self._dp(t),
                  self._dp(s))
        return v, t, s

    def mutate(self, value, ntype, size, mut_idx, nvalue, ntype_new, nsize):
        import torch
        P, L = value.shape
        v, t, s = torch.zeros_like(value), torch.zeros_like(ntype), torch.zeros_like(size)
        self._mut(C.c_int(P), C.c_int(L), self._dp(value), self._dp(ntype), self._dp(size), self._dp(mut_idx),
                  self._dp(nvalue), self._dp(ntype_new), self._dp(nsize), self._dp(v), self._dp(t), self._dp(s))
        return v, t, s

    def crossover(self, value, ntype, size, left_idx, right_idx, left_node, right_node):
        import torch
        P, L = value.shape
        Pn = left_idx.shape[0]
        dev = value.device
        v = torch.zeros((Pn, L), dtype=torch.float32, device=dev)
        t = torch.zeros((Pn, L), dtype=torch.int16, device=dev)
        s = torch.zeros((Pn, L), dtype=torch.int16, device=dev)
        self._cx(C.c_int(P), C.c_int(Pn), C.c_int(L), self._dp(value), self._dp(ntype), self._dp(size),
                 self._dp(left_idx), self._dp(right_idx), self._dp(left_node), self._dp(right_node), self._dp(v),
                 self._dp(t), self._dp(s))
        return v, t, s


def ref_gpu():
    global _ref
    if _ref is None:
        if not os.path.exists(_REF_SO):
            raise FileNotFoundError(f"{_REF_SO} missing: run oracle/build_ref.sh where /root/reference exists")
        _ref = _RefGPU(_REF_SO)
    return _ref
