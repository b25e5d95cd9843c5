"""Loads the two native libraries and exposes the C ABI through ctypes.

Fails loudly (RuntimeError) when a library is missing — the product has no CPU or
pure-PyTorch fallback for any operator.
"""
import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# EVOGP_B200_LIB: another build of the kernel library (developer A/B runs: tools/time_eval.py)
LIB_PATH = os.environ.get("EVOGP_B200_LIB") or os.path.join(_PKG, "lib", "libevogp_b200.so")
OPS_PATH = os.path.join(_PKG, "lib", "evogp_cuda_ops.so")

_abi = None
_ops_loaded = False

# every symbol include/evogp_b200.h declares
ABI_SYMBOLS = (
    "evogp_version", "evogp_last_error", "evogp_launch_count", "evogp_generate", "evogp_mutate", "evogp_crossover",
    "evogp_eval_workspace_bytes", "evogp_eval_set_timing_events", "evogp_evaluate", "evogp_SR_fitness", "evogp_batch_forward",
    "evogp_SR_fitness_host", "evogp_host_release", "evogp_next_generation", "evogp_SR_fitness_scatter", "evogp_debug_lower", "evogp_classification_accuracy", "evogp_generate_philox", "evogp_extract_subtree",
    "evogp_tournament_select", "evogp_push_fitness", "evogp_eval_set_replay_width",
)


def _missing(path):
    return RuntimeError(
        f"{path} not found: build the native libraries first (python -m evogp_b200.build). "
        "evogp_b200 has no CPU / PyTorch fallback."
    )


def abi():
    """ctypes handle of libevogp_b200.so with argtypes set."""
    global _abi
    if _abi is not None:
        return _abi
    if not os.path.exists(LIB_PATH):
        raise _missing(LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, u, i, f, sz = C.c_void_p, C.c_uint, C.c_int, C.c_float, C.c_size_t
    L.evogp_version.restype = i
    L.evogp_last_error.restype = C.c_char_p
    L.evogp_launch_count.restype = C.c_ulonglong
    L.evogp_eval_workspace_bytes.restype = sz
    L.evogp_eval_workspace_bytes.argtypes = [u, u]
    L.evogp_generate.argtypes = [u, u, u, u, u, f, f, vp, vp, vp, vp, vp, vp, vp, vp]
    L.evogp_mutate.argtypes = [i, i] + [vp] * 11
    L.evogp_crossover.argtypes = [i, i, i] + [vp] * 11
    L.evogp_evaluate.argtypes = [u, u, u, u, vp, vp, vp, vp, vp, vp, sz, vp]
    L.evogp_SR_fitness.argtypes = [u, u, u, u, u, i, vp, vp, vp, vp, vp, vp, u, vp, sz, vp]
    L.evogp_batch_forward.argtypes = [u, u, u, u, u, vp, vp, vp, vp, vp, vp, sz, vp]
    L.evogp_SR_fitness_scatter.argtypes = [u, u, u, u, u, i, vp, vp, vp, vp, vp, vp, vp, u, u, vp, sz, vp]
    L.evogp_SR_fitness_scatter.restype = i
    L.evogp_SR_fitness_host.argtypes = [u, u, u, u, u, i, vp, vp, vp, vp, vp, vp, i]
    L.evogp_host_release.restype = None
    L.evogp_next_generation.restype = i
    L.evogp_next_generation.argtypes = [i, i, vp, vp, vp, vp, i, i, f, u, u, u, f, f, vp, vp, vp, vp, vp, vp, vp, vp]
    L.evogp_push_fitness.restype = i
    L.evogp_push_fitness.argtypes = [vp, u, vp, u, u, vp]
    L.evogp_generate_philox.restype = i
    L.evogp_generate_philox.argtypes = [u, u, u, u, u, f, f, vp, vp, vp, vp, vp, vp, vp, vp]
    L.evogp_extract_subtree.restype = i
    L.evogp_extract_subtree.argtypes = [i, i] + [vp] * 8
    L.evogp_tournament_select.restype = i
    L.evogp_tournament_select.argtypes = [i, vp, i, f, i, i, vp, vp, vp]
    L.evogp_classification_accuracy.restype = i
    L.evogp_classification_accuracy.argtypes = [u, u, u, u, u, vp, vp, vp, vp, vp, f, vp, vp, sz, vp]
    L.evogp_debug_lower.restype = i
    L.evogp_debug_lower.argtypes = [u, u, u, u, vp, vp, vp, i, i, vp, sz, vp, vp]
    if hasattr(L, "evogp_eval_set_replay_width") or not os.environ.get("EVOGP_B200_LIB"):   # older A/B builds lack it
        L.evogp_eval_set_replay_width.restype = i
        L.evogp_eval_set_replay_width.argtypes = [i]
    L.evogp_eval_set_timing_events.restype = None
    L.evogp_eval_set_timing_events.argtypes = [vp, vp]
    for name in ("evogp_generate", "evogp_mutate", "evogp_crossover", "evogp_evaluate", "evogp_SR_fitness",
                 "evogp_batch_forward", "evogp_SR_fitness_host"):
        getattr(L, name).restype = i
    _abi = L
    return L


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what}: {abi().evogp_last_error().decode()}")


def load_ops():
    """Registers torch.ops.evogp_cuda.* (idempotent)."""
    global _ops_loaded
    if _ops_loaded:
        return
    if not os.path.exists(OPS_PATH):
        raise _missing(OPS_PATH)
    abi()
    torch.ops.load_library(OPS_PATH)
    _ops_loaded = True


# functions whose bodies the 16-datapoints-per-lane kernel keeps next to its dispatch loop (csrc/gen_fastpath.py HOT_*)
_HOT_FUNCS = frozenset({"+", "-", "*", "/", "neg", "sin", "cos"})


def set_replay_width(datapoints_per_lane: int):
    """0 = automatic, 8 or 16 datapoints per lane in the single-output evaluation kernel (include/evogp_b200.h)."""
    check(abi().evogp_eval_set_replay_width(int(datapoints_per_lane)), "evogp_eval_set_replay_width")


def hint_function_set(names):
    """Called with the function names of the descriptor a forest is generated from: populations that use anything beyond
    + - * / neg sin cos evaluate faster with 8 datapoints per lane (smaller operator bodies).
    EVOGP_REPLAY_K in the environment overrides the hint."""
    if names is None or os.environ.get("EVOGP_REPLAY_K") or not hasattr(abi(), "evogp_eval_set_replay_width"):
        return
    set_replay_width(0 if set(names) <= _HOT_FUNCS else 8)


def launch_count():
    return int(abi().evogp_launch_count())


def device():
    """The device Forest tensors live on: the current CUDA device of this process
    (one process per GPU; torch.cuda.set_device(local_rank) selects it)."""
    return torch.device("cuda", torch.cuda.current_device())
