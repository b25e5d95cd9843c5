"""Constants shared with the kernels (reference: src/evogp/tree/utils.py:7-136,
src/evogp/cuda/defs.h) and small tensor helpers."""
import torch

DELTA = 1e-9
MAXVAL = 1e9
MAX_STACK = 1024
MAX_FULL_DEPTH = 10


class NType:
    """node_type codes (defs.h:10-22)."""
    VAR, CONST, UFUNC, BFUNC, TFUNC = 0, 1, 2, 3, 4
    TYPE_MASK = 0x7F
    OUT_NODE = 0x80
    UFUNC_OUT, BFUNC_OUT, TFUNC_OUT = UFUNC + OUT_NODE, BFUNC + OUT_NODE, TFUNC + OUT_NODE


# function ids in kernel order (defs.h:24-57): 1 ternary, 13 binary, 15 unary
FUNCS_NAMES = [
    "if",
    "+", "-", "*", "/", "loose_div", "pow", "loose_pow", "max", "min", "<", ">", "<=", ">=",
    "sin", "cos", "tan", "sinh", "cosh", "tanh", "log", "loose_log", "exp", "inv", "loose_inv", "neg", "abs",
    "sqrt", "loose_sqrt",
]


class Func:
    TF_START, BF_START, UF_START, END = 0, 1, 14, 29
    (IF, ADD, SUB, MUL, DIV, LOOSE_DIV, POW, LOOSE_POW, MAX, MIN, LT, GT, LE, GE, SIN, COS, TAN, SINH, COSH, TANH, LOG,
     LOOSE_LOG, EXP, INV, LOOSE_INV, NEG, ABS, SQRT, LOOSE_SQRT) = range(29)


FUNCS = list(range(Func.END))


def func_arity(fid):
    return 3 if fid < Func.BF_START else (2 if fid < Func.UF_START else 1)


def dict2prob(weights):
    """{function name: weight} -> normalised f32[29] probability vector."""
    prob = torch.zeros(Func.END, dtype=torch.float32)
    for name, w in weights.items():
        if name not in FUNCS_NAMES:
            raise AssertionError(f"Unknown function name: {name}, total functions are {FUNCS_NAMES}")
        prob[FUNCS_NAMES.index(name)] = float(w)
    return prob / prob.sum()


def to_cuda_f32(x, device=None):
    """Tensor-or-array-like -> detached CUDA tensor (float32 when converted from python data)."""
    from .. import _native

    device = device or _native.device()
    if isinstance(x, torch.Tensor):
        return x.to(device).detach().requires_grad_(False)
    return torch.tensor(x, dtype=torch.float32, device=device, requires_grad=False)


check_tensor = to_cuda_f32


def randint(size, low, high, dtype=torch.int32, device="cuda", requires_grad=False):
    """low + U[0,1) * (high - low), truncated (reference utils.py:306-310)."""
    r = low + torch.rand(size, device=device, requires_grad=requires_grad) * (high - low)
    return r.to(dtype=dtype)


def infix(value, node_type, subtree_size, var_names=None):
    """Readable infix string of one packed row (CPU lists / numpy)."""
    pos = 0

    def rec():
        nonlocal pos
        t = int(node_type[pos]) & NType.TYPE_MASK
        is_out = bool(int(node_type[pos]) & NType.OUT_NODE)
        v = value[pos]
        pos += 1
        if t == NType.CONST:
            return f"{float(v):.4g}"
        if t == NType.VAR:
            k = int(v)
            return var_names[k] if var_names else f"x{k}"
        if is_out:
            import numpy as np
            bits = int(np.float32(v).view(np.uint32))
            fid, oidx = bits & 0xFFFF, bits >> 16
        else:
            fid, oidx = int(v), None
        name = FUNCS_NAMES[fid] if 0 <= fid < Func.END else f"f{fid}"
        args = [rec() for _ in range(t - 1)]
        if t == NType.BFUNC and name in ("+", "-", "*", "/", "<", ">", "<=", ">="):
            s = f"({args[0]} {name} {args[1]})"
        else:
            s = f"{name}({', '.join(args)})"
        return f"out{oidx}[{s}]" if is_out else s

    return rec()
