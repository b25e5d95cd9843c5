"""Tree — one row of a Forest (reference: src/evogp/tree/tree.py:9-324).  Evaluation goes
through the same operators as Forest; presentation helpers stay on the host."""
import torch
from torch import Tensor

from .utils import NType, Func, FUNCS_NAMES, to_cuda_f32, infix

_MODES = {"hybrid parallel": 3, "data parallel": 1, "tree parallel": 2, "auto": 4}   # tree.py:94-101


class Tree:
    def __init__(self, input_len, output_len, node_value: Tensor, node_type: Tensor, subtree_size: Tensor):
        self.input_len = input_len
        self.output_len = output_len
        self.max_tree_len = node_value.shape[0]
        for name, t in (("node_value", node_value), ("node_type", node_type), ("subtree_size", subtree_size)):
            assert t.shape == (self.max_tree_len,), f"{name} shape should be {self.max_tree_len}, but got {t.shape}"
        self.node_value = node_value
        self.node_type = node_type
        self.subtree_size = subtree_size

    @staticmethod
    def random_generate(descriptor):
        from .forest import Forest

        return Forest.random_generate(pop_size=1, descriptor=descriptor)[0]

    def to_forest(self):
        from .forest import Forest

        return Forest(self.input_len, self.output_len, self.node_value[None, :].contiguous(),
                      self.node_type[None, :].contiguous(), self.subtree_size[None, :].contiguous())

    def forward(self, x: Tensor):
        """x: [input_len] -> [output_len], or [N, input_len] -> [N, output_len]."""
        x = to_cuda_f32(x, self.node_value.device)
        assert x.dim() <= 2, f"x dim should be <= 2, but got {x.dim()}"
        single = x.dim() == 1
        if single:
            x = x.unsqueeze(0)
        assert x.shape[1] == self.input_len, f"x shape should be {self.input_len}, but got {x.shape[1]}"
        res = self.to_forest().batch_forward(x)[0]
        return res[0] if single else res

    def SR_fitness(self, inputs: Tensor, labels: Tensor, use_MSE: bool = True, execute_mode: str = "auto"):
        assert execute_mode in _MODES, f"execute_mode should be one of {list(_MODES)}, but got {execute_mode}"
        return self.to_forest().SR_fitness(inputs, labels, use_MSE, execute_mode)

    # ---- presentation -----------------------------------------------------
    def _host(self):
        n = int(self.subtree_size[0])
        return (self.node_value[:n].cpu().numpy(), self.node_type[:n].cpu().numpy(), self.subtree_size[:n].cpu().numpy())

    def to_infix(self, var_names=None):
        v, t, s = self._host()
        return infix(v, t, s, var_names)

    def to_sympy_expr(self, symbol_names=None):
        """sympy expression (list of expressions when output_len > 1: sum of the out nodes per index)."""
        import numpy as np
        import sympy as sp

        v, t, _ = self._host()
        names = symbol_names or [f"x{i}" for i in range(self.input_len)]
        syms = [sp.Symbol(n) for n in names]
        outs = [sp.Integer(0)] * self.output_len
        pos = 0
        un = {Func.SIN: sp.sin, Func.COS: sp.cos, Func.TAN: sp.tan, Func.SINH: sp.sinh, Func.COSH: sp.cosh,
              Func.TANH: sp.tanh, Func.LOG: sp.log, Func.LOOSE_LOG: lambda a: sp.log(sp.Abs(a)), Func.EXP: sp.exp,
              Func.INV: lambda a: 1 / a, Func.LOOSE_INV: lambda a: 1 / a, Func.NEG: lambda a: -a, Func.ABS: sp.Abs,
              Func.SQRT: sp.sqrt, Func.LOOSE_SQRT: lambda a: sp.sqrt(sp.Abs(a))}
        rel = {Func.LT: sp.Lt, Func.GT: sp.Gt, Func.LE: sp.Le, Func.GE: sp.Ge}
        bi = {Func.ADD: lambda a, b: a + b, Func.SUB: lambda a, b: a - b, Func.MUL: lambda a, b: a * b,
              Func.DIV: lambda a, b: a / b, Func.LOOSE_DIV: lambda a, b: a / b, Func.POW: lambda a, b: a**b,
              Func.LOOSE_POW: lambda a, b: sp.Abs(a) ** b, Func.MAX: sp.Max, Func.MIN: sp.Min}

        def rec():
            nonlocal pos
            ty = int(t[pos])
            base, is_out = ty & NType.TYPE_MASK, bool(ty & NType.OUT_NODE)
            val = v[pos]
            pos += 1
            if base == NType.CONST:
                return sp.Float(float(val))
            if base == NType.VAR:
                return syms[int(val)]
            if is_out:
                bits = int(np.float32(val).view(np.uint32))
                fid, oidx = bits & 0xFFFF, bits >> 16
            else:
                fid, oidx = int(val), None
            args = [rec() for _ in range(base - 1)]
            if base == NType.UFUNC:
                e = un.get(fid, lambda a: sp.Integer(0))(args[0])
            elif base == NType.BFUNC:
                if fid in rel:
                    e = sp.Piecewise((1, rel[fid](args[0], args[1])), (-1, True))
                else:
                    e = bi.get(fid, lambda a, b: sp.Integer(0))(args[0], args[1])
            else:
                e = sp.Piecewise((args[1], args[0] > 0), (args[2], True))
            if is_out:
                if oidx < self.output_len:
                    outs[oidx] = outs[oidx] + e
                return args[-1]
            return e

        root = rec()
        return root if self.output_len == 1 else outs

    def __str__(self):
        return self.to_infix()

    def __repr__(self):
        return f"Tree({self.to_infix()})"
