"""GenerateDescriptor — the inputs of random tree generation.

Same constructor signature, attributes and `update()` semantics as the reference
(src/evogp/tree/descriptor.py:42-188); built fresh.
"""
import warnings
from typing import Optional, Tuple, Union

import torch
from torch import Tensor

from .. import _native
from .utils import MAX_STACK, MAX_FULL_DEPTH, FUNCS_NAMES, Func, dict2prob, func_arity, to_cuda_f32


def full_tree_len(max_arity, layers):
    """Nodes of a complete `max_arity`-ary tree with `layers` layers."""
    return layers if max_arity <= 1 else (max_arity**layers - 1) // (max_arity - 1)


def depth_schedule(max_tree_len, using_funcs, max_layer_cnt, layer_leaf_prob, device):
    """depth2leaf_probs: `layer_leaf_prob` on every non-final layer, 1.0 afterwards
    (descriptor.py:8-39), after checking the worst-case tree fits `max_tree_len`."""
    arity = max((func_arity(FUNCS_NAMES.index(f)) for f in using_funcs), default=0)
    need = full_tree_len(arity, max_layer_cnt)
    assert max_tree_len >= need, (
        f"a full tree of {max_layer_cnt} layers over functions of arity up to {arity} has {need} nodes; "
        f"max_tree_len {max_tree_len} cannot hold it"
    )
    inner = max_layer_cnt - 1
    return torch.tensor([layer_leaf_prob] * inner + [1.0] * (MAX_FULL_DEPTH - inner), dtype=torch.float32, device=device)


class GenerateDescriptor:
    def __init__(
        self,
        max_tree_len: int,
        input_len: int,
        output_len: int,
        const_prob: float = 0.5,
        out_prob: float = 0.5,
        depth2leaf_probs: Optional[Tensor] = None,
        roulette_funcs: Optional[Tensor] = None,
        const_samples: Optional[Union[list, Tensor]] = None,
        using_funcs: Optional[Union[dict, list]] = None,
        max_layer_cnt: Optional[int] = None,
        layer_leaf_prob: Optional[float] = 0.2,
        const_range: Optional[Tuple[float, float]] = None,
        sample_cnt: Optional[int] = None,
    ):
        self._ctor_kwargs = {k: v for k, v in locals().items() if k not in ("self", "__class__")}
        dev = _native.device()

        assert max_tree_len <= MAX_STACK, f"max_tree_len {max_tree_len} exceeds the kernels' row limit ({MAX_STACK})"
        assert isinstance(input_len, int) and input_len > 0, "input_len: expected a positive int"
        assert isinstance(output_len, int) and output_len > 0, "output_len: expected a positive int"
        assert 0.0 <= const_prob <= 1.0, "const_prob is a probability: 0 <= const_prob <= 1"
        assert 0.0 <= out_prob <= 1.0, "out_prob is a probability: 0 <= out_prob <= 1"
        if output_len > 1 and out_prob == 0.0:
            warnings.warn(f"output_len={output_len} > 1, but out_prob={out_prob} is 0.0.")

        if depth2leaf_probs is None:
            assert max_layer_cnt is not None, "give either depth2leaf_probs or max_layer_cnt (+ layer_leaf_prob)"
            assert layer_leaf_prob is not None, "layer_leaf_prob is needed to build depth2leaf_probs from max_layer_cnt"
            funcs = list(using_funcs) if using_funcs is not None else []
            depth2leaf_probs = depth_schedule(max_tree_len, funcs, max_layer_cnt, layer_leaf_prob, dev)

        self.roulette_ufuncs = self.roulette_bfuncs = self.roulette_tfuncs = None
        self.func_names = None     # names of the functions this descriptor can draw (None: unknown)
        if roulette_funcs is not None:
            cum = torch.as_tensor(roulette_funcs).detach().float().cpu()
            steps = torch.diff(cum, prepend=torch.zeros(1))
            self.func_names = tuple(FUNCS_NAMES[k] for k in range(min(len(FUNCS_NAMES), steps.numel())) if steps[k] > 0)
        if roulette_funcs is None:
            assert using_funcs is not None, "give either roulette_funcs or using_funcs"
            assert isinstance(using_funcs, (dict, list)), "using_funcs: a list of function names, or a dict name -> weight"
            weights = using_funcs if isinstance(using_funcs, dict) else {f: 1.0 for f in using_funcs}
            self.func_names = tuple(f for f, wgt in weights.items() if wgt > 0)
            prob = dict2prob(weights)
            roulette_funcs = torch.cumsum(prob, dim=0, dtype=torch.float32).to(dev)

            def class_roulette(lo, hi):
                p = torch.zeros_like(prob)
                p[lo:hi] = prob[lo:hi]
                return torch.cumsum(p, dim=0, dtype=torch.float32).to(dev)

            self.roulette_tfuncs = class_roulette(Func.TF_START, Func.BF_START)
            self.roulette_bfuncs = class_roulette(Func.BF_START, Func.UF_START)
            self.roulette_ufuncs = class_roulette(Func.UF_START, Func.END)

        if const_samples is None:
            assert const_range is not None, "give either const_samples or const_range (+ sample_cnt)"
            assert sample_cnt is not None, "sample_cnt is needed to draw constants from const_range"
            lo, hi = const_range
            const_samples = torch.rand(sample_cnt, device=dev) * (hi - lo) + lo
        if isinstance(const_samples, list):
            const_samples = torch.tensor(const_samples, dtype=torch.float32, device=dev, requires_grad=False)

        depth2leaf_probs = to_cuda_f32(depth2leaf_probs, dev).to(torch.float32).contiguous()
        roulette_funcs = to_cuda_f32(roulette_funcs, dev).to(torch.float32).contiguous()
        const_samples = to_cuda_f32(const_samples, dev).to(torch.float32).contiguous()
        assert depth2leaf_probs.shape == (MAX_FULL_DEPTH,), \
            f"depth2leaf_probs must hold {MAX_FULL_DEPTH} probabilities, got shape {tuple(depth2leaf_probs.shape)}"
        assert roulette_funcs.shape == (Func.END,), f"roulette_funcs must hold {Func.END} cumulative probabilities, got shape {tuple(roulette_funcs.shape)}"
        assert const_samples.dim() == 1, f"const_samples must be one-dimensional, got {const_samples.dim()} dimensions"

        self.max_tree_len = max_tree_len
        self.input_len = input_len
        self.output_len = output_len
        self.const_prob = const_prob
        self.out_prob = out_prob
        self.depth2leaf_probs = depth2leaf_probs
        self.roulette_funcs = roulette_funcs
        self.const_samples = const_samples

    def update(self, **kwargs):
        """A NEW descriptor built from the original constructor arguments overridden by kwargs."""
        merged = dict(self._ctor_kwargs)
        merged.update(kwargs)
        return type(self)(**merged)

    def __str__(self):
        keys = ("max_tree_len", "input_len", "output_len", "const_prob", "out_prob", "depth2leaf_probs",
                "roulette_funcs", "const_samples")
        return "".join(f"{k}: {getattr(self, k)}\n" for k in keys)
