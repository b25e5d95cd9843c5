"""Forest — the packed population and the caller of every native operator.

API-compatible with the reference's Forest (src/evogp/tree/forest.py:13-499): same
constructor, static generators, forward / batch_forward / mutate / crossover / SR_fitness,
indexing, concatenation and pickling.  Differences: tensors live on the process's current
CUDA device (not a hard-coded "cuda"), batch_forward is one fused kernel instead of a
P*N-row replication, and index/dtype mistakes raise instead of reinterpreting memory.
"""
import numpy as np
import torch
from torch import Tensor

from .. import _native
from .descriptor import GenerateDescriptor
from .tree import Tree
from .utils import NType, to_cuda_f32

_ops = torch.ops.evogp_cuda
_SR_MODES = {"hybrid parallel": 0, "data parallel": 1, "tree parallel": 2, "auto": 4}   # forest.py:340-347


def _i32(x, device):
    x = to_cuda_f32(x, device)
    return x if x.dtype == torch.int32 else x.to(torch.int32)


class Forest:
    def __init__(self, input_len, output_len, batch_node_value: Tensor, batch_node_type: Tensor,
                 batch_subtree_size: Tensor):
        self.input_len = input_len
        self.output_len = output_len
        self.pop_size, self.max_tree_len = batch_node_value.shape
        want = (self.pop_size, self.max_tree_len)
        assert batch_node_type.shape == want, f"node_type shape should be {want}, but got {batch_node_type.shape}"
        assert batch_subtree_size.shape == want, f"subtree_size shape should be {want}, but got {batch_subtree_size.shape}"
        self.batch_node_value = batch_node_value
        self.batch_node_type = batch_node_type
        self.batch_subtree_size = batch_subtree_size

    # ---- construction -----------------------------------------------------
    def _like(self, value, ntype, size):
        return Forest(self.input_len, self.output_len, value, ntype, size)

    def _arrays(self):
        return (self.batch_node_value.contiguous(), self.batch_node_type.contiguous(),
                self.batch_subtree_size.contiguous())

    @staticmethod
    def random_generate(pop_size: int, descriptor: GenerateDescriptor) -> "Forest":
        assert isinstance(pop_size, int) and pop_size > 0, "pop_size should be a positive integer"
        # same draw as the reference (forest.py:51-58) so seeded runs pick the same keys
        keys = torch.randint(low=0, high=1000000, size=(2,), dtype=torch.uint32, device=_native.device(),
                             requires_grad=False)
        return Forest.generate_with_keys(pop_size, descriptor, keys)

    @staticmethod
    def generate_with_keys(pop_size: int, descriptor: GenerateDescriptor, keys: Tensor, rng: str = "taus88") -> "Forest":
        """rng="taus88": the reference's per-tree generator, bit-identical trees (generate.cu:40-41);
        rng="philox": counter-based Philox4x32-10 draws (`tree_generate_philox`) - same growth rules, other trees."""
        assert rng in ("taus88", "philox"), f"rng should be 'taus88' or 'philox', but got {rng}"
        d = descriptor
        _native.hint_function_set(getattr(d, "func_names", None))      # evaluation kernel width for this function set
        op = _ops.tree_generate if rng == "taus88" else _ops.tree_generate_philox
        v, t, s = op(pop_size, d.max_tree_len, d.input_len, d.output_len, d.const_samples.shape[0],
                                     d.out_prob, d.const_prob, keys, d.depth2leaf_probs, d.roulette_funcs,
                                     d.const_samples)
        return Forest(d.input_len, d.output_len, v, t, s)

    @staticmethod
    def zero_generate(pop_size: int, max_tree_len: int, input_len: int, output_len: int) -> "Forest":
        dev = _native.device()
        v = torch.zeros((pop_size, max_tree_len), dtype=torch.float32, device=dev)
        t = torch.zeros((pop_size, max_tree_len), dtype=torch.int16, device=dev)
        s = torch.zeros((pop_size, max_tree_len), dtype=torch.int16, device=dev)
        t[:, 0] = NType.CONST
        s[:, 0] = 1
        return Forest(input_len, output_len, v, t, s)

    # ---- evaluation -------------------------------------------------------
    def forward(self, x: Tensor) -> Tensor:
        """Tree i on its own input row: x [pop_size, input_len] -> [pop_size, output_len]."""
        x = to_cuda_f32(x, self.batch_node_value.device).contiguous()
        assert x.shape == (self.pop_size, self.input_len), \
            f"x shape should be ({self.pop_size}, {self.input_len}), but got {x.shape}"
        v, t, s = self._arrays()
        return _ops.tree_evaluate(self.pop_size, self.max_tree_len, self.input_len, self.output_len, v, t, s, x)

    def batch_forward(self, x: Tensor) -> Tensor:
        """Every tree on every row: x [N, input_len] -> [pop_size, N, output_len]."""
        x = to_cuda_f32(x, self.batch_node_value.device).contiguous()
        assert x.dim() == 2 and x.shape[1] == self.input_len, f"x shape[1] should be {self.input_len}, but got {tuple(x.shape)}"
        v, t, s = self._arrays()
        return _ops.tree_batch_forward(self.pop_size, x.shape[0], self.max_tree_len, self.input_len, self.output_len,
                                       v, t, s, x)

    def SR_fitness(self, inputs: Tensor, labels: Tensor, use_MSE: bool = True, execute_mode: str = "auto") -> Tensor:
        """(1/N) sum_n sum_o loss(labels - output): [pop_size].  `execute_mode` is accepted for
        compatibility; one kernel serves every mode."""
        dev = self.batch_node_value.device
        inputs = to_cuda_f32(inputs, dev).contiguous()
        labels = to_cuda_f32(labels, dev).contiguous()
        n = inputs.shape[0]
        assert inputs.shape == (n, self.input_len), f"inputs shape should be ({n}, {self.input_len}), but got {inputs.shape}"
        assert labels.shape == (n, self.output_len), f"outputs shape should be ({n}, {self.output_len}), but got {labels.shape}"
        assert execute_mode in _SR_MODES, f"execute_mode should be one of {list(_SR_MODES)}, but got {execute_mode}"
        v, t, s = self._arrays()
        return _ops.tree_SR_fitness(self.pop_size, n, self.max_tree_len, self.input_len, self.output_len, use_MSE,
                                    v, t, s, inputs, labels, _SR_MODES[execute_mode])

    # ---- genetic operators ------------------------------------------------
    def mutate(self, replace_pos: Tensor, new_sub_forest: "Forest") -> "Forest":
        """Row i: subtree at replace_pos[i] replaced by the whole tree new_sub_forest[i]."""
        dev = self.batch_node_value.device
        replace_pos = _i32(replace_pos, dev).contiguous()
        assert replace_pos.shape == (self.pop_size,), f"replace_pos shape should be ({self.pop_size}, ), but got {replace_pos.shape}"
        for attr in ("pop_size", "input_len", "output_len", "max_tree_len"):
            assert getattr(self, attr) == getattr(new_sub_forest, attr), \
                f"{attr} should be {getattr(self, attr)}, but got {getattr(new_sub_forest, attr)}"
        v, t, s = self._arrays()
        nv, nt, ns = new_sub_forest._arrays()
        return self._like(*_ops.tree_mutate(self.pop_size, self.max_tree_len, v, t, s, replace_pos, nv, nt, ns))

    def crossover(self, left_indices: Tensor, right_indices: Tensor, left_pos: Tensor, right_pos: Tensor) -> "Forest":
        """Child i = self[left_indices[i]] with the subtree at left_pos[i] replaced by the
        subtree at right_pos[i] of self[right_indices[i]]."""
        dev = self.batch_node_value.device
        idx = [_i32(a, dev).contiguous() for a in (left_indices, right_indices, left_pos, right_pos)]
        n = idx[0].shape[0]
        for name, a in zip(("left_indices", "right_indices", "left_pos", "right_pos"), idx):
            assert a.shape == (n,), f"{name} shape should be ({n}, ), but got {a.shape}"
        v, t, s = self._arrays()
        return self._like(*_ops.tree_crossover(self.pop_size, n, self.max_tree_len, v, t, s, *idx))

    # ---- container protocol ----------------------------------------------
    def __getitem__(self, index):
        if isinstance(index, int) or (hasattr(index, "shape") and tuple(index.shape) == ()):
            return Tree(self.input_len, self.output_len, self.batch_node_value[index], self.batch_node_type[index],
                        self.batch_subtree_size[index])
        if isinstance(index, (slice, Tensor, np.ndarray)):
            if isinstance(index, Tensor) and index.device != self.batch_node_value.device:
                index = index.to(self.batch_node_value.device)
            return self._like(self.batch_node_value[index], self.batch_node_type[index], self.batch_subtree_size[index])
        raise Exception("Do not support index type {}".format(type(index)))

    def __setitem__(self, index, value):
        if isinstance(index, int):
            assert isinstance(value, Tree), f"value should be Tree when index is int, but got {type(value)}"
            self.batch_node_value[index] = value.node_value
            self.batch_node_type[index] = value.node_type
            self.batch_subtree_size[index] = value.subtree_size
        elif isinstance(index, (slice, Tensor, np.ndarray)):
            assert isinstance(value, Forest), f"value should be Forest when index is slice, but got {type(value)}"
            if isinstance(index, Tensor) and index.device != self.batch_node_value.device:
                index = index.to(self.batch_node_value.device)
            self.batch_node_value[index] = value.batch_node_value
            self.batch_node_type[index] = value.batch_node_type
            self.batch_subtree_size[index] = value.batch_subtree_size
        else:
            raise NotImplementedError

    def __iter__(self):
        return (self[i] for i in range(self.pop_size))

    def __len__(self):
        return self.pop_size

    def __add__(self, other):
        assert other.input_len == self.input_len and other.output_len == self.output_len
        if isinstance(other, Forest):
            parts = (other.batch_node_value, other.batch_node_type, other.batch_subtree_size)
        elif isinstance(other, Tree):
            parts = (other.node_value.unsqueeze(0), other.node_type.unsqueeze(0), other.subtree_size.unsqueeze(0))
        else:
            raise NotImplementedError
        mine = (self.batch_node_value, self.batch_node_type, self.batch_subtree_size)
        return self._like(*(torch.cat([a, b], dim=0) for a, b in zip(mine, parts)))

    def __radd__(self, other):
        return self.__add__(other)

    def __str__(self):
        rows = "".join(f"  {tree}, \n" for tree in self)
        return f"Forest(pop size: {self.pop_size})\n[\n{rows}]"

    __repr__ = __str__

    def __getstate__(self):
        return {"input_len": self.input_len, "output_len": self.output_len,
                "batch_node_value": self.batch_node_value.cpu().numpy(),
                "batch_node_type": self.batch_node_type.cpu().numpy(),
                "batch_subtree_size": self.batch_subtree_size.cpu().numpy()}

    def __setstate__(self, state):
        dev = _native.device()
        self.input_len, self.output_len = state["input_len"], state["output_len"]
        self.pop_size, self.max_tree_len = state["batch_node_value"].shape
        self.batch_node_value = torch.from_numpy(state["batch_node_value"]).to(dev)
        self.batch_node_type = torch.from_numpy(state["batch_node_type"]).to(dev)
        self.batch_subtree_size = torch.from_numpy(state["batch_subtree_size"]).to(dev)
