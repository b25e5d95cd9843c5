"""Stand-in for the reference's extension module `evogp.evogp_cuda` (src/evogp/cuda/torch_wrapper.cu:287-307):
importing it registers the `torch.ops.evogp_cuda.tree_*` operators, here from the sm_100a operator library.
The reference front-end does `import evogp.evogp_cuda` in src/evogp/tree/__init__.py:2 and nothing else with it."""
from evogp_b200 import _native

_native.load_ops()
