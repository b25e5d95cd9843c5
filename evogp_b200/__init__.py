"""evogp_b200 — B200-native implementation of EvoGP's packed-forest hot path.

Public surface mirrors the reference's (`evogp.tree`, `evogp.algorithm`, `evogp.problem`,
`evogp.pipeline`); the native side is `lib/libevogp_b200.so` (C ABI, include/evogp_b200.h)
and `lib/evogp_cuda_ops.so` (the `torch.ops.evogp_cuda.*` operator library).  There is no
CPU fallback: importing the tree package without the built libraries raises.
"""
__version__ = "0.1.0"
