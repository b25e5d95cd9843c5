"""evogp_b200 — B200-native implementation of EvoGP's packed-forest hot path.

Public surface mirrors the reference's (`evogp.tree`, `evogp.algorithm`, `evogp.problem`,
`evogp.pipeline`); the native side is `lib/libevogp_b200.so` (C ABI, include/evogp_b200.h)
and `lib/evogp_cuda_ops.so` (the `torch.ops.evogp_cuda.*` operator library).  There is no
CPU fallback: importing the tree package without the built libraries raises.
"""
__version__ = "0.1.0"


def set_replay_width(datapoints_per_lane: int):
    """Evaluation kernel width: 0 automatic (default), 8 or 16 datapoints per lane; see include/evogp_b200.h
    `evogp_eval_set_replay_width`.  `Forest.random_generate` sets it from the descriptor's function set."""
    from . import _native
    _native.set_replay_width(datapoints_per_lane)
