// eval_acc.cu — the evaluation kernels with the classification-accuracy epilogue compiled in (FEAT_ACC, MODE_ACC):
// prediction and comparison with the label happen inside the kernel (problem/classification.py:54-67), replay.cuh.
#include "replay.cuh"

namespace evogp {
EVOGP_DEFINE_REPLAY_DISPATCH(launch_replay_acc, FEAT_ACC)
}
