// eval_exchange.cu — the evaluation kernels with the multi-GPU fitness exchange compiled in (FEAT_EXCHANGE): chunks of
// finished trees are pushed to every rank through peer-mapped memory by the warp that completes them (replay.cuh).
#include "replay.cuh"

namespace evogp {
EVOGP_DEFINE_REPLAY_DISPATCH(launch_replay_exchange, FEAT_EXCHANGE)
}
