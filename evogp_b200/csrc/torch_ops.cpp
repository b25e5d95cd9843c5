// torch_ops.cpp — the `evogp_cuda` operator library, rebuilt on libevogp_b200.so.
//
// Same five schemas the reference registers (src/evogp/cuda/torch_wrapper.cu:291-299),
// CUDA dispatch key only (:301-307), same argument checks (:7-17, :48-60, ...), so
// `torch.ops.evogp_cuda.tree_*` call sites in a Forest front-end keep working.
// Differences, all deliberate:
//   * kernels are enqueued on torch's CURRENT stream of the tensors' device (the
//     reference uses the legacy default stream and device 0 implicitly);
//   * a failing launch raises (the reference's checks are commented out, :84,135,...);
//   * index tensors must be int32 and node arrays f32/i16 — checked, not assumed;
//   * one extra op, tree_batch_forward, the fused form of Forest.batch_forward;
//   * tree_generate reads the function roulette it is given once (29 floats, cached by tensor identity) and sets the
//     evaluation kernel's width for that function set (evogp_eval_set_replay_width) - the one place where a front-end
//     that knows nothing of this library (the reference's) tells it which functions a run uses.
// This file is plain C++ (g++): all device code lives behind the C ABI.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAGraphsC10Utils.h>
#include <ATen/ATen.h>
#include <torch/library.h>
#include <cstdlib>
#include <mutex>
#include <tuple>
#include "../../include/evogp_b200.h"

namespace {

using at::Tensor;
using Tensor3 = std::tuple<Tensor, Tensor, Tensor>;

void check_tensor(const Tensor &t, c10::IntArrayRef shape, c10::ScalarType dtype, const char *name) {
    TORCH_CHECK(t.is_cuda() && t.is_contiguous(), name, " must be a contiguous CUDA tensor");
    TORCH_CHECK(t.sizes() == shape, name, " must have shape ", shape, ", but got shape ", t.sizes());
    TORCH_CHECK(t.scalar_type() == dtype, name, " must have dtype ", dtype, ", but got ", t.scalar_type());
}

void check_rc(int rc, const char *op) { TORCH_CHECK(rc == EVOGP_OK, "evogp_cuda::", op, ": ", evogp_last_error()); }

void *cur_stream(const Tensor &t) { return at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

Tensor3 alloc_forest(int64_t rows, int64_t len, const Tensor &like) {
    auto f32 = at::TensorOptions().dtype(at::kFloat).device(like.device()).requires_grad(false);
    return {at::empty({rows, len}, f32), at::empty({rows, len}, f32.dtype(at::kShort)),
            at::empty({rows, len}, f32.dtype(at::kShort))};
}

Tensor eval_workspace(int64_t pop, int64_t len, const Tensor &like, size_t &bytes) {
    bytes = evogp_eval_workspace_bytes((unsigned)pop, (unsigned)len);
    return at::empty({(int64_t)bytes}, at::TensorOptions().dtype(at::kByte).device(like.device()));
}

// Populations of + - * / neg sin cos evaluate fastest 16 datapoints per lane, every other function set 8 per lane
// (include/evogp_b200.h evogp_eval_set_replay_width).  The roulette is a cumulative sum: function k is in use when it
// rises at k.  One device-to-host copy per distinct roulette tensor (identity = storage address + version counter);
// EVOGP_REPLAY_K in the environment keeps the width fixed.
void hint_replay_width(const Tensor &roulette) {
    static std::mutex mu;
    static const void *seen_ptr = nullptr;
    static uint32_t seen_version = 0;
    if (std::getenv("EVOGP_REPLAY_K") != nullptr) return;
    if (c10::cuda::currentStreamCaptureStatusMayInitCtx() != c10::cuda::CaptureStatus::None) return;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (roulette.data_ptr() == seen_ptr && roulette._version() == seen_version) return;
        seen_ptr = roulette.data_ptr();
        seen_version = roulette._version();
    }
    const Tensor host = roulette.to(at::kCPU);
    const float *cum = host.data_ptr<float>();
    bool rare = false;
    float below = 0.0f;
    for (int k = 0; k < EVOGP_FUNC_END; ++k) {
        const bool hot = (k >= 1 && k <= 4) || k == 14 || k == 15 || k == 25;      // + - * /, sin, cos, neg (kernel.h Function)
        if (cum[k] > below && !hot) rare = true;
        below = cum[k] > below ? cum[k] : below;
    }
    evogp_eval_set_replay_width(rare ? 8 : 0);
}

Tensor3 generate_impl(bool philox, int64_t pop_size, int64_t gp_len, int64_t var_len, int64_t out_len, int64_t const_samples_len,
                      double out_prob, double const_prob, Tensor keys, Tensor depth2leaf_probs, Tensor roulette_funcs,
                      Tensor const_samples) {
    TORCH_CHECK(pop_size > 0, "pop_size must larger than 0, but got ", pop_size);
    TORCH_CHECK(0 < gp_len && gp_len <= EVOGP_MAX_STACK, "gp_len must be in range (0, ", EVOGP_MAX_STACK, "], but got ", gp_len);
    TORCH_CHECK(0 < var_len, "var_len must larger than 0, but got ", var_len);
    TORCH_CHECK(0 < out_len, "out_len must larger than 0, but got ", out_len);
    TORCH_CHECK(0 < const_samples_len, "const_samples_len must larger than 0, but got ", const_samples_len);
    TORCH_CHECK(0 <= out_prob && out_prob <= 1, "out_prob must be in range [0, 1], but got ", out_prob);
    TORCH_CHECK(0 <= const_prob && const_prob <= 1, "const_prob must be in range [0, 1], but got ", const_prob);
    check_tensor(keys, {2}, at::kUInt32, "keys");
    check_tensor(depth2leaf_probs, {EVOGP_MAX_FULL_DEPTH}, at::kFloat, "depth2leaf_probs");
    check_tensor(roulette_funcs, {EVOGP_FUNC_END}, at::kFloat, "roulette_funcs");
    check_tensor(const_samples, {const_samples_len}, at::kFloat, "const_samples");
    c10::cuda::CUDAGuard guard(keys.device());
    hint_replay_width(roulette_funcs);
    auto out = alloc_forest(pop_size, gp_len, keys);
    auto fn = philox ? &evogp_generate_philox : &evogp_generate;
    check_rc(fn((unsigned)pop_size, (unsigned)gp_len, (unsigned)var_len, (unsigned)out_len, (unsigned)const_samples_len,
                (float)out_prob, (float)const_prob, static_cast<const unsigned *>(keys.data_ptr()),
                depth2leaf_probs.data_ptr<float>(), roulette_funcs.data_ptr<float>(), const_samples.data_ptr<float>(),
                std::get<0>(out).data_ptr<float>(), std::get<1>(out).data_ptr<int16_t>(), std::get<2>(out).data_ptr<int16_t>(),
                cur_stream(keys)),
             philox ? "tree_generate_philox" : "tree_generate");
    return out;
}

Tensor3 tree_generate(int64_t pop_size, int64_t gp_len, int64_t var_len, int64_t out_len, int64_t const_samples_len,
                      double out_prob, double const_prob, Tensor keys, Tensor depth2leaf_probs, Tensor roulette_funcs,
                      Tensor const_samples) {
    return generate_impl(false, pop_size, gp_len, var_len, out_len, const_samples_len, out_prob, const_prob, keys, depth2leaf_probs,
                         roulette_funcs, const_samples);
}

Tensor3 tree_generate_philox(int64_t pop_size, int64_t gp_len, int64_t var_len, int64_t out_len, int64_t const_samples_len,
                             double out_prob, double const_prob, Tensor keys, Tensor depth2leaf_probs, Tensor roulette_funcs,
                             Tensor const_samples) {
    return generate_impl(true, pop_size, gp_len, var_len, out_len, const_samples_len, out_prob, const_prob, keys, depth2leaf_probs,
                         roulette_funcs, const_samples);
}

Tensor3 tree_extract_subtree(int64_t pop_size, int64_t gp_len, Tensor value, Tensor node_type, Tensor subtree_size, Tensor positions) {
    TORCH_CHECK(pop_size > 0, "pop_size must larger than 0, but got ", pop_size);
    TORCH_CHECK(0 < gp_len && gp_len <= EVOGP_MAX_STACK, "gp_len must be in range (0, ", EVOGP_MAX_STACK, "], but got ", gp_len);
    check_tensor(value, {pop_size, gp_len}, at::kFloat, "value");
    check_tensor(node_type, {pop_size, gp_len}, at::kShort, "node_type");
    check_tensor(subtree_size, {pop_size, gp_len}, at::kShort, "subtree_size");
    check_tensor(positions, {pop_size}, at::kInt, "positions");
    c10::cuda::CUDAGuard guard(value.device());
    auto out = alloc_forest(pop_size, gp_len, value);
    check_rc(evogp_extract_subtree((int)pop_size, (int)gp_len, value.data_ptr<float>(), node_type.data_ptr<int16_t>(),
                                   subtree_size.data_ptr<int16_t>(), positions.data_ptr<int>(), std::get<0>(out).data_ptr<float>(),
                                   std::get<1>(out).data_ptr<int16_t>(), std::get<2>(out).data_ptr<int16_t>(), cur_stream(value)),
             "tree_extract_subtree");
    return out;
}

Tensor tree_tournament_select(Tensor fitness, int64_t tournament_size, double best_probability, bool replace, int64_t winner_cnt,
                              Tensor keys) {
    TORCH_CHECK(fitness.is_cuda() && fitness.is_contiguous() && fitness.dim() == 1 && fitness.scalar_type() == at::kFloat,
                "fitness must be a contiguous 1-D float CUDA tensor");
    check_tensor(keys, {2}, at::kUInt32, "keys");
    TORCH_CHECK(winner_cnt > 0, "winner_cnt must larger than 0, but got ", winner_cnt);
    c10::cuda::CUDAGuard guard(fitness.device());
    auto winners = at::empty({winner_cnt}, fitness.options().dtype(at::kInt));
    check_rc(evogp_tournament_select((int)fitness.numel(), fitness.data_ptr<float>(), (int)tournament_size, (float)best_probability,
                                     replace ? 1 : 0, (int)winner_cnt, static_cast<const unsigned *>(keys.data_ptr()),
                                     winners.data_ptr<int>(), cur_stream(fitness)),
             "tree_tournament_select");
    return winners;
}

Tensor3 tree_mutate(int64_t pop_size, int64_t gp_len, Tensor value_ori, Tensor type_ori, Tensor subtree_size_ori,
                    Tensor mutateIndices, Tensor value_new, Tensor type_new, Tensor subtree_size_new) {
    TORCH_CHECK(pop_size > 0, "pop_size must larger than 0, but got ", pop_size);
    TORCH_CHECK(0 < gp_len && gp_len <= EVOGP_MAX_STACK, "gp_len must be in range (0, ", EVOGP_MAX_STACK, "], but got ", gp_len);
    check_tensor(value_ori, {pop_size, gp_len}, at::kFloat, "value_ori");
    check_tensor(type_ori, {pop_size, gp_len}, at::kShort, "type_ori");
    check_tensor(subtree_size_ori, {pop_size, gp_len}, at::kShort, "subtree_size_ori");
    check_tensor(mutateIndices, {pop_size}, at::kInt, "mutateIndices");
    check_tensor(value_new, {pop_size, gp_len}, at::kFloat, "value_new");
    check_tensor(type_new, {pop_size, gp_len}, at::kShort, "type_new");
    check_tensor(subtree_size_new, {pop_size, gp_len}, at::kShort, "subtree_size_new");
    c10::cuda::CUDAGuard guard(value_ori.device());
    auto out = alloc_forest(pop_size, gp_len, value_ori);
    check_rc(evogp_mutate((int)pop_size, (int)gp_len, value_ori.data_ptr<float>(), type_ori.data_ptr<int16_t>(),
                          subtree_size_ori.data_ptr<int16_t>(), mutateIndices.data_ptr<int>(),
                          value_new.data_ptr<float>(), type_new.data_ptr<int16_t>(),
                          subtree_size_new.data_ptr<int16_t>(), std::get<0>(out).data_ptr<float>(),
                          std::get<1>(out).data_ptr<int16_t>(), std::get<2>(out).data_ptr<int16_t>(),
                          cur_stream(value_ori)),
             "tree_mutate");
    return out;
}

Tensor3 tree_crossover(int64_t pop_size_ori, int64_t pop_size_new, int64_t gp_len, Tensor value_ori, Tensor type_ori,
                       Tensor subtree_size_ori, Tensor left_idx, Tensor right_idx, Tensor left_node_idx,
                       Tensor right_node_idx) {
    TORCH_CHECK(pop_size_ori > 0, "pop_size_ori must larger than 0, but got ", pop_size_ori);
    TORCH_CHECK(pop_size_new > 0, "pop_size_new must larger than 0, but got ", pop_size_new);
    TORCH_CHECK(0 < gp_len && gp_len <= EVOGP_MAX_STACK, "gp_len must be in range (0, ", EVOGP_MAX_STACK, "], but got ", gp_len);
    check_tensor(value_ori, {pop_size_ori, gp_len}, at::kFloat, "value_ori");
    check_tensor(type_ori, {pop_size_ori, gp_len}, at::kShort, "type_ori");
    check_tensor(subtree_size_ori, {pop_size_ori, gp_len}, at::kShort, "subtree_size_ori");
    check_tensor(left_idx, {pop_size_new}, at::kInt, "left_idx");
    check_tensor(right_idx, {pop_size_new}, at::kInt, "right_idx");
    check_tensor(left_node_idx, {pop_size_new}, at::kInt, "left_node_idx");
    check_tensor(right_node_idx, {pop_size_new}, at::kInt, "right_node_idx");
    c10::cuda::CUDAGuard guard(value_ori.device());
    auto out = alloc_forest(pop_size_new, gp_len, value_ori);
    check_rc(evogp_crossover((int)pop_size_ori, (int)pop_size_new, (int)gp_len, value_ori.data_ptr<float>(),
                             type_ori.data_ptr<int16_t>(), subtree_size_ori.data_ptr<int16_t>(),
                             left_idx.data_ptr<int>(), right_idx.data_ptr<int>(), left_node_idx.data_ptr<int>(),
                             right_node_idx.data_ptr<int>(), std::get<0>(out).data_ptr<float>(),
                             std::get<1>(out).data_ptr<int16_t>(), std::get<2>(out).data_ptr<int16_t>(),
                             cur_stream(value_ori)),
             "tree_crossover");
    return out;
}

void check_forest_args(int64_t pop_size, int64_t gp_len, int64_t var_len, int64_t out_len, const Tensor &value,
                       const Tensor &node_type, const Tensor &subtree_size) {
    TORCH_CHECK(pop_size > 0, "pop_size must larger than 0, but got ", pop_size);
    TORCH_CHECK(0 < gp_len && gp_len <= EVOGP_MAX_STACK, "gp_len must be in range (0, ", EVOGP_MAX_STACK, "], but got ", gp_len);
    TORCH_CHECK(0 < var_len, "var_len must larger than 0, but got ", var_len);
    TORCH_CHECK(0 < out_len, "out_len must larger than 0, but got ", out_len);
    check_tensor(value, {pop_size, gp_len}, at::kFloat, "value");
    check_tensor(node_type, {pop_size, gp_len}, at::kShort, "node_type");
    check_tensor(subtree_size, {pop_size, gp_len}, at::kShort, "subtree_size");
}

Tensor tree_evaluate(int64_t pop_size, int64_t gp_len, int64_t var_len, int64_t out_len, Tensor value,
                     Tensor node_type, Tensor subtree_size, Tensor variables) {
    check_forest_args(pop_size, gp_len, var_len, out_len, value, node_type, subtree_size);
    check_tensor(variables, {pop_size, var_len}, at::kFloat, "variables");
    c10::cuda::CUDAGuard guard(value.device());
    auto results = at::empty({pop_size, out_len}, value.options());
    size_t wsb = 0;
    auto ws = eval_workspace(pop_size, gp_len, value, wsb);
    check_rc(evogp_evaluate((unsigned)pop_size, (unsigned)gp_len, (unsigned)var_len, (unsigned)out_len,
                            value.data_ptr<float>(), node_type.data_ptr<int16_t>(), subtree_size.data_ptr<int16_t>(),
                            variables.data_ptr<float>(), results.data_ptr<float>(), ws.data_ptr(), wsb,
                            cur_stream(value)),
             "tree_evaluate");
    return results;
}

Tensor tree_SR_fitness(int64_t pop_size, int64_t data_points, int64_t gp_len, int64_t var_len, int64_t out_len,
                       bool useMSE, Tensor value, Tensor node_type, Tensor subtree_size, Tensor variables,
                       Tensor labels, int64_t kernel_type) {
    check_forest_args(pop_size, gp_len, var_len, out_len, value, node_type, subtree_size);
    TORCH_CHECK(data_points > 0, "data_points must larger than 0, but got ", data_points);
    check_tensor(variables, {data_points, var_len}, at::kFloat, "variables");
    check_tensor(labels, {data_points, out_len}, at::kFloat, "labels");
    c10::cuda::CUDAGuard guard(value.device());
    auto fitness = at::empty({pop_size}, value.options());
    size_t wsb = 0;
    auto ws = eval_workspace(pop_size, gp_len, value, wsb);
    check_rc(evogp_SR_fitness((unsigned)pop_size, (unsigned)data_points, (unsigned)gp_len, (unsigned)var_len,
                              (unsigned)out_len, useMSE ? 1 : 0, value.data_ptr<float>(),
                              node_type.data_ptr<int16_t>(), subtree_size.data_ptr<int16_t>(),
                              variables.data_ptr<float>(), labels.data_ptr<float>(), fitness.data_ptr<float>(),
                              (unsigned)kernel_type, ws.data_ptr(), wsb, cur_stream(value)),
             "tree_SR_fitness");
    return fitness;
}

Tensor tree_batch_forward(int64_t pop_size, int64_t data_points, int64_t gp_len, int64_t var_len, int64_t out_len,
                          Tensor value, Tensor node_type, Tensor subtree_size, Tensor variables) {
    check_forest_args(pop_size, gp_len, var_len, out_len, value, node_type, subtree_size);
    TORCH_CHECK(data_points > 0, "data_points must larger than 0, but got ", data_points);
    check_tensor(variables, {data_points, var_len}, at::kFloat, "variables");
    c10::cuda::CUDAGuard guard(value.device());
    auto results = at::empty({pop_size, data_points, out_len}, value.options());
    size_t wsb = 0;
    auto ws = eval_workspace(pop_size, gp_len, value, wsb);
    check_rc(evogp_batch_forward((unsigned)pop_size, (unsigned)data_points, (unsigned)gp_len, (unsigned)var_len,
                                 (unsigned)out_len, value.data_ptr<float>(), node_type.data_ptr<int16_t>(),
                                 subtree_size.data_ptr<int16_t>(), variables.data_ptr<float>(),
                                 results.data_ptr<float>(), ws.data_ptr(), wsb, cur_stream(value)),
             "tree_batch_forward");
    return results;
}

Tensor tree_classification_accuracy(int64_t pop_size, int64_t data_points, int64_t gp_len, int64_t var_len, int64_t out_len,
                                    Tensor value, Tensor node_type, Tensor subtree_size, Tensor variables, Tensor class_labels,
                                    double max_class) {
    check_forest_args(pop_size, gp_len, var_len, out_len, value, node_type, subtree_size);
    TORCH_CHECK(data_points > 0, "data_points must larger than 0, but got ", data_points);
    check_tensor(variables, {data_points, var_len}, at::kFloat, "variables");
    check_tensor(class_labels, {data_points}, at::kFloat, "class_labels");
    c10::cuda::CUDAGuard guard(value.device());
    auto accuracy = at::empty({pop_size}, value.options());
    size_t wsb = 0;
    auto ws = eval_workspace(pop_size, gp_len, value, wsb);
    check_rc(evogp_classification_accuracy((unsigned)pop_size, (unsigned)data_points, (unsigned)gp_len, (unsigned)var_len,
                                           (unsigned)out_len, value.data_ptr<float>(), node_type.data_ptr<int16_t>(),
                                           subtree_size.data_ptr<int16_t>(), variables.data_ptr<float>(),
                                           class_labels.data_ptr<float>(), (float)max_class, accuracy.data_ptr<float>(),
                                           ws.data_ptr(), wsb, cur_stream(value)),
             "tree_classification_accuracy");
    return accuracy;
}

Tensor3 tree_next_generation(int64_t pop_size, int64_t gp_len, Tensor value, Tensor node_type, Tensor subtree_size, Tensor order,
                             int64_t elite_cnt, int64_t survivor_cnt, double mutation_rate, int64_t var_len, int64_t out_len,
                             double out_prob, double const_prob, Tensor depth2leaf_probs, Tensor roulette_funcs,
                             Tensor const_samples, Tensor keys) {
    check_forest_args(pop_size, gp_len, var_len, out_len, value, node_type, subtree_size);
    check_tensor(order, {pop_size}, at::kLong, "order");
    check_tensor(keys, {2}, at::kUInt32, "keys");
    check_tensor(depth2leaf_probs, {EVOGP_MAX_FULL_DEPTH}, at::kFloat, "depth2leaf_probs");
    check_tensor(roulette_funcs, {EVOGP_FUNC_END}, at::kFloat, "roulette_funcs");
    TORCH_CHECK(const_samples.is_cuda() && const_samples.is_contiguous() && const_samples.dim() == 1 &&
                    const_samples.scalar_type() == at::kFloat && const_samples.numel() > 0,
                "const_samples must be a non-empty contiguous 1-D float CUDA tensor");
    c10::cuda::CUDAGuard guard(value.device());
    auto out = alloc_forest(pop_size, gp_len, value);
    check_rc(evogp_next_generation((int)pop_size, (int)gp_len, value.data_ptr<float>(), node_type.data_ptr<int16_t>(),
                                   subtree_size.data_ptr<int16_t>(), reinterpret_cast<const long long *>(order.data_ptr<int64_t>()),
                                   (int)elite_cnt, (int)survivor_cnt, (float)mutation_rate, (unsigned)var_len,
                                   (unsigned)out_len, (unsigned)const_samples.numel(), (float)out_prob, (float)const_prob,
                                   depth2leaf_probs.data_ptr<float>(), roulette_funcs.data_ptr<float>(),
                                   const_samples.data_ptr<float>(), static_cast<const unsigned *>(keys.data_ptr()),
                                   std::get<0>(out).data_ptr<float>(), std::get<1>(out).data_ptr<int16_t>(),
                                   std::get<2>(out).data_ptr<int16_t>(), cur_stream(value)),
             "tree_next_generation");
    return out;
}

}  // namespace

TORCH_LIBRARY(evogp_cuda, m) {
    m.def("tree_generate(int pop_size, int gp_len, int var_len, int out_len, int const_samples_len, float out_prob, float const_prob, Tensor keys, Tensor depth2leaf_probs, Tensor roulette_funcs, Tensor const_samples) -> (Tensor, Tensor, Tensor)");
    m.def("tree_mutate(int pop_size, int gp_len, Tensor value_ori, Tensor type_ori, Tensor subtree_size_ori, Tensor mutateIndices, Tensor value_new, Tensor type_new, Tensor subtree_size_new) -> (Tensor, Tensor, Tensor)");
    m.def("tree_crossover(int pop_size_ori, int pop_size_new, int gp_len, Tensor value_ori, Tensor type_ori, Tensor subtree_size_ori, Tensor left_idx, Tensor right_idx, Tensor left_node_idx, Tensor right_node_idx) -> (Tensor, Tensor, Tensor)");
    m.def("tree_evaluate(int pop_size, int gp_len, int var_len, int out_len, Tensor value, Tensor node_type, Tensor subtree_size, Tensor variables) -> Tensor");
    m.def("tree_SR_fitness(int pop_size, int data_points, int gp_len, int var_len, int out_len, bool useMSE, Tensor value, Tensor node_type, Tensor subtree_size, Tensor variables, Tensor labels, int kernel_type) -> Tensor");
    m.def("tree_next_generation(int pop_size, int gp_len, Tensor value, Tensor node_type, Tensor subtree_size, Tensor order, int elite_cnt, int survivor_cnt, float mutation_rate, int var_len, int out_len, float out_prob, float const_prob, Tensor depth2leaf_probs, Tensor roulette_funcs, Tensor const_samples, Tensor keys) -> (Tensor, Tensor, Tensor)");
    m.def("tree_generate_philox(int pop_size, int gp_len, int var_len, int out_len, int const_samples_len, float out_prob, float const_prob, Tensor keys, Tensor depth2leaf_probs, Tensor roulette_funcs, Tensor const_samples) -> (Tensor, Tensor, Tensor)");
    m.def("tree_extract_subtree(int pop_size, int gp_len, Tensor value, Tensor node_type, Tensor subtree_size, Tensor positions) -> (Tensor, Tensor, Tensor)");
    m.def("tree_tournament_select(Tensor fitness, int tournament_size, float best_probability, bool replace, int winner_cnt, Tensor keys) -> Tensor");
    m.def("tree_classification_accuracy(int pop_size, int data_points, int gp_len, int var_len, int out_len, Tensor value, Tensor node_type, Tensor subtree_size, Tensor variables, Tensor class_labels, float max_class) -> Tensor");
    m.def("tree_batch_forward(int pop_size, int data_points, int gp_len, int var_len, int out_len, Tensor value, Tensor node_type, Tensor subtree_size, Tensor variables) -> Tensor");
}

TORCH_LIBRARY_IMPL(evogp_cuda, CUDA, m) {
    m.impl("tree_generate", &tree_generate);
    m.impl("tree_mutate", &tree_mutate);
    m.impl("tree_crossover", &tree_crossover);
    m.impl("tree_evaluate", &tree_evaluate);
    m.impl("tree_SR_fitness", &tree_SR_fitness);
    m.impl("tree_batch_forward", &tree_batch_forward);
    m.impl("tree_classification_accuracy", &tree_classification_accuracy);
    m.impl("tree_generate_philox", &tree_generate_philox);
    m.impl("tree_extract_subtree", &tree_extract_subtree);
    m.impl("tree_tournament_select", &tree_tournament_select);
    m.impl("tree_next_generation", &tree_next_generation);
}
