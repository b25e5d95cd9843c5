/*
 * evogp_b200.h — C ABI of the B200-native EvoGP hot path (libevogp_b200.so).
 *
 * Drop-in boundary.  The reference's FFI for this path is the torch operator
 * library `evogp_cuda` (src/evogp/cuda/torch_wrapper.cu:291-307) sitting on
 * five plain host functions that take raw device pointers
 * (src/evogp/cuda/kernel.h:23-97).  The entry points below are those five
 * functions — same argument order and meaning — as `extern "C"` symbols, with
 * three additions the reference lacks: an explicit CUDA stream, an int status
 * (0 = ok, else see evogp_last_error()), and caller-provided scratch for the
 * fitness evaluator.  evogp_b200/csrc/torch_ops.cpp re-registers the same five
 * `evogp_cuda::tree_*` schemas on top of them.
 *
 * All pointers are DEVICE pointers unless the function name ends in `_host`.
 * Inputs are borrowed and never written; outputs must be preallocated by the
 * caller ([P, L] arrays are row-major, contiguous).  Calls are asynchronous on
 * `stream` (a cudaStream_t; NULL = legacy default stream).  There is no CPU
 * fallback: every call fails with EVOGP_ERR_CUDA when no sm_100 device is
 * usable.
 *
 * Packed forest layout (reference: src/evogp/tree/forest.py:13-40,
 * src/evogp/cuda/defs.h:10-22): three [P, L] arrays, one tree per row in
 * prefix order; node_value f32, node_type i16, subtree_size i16; valid prefix
 * length = subtree_size[i, 0].  Unlike the reference (which leaves row tails
 * uninitialised), every producer here zero-fills the tail.
 */
#ifndef EVOGP_B200_H
#define EVOGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVOGP_MAX_STACK 1024      /* defs.h:5  — upper bound of max_tree_len   */
#define EVOGP_MAX_FULL_DEPTH 10   /* defs.h:5  — length of depth2leaf_probs    */
#define EVOGP_FUNC_END 29         /* defs.h:56 — length of roulette_funcs      */

enum {
    EVOGP_OK = 0,
    EVOGP_ERR_ARG = 1,        /* scalar argument out of range (torch_wrapper.cu:48-60 checks) */
    EVOGP_ERR_CUDA = 2,       /* launch / runtime failure, or no usable device              */
    EVOGP_ERR_WORKSPACE = 3,  /* scratch buffer too small                                   */
    EVOGP_ERR_UNSUPPORTED = 4 /* shape beyond what the kernels stage in shared memory       */
};

/* Library identity / diagnostics. */
int evogp_version(void);
const char *evogp_last_error(void);
/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
unsigned long long evogp_launch_count(void);

/* replaces generate(), kernel.h:23-38 (generate.cu:210-234).  keys: uint32[2];
 * depth2leafProbs: f32[10]; rouletteFuncs: f32[29] cumulative; constSamples: f32[constSamplesLen].
 * RNG: taus88 seeded with hash(n, keys[0], keys[1]) — bit-identical trees to the reference. */
int evogp_generate(unsigned popSize, unsigned maxGPLen, unsigned varLen, unsigned outLen, unsigned constSamplesLen,
                   float outProb, float constProb, const unsigned *keys, const float *depth2leafProbs,
                   const float *rouletteFuncs, const float *constSamples, float *value_res, int16_t *type_res,
                   int16_t *subtree_size_res, void *stream);

/* Philox mode of evogp_generate (BASELINE.json north_star: "cuRAND/Philox per-thread state"; the reference's generator
 * site is generate.cu:40-41).  Same growth rules and draw order; draw j of tree n is word j % 4 of
 * Philox4x32-10(counter = (n, 0x10000 + j / 4), key = keys): counter-based, nothing to seed or store.  Trees differ
 * from the taus88 mode (which stays the bit-exact-with-the-reference default); parity: oracle_generate_philox. */
int evogp_generate_philox(unsigned popSize, unsigned maxGPLen, unsigned varLen, unsigned outLen, unsigned constSamplesLen,
                          float outProb, float constProb, const unsigned *keys, const float *depth2leafProbs,
                          const float *rouletteFuncs, const float *constSamples, float *value_res, int16_t *type_res,
                          int16_t *subtree_size_res, void *stream);

/* replaces mutate(), kernel.h:40-53 (mutation.cu:186-219). */
int evogp_mutate(int popSize, int gpLen, const float *value_ori, const int16_t *type_ori,
                 const int16_t *subtree_size_ori, const int *mutateIndices, const float *value_new,
                 const int16_t *type_new, const int16_t *subtree_size_new, float *value_res, int16_t *type_res,
                 int16_t *subtree_size_res, void *stream);

/* replaces crossover(), kernel.h:55-69 (mutation.cu:312-347). */
int evogp_crossover(int pop_size_ori, int pop_size_new, int gpLen, const float *value_ori, const int16_t *type_ori,
                    const int16_t *subtree_size_ori, const int *left_idx, const int *right_idx,
                    const int *left_node_idx, const int *right_node_idx, float *value_res, int16_t *type_res,
                    int16_t *subtree_size_res, void *stream);

/* Scratch needed by evogp_evaluate / evogp_SR_fitness / evogp_batch_forward for a
 * population of popSize rows of width maxGPLen (compiled programs + scheduler words). */
size_t evogp_eval_workspace_bytes(unsigned popSize, unsigned maxGPLen);

/* Measurement hook: when both are non-NULL cudaEvent_t handles, every evaluation entry point
 * records `begin_event` immediately before and `end_event` immediately after the replay kernel
 * launch on the caller's stream (so the tree-evaluation kernel can be timed apart from the
 * lowering pass).  Pass NULLs to switch it off. */
void evogp_eval_set_timing_events(void *begin_event, void *end_event);

/* Datapoints per lane of the single-output evaluation kernel: 16 (two passes over 1024 datapoints; the code of its
 * + - * / neg sin cos bodies just fits the instruction cache in front of the interpreter loop), 8 (half the code per
 * operator body: faster as soon as the population uses any other function - 15 % with + - * / sin cos tan, 2.5x with
 * all of them) or 0 = chosen per launch from the shapes alone (the default; EVOGP_REPLAY_K presets it).
 * The kernels cannot know the function set of a population before they have run, the caller does: the Python front-end
 * calls this from the GenerateDescriptor a forest is generated with.  The width is a speed knob: a lane owns the same
 * datapoints at either width and adds their errors in the same order (bit-equal fitness when dataPoints is a multiple
 * of 512; a ragged last pass differs in the last bits).
 * Process-wide; returns EVOGP_ERR_INVALID_ARGUMENT otherwise. */
int evogp_eval_set_replay_width(int datapoints_per_lane);

/* Diagnostics: run only the lowering pass (packed rows -> accumulator-machine programs) and copy the
 * programs out: programs = DEVICE u64[popSize][(maxGPLen + 2) & ~1].  use_fast: 1 = the register-resident pass where it
 * applies (single-output, maxGPLen <= 64), 0 = the generic pass.  deep_from: operand-stack slots >= deep_from are marked
 * with the deep opcodes (0 = none).  The two passes must produce identical programs (tests/test_gpu_lowering.py). */
int evogp_debug_lower(unsigned popSize, unsigned maxGPLen, unsigned varLen, unsigned outLen, const float *value,
                      const int16_t *type, const int16_t *subtree_size, int use_fast, int deep_from, void *workspace,
                      size_t workspace_bytes, unsigned long long *programs, void *stream);

/* replaces evaluate(), kernel.h:71-81 (forward.cu:353-371): tree n on variables[n, :] -> results[n, :outLen]. */
int evogp_evaluate(unsigned popSize, unsigned maxGPLen, unsigned varLen, unsigned outLen, const float *value,
                   const int16_t *type, const int16_t *subtree_size, const float *variables, float *results,
                   void *workspace, size_t workspace_bytes, void *stream);

/* replaces SR_fitness(), kernel.h:83-97 (forward.cu:827-856).
 * fitnesses[i] = (1/dataPoints) * sum_n sum_o loss(labels[n,o] - out_o(tree_i, variables[n,:])),
 * loss = square (useMSE) or abs.  kernel_type (the reference's execute_mode code 0..4) is accepted
 * and ignored: one kernel serves every mode. */
int evogp_SR_fitness(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen, unsigned outLen,
                     int useMSE, const float *value, const int16_t *type, const int16_t *subtree_size,
                     const float *variables, const float *labels, float *fitnesses, unsigned kernel_type,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Multi-GPU form of evogp_SR_fitness (SURVEY.md row e; no counterpart in the reference, which is single-GPU): this
 * rank evaluates its shard of popSize trees and the kernel stores each fitness straight into EVERY rank's
 * full-population buffer at [row_offset + i] through peer-mapped memory (NVLink) - the all-gather of the fitness
 * scalars is fused into the evaluation kernel.  peer_fitnesses: DEVICE array of `world` (<= 32) device pointers, entry r
 * = rank r's buffer (this rank's own included), mapped into this process (CUDA IPC / symmetric memory).  fitnesses:
 * local [popSize] (also written).  The caller synchronises the ranks afterwards (any inter-GPU barrier on the stream). */
int evogp_SR_fitness_scatter(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen, unsigned outLen,
                             int useMSE, const float *value, const int16_t *type, const int16_t *subtree_size,
                             const float *variables, const float *labels, float *fitnesses,
                             float *const *peer_fitnesses, unsigned world, unsigned row_offset, void *workspace,
                             size_t workspace_bytes, void *stream);

/* The exchange of evogp_SR_fitness_scatter as a kernel of its own: copies this rank's fitness slice local_fitness[count]
 * into EVERY rank's full-population buffer at [row_offset, row_offset + count) through peer-mapped memory (coalesced
 * 128-byte stores over NVLink).  Measured on B200s this costs less than carrying the exchange protocol
 * inside the evaluation kernel, so parallel.FitnessExchange uses it by default.  Arguments as evogp_SR_fitness_scatter. */
int evogp_push_fitness(const float *local_fitness, unsigned count, float *const *peer_fitnesses, unsigned world,
                       unsigned row_offset, void *stream);

/* Fused form of Forest.batch_forward (tree/forest.py:143-176), which the reference
 * implements by replicating the forest dataPoints times: results[P, N, O]. */
int evogp_batch_forward(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen, unsigned outLen,
                        const float *value, const int16_t *type, const int16_t *subtree_size,
                        const float *variables, float *results, void *workspace, size_t workspace_bytes,
                        void *stream);

/* Native form of vmap_subtree (src/evogp/algorithm/mutation/mutation_utils.py:6-48), the gather the reference's Hoist /
 * Insert / Delete mutations build from torch ops: row n of the result = the subtree of tree n rooted at positions[n]
 * (int32[popSize]), moved to the front, tail zero-filled; a position outside [0, gpLen) gives an all-zero row. */
int evogp_extract_subtree(int popSize, int gpLen, const float *value, const int16_t *type, const int16_t *subtree_size,
                          const int *positions, float *value_res, int16_t *type_res, int16_t *subtree_size_res, void *stream);

/* Native form of TournamentSelection (src/evogp/algorithm/selection/tournament.py:59-133; the reference draws contenders
 * with torch.multinomial over a [k_times, P] matrix under vmap, :73-79): winners[j], j < winnerCnt, is the nth best of
 * tournamentSize contenders, nth geometric in bestProbability (:97-101; 1 = always the best), NaN fitness ranking last.
 * replace != 0: contenders are independent uniform draws; else tournaments are consecutive slices of a pseudo-random
 * permutation of the population, a fresh one every popSize / tournamentSize tournaments.  Draws: Philox4x32-10 keyed by
 * keys[2] (counter-based: deterministic in keys, not torch's stream).  winners: int32[winnerCnt] row indices. */
int evogp_tournament_select(int popSize, const float *fitness, int tournamentSize, float bestProbability, int replace,
                            int winnerCnt, const unsigned *keys, int *winners, void *stream);

/* Fused form of Classification.evaluate (problem/classification.py:54-67), which the reference computes from
 * batch_forward's [P, N, O] output with torch softmax / argmax: accuracy[i] = (1/N) * #{n : pred_i(n) == class_labels[n]}.
 * outLen > 1: pred = arg-max over the outputs (first maximum; 0 when an output is NaN or the maximum infinite - torch's
 * softmax makes every probability NaN there and argmax returns index 0).  outLen == 1: pred = clamp(round(out +
 * max_class / 2), 0, max_class), round half to even.  class_labels: f32[dataPoints] class ids.  Nothing but one float
 * per tree leaves the SM (BASELINE configs[3]: the batch_forward round trip would be 9.8 GB). */
int evogp_classification_accuracy(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen, unsigned outLen,
                                  const float *value, const int16_t *type, const int16_t *subtree_size,
                                  const float *variables, const float *class_labels, float max_class, float *accuracy,
                                  void *workspace, size_t workspace_bytes, void *stream);

/* One whole generation step in one kernel ("next" row f-1; no counterpart in the reference's kernel.h — it fuses
 * what algorithm/genetic_programming.py:110-118, crossover/default.py and mutation/default.py do with ~14 torch
 * calls around three kernels):
 *   next[n] = current[order[n]]                                                      n <  eliteCnt
 *   next[n] = mutate?( crossover(current[order[a]], current[order[b]]) ),  a, b < survivorCnt, otherwise
 * order: int64[popSize] row indices, best first (torch.sort of the fitness).  Parents, splice positions and the
 * mutation coin come from Philox4x32-10 keyed by keys[2] and the child index; mutation donors are grown in the
 * kernel by the generator of evogp_generate (descriptor arguments as there).  Splice rules and fallbacks are the
 * reference's; the random stream is not torch's, so results are deterministic in (keys, inputs) but not
 * bit-comparable with the unfused operator sequence. */
int evogp_next_generation(int popSize, int gpLen, const float *value, const int16_t *type, const int16_t *subtree_size,
                          const long long *order, int eliteCnt, int survivorCnt, float mutationRate, unsigned varLen,
                          unsigned outLen, unsigned constSamplesLen, float outProb, float constProb,
                          const float *depth2leafProbs, const float *rouletteFuncs, const float *constSamples,
                          const unsigned *keys, float *value_res, int16_t *type_res, int16_t *subtree_size_res,
                          void *stream);

/* Host-buffer form of evogp_SR_fitness: every pointer is HOST memory (pinned memory
 * makes the copies asynchronous).  Uploads the forest in row chunks on two streams so
 * H2D overlaps evaluation, downloads fitnesses[popSize], and returns after the result
 * is on the host.  Device staging buffers are cached inside the library.  device = CUDA
 * device ordinal. */
int evogp_SR_fitness_host(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen, unsigned outLen,
                          int useMSE, const float *value, const int16_t *type, const int16_t *subtree_size,
                          const float *variables, const float *labels, float *fitnesses, int device);
/* Free the cached staging buffers of evogp_SR_fitness_host. */
void evogp_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* EVOGP_B200_H */
