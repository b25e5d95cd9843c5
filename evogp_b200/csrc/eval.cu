// eval.cu — batched expression-tree evaluation on sm_100a.
//
// Replaces the reference's evaluate / SR_fitness host functions
// (src/evogp/cuda/forward.cu:353-371, :827-856) and the four kernels behind them
// (:304-351, :402-479, :591-647, :738-779), and fuses Forest.batch_forward
// (tree/forest.py:143-176).
//
// Two kernels (a single fused kernel — each warp lowering its own tree into shared memory before replaying
// it — was built and measured: same total time at best, because lowering + replay code together overflow the
// SM instruction cache: sm__icc_request_hit_rate 97 % -> 70 %, profiles/r1_fused_kernel_icache.txt):
//   lower_kernel / lower_fast_kernel (lower.cuh, lower_fast.cuh)  packed rows -> accumulator-machine programs, 8 B/slot;
//        one warp per tree; the register-resident fast pass serves single-output rows of max_tree_len <= 64
//   replay_kernel (replay.cuh; this file instantiates the plain flavour, eval_exchange.cu / eval_acc.cu the others)
//        persistent CTAs; each WARP owns one tree at a time,
//        lanes own datapoints (K per lane, float4-vectorised), so the opcode
//        dispatch is warp-uniform.  The next tree's program row is pulled into
//        shared memory by a 1-D TMA bulk copy (cp.async.bulk + mbarrier) while the
//        current one replays; trees are handed out by an atomic ticket so ragged
//        tree lengths balance.  The dataset is staged once per CTA, transposed to
//        [var][datapoint].  Squared/absolute error is accumulated in registers and
//        reduced with warp shuffles: one plain store per tree — no memset, no
//        atomics on fitness, no averaging kernel.
//        Single-output programs keep their operand stack in TENSOR MEMORY (TSTK):
//        tcgen05.st / tcgen05.ld of one K-column slot per save / restore, no MMA
//        involved; K = 16 datapoints per lane in one 32-warp CTA per SM, K = 8 in
//        four 8-warp CTAs.  Multi-output programs: their own PTX loop (K = 8), outs[] in shared
//        memory.  Multi-GPU: evogp_push_fitness after this kernel, or the exchange flavour of it.
#include <atomic>
#include "replay.cuh"
#include "lower_fast.cuh"

namespace evogp {

int run_eval(int mode, unsigned P, unsigned N, unsigned L, unsigned V, unsigned O, const float *value, const int16_t *type,
             const int16_t *size, int len_stride, const float *X, const float *labels, float *out, void *workspace,
             size_t workspace_bytes, void *stream, Scatter scatter = Scatter(), float acc_half = 0.0f, float acc_max = 0.0f);

EVOGP_DEFINE_REPLAY_DISPATCH(launch_replay_plain, FEAT_PLAIN)

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// device properties of the CURRENT device (refreshed whenever the calling thread's device changes: one process may
// drive several GPUs; cudaFuncSetAttribute and occupancy answers are per device too, so nothing below caches them
// across devices)
static thread_local int g_props_dev = -1;
static thread_local int g_sm_count = 0, g_max_smem = 0, g_smem_per_sm = 0;
// EVOGP_TMEM_STACK=0 keeps the operand stack in shared memory (the A/B switch of profiles/; default on)
static const bool g_use_tmem_stack = []() { const char *e = getenv("EVOGP_TMEM_STACK"); return !(e && e[0] == '0'); }();
// EVOGP_K16_SPLIT=1 feeds the K = 16 kernel split-mode programs (lower.cuh; the A/B switch of profiles/; default off)
static const bool g_k16_split = []() { const char *e = getenv("EVOGP_K16_SPLIT"); return e && e[0] == '1'; }();
// EVOGP_FOLD=0 lowers without constant folding (the A/B switch of profiles/; default on)
static const bool g_fold = []() { const char *e = getenv("EVOGP_FOLD"); return !(e && e[0] == '0'); }();
// datapoints per lane of the single-output replay kernel: 0 = by the cost model (choose_replay), 8 or 16 forced.
// EVOGP_REPLAY_K sets the start value, evogp_eval_set_replay_width() changes it at run time (the front-end does, from the
// function set of the run: DESIGN.md 3.2 "function sets and the instruction cache").
static std::atomic<int> g_force_k{[]() { const char *e = getenv("EVOGP_REPLAY_K"); return e ? atoi(e) : 0; }()};
// optional cudaEvent_t pair recorded around the replay launch (bench.py's per-kernel timing)
static cudaEvent_t g_ev_replay_begin = nullptr, g_ev_replay_end = nullptr;

static int device_props() {
    int dev = 0;
    EVOGP_CUDA(cudaGetDevice(&dev));
    if (dev == g_props_dev) return EVOGP_OK;
    EVOGP_CUDA(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev));
    EVOGP_CUDA(cudaDeviceGetAttribute(&g_max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    EVOGP_CUDA(cudaDeviceGetAttribute(&g_smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    g_props_dev = dev;
    return EVOGP_OK;
}

// row pitch in slots: one spare slot so that every program ends in C_END (the fast path never
// checks a bound) and stays a multiple of 16 bytes for the bulk copy
static inline int prog_pitch(unsigned L) { return (int)((L + 2) & ~1u); }

struct Workspace {
    uint2 *prog;
    unsigned *sched;   // 64 words
    unsigned *chunk_done;   // ceil(P / kPushChunk) words behind the programs
};
static size_t chunk_words(unsigned P) { return ((size_t)P + kPushChunk - 1) / kPushChunk; }
static size_t prog_bytes(unsigned P, unsigned L) { return (size_t)P * prog_pitch(L) * sizeof(uint2); }
static Workspace carve(void *ws, unsigned P, unsigned L) {
    Workspace w;
    unsigned char *b = static_cast<unsigned char *>(ws);
    w.sched = reinterpret_cast<unsigned *>(b);
    w.prog = reinterpret_cast<uint2 *>(b + 256);
    w.chunk_done = reinterpret_cast<unsigned *>(b + 256 + ((prog_bytes(P, L) + 15) & ~(size_t)15));
    return w;
}

template <bool MULTI, bool SPLIT>
static int launch_lower_t(const Workspace &w, unsigned P, unsigned L, unsigned V, unsigned O, const float *value,
                          const int16_t *type, const int16_t *size, int len_stride, int depth, int deep_from, cudaStream_t st) {
    auto kern = lower_kernel<MULTI, SPLIT>;
    // one warp per tree; per-warp scratch is 18 B per node slot (lower.cuh)
    const size_t per_warp = (lower_scratch_bytes((int)L) + 15) & ~(size_t)15;
    int warps = 8;
    while (warps > 1 && warps * per_warp > 96 * 1024) warps >>= 1;
    const size_t smem = warps * per_warp;
    if (smem > 48 * 1024) EVOGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device: not cached
    LowerArgs a;
    a.value = value; a.type = type; a.size = size;
    a.prog = w.prog; a.sched = w.sched;
    a.P = (int)P; a.L = (int)L; a.Lp = prog_pitch(L); a.V = (int)V; a.O = (int)O; a.depth_budget = depth;
    a.rows_have_sizes = len_stride > 1;                       // 1: one length per tree; 0: compact rows, `size` is the offsets array
    a.offsets = len_stride == 0 ? reinterpret_cast<const unsigned *>(size) : nullptr;
    a.deep_from = deep_from;
    a.fold = g_fold ? 1 : 0;
    a.chunk_done = w.chunk_done; a.nchunks = (int)chunk_words(P);
    // exactly one resident wave: the kernel strides over the population, so CTAs beyond what the SMs hold at once
    // would only run as a second, half-empty wave (measured: 46 % -> 60 % warps active)
    static thread_local int per_sm_cached = 0, per_sm_dev = -1;
    static thread_local size_t per_sm_smem = ~(size_t)0;
    if (per_sm_smem != smem || per_sm_dev != g_props_dev) {
        int n = 0;
        EVOGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, warps * 32, smem));
        per_sm_cached = n < 1 ? 1 : n;
        per_sm_smem = smem;
        per_sm_dev = g_props_dev;
    }
    long long grid = ((long long)P + warps - 1) / warps;
    const long long cap = (long long)g_sm_count * per_sm_cached;
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, warps * 32, smem, st>>>(a);
    count_launch();
    return check_launch("lower_kernel");
}
// EVOGP_LOWER_FAST=0 keeps every row on the generic pass (the A/B switch of profiles/; default on)
static const bool g_lower_fast = []() { const char *e = getenv("EVOGP_LOWER_FAST"); return !(e && e[0] == '0'); }();

// the register-resident pass of lower_fast.cuh: single-output, max_tree_len <= 64, packed size rows
template <int NSETS>
static int launch_lower_fast(const Workspace &w, unsigned P, unsigned L, unsigned V, unsigned O, const float *value,
                             const int16_t *type, const int16_t *size, int depth, int deep_from, cudaStream_t st) {
    auto kern = lower_fast_kernel<NSETS>;
    const int warps = 8;
    const size_t smem = warps * lower_fast_per_warp<NSETS>((int)L);
    LowerArgs a;
    a.value = value; a.type = type; a.size = size;
    a.prog = w.prog; a.sched = w.sched;
    a.P = (int)P; a.L = (int)L; a.Lp = prog_pitch(L); a.V = (int)V; a.O = (int)O; a.depth_budget = depth;
    a.rows_have_sizes = 1;
    a.offsets = nullptr;
    a.deep_from = deep_from;
    a.fold = g_fold ? 1 : 0;
    a.chunk_done = w.chunk_done; a.nchunks = (int)chunk_words(P);
    static thread_local int per_sm_cached = 0, per_sm_dev = -1;
    if (per_sm_dev != g_props_dev * 4 + NSETS) {
        int n = 0;
        EVOGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, warps * 32, smem));
        per_sm_cached = n < 1 ? 1 : n;
        per_sm_dev = g_props_dev * 4 + NSETS;
    }
    long long grid = ((long long)P + warps - 1) / warps;
    const long long cap = (long long)g_sm_count * per_sm_cached;       // one resident wave, grid-stride (as lower_kernel)
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, warps * 32, smem, st>>>(a);
    count_launch();
    return check_launch("lower_fast_kernel");
}

// split: single-output programs for the K = 16 replay kernel (LOAD + acc-form for operators on leaves)
template <bool MULTI>
static int launch_lower(const Workspace &w, unsigned P, unsigned L, unsigned V, unsigned O, const float *value,
                        const int16_t *type, const int16_t *size, int len_stride, int depth, bool split, int deep_from,
                        cudaStream_t st) {
    if constexpr (!MULTI) {
        if (split) return launch_lower_t<false, true>(w, P, L, V, O, value, type, size, len_stride, depth, deep_from, st);
        if (g_lower_fast && len_stride > 1 && L <= 64)
            return L <= 32 ? launch_lower_fast<1>(w, P, L, V, O, value, type, size, depth, deep_from, st)
                           : launch_lower_fast<2>(w, P, L, V, O, value, type, size, depth, deep_from, st);
    }
    return launch_lower_t<MULTI, false>(w, P, L, V, O, value, type, size, len_stride, depth, deep_from, st);
}

// does ONE pass (32 * K datapoints of `per_dp` bytes each) fit next to four warps?  Wide datasets (hundreds of
// inputs) push the choice down to K = 4 / 1 instead of failing (the reference accepts var_len <= 512)
static bool replay_fits(int K, bool multi, bool tmem, int depth, int Lp, int O, size_t per_dp) {
    return (size_t)32 * K * per_dp + 4 * replay_per_warp(K, multi, tmem, depth, Lp, O) <= (size_t)g_max_smem;
}

// cost model for choosing K: issue slots per datapoint ~ (dispatch overhead + K) / K, times padding waste
static int choose_k(int N, bool multi, int depth, int Lp, int O, size_t per_dp) {
    const int ks[3] = {8, 4, 1};
    int best = 1;
    double best_cost = 1e30;
    for (int i = 0; i < 3; ++i) {
        const int K = ks[i];
        if (K > 1 && !replay_fits(K, multi, false, depth, Lp, O, per_dp)) continue;
        const int npass = (N + 32 * K - 1) / (32 * K);
        const double cost = (double)npass * (14.0 + 2.0 * K);
        if (cost < best_cost) { best_cost = cost; best = K; }
    }
    return best;
}

// Which single-output replay kernel serves a launch: 16 or 8 datapoints per lane with the operand stack in tensor
// memory, or the shared-memory-stack kernels (K = 8 / 4 / 1).  Decided before the lowering pass because K = 16
// wants split-mode programs (lower.cuh).
static ReplayChoice choose_replay(bool multi, int mode, int N, int depth, int Lp, int O, size_t per_dp) {
    ReplayChoice c{1, false};
    if (mode == MODE_ROWWISE) return c;
    const size_t dataset_bytes = (size_t)N * per_dp;
    c.K = choose_k(N, multi, depth, Lp, O, per_dp);
    if (multi || c.K != 8 || !g_use_tmem_stack) return c;
    // tensor-memory stack while two warps per lane quarter fit the columns (depth <= 8, i.e. max_tree_len <= 256);
    // 16 datapoints per lane when that does not waste passes on padding (cost model of choose_k;
    // EVOGP_REPLAY_K=8|16 overrides)
    const int d = depth > 0 ? depth : 1;
    const double c8 = (double)((N + 255) / 256) * (14.0 + 16.0), c16 = (double)((N + 511) / 512) * (14.0 + 32.0);
    const int force_k = g_force_k.load(std::memory_order_relaxed);
    const bool want16 = force_k ? force_k == 16 : c16 < c8;
    // K = 16 needs room for >= 16 warps' program buffers and deep slots next to the dataset (very wide rows do not
    // leave it: they run on the 8-datapoint kernels)
    const size_t per_warp16 = (size_t)2 * Lp * 8 + (size_t)(d > kTmemSlots16 ? d - kTmemSlots16 : 0) * 512 * 4 + 16;
    const size_t staged = dataset_bytes < 48 * 1024 ? dataset_bytes : 48 * 1024;   // larger datasets are tiled to <= 48 KB
    const bool fits16 = staged + 16 * per_warp16 <= (size_t)g_max_smem && (size_t)512 * per_dp + 16 * per_warp16 <= (size_t)g_max_smem;
    if (want16 && (fits16 || force_k == 16)) { c.K = 16; c.tmem = true; }
    else if (tmem_stack_cols(8, d, 8) <= 128) c.tmem = true;
    return c;
}

int run_eval(int mode, unsigned P, unsigned N, unsigned L, unsigned V, unsigned O, const float *value,
             const int16_t *type, const int16_t *size, int len_stride, const float *X, const float *labels, float *out,
             void *workspace, size_t workspace_bytes, void *stream, Scatter scatter, float acc_half, float acc_max) {
    EVOGP_REQUIRE(P > 0, "popSize must be larger than 0, got %u", P);
    EVOGP_REQUIRE(scatter.peers == nullptr || (scatter.world >= 1 && scatter.world <= 32 && is_reduce_mode(mode)),
                  "fitness scatter: world must be in [1, 32] (got %d) and the mode a loss mode", scatter.world);
    EVOGP_REQUIRE(L > 0 && L <= (unsigned)kMaxStack, "gp_len must be in (0, %d], got %u", kMaxStack, L);
    EVOGP_REQUIRE(V > 0 && V <= 512, "var_len must be in (0, 512], got %u", V);   // forward.cu:320 asserts the same bound
    EVOGP_REQUIRE(O > 0 && O <= 256, "out_len must be in (0, 256], got %u", O);
    EVOGP_REQUIRE(N > 0, "data_points must be larger than 0, got %u", N);
    EVOGP_REQUIRE((unsigned long long)P * prog_pitch(L) < (1ull << 40), "population too large");
    if (workspace_bytes < evogp_eval_workspace_bytes(P, L) || workspace == nullptr) {
        set_error("workspace too small: need %zu bytes, got %zu", evogp_eval_workspace_bytes(P, L), workspace_bytes);
        return EVOGP_ERR_WORKSPACE;
    }
    int rc = ensure_device_ok();
    if (rc) return rc;
    rc = device_props();
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Workspace w = carve(workspace, P, L);
    const int depth = stack_depth_bound((int)L);
    const bool multi = O > 1;
    const ReplayChoice choice = choose_replay(multi, mode, (int)N, depth, prog_pitch(L), (int)O, ((size_t)V + label_columns(mode, (int)O)) * 4);
    const int deep_from = choice.K == 16 ? kTmemSlots16 : kNoDeepSlots;
    rc = multi ? launch_lower<true>(w, P, L, V, O, value, type, size, len_stride, depth, false, kNoDeepSlots, st)
               : launch_lower<false>(w, P, L, V, O, value, type, size, len_stride, depth, choice.K == 16 && g_k16_split, deep_from, st);
    if (rc) return rc;
    ReplayArgs a;
    a.prog = w.prog; a.sched = w.sched; a.X = X; a.labels = labels; a.out = out;
    a.P = (int)P; a.Lp = prog_pitch(L); a.N = (int)N; a.V = (int)V; a.O = (int)O;
    a.NP = 0; a.npass = 0; a.depth = depth; a.mode = mode;
    a.peers = scatter.peers; a.world = scatter.world; a.row_offset = scatter.row_offset;
    a.chunk_done = w.chunk_done;
    a.acc_half = acc_half; a.acc_max = acc_max;
    const ReplayEnv env{g_sm_count, g_max_smem, g_smem_per_sm, g_ev_replay_begin, g_ev_replay_end};
    if (scatter.peers != nullptr) return launch_replay_exchange(multi, a, depth, choice, st, env);
    if (mode == MODE_ACC) return launch_replay_acc(multi, a, depth, choice, st, env);
    return launch_replay_plain(multi, a, depth, choice, st, env);
}

}  // namespace evogp

using namespace evogp;

// internal (host_api.cu): SR fitness over a COMPACT forest - valid prefixes of value / type back to back, tree n at
// [offsets[n], offsets[n + 1]) - as the host path uploads it
int evogp_sr_fitness_packed(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen, unsigned outLen, int useMSE,
                            const float *value, const int16_t *type, const unsigned *offsets, const float *variables,
                            const float *labels, float *fitnesses, void *workspace, size_t workspace_bytes, void *stream) {
    return run_eval(useMSE ? MODE_MSE : MODE_ABS, popSize, dataPoints, gpLen, varLen, outLen, value, type,
                    reinterpret_cast<const int16_t *>(offsets), 0, variables, labels, fitnesses, workspace, workspace_bytes, stream);
}

// internal: SR fitness with tree lengths given as a compact int16[popSize] array
int evogp_sr_fitness_compact_len(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen, unsigned outLen,
                                 int useMSE, const float *value, const int16_t *type, const int16_t *lengths,
                                 const float *variables, const float *labels, float *fitnesses, void *workspace,
                                 size_t workspace_bytes, void *stream) {
    return run_eval(useMSE ? MODE_MSE : MODE_ABS, popSize, dataPoints, gpLen, varLen, outLen, value, type, lengths, 1,
                    variables, labels, fitnesses, workspace, workspace_bytes, stream);
}

// Diagnostics: run only the lowering pass and hand back the programs (tests compare the fast and the generic pass).
extern "C" int evogp_debug_lower(unsigned popSize, unsigned gpLen, unsigned varLen, unsigned outLen, const float *value,
                                 const int16_t *type, const int16_t *subtree_size, int use_fast, int deep_from,
                                 void *workspace, size_t workspace_bytes, unsigned long long *programs, void *stream) {
    EVOGP_REQUIRE(popSize > 0 && gpLen > 0 && gpLen <= (unsigned)kMaxStack && varLen > 0 && outLen > 0, "bad shape");
    if (workspace_bytes < evogp_eval_workspace_bytes(popSize, gpLen) || workspace == nullptr) {
        set_error("workspace too small");
        return EVOGP_ERR_WORKSPACE;
    }
    int rc = ensure_device_ok();
    if (rc) return rc;
    rc = device_props();
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const Workspace w = carve(workspace, popSize, gpLen);
    const int depth = stack_depth_bound((int)gpLen);
    const int df = deep_from > 0 ? deep_from : kNoDeepSlots;
    EVOGP_CUDA(cudaMemsetAsync(w.prog, 0, prog_bytes(popSize, gpLen), st));
    if (outLen > 1) rc = launch_lower_t<true, false>(w, popSize, gpLen, varLen, outLen, value, type, subtree_size, (int)gpLen, depth, kNoDeepSlots, st);
    else if (use_fast && gpLen <= 64)
        rc = gpLen <= 32 ? launch_lower_fast<1>(w, popSize, gpLen, varLen, outLen, value, type, subtree_size, depth, df, st)
                         : launch_lower_fast<2>(w, popSize, gpLen, varLen, outLen, value, type, subtree_size, depth, df, st);
    else rc = launch_lower_t<false, false>(w, popSize, gpLen, varLen, outLen, value, type, subtree_size, (int)gpLen, depth, df, st);
    if (rc) return rc;
    EVOGP_CUDA(cudaMemcpyAsync(programs, w.prog, prog_bytes(popSize, gpLen), cudaMemcpyDeviceToDevice, st));
    return EVOGP_OK;
}

extern "C" void evogp_eval_set_timing_events(void *begin_event, void *end_event) {
    g_ev_replay_begin = static_cast<cudaEvent_t>(begin_event);
    g_ev_replay_end = static_cast<cudaEvent_t>(end_event);
}

extern "C" int evogp_eval_set_replay_width(int datapoints_per_lane) {
    EVOGP_REQUIRE(datapoints_per_lane == 0 || datapoints_per_lane == 8 || datapoints_per_lane == 16,
                  "replay width must be 0 (automatic), 8 or 16, got %d", datapoints_per_lane);
    g_force_k.store(datapoints_per_lane, std::memory_order_relaxed);
    return EVOGP_OK;
}

extern "C" size_t evogp_eval_workspace_bytes(unsigned popSize, unsigned maxGPLen) {
    return 256 + ((prog_bytes(popSize, maxGPLen) + 15) & ~(size_t)15) + ((chunk_words(popSize) * 4 + 255) & ~(size_t)255);
}

extern "C" int evogp_evaluate(unsigned popSize, unsigned maxGPLen, unsigned varLen, unsigned outLen, const float *value,
                              const int16_t *type, const int16_t *subtree_size, const float *variables, float *results,
                              void *workspace, size_t workspace_bytes, void *stream) {
    return run_eval(MODE_ROWWISE, popSize, 1, maxGPLen, varLen, outLen, value, type, subtree_size, (int)maxGPLen, variables, nullptr,
                    results, workspace, workspace_bytes, stream);
}

extern "C" int evogp_SR_fitness(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen, unsigned outLen,
                                int useMSE, const float *value, const int16_t *type, const int16_t *subtree_size,
                                const float *variables, const float *labels, float *fitnesses, unsigned kernel_type,
                                void *workspace, size_t workspace_bytes, void *stream) {
    (void)kernel_type;   // reference execute_mode 0..4 (forest.py:340-347): one kernel serves all
    return run_eval(useMSE ? MODE_MSE : MODE_ABS, popSize, dataPoints, gpLen, varLen, outLen, value, type,
                    subtree_size, (int)gpLen, variables, labels, fitnesses, workspace, workspace_bytes, stream);
}

extern "C" int evogp_SR_fitness_scatter(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen,
                                        unsigned outLen, int useMSE, const float *value, const int16_t *type,
                                        const int16_t *subtree_size, const float *variables, const float *labels,
                                        float *fitnesses, float *const *peer_fitnesses, unsigned world, unsigned row_offset,
                                        void *workspace, size_t workspace_bytes, void *stream) {
    EVOGP_REQUIRE(peer_fitnesses != nullptr, "peer_fitnesses must not be NULL");
    Scatter sc;
    sc.peers = peer_fitnesses; sc.world = (int)world; sc.row_offset = row_offset;
    return run_eval(useMSE ? MODE_MSE : MODE_ABS, popSize, dataPoints, gpLen, varLen, outLen, value, type,
                    subtree_size, (int)gpLen, variables, labels, fitnesses, workspace, workspace_bytes, stream, sc);
}

extern "C" int evogp_classification_accuracy(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen,
                                             unsigned outLen, const float *value, const int16_t *type,
                                             const int16_t *subtree_size, const float *variables, const float *class_labels,
                                             float max_class, float *accuracy, void *workspace, size_t workspace_bytes,
                                             void *stream) {
    EVOGP_REQUIRE(max_class >= 0.0f, "max_class must be non-negative, got %f", max_class);
    return run_eval(MODE_ACC, popSize, dataPoints, gpLen, varLen, outLen, value, type, subtree_size, (int)gpLen, variables,
                    class_labels, accuracy, workspace, workspace_bytes, stream, Scatter(), max_class / 2, max_class);
}

extern "C" int evogp_batch_forward(unsigned popSize, unsigned dataPoints, unsigned gpLen, unsigned varLen,
                                   unsigned outLen, const float *value, const int16_t *type,
                                   const int16_t *subtree_size, const float *variables, float *results,
                                   void *workspace, size_t workspace_bytes, void *stream) {
    return run_eval(MODE_OUTPUT, popSize, dataPoints, gpLen, varLen, outLen, value, type, subtree_size, (int)gpLen, variables,
                    nullptr, results, workspace, workspace_bytes, stream);
}
