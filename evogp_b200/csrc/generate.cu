// generate.cu — random tree generation into the packed arrays.
//
// Replaces generate() (src/evogp/cuda/generate.cu:210-234) and treeGPGenerate
// (:16-173).  Trees are bit-identical to the reference's for the same `keys`:
// per-tree taus88 seeded with the reference's FNV-1a hash (kernel.h:157-180), the
// same draw order, the same roulette scan.
//
// What is different is everything around the draws.  The reference keeps a
// 1024-entry node array plus a 1024-entry frame stack per thread in local memory
// (12 KB/thread) and writes rows one thread at a time.  Here
//   * the frame stack is a register: frames on the stack have strictly increasing
//     depth, so "children still owed at depth d" is a 4-bit field of one 64-bit word;
//   * nodes are built in shared memory (row pitch odd, so lanes at different
//     positions rarely collide on a bank);
//   * subtree sizes need no stack either: scanning the prefix backwards,
//     size[i] = 1 + size[c1] + size[c2] + ... with c1 = i+1, c2 = c1 + size[c1];
//   * rows leave the SM through warp-cooperative, coalesced, zero-filled stores.
#include <cstdlib>
#include "gen_tree.cuh"

namespace evogp {

struct GenArgs {
    const unsigned *keys;
    const float *depth2leaf;   // [10]
    const float *roulette;     // [29]
    const float *consts;       // [S]
    float *ovalue;
    int16_t *otype;
    int16_t *osize;
    unsigned P, L, V, O, S;
    float outProb, constProb;
    int trees_per_block, pitch;
    unsigned long long magicV, magicS;   // floor(2^64 / V) + 1, floor(2^64 / S) + 1 (generate_fast_kernel's fastmod)
};

template <bool MULTI, bool PHILOX>
__global__ void __launch_bounds__(128) generate_kernel(GenArgs g) {
    extern __shared__ uint32_t gsm[];
    __shared__ float s_leaf[kMaxFullDepth];
    __shared__ float s_roul[F_END];
    const int T = g.trees_per_block, pitch = g.pitch;
    uint32_t *s_val = gsm;                      // [T][pitch] value bits
    uint32_t *s_ts = gsm + (size_t)T * pitch;   // [T][pitch] type | size << 16
    if (threadIdx.x < kMaxFullDepth) s_leaf[threadIdx.x] = g.depth2leaf[threadIdx.x];
    if (threadIdx.x < F_END) s_roul[threadIdx.x] = g.roulette[threadIdx.x];
    __syncthreads();

    const unsigned first = blockIdx.x * T;
    const unsigned n = first + threadIdx.x;
    int len = 0;
    if ((int)threadIdx.x < T && n < g.P) {
        uint32_t *val = s_val + (size_t)threadIdx.x * pitch;
        uint32_t *ts = s_ts + (size_t)threadIdx.x * pitch;
        GrowParams gp;
        gp.leaf = s_leaf; gp.roul = s_roul; gp.consts = g.consts;
        gp.L = g.L; gp.V = g.V; gp.O = g.O; gp.S = g.S; gp.outProb = g.outProb; gp.constProb = g.constProb;
        int cnt;
        if constexpr (PHILOX) {
            PhiloxStream rng(n, g.keys[0], g.keys[1]);
            cnt = grow_tree<MULTI>(rng, gp, val, ts);
        } else {
            Taus88 rng(tree_seed(n, g.keys[0], g.keys[1]));
            cnt = grow_tree<MULTI>(rng, gp, val, ts);
        }
        len = cnt > 0 ? (int)(ts[0] >> 16) : 0;
        if (cnt < pitch) ts[cnt] = 0;
        // remember the valid length in the padding word of the row (pitch > L always)
        val[pitch - 1] = (uint32_t)len;
    }
    __syncthreads();

    // cooperative, coalesced, zero-filled write-out: one warp per row
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int L = (int)g.L;
    for (int r = warp; r < T; r += nwarp) {
        const unsigned row = first + r;
        if (row >= g.P) break;
        const uint32_t *val = s_val + (size_t)r * pitch;
        const uint32_t *ts = s_ts + (size_t)r * pitch;
        const int rl = (int)val[pitch - 1];
        float *ov = g.ovalue + (size_t)row * L;
        int16_t *ot = g.otype + (size_t)row * L;
        int16_t *os = g.osize + (size_t)row * L;
        if ((L & 1) == 0) {
            for (int j = lane * 2; j < L; j += 64) {
                const uint32_t v0 = j < rl ? val[j] : 0u, v1 = j + 1 < rl ? val[j + 1] : 0u;
                const uint32_t a0 = j < rl ? ts[j] : 0u, a1 = j + 1 < rl ? ts[j + 1] : 0u;
                *reinterpret_cast<uint2 *>(ov + j) = make_uint2(v0, v1);
                *reinterpret_cast<uint32_t *>(ot + j) = (a0 & 0xFFFFu) | (a1 << 16);
                *reinterpret_cast<uint32_t *>(os + j) = (a0 >> 16) | (a1 & 0xFFFF0000u);
            }
        } else {
            for (int j = lane; j < L; j += 32) {
                const uint32_t v = j < rl ? val[j] : 0u, a = j < rl ? ts[j] : 0u;
                ov[j] = __uint_as_float(v);
                ot[j] = (int16_t)(a & 0xFFFFu);
                os[j] = (int16_t)(a >> 16);
            }
        }
    }
}

}  // namespace evogp

using namespace evogp;

// ---------------------------------------------------------------------------
// generate_fast_kernel — the taus88 mode at full residency.
//
// generate_kernel above is bound by its own shape: 8 bytes of shared memory per node slot cap it at 12 warps per SM, the
// roulette is scanned downwards entry by entry (~25 iterations per function node with four functions in use), every
// lane's function / leaf branch serialises the warp, and one thread per tree walks the row backwards again for the
// subtree sizes (154 us for 100000 trees = 5 % of the HBM roofline for the 51 MB it writes, BENCH r2).  Same draws,
// same trees (tests/test_gpu_parity.py::test_generate_bit_exact, tests/golden), different everything else:
//   * a node is ONE 32-bit word in shared memory (variable / constant-sample index or function id, 3-bit type, 12-bit
//     size) -> 24 warps per SM at max_tree_len 64; values are decoded at write-out;
//   * the body of the growth loop is branch-free: the draws a leaf needs beyond a function's are made speculatively and
//     the generator state is committed with selects; the roulette is a 5-step binary search (it is a cumulative sum;
//     a non-monotone table, possible through the raw-tensor argument, keeps the reference's downward scan);
//   * subtree sizes are finalised as frames pop (a frame's function node spans [start, cnt); the open frames are a
//     linked list threaded through the spare bits of the function words), no second pass; the frame counters are
//     2 bits per depth in one 32-bit register; `raw % V` / `raw % S` use precomputed 64-bit reciprocals;
//   * each warp writes its own 32 rows as soon as its longest tree is done (no CTA barrier).
// ---------------------------------------------------------------------------
// single-output trees (out_len == 1; multi-output populations keep generate_kernel)
__global__ void __launch_bounds__(256, 3) generate_fast_kernel(GenArgs g) {
    extern __shared__ uint32_t gsm[];
    __shared__ float s_leaf[16];
    __shared__ float s_roul[32];
    __shared__ int s_mono;
    const int pitch = g.pitch, L = (int)g.L;
    if (threadIdx.x < 16) s_leaf[threadIdx.x] = threadIdx.x < kMaxFullDepth ? g.depth2leaf[threadIdx.x] : 2.0f;   // depth >= 10: a leaf (the reference reads out of bounds there)
    if (threadIdx.x < 32) s_roul[threadIdx.x] = threadIdx.x < F_END ? g.roulette[threadIdx.x] : __int_as_float(0x7f800000);
    __syncthreads();
    if (threadIdx.x == 0) {
        int mono = 1;
        for (int i = 1; i < F_END; ++i) mono &= s_roul[i] >= s_roul[i - 1];
        s_mono = mono;
    }
    __syncthreads();
    const bool mono = s_mono != 0;
    const int lane = threadIdx.x & 31;
    uint32_t *row = gsm + (size_t)threadIdx.x * pitch;
    const unsigned n = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t V = g.V, S = g.S;
    const uint64_t MV = g.magicV, MS = g.magicS;
    const float constProb = g.constProb;

    const int len = grow_tree_packed(tree_seed(n, g.keys[0], g.keys[1]), n < g.P, s_leaf, s_roul, mono, V, S, MV, MS, constProb, L, row);
    __syncwarp();

    // ---- this warp's 32 rows leave through coalesced, zero-filled stores ----
    const unsigned first = blockIdx.x * blockDim.x + (threadIdx.x & ~31);
    const uint32_t *rows = gsm + (size_t)(threadIdx.x & ~31) * pitch;
    for (int r = 0; r < 32; ++r) {
        const unsigned tree = first + r;
        if (tree >= g.P) break;
        const int rl = __shfl_sync(0xffffffffu, len, r);
        write_packed_row(rows + (size_t)r * pitch, rl, lane, L, g.consts, g.ovalue + (size_t)tree * L, g.otype + (size_t)tree * L,
                         g.osize + (size_t)tree * L);
    }
}

// ---------------------------------------------------------------------------
// generate_balanced_kernel - the same growth, lanes re-armed.
//
// In generate_fast_kernel a warp runs until the longest of its 32 trees is done: tree lengths spread from 1 to
// max_tree_len, so 14.6 of 32 lanes are active on average (ncu, profiles/r2_genetic_final_ncu.txt) and, at 100000 trees,
// the whole launch is one wave whose duration is that of its slowest warp.  Here a warp owns a contiguous span of
// `per_warp` trees (more than 32): the 32 lanes start the first 32; whenever lanes finish (one ballot per node step) the
// warp writes their rows out - coalesced, one row at a time, as before - and hands them the next trees of the span.  The
// grid is a multiple of the SM count, so every SM carries the same number of spans.  A tree's nodes depend on its index
// alone (seed = f(index, keys)): the output is bit-identical to the other kernels'.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 3) generate_balanced_kernel(GenArgs g) {
    extern __shared__ uint32_t gsm[];
    __shared__ float s_leaf[16];
    __shared__ float s_roul[32];
    __shared__ int s_mono;
    const int pitch = g.pitch, L = (int)g.L;
    if (threadIdx.x < 16) s_leaf[threadIdx.x] = threadIdx.x < kMaxFullDepth ? g.depth2leaf[threadIdx.x] : 2.0f;
    if (threadIdx.x < 32) s_roul[threadIdx.x] = threadIdx.x < F_END ? g.roulette[threadIdx.x] : __int_as_float(0x7f800000);
    __syncthreads();
    if (threadIdx.x == 0) {
        int mono = 1;
        for (int i = 1; i < F_END; ++i) mono &= s_roul[i] >= s_roul[i - 1];
        s_mono = mono;
    }
    __syncthreads();
    const bool mono = s_mono != 0;
    const int lane = threadIdx.x & 31;
    uint32_t *row = gsm + (size_t)threadIdx.x * pitch;
    const uint32_t *rows = gsm + (size_t)(threadIdx.x & ~31) * pitch;
    const uint32_t V = g.V, S = g.S;
    const uint64_t MV = g.magicV, MS = g.magicS;
    const float constProb = g.constProb;
    const uint32_t k0 = g.keys[0], k1 = g.keys[1];

    const unsigned span = (unsigned)g.trees_per_block;                       // trees per WARP in this kernel
    const unsigned wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const unsigned long long b0 = (unsigned long long)wid * span;
    const unsigned t0 = b0 < g.P ? (unsigned)b0 : g.P;
    const unsigned t1 = g.P - t0 < span ? g.P : t0 + span;
    unsigned next = t0 + 32 < t1 ? t0 + 32 : t1;                             // first tree of the span not handed out yet
    unsigned mine = t0 + (unsigned)lane;                                     // the tree this lane is growing
    bool busy = mine < t1;
    PackedGrowth t;
    t.start(tree_seed(mine, k0, k1), busy);
    while (__any_sync(0xffffffffu, busy)) {
        if (t.growing(L)) t.step(s_leaf, s_roul, mono, V, S, MV, MS, constProb, row);
        const bool fin = busy && !t.growing(L);
        unsigned done = __ballot_sync(0xffffffffu, fin);
        if (done == 0u) continue;
        const int len = fin ? t.finish(row) : 0;
        __syncwarp();
        const unsigned rank = __popc(done & ((1u << lane) - 1u));
        const unsigned handed = __popc(done);
        while (done) {                                                      // rows of the finished lanes, one at a time
            const int r = __ffs(done) - 1;
            done &= done - 1u;
            const unsigned tree = __shfl_sync(0xffffffffu, mine, r);
            const int rl = __shfl_sync(0xffffffffu, len, r);
            write_packed_row(rows + (size_t)r * pitch, rl, lane, L, g.consts, g.ovalue + (size_t)tree * L, g.otype + (size_t)tree * L,
                             g.osize + (size_t)tree * L);
        }
        __syncwarp();                                                       // the rows were read; their lanes may overwrite them
        if (fin) {
            mine = next + rank;
            busy = mine < t1;
            t.start(tree_seed(mine, k0, k1), busy);
        }
        next = next + handed < t1 ? next + handed : t1;
    }
}

// EVOGP_GENERATE_FAST=0 keeps the one-size-fits-all kernel (the A/B switch of profiles/; default on)
static const int g_generate_balanced = []() { const char *e = getenv("EVOGP_GENERATE_BALANCED"); return e ? atoi(e) : -1; }();
static int g_sm_count_gen() {
    thread_local int dev_cached = -1, sms = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != dev_cached) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        dev_cached = dev;
    }
    return sms > 0 ? sms : 148;
}
static const bool g_generate_fast = []() { const char *e = getenv("EVOGP_GENERATE_FAST"); return !(e && e[0] == '0'); }();

static int generate_impl(bool philox, unsigned popSize, unsigned maxGPLen, unsigned varLen, unsigned outLen,
                         unsigned constSamplesLen, float outProb, float constProb, const unsigned *keys,
                         const float *depth2leafProbs, const float *rouletteFuncs, const float *constSamples,
                         float *value_res, int16_t *type_res, int16_t *subtree_size_res, void *stream) {
    // torch_wrapper.cu:48-54
    EVOGP_REQUIRE(popSize > 0, "pop_size must be larger than 0, got %u", popSize);
    EVOGP_REQUIRE(maxGPLen > 0 && maxGPLen <= (unsigned)kMaxStack, "gp_len must be in (0, %d], got %u", kMaxStack, maxGPLen);
    EVOGP_REQUIRE(varLen > 0, "var_len must be larger than 0, got %u", varLen);
    EVOGP_REQUIRE(outLen > 0, "out_len must be larger than 0, got %u", outLen);
    EVOGP_REQUIRE(constSamplesLen > 0, "const_samples_len must be larger than 0, got %u", constSamplesLen);
    EVOGP_REQUIRE(outProb >= 0.f && outProb <= 1.f, "out_prob must be in [0, 1], got %f", outProb);
    EVOGP_REQUIRE(constProb >= 0.f && constProb <= 1.f, "const_prob must be in [0, 1], got %f", constProb);
    int rc = ensure_device_ok();
    if (rc) return rc;
    GenArgs a;
    a.keys = keys; a.depth2leaf = depth2leafProbs; a.roulette = rouletteFuncs; a.consts = constSamples;
    a.ovalue = value_res; a.otype = type_res; a.osize = subtree_size_res;
    a.P = popSize; a.L = maxGPLen; a.V = varLen; a.O = outLen; a.S = constSamplesLen;
    a.outProb = outProb; a.constProb = constProb;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the packed-word kernel: taus88 mode, codes that fit 16 bits (variable / constant-sample index, out index << 5)
    if (!philox && g_generate_fast && outLen == 1 && varLen <= 65535 && constSamplesLen <= 65535 && maxGPLen <= 1024) {
        a.magicV = ~0ull / varLen + 1ull;
        a.magicS = ~0ull / constSamplesLen + 1ull;
        a.pitch = (int)(maxGPLen | 1u);                            // words per row, odd
        int lanes = (int)((200 * 1024) / ((size_t)a.pitch * 4) / 3);   // three CTAs per SM
        lanes = lanes >= 256 ? 256 : (lanes / 32) * 32;
        if (lanes < 32) lanes = (size_t)a.pitch * 4 * 32 <= 220 * 1024 ? 32 : 0;
        // EVOGP_GENERATE_BALANCED=<CTAs per SM, 1..3> (default: by population size) runs the lane-re-arming kernel with a
        // grid of that many CTAs per SM; 0 keeps one tree per lane
        // (measured, profiles/README.md: 100000 trees - one wave of the one-tree-per-lane kernel - 64 us against 71 us re-armed;
        //  500000 trees 280 -> 242 us, 1000000 trees 510 -> 430 us)
        const bool many = popSize >= 5u * 768u * (unsigned)g_sm_count_gen() / 2u;      // 2.5 trees per resident lane and more
        if (lanes == 256 && g_generate_balanced != 0 && (g_generate_balanced > 0 ? popSize >= 64u * 8u * (unsigned)g_sm_count_gen() : many)) {
            const unsigned sms = (unsigned)g_sm_count_gen();
            unsigned c = g_generate_balanced > 0 ? (unsigned)g_generate_balanced : (popSize + sms * 8u * 24u) / (sms * 8u * 48u);
            c = c < 1u ? 1u : (c > 3u ? 3u : c);
            const unsigned warps = sms * c * 8u;
            a.trees_per_block = (int)((popSize + warps - 1u) / warps);       // trees per warp
            const size_t smem = (size_t)256 * a.pitch * 4;
            EVOGP_CUDA(cudaFuncSetAttribute(generate_balanced_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            generate_balanced_kernel<<<sms * c, 256, smem, st>>>(a);
            count_launch();
            return check_launch("generate");
        }
        if (lanes >= 32) {
            a.trees_per_block = lanes;
            const size_t smem = (size_t)lanes * a.pitch * 4;
            const unsigned grid = (popSize + lanes - 1) / lanes;
            EVOGP_CUDA(cudaFuncSetAttribute(generate_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            generate_fast_kernel<<<grid, lanes, smem, st>>>(a);
            count_launch();
            return check_launch("generate");
        }
    }
    a.pitch = (int)(maxGPLen | 1u) + ((maxGPLen & 1u) ? 2 : 0);   // odd and > L
    const size_t per_tree = (size_t)a.pitch * 8;
    int T = (int)((192 * 1024) / per_tree);
    if (T > 128) T = 128;
    if (T < 1) T = 1;
    a.trees_per_block = T;
    const int threads = ((T + 31) / 32) * 32;
    const size_t smem = per_tree * T;
    const unsigned grid = (popSize + T - 1) / T;
    auto launch = [&](auto kern) -> int {
        EVOGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, threads, smem, st>>>(a);
        return EVOGP_OK;
    };
    if (outLen > 1) rc = philox ? launch(generate_kernel<true, true>) : launch(generate_kernel<true, false>);
    else rc = philox ? launch(generate_kernel<false, true>) : launch(generate_kernel<false, false>);
    if (rc) return rc;
    count_launch();
    return check_launch("generate");
}

extern "C" int evogp_generate(unsigned popSize, unsigned maxGPLen, unsigned varLen, unsigned outLen,
                              unsigned constSamplesLen, float outProb, float constProb, const unsigned *keys,
                              const float *depth2leafProbs, const float *rouletteFuncs, const float *constSamples,
                              float *value_res, int16_t *type_res, int16_t *subtree_size_res, void *stream) {
    return generate_impl(false, popSize, maxGPLen, varLen, outLen, constSamplesLen, outProb, constProb, keys, depth2leafProbs,
                         rouletteFuncs, constSamples, value_res, type_res, subtree_size_res, stream);
}

extern "C" int evogp_generate_philox(unsigned popSize, unsigned maxGPLen, unsigned varLen, unsigned outLen,
                                     unsigned constSamplesLen, float outProb, float constProb, const unsigned *keys,
                                     const float *depth2leafProbs, const float *rouletteFuncs, const float *constSamples,
                                     float *value_res, int16_t *type_res, int16_t *subtree_size_res, void *stream) {
    return generate_impl(true, popSize, maxGPLen, varLen, outLen, constSamplesLen, outProb, constProb, keys, depth2leafProbs,
                         rouletteFuncs, constSamples, value_res, type_res, subtree_size_res, stream);
}
