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
    a.pitch = (int)(maxGPLen | 1u) + ((maxGPLen & 1u) ? 2 : 0);   // odd and > L
    const size_t per_tree = (size_t)a.pitch * 8;
    int T = (int)((192 * 1024) / per_tree);
    if (T > 128) T = 128;
    if (T < 1) T = 1;
    a.trees_per_block = T;
    const int threads = ((T + 31) / 32) * 32;
    const size_t smem = per_tree * T;
    const unsigned grid = (popSize + T - 1) / T;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
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
