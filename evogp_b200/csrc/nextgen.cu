// nextgen.cu — one kernel for a whole generation step over the packed arrays:
//     next[n] = current[order[n]]                                   n <  elite   (elitism)
//     next[n] = mutate?(crossover(survivor a, survivor b))          n >= elite
// This is the "next" row f-1 of SURVEY.md §8: in the reference (and in this repo's
// reference-equivalent path, evogp_b200/algorithm/*.py) a generation is ~14 torch calls around three
// kernels — gather the survivors (algorithm/crossover/default.py:37), draw indices and positions,
// crossover, draw a mutation mask on the CPU, gather the mutants, generate donors, mutate, scatter
// back, concatenate the elites (genetic_programming.py:110-118): ~1.0 ms of glue around 0.15 ms of
// kernels at pop 100000.  Here the survivors are addressed through `order` (no gather), parents and
// positions come from a counter-based generator (Philox4x32-10 keyed by the generation's keys and
// the child index), the child is spliced in shared memory, the donor of a mutation is grown in
// place by the same taus88 routine evogp_generate uses, and the row is written once.
//
// The operators are the reference's (mutation.cu:5-115 splice rules and fallbacks, generate.cu grow
// method, selection/default.py truncation + elitism); the RANDOM STREAM is not (torch's generators
// cannot be reproduced inside a kernel), so this path is statistically — not bit-for-bit —
// equivalent to DefaultSelection/DefaultCrossover/DefaultMutation.  It is deterministic in
// (keys, inputs), which is what replicated multi-GPU populations need.
#include <cstdlib>
#include "gen_tree.cuh"

namespace evogp {

// draws of child n: Philox4x32-10 blocks (n, 0) and (n, 1) under the generation's keys (gen_tree.cuh)
__device__ __forceinline__ uint4 philox_block(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
    uint32_t o[4];
    philox4x32_10(c0, c1, k0, k1, o);
    return make_uint4(o[0], o[1], o[2], o[3]);
}

struct NextGenArgs {
    const float *value;      // current population [P][L]
    const int16_t *type;
    const int16_t *size;
    const long long *order;  // [P] row indices, best first
    const unsigned *keys;    // [2]
    const float *depth2leaf, *roulette, *consts;
    float *ovalue;
    int16_t *otype;
    int16_t *osize;
    int P, L, elite, survivors;
    unsigned V, O, S;
    float mutationRate, outProb, constProb;
};

// source of output slot j of a splice (see splice.cu): recipient prefix / donor subtree / shifted tail
struct SplicePlan {
    int pos, dpos, dsub, diff, newlen;
};
__device__ __forceinline__ SplicePlan plan_splice(int rlen, int pos, int rsub, int dpos, int dsub, int L, bool ok) {
    SplicePlan s;
    ok = ok && dsub >= 1 && rlen + dsub - rsub <= L;        // mutation.cu:163, :279: too long -> keep the recipient
    s.pos = ok ? pos : rlen;
    s.dpos = dpos;
    s.dsub = ok ? dsub : 0;
    s.diff = ok ? dsub - rsub : 0;
    s.newlen = rlen + s.diff;
    return s;
}

template <bool MULTI>
__global__ void __launch_bounds__(256) nextgen_kernel(NextGenArgs g) {
    extern __shared__ __align__(16) uint32_t ng_smem[];
    __shared__ float s_leaf[kMaxFullDepth];
    __shared__ float s_roul[F_END];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int L = g.L;
    if (threadIdx.x < kMaxFullDepth) s_leaf[threadIdx.x] = g.depth2leaf[threadIdx.x];
    if (threadIdx.x < F_END) s_roul[threadIdx.x] = g.roulette[threadIdx.x];
    __syncthreads();
    // per warp: child row (value bits, type|size<<16) and donor row, L words each
    uint32_t *cv = ng_smem + (size_t)warp * 4 * L, *cts = cv + L, *dv = cts + L, *dts = dv + L;
    const uint32_t k0 = g.keys[0], k1 = g.keys[1];

    for (int n = blockIdx.x * nwarp + warp; n < g.P; n += gridDim.x * nwarp) {
        float *ov = g.ovalue + (size_t)n * L;
        int16_t *ot = g.otype + (size_t)n * L;
        int16_t *os = g.osize + (size_t)n * L;
        if (n < g.elite) {   // elitism: verbatim copy of the n-th best row
            const size_t src = (size_t)g.order[n] * L;
            for (int j = lane; j < L; j += 32) {
                ov[j] = g.value[src + j];
                ot[j] = g.type[src + j];
                os[j] = g.size[src + j];
            }
            continue;
        }
        // ---- draws (warp-uniform: every lane computes the same Philox block) ----
        const uint4 r0 = philox_block((uint32_t)n, 0u, k0, k1), r1 = philox_block((uint32_t)n, 1u, k0, k1);
        const size_t lrow = (size_t)g.order[r0.x % (uint32_t)g.survivors] * L;
        const size_t rrow = (size_t)g.order[r0.y % (uint32_t)g.survivors] * L;
        const int llen = g.size[lrow], rlen = g.size[rrow];
        const int lpos = (int)(r0.z % (uint32_t)max(llen, 1)), rpos = (int)(r0.w % (uint32_t)max(rlen, 1));
        const bool rows_ok = llen >= 1 && llen <= L && rlen >= 1 && rlen <= L;
        const SplicePlan cx = plan_splice(llen, lpos, rows_ok ? g.size[lrow + lpos] : 0, rpos,
                                          rows_ok ? g.size[rrow + rpos] : 0, L, rows_ok);
        // ---- crossover into shared memory ----
        for (int j = lane; j < L; j += 32) {
            uint32_t v = 0, ts = 0;
            if (j < cx.newlen) {
                size_t q;
                int add = 0;
                if (j < cx.pos) {
                    q = lrow + j;
                    if (j + g.size[q] > cx.pos) add = cx.diff;     // ancestor of the splice point
                } else if (j < cx.pos + cx.dsub) q = rrow + cx.dpos + j - cx.pos;
                else q = lrow + j - cx.diff;
                v = __float_as_uint(g.value[q]);
                ts = ((uint32_t)(uint16_t)g.type[q]) | ((uint32_t)(uint16_t)(g.size[q] + add) << 16);
            }
            cv[j] = v;
            cts[j] = ts;
        }
        __syncwarp();
        // ---- mutation: with probability mutationRate replace a random subtree by a freshly grown one ----
        const float u = __uint2float_rn(r1.x) * 2.3283064365386963e-10f;
        SplicePlan mu = plan_splice(cx.newlen, cx.newlen, 0, 0, 0, L, false);      // identity
        if (u < g.mutationRate) {
            int dlen = 0;
            if (lane == 0) {
                Taus88 rng(tree_seed((uint32_t)n, k0 ^ 0x5bd1e995u, k1));
                GrowParams gp;
                gp.leaf = s_leaf; gp.roul = s_roul; gp.consts = g.consts;
                gp.L = (unsigned)L; gp.V = g.V; gp.O = g.O; gp.S = g.S; gp.outProb = g.outProb; gp.constProb = g.constProb;
                const int cnt = grow_tree<MULTI>(rng, gp, dv, dts);
                dlen = cnt > 0 ? (int)(dts[0] >> 16) : 0;
            }
            dlen = __shfl_sync(0xffffffffu, dlen, 0);
            const int mpos = (int)(r1.y % (uint32_t)max(cx.newlen, 1));
            mu = plan_splice(cx.newlen, mpos, (int)(cts[mpos] >> 16), 0, dlen, L, dlen >= 1);
            __syncwarp();   // the donor row written by lane 0 is visible to every lane
        }
        // ---- final row: child (or mutated child) -> global, zero-filled tail ----
        for (int j = lane; j < L; j += 32) {
            uint32_t v = 0, t = 0, s = 0;
            if (j < mu.newlen) {
                if (j < mu.pos) {
                    v = cv[j]; t = cts[j] & 0xFFFFu; s = cts[j] >> 16;
                    if (j + (int)s > mu.pos) s += mu.diff;
                } else if (j < mu.pos + mu.dsub) {
                    const int q = j - mu.pos;
                    v = dv[q]; t = dts[q] & 0xFFFFu; s = dts[q] >> 16;
                } else {
                    const int q = j - mu.diff;
                    v = cv[q]; t = cts[q] & 0xFFFFu; s = cts[q] >> 16;
                }
            }
            ov[j] = __uint_as_float(v);
            ot[j] = (int16_t)t;
            os[j] = (int16_t)s;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// nextgen_batch_kernel — single-output populations.  nextgen_kernel above grows the donor of a mutation on lane 0 while
// 31 lanes wait: with mutation_rate 0.2 that serial growth is 420 of the 579 warp-instructions a child costs
// (profiles/r2_genetic_ncu.txt).  Here a warp takes 32 children at a time: first every lane decides its own child's
// mutation coin and grows that child's donor (grow_tree_packed: branch-free, 32 donors at once), then the warp splices
// the 32 children one after the other exactly as before.  Same draws, same children (oracle_next_generation).
// ---------------------------------------------------------------------------
struct NextGenBatchArgs {
    NextGenArgs a;
    unsigned long long magicV, magicS;
    int pitch;     // words per donor row (odd)
};

// NJ = ceil(max_tree_len / 32) when that is 1 or 2: the row loops are unrolled and phase B is software-pipelined (below);
// NJ = 0: any row width, children strictly one after the other.
template <int NJ>
__global__ void __launch_bounds__(256) nextgen_batch_kernel(NextGenBatchArgs gb) {
    const NextGenArgs &g = gb.a;
    extern __shared__ __align__(16) uint32_t ng_smem[];
    __shared__ float s_leaf[16];
    __shared__ float s_roul[32];
    __shared__ int s_mono;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int L = g.L, pitch = gb.pitch;
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
    // per warp: child row (value bits, type | size << 16): 2 L words; then 32 donor rows of `pitch` packed words
    uint32_t *cv = ng_smem + (size_t)warp * (2 * L + 32 * pitch), *cts = cv + L, *donors = cts + L;
    uint32_t *my_donor = donors + (size_t)lane * pitch;
    const uint32_t k0 = g.keys[0], k1 = g.keys[1];
    const int nbatch = (g.P + 31) / 32;

    for (int batch = blockIdx.x * nwarp + warp; batch < nbatch; batch += gridDim.x * nwarp) {
        // ---- phase A: lane i looks after child batch * 32 + i: all its draws, its parents' header (the chain of dependent
        //      loads order -> length -> subtree size runs ONCE per batch, 32 children wide, instead of once per child),
        //      its mutation coin and its donor ----
        const int mine = batch * 32 + lane;
        const bool child = mine >= g.elite && mine < g.P;
        const uint4 q0 = philox_block((uint32_t)mine, 0u, k0, k1), q1 = philox_block((uint32_t)mine, 1u, k0, k1);
        long long my_l = 0, my_r = 0;         // parents' rows (elites: the row to copy)
        int my_llen = 0, my_rlen = 0, my_lpos = 0, my_rpos = 0, my_lsub = 0, my_rsub = 0;
        if (mine < g.P) {
            if (child) {
                my_l = g.order[q0.x % (uint32_t)g.survivors];
                my_r = g.order[q0.y % (uint32_t)g.survivors];
                my_llen = g.size[(size_t)my_l * L];
                my_rlen = g.size[(size_t)my_r * L];
                my_lpos = (int)(q0.z % (uint32_t)max(my_llen, 1));
                my_rpos = (int)(q0.w % (uint32_t)max(my_rlen, 1));
                if (my_llen >= 1 && my_llen <= L && my_rlen >= 1 && my_rlen <= L) {
                    my_lsub = g.size[(size_t)my_l * L + my_lpos];
                    my_rsub = g.size[(size_t)my_r * L + my_rpos];
                }
            } else {
                my_l = g.order[mine];
            }
        }
        const bool mutate = child && __uint2float_rn(q1.x) * 2.3283064365386963e-10f < g.mutationRate;
        const int my_dlen = grow_tree_packed(tree_seed((uint32_t)mine, k0 ^ 0x5bd1e995u, k1), mutate, s_leaf, s_roul, mono, g.V, g.S,
                                             gb.magicV, gb.magicS, g.constProb, L, my_donor);
        __syncwarp();
        // ---- phase B: the warp builds the 32 children one after the other.  Pipelined form (rows of <= 64 slots, no elite
        //      in the batch): the gather of child i + 1 - its two parents' spans, 3 loads per slot - is issued before child i
        //      is assembled and written, so its latency hides behind that work instead of stalling the warp twice per child
        //      (ncu before: long_scoreboard 5.3 warps per issue, 16 % of the DRAM bandwidth) ----
        bool pipelined = false;
        if constexpr (NJ > 0) pipelined = batch * 32 >= g.elite;
        if constexpr (NJ > 0) {
            if (pipelined) {
                struct Gather {
                    uint32_t v[NJ];
                    int t[NJ], s[NJ];
                    SplicePlan cx;
                };
                auto gather = [&](int i, Gather &G) {
                    const size_t lrow = (size_t)__shfl_sync(0xffffffffu, my_l, i) * L, rrow = (size_t)__shfl_sync(0xffffffffu, my_r, i) * L;
                    const int llen = __shfl_sync(0xffffffffu, my_llen, i), rlen = __shfl_sync(0xffffffffu, my_rlen, i);
                    const int lpos = __shfl_sync(0xffffffffu, my_lpos, i), rpos = __shfl_sync(0xffffffffu, my_rpos, i);
                    const int lsub = __shfl_sync(0xffffffffu, my_lsub, i), rsub = __shfl_sync(0xffffffffu, my_rsub, i);
                    const bool rows_ok = llen >= 1 && llen <= L && rlen >= 1 && rlen <= L;
                    G.cx = plan_splice(llen, lpos, lsub, rpos, rsub, L, rows_ok);
#pragma unroll
                    for (int k = 0; k < NJ; ++k) {
                        const int j = lane + 32 * k;
                        G.v[k] = 0; G.t[k] = 0; G.s[k] = 0;
                        if (j < G.cx.newlen && batch * 32 + i < g.P) {
                            const size_t q = j < G.cx.pos ? lrow + j : (j < G.cx.pos + G.cx.dsub ? rrow + G.cx.dpos + j - G.cx.pos : lrow + j - G.cx.diff);
                            G.v[k] = __float_as_uint(g.value[q]);
                            G.t[k] = g.type[q];
                            G.s[k] = g.size[q];
                        }
                    }
                };
                Gather cur;
                gather(0, cur);
                for (int i = 0; i < 32; ++i) {
                    const int n = batch * 32 + i;
                    if (n >= g.P) break;
                    Gather nxt;
                    if (i + 1 < 32) gather(i + 1, nxt);            // in flight while child i is assembled and written
                    const SplicePlan cx = cur.cx;
#pragma unroll
                    for (int k = 0; k < NJ; ++k) {                  // the crossover child, in shared memory
                        const int j = lane + 32 * k;
                        if (j < L) {
                            int sz = cur.s[k];
                            if (j < cx.pos && j + sz > cx.pos) sz += cx.diff;            // ancestor of the splice point
                            cv[j] = cur.v[k];
                            cts[j] = (uint32_t)(uint16_t)cur.t[k] | ((uint32_t)(uint16_t)sz << 16);
                        }
                    }
                    __syncwarp();
                    const uint32_t mut_word = __shfl_sync(0xffffffffu, q1.y, i);
                    const int dlen = __shfl_sync(0xffffffffu, mutate ? my_dlen : -1, i);      // -1: no mutation for this child
                    SplicePlan mu = plan_splice(cx.newlen, cx.newlen, 0, 0, 0, L, false);      // identity
                    if (dlen >= 0) {
                        const int mpos = (int)(mut_word % (uint32_t)max(cx.newlen, 1));
                        mu = plan_splice(cx.newlen, mpos, (int)(cts[mpos] >> 16), 0, dlen, L, dlen >= 1);
                    }
                    const uint32_t *drow = donors + (size_t)i * pitch;
                    float *ov = g.ovalue + (size_t)n * L;
                    int16_t *ot = g.otype + (size_t)n * L;
                    int16_t *os = g.osize + (size_t)n * L;
#pragma unroll
                    for (int k = 0; k < NJ; ++k) {                  // child (or mutated child) -> global, zero-filled tail
                        const int j = lane + 32 * k;
                        if (j >= L) break;
                        uint32_t v = 0, t = 0, sz = 0;
                        if (j < mu.newlen) {
                            if (j < mu.pos) {
                                v = cv[j]; t = cts[j] & 0xFFFFu; sz = cts[j] >> 16;
                                if (j + (int)sz > mu.pos) sz += mu.diff;
                            } else if (j < mu.pos + mu.dsub) {
                                decode_packed_node(drow[j - mu.pos], g.consts, v, t, sz);
                            } else {
                                const int q = j - mu.diff;
                                v = cv[q]; t = cts[q] & 0xFFFFu; sz = cts[q] >> 16;
                            }
                        }
                        ov[j] = __uint_as_float(v);
                        ot[j] = (int16_t)t;
                        os[j] = (int16_t)sz;
                    }
                    __syncwarp();
                    cur = nxt;
                }
            }
        }
        if (!pipelined)
        for (int i = 0; i < 32; ++i) {
            const int n = batch * 32 + i;
            if (n >= g.P) break;
            float *ov = g.ovalue + (size_t)n * L;
            int16_t *ot = g.otype + (size_t)n * L;
            int16_t *os = g.osize + (size_t)n * L;
            const size_t lrow = (size_t)__shfl_sync(0xffffffffu, my_l, i) * L;
            if (n < g.elite) {   // elitism: verbatim copy of the n-th best row
                for (int j = lane; j < L; j += 32) {
                    ov[j] = g.value[lrow + j];
                    ot[j] = g.type[lrow + j];
                    os[j] = g.size[lrow + j];
                }
                continue;
            }
            const size_t rrow = (size_t)__shfl_sync(0xffffffffu, my_r, i) * L;
            const int llen = __shfl_sync(0xffffffffu, my_llen, i), rlen = __shfl_sync(0xffffffffu, my_rlen, i);
            const int lpos = __shfl_sync(0xffffffffu, my_lpos, i), rpos = __shfl_sync(0xffffffffu, my_rpos, i);
            const int lsub = __shfl_sync(0xffffffffu, my_lsub, i), rsub = __shfl_sync(0xffffffffu, my_rsub, i);
            const uint32_t mut_word = __shfl_sync(0xffffffffu, q1.y, i);
            const int dlen = __shfl_sync(0xffffffffu, mutate ? my_dlen : -1, i);      // -1: no mutation for this child
            const bool rows_ok = llen >= 1 && llen <= L && rlen >= 1 && rlen <= L;
            const SplicePlan cx = plan_splice(llen, lpos, lsub, rpos, rsub, L, rows_ok);
            for (int j = lane; j < L; j += 32) {       // crossover into shared memory
                uint32_t v = 0, ts = 0;
                if (j < cx.newlen) {
                    size_t q;
                    int add = 0;
                    if (j < cx.pos) {
                        q = lrow + j;
                        if (j + g.size[q] > cx.pos) add = cx.diff;     // ancestor of the splice point
                    } else if (j < cx.pos + cx.dsub) q = rrow + cx.dpos + j - cx.pos;
                    else q = lrow + j - cx.diff;
                    v = __float_as_uint(g.value[q]);
                    ts = ((uint32_t)(uint16_t)g.type[q]) | ((uint32_t)(uint16_t)(g.size[q] + add) << 16);
                }
                cv[j] = v;
                cts[j] = ts;
            }
            __syncwarp();
            SplicePlan mu = plan_splice(cx.newlen, cx.newlen, 0, 0, 0, L, false);      // identity
            if (dlen >= 0) {
                const int mpos = (int)(mut_word % (uint32_t)max(cx.newlen, 1));
                mu = plan_splice(cx.newlen, mpos, (int)(cts[mpos] >> 16), 0, dlen, L, dlen >= 1);
            }
            const uint32_t *drow = donors + (size_t)i * pitch;
            for (int j = lane; j < L; j += 32) {       // child (or mutated child) -> global, zero-filled tail
                uint32_t v = 0, t = 0, sz = 0;
                if (j < mu.newlen) {
                    if (j < mu.pos) {
                        v = cv[j]; t = cts[j] & 0xFFFFu; sz = cts[j] >> 16;
                        if (j + (int)sz > mu.pos) sz += mu.diff;
                    } else if (j < mu.pos + mu.dsub) {
                        decode_packed_node(drow[j - mu.pos], g.consts, v, t, sz);
                    } else {
                        const int q = j - mu.diff;
                        v = cv[q]; t = cts[q] & 0xFFFFu; sz = cts[q] >> 16;
                    }
                }
                ov[j] = __uint_as_float(v);
                ot[j] = (int16_t)t;
                os[j] = (int16_t)sz;
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

}  // namespace evogp

using namespace evogp;

extern "C" int evogp_next_generation(int popSize, int gpLen, const float *value, const int16_t *type, const int16_t *subtree_size,
                                     const long long *order, int eliteCnt, int survivorCnt, float mutationRate,
                                     unsigned varLen, unsigned outLen, unsigned constSamplesLen, float outProb,
                                     float constProb, const float *depth2leafProbs, const float *rouletteFuncs,
                                     const float *constSamples, const unsigned *keys, float *value_res,
                                     int16_t *type_res, int16_t *subtree_size_res, void *stream) {
    EVOGP_REQUIRE(popSize > 0, "pop_size must be larger than 0, got %d", popSize);
    EVOGP_REQUIRE(gpLen > 0 && gpLen <= kMaxStack, "gp_len must be in (0, %d], got %d", kMaxStack, gpLen);
    EVOGP_REQUIRE(eliteCnt >= 0 && eliteCnt <= popSize, "elite_cnt must be in [0, pop_size], got %d", eliteCnt);
    EVOGP_REQUIRE(survivorCnt > 0 && survivorCnt <= popSize, "survivor_cnt must be in (0, pop_size], got %d", survivorCnt);
    EVOGP_REQUIRE(mutationRate >= 0.f && mutationRate <= 1.f, "mutation_rate must be in [0, 1], got %f", mutationRate);
    EVOGP_REQUIRE(varLen > 0 && outLen > 0 && constSamplesLen > 0, "var_len, out_len, const_samples_len must be positive");
    EVOGP_REQUIRE(outProb >= 0.f && outProb <= 1.f && constProb >= 0.f && constProb <= 1.f, "probabilities must be in [0, 1]");
    int rc = ensure_device_ok();
    if (rc) return rc;
    NextGenArgs a;
    a.value = value; a.type = type; a.size = subtree_size; a.order = order; a.keys = keys;
    a.depth2leaf = depth2leafProbs; a.roulette = rouletteFuncs; a.consts = constSamples;
    a.ovalue = value_res; a.otype = type_res; a.osize = subtree_size_res;
    a.P = popSize; a.L = gpLen; a.elite = eliteCnt; a.survivors = survivorCnt;
    a.V = varLen; a.O = outLen; a.S = constSamplesLen;
    a.mutationRate = mutationRate; a.outProb = outProb; a.constProb = constProb;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static const bool use_batch = []() { const char *e = getenv("EVOGP_NEXTGEN_BATCH"); return !(e && e[0] == '0'); }();   // A/B switch
    if (use_batch && outLen == 1 && varLen <= 65535 && constSamplesLen <= 65535) {
        NextGenBatchArgs gb;
        gb.a = a;
        gb.magicV = ~0ull / varLen + 1ull;
        gb.magicS = ~0ull / constSamplesLen + 1ull;
        gb.pitch = gpLen | 1;
        const size_t per_warp = ((size_t)2 * gpLen + (size_t)32 * gb.pitch) * 4;
        int warps = 8;
        while (warps > 1 && warps * per_warp > 72 * 1024) warps >>= 1;      // three CTAs per SM at max_tree_len 64
        const size_t smem = warps * per_warp;
        if (smem <= 200 * 1024) {
            static const bool pipeline = []() { const char *e = getenv("EVOGP_NEXTGEN_PIPELINE"); return !(e && e[0] == '0'); }();   // A/B switch
            auto kern = !pipeline || gpLen > 64 ? nextgen_batch_kernel<0> : (gpLen > 32 ? nextgen_batch_kernel<2> : nextgen_batch_kernel<1>);
            if (smem > 48 * 1024) EVOGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            long long grid = (((long long)popSize + 31) / 32 + warps - 1) / warps;
            const long long cap = (long long)sms * 3;
            if (grid > cap) grid = cap;
            kern<<<(unsigned)grid, warps * 32, smem, st>>>(gb);
            count_launch();
            return check_launch("next_generation");
        }
    }
    const size_t per_warp = (size_t)gpLen * 16;
    int warps = 8;
    while (warps > 1 && warps * per_warp > 160 * 1024) warps >>= 1;
    const size_t smem = warps * per_warp;
    long long grid = ((long long)popSize + warps - 1) / warps;
    const long long cap = (long long)sms * 8;
    if (grid > cap) grid = cap;
    if (outLen > 1) {
        if (smem > 48 * 1024) EVOGP_CUDA(cudaFuncSetAttribute(nextgen_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nextgen_kernel<true><<<(unsigned)grid, warps * 32, smem, st>>>(a);
    } else {
        if (smem > 48 * 1024) EVOGP_CUDA(cudaFuncSetAttribute(nextgen_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        nextgen_kernel<false><<<(unsigned)grid, warps * 32, smem, st>>>(a);
    }
    count_launch();
    return check_launch("next_generation");
}
