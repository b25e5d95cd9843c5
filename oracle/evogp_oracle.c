/*
 * evogp_oracle.c — CPU restatement of EvoGP's packed-forest hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the sm_100a
 * kernels in evogp_b200/csrc and the CPU baseline timed by bench.py.  Nothing
 * in the product package (evogp_b200/) may import, link or call it; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs do.
 *
 * The reference (EMI-Group/evogp) ships no CPU implementation of this path:
 * only the CUDA dispatch key is registered (src/evogp/cuda/torch_wrapper.cu:301-307).
 * Every function below therefore restates one reference CUDA routine, cited
 * by file:line relative to /root/reference/src/evogp/cuda/.
 *
 * Parity pins (see tests/test_oracle.py, tests/golden/):
 *   - RNG: thrust KAT (10000th draw of taus88(341) == 3535848941) and the
 *     keys=(42,0) seeds/draws probed from the reference's kernel.h hash();
 *   - fitness: the hand tree of reference test/fix_bug.py:7-10 -> 0.5;
 *   - golden vectors produced by the reference's own CUDA kernels
 *     (oracle/_ref/libevogp_ref.so, built from /root/reference by
 *     oracle/build_ref.sh) on a B200: tests/golden/ref_*.npz.
 *
 * Numeric contract: the reference is compiled with -use_fast_math
 * (setup.py:55): flush-to-zero everywhere and approximate MUFU intrinsics.
 * The CPU cannot reproduce MUFU bit patterns; this file reproduces the
 * *semantics* (special values, NaN rules, pow = exp2(y*log2 x)) and runs with
 * FTZ/DAZ enabled.  Integer results (generate / crossover / mutate) are
 * bit-exact.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <xmmintrin.h>
#include <pmmintrin.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* defs.h:5-8 */
#define ORC_MAX_STACK 1024
#define ORC_MAX_FULL_DEPTH 10
#define ORC_DELTA 1e-9f
#define ORC_MAX_VAL 1e9f
/* defs.h:10-22 */
enum { NT_VAR = 0, NT_CONST = 1, NT_UFUNC = 2, NT_BFUNC = 3, NT_TFUNC = 4, NT_MASK = 0x7F, NT_OUT = 0x80 };
/* defs.h:24-57 */
enum {
    F_IF, F_ADD, F_SUB, F_MUL, F_DIV, F_LOOSE_DIV, F_POW, F_LOOSE_POW, F_MAX, F_MIN, F_LT, F_GT, F_LE, F_GE,
    F_SIN, F_COS, F_TAN, F_SINH, F_COSH, F_TANH, F_LOG, F_LOOSE_LOG, F_EXP, F_INV, F_LOOSE_INV, F_NEG, F_ABS,
    F_SQRT, F_LOOSE_SQRT, F_END
};

/* ------------------------------------------------------------------ */
/* fast-math flavoured scalar ops               */
/* ------------------------------------------------------------------ */

/* div.approx.ftz.f32: a * rcp(b); for 2^126 < |b| < 2^128 the PTX ISA
 * defines the result as 0 (NaN when a is infinite). */
static inline float fm_div(float a, float b) {
    float ab = fabsf(b);
    if (ab > 8.507059173023462e37f /* 2^126 */ && !isinf(b) && !isnan(b)) {
        if (isnan(a)) return a;
        return isinf(a) ? NAN : copysignf(0.0f, a) * copysignf(1.0f, b);
    }
    return a / b;
}
/* __powf: ex2.approx(y * lg2.approx(x)) — NaN for x<0, NaN for 0^0. */
static inline float fm_pow(float a, float b) { return exp2f(b * log2f(a)); }

/* forward.cu:125-168 */
static inline float op_unary(unsigned f, float a) {
    switch (f) {
    case F_SIN: return sinf(a);
    case F_COS: return cosf(a);
    case F_TAN: return tanf(a);
    case F_SINH: return sinhf(a);
    case F_COSH: return coshf(a);
    case F_TANH: return tanhf(a);
    case F_LOG: return logf(a);
    case F_LOOSE_LOG: return a == 0.0f ? -ORC_MAX_VAL : logf(fabsf(a));
    case F_EXP: return expf(a);
    case F_INV: return a == 0.0f ? NAN : fm_div(1.0f, a);
    case F_LOOSE_INV:
        if (fabsf(a) <= ORC_DELTA) a = copysignf(ORC_DELTA, a);
        return fm_div(1.0f, a);
    case F_NEG: return -a;
    case F_ABS: return fabsf(a);
    case F_SQRT: return sqrtf(a);
    case F_LOOSE_SQRT: return sqrtf(fabsf(a));
    default: return 0.0f; /* unknown id leaves top_val{} == 0 (forward.cu:120) */
    }
}
/* forward.cu:169-213 */
static inline float op_binary(unsigned f, float a, float b) {
    switch (f) {
    case F_ADD: return a + b;
    case F_SUB: return a - b;
    case F_MUL: return a * b;
    case F_DIV: return b == 0.0f ? NAN : fm_div(a, b);
    case F_LOOSE_DIV:
        if (fabsf(b) <= ORC_DELTA) b = copysignf(ORC_DELTA, b);
        return fm_div(a, b);
    case F_POW: return fm_pow(a, b);
    case F_LOOSE_POW: return (a == 0.0f && b == 0.0f) ? 0.0f : fm_pow(fabsf(a), b);
    case F_MAX: return a >= b ? a : b;
    case F_MIN: return a <= b ? a : b;
    case F_LT: return a < b ? 1.0f : -1.0f;
    case F_GT: return a > b ? 1.0f : -1.0f;
    case F_LE: return a <= b ? 1.0f : -1.0f;
    case F_GE: return a >= b ? 1.0f : -1.0f;
    default: return 0.0f;
    }
}

/* float -> unsigned as cvt.rzi.u32.f32 does it (saturating, NaN -> 0). */
static inline unsigned f2u(float v) {
    if (!(v > 0.0f)) return 0u;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (unsigned)v;
}
static inline int f2i(float v) {
    if (isnan(v)) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
}

/*
 * One tree on one input row.  forward.cu:246-302 (_treeGPEvalByStack) drives
 * forward.cu:79-244 (_process_node): the prefix row is consumed from its LAST
 * valid node to its first; leaves push, functions pop (first pop = left-most
 * child).  multi != 0 selects the multiOutput template branch
 * (forward.cu:88-93,107-114,236-242).  Returns the final stack height
 * (reference asserts == 1); *root receives the value left on the stack.
 */
static int eval_tree(const float *value, const int16_t *type, int len, const float *vars, int multi, int outLen,
                     float *outs, float *stack, float *root) {
    int top = 0;
    if (multi)
        for (int o = 0; o < outLen; o++) outs[o] = 0.0f;
    for (int i = len - 1; i >= 0; i--) {
        int16_t t = type[i];
        float v = value[i];
        int is_out = 0;
        if (multi) {
            is_out = t & NT_OUT;
            t &= NT_MASK;
        }
        if (t == NT_CONST) {
            stack[top++] = v;
            continue;
        }
        if (t == NT_VAR) {
            stack[top++] = vars[f2i(v)];
            continue;
        }
        unsigned f = f2u(v), oidx = 0;
        if (multi && is_out) { /* kernel.h:105-113 OutNodeValue */
            uint32_t bits;
            memcpy(&bits, &v, 4);
            f = (unsigned)(int)(int16_t)(bits & 0xFFFF);
            oidx = (unsigned)(int)(int16_t)(bits >> 16);
        }
        float r, right;
        if (t == NT_UFUNC) {
            float a = stack[--top];
            right = a;
            r = op_unary(f, a);
        } else if (t == NT_BFUNC) {
            float a = stack[--top], b = stack[--top];
            right = b;
            r = op_binary(f, a, b);
        } else { /* everything else is treated as IF (forward.cu:214-224) */
            float a = stack[--top], b = stack[--top], c = stack[--top];
            right = c;
            r = a > 0.0f ? b : c;
        }
        if (multi) {
            if (is_out && oidx < (unsigned)outLen) outs[oidx] += r;
            r = right;
        }
        stack[top++] = r;
    }
    *root = top > 0 ? stack[top - 1] : 0.0f;
    return top;
}

static inline unsigned ftz_on(void) {
    unsigned old = _mm_getcsr();
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
    return old;
}

/* forward.cu:304-371 (treeGPEvalKernel / evaluate): tree n on variables[n,:]. */
void oracle_evaluate(unsigned P, unsigned L, unsigned V, unsigned O, const float *value, const int16_t *type,
                     const int16_t *size, const float *variables, float *results, int nthreads) {
    const int multi = O > 1;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        unsigned csr = ftz_on();
        float *stack = (float *)malloc(sizeof(float) * (ORC_MAX_STACK + 8));
        float *outs = (float *)malloc(sizeof(float) * (O + 1));
#pragma omp for schedule(dynamic, 64)
        for (long n = 0; n < (long)P; n++) {
            float root;
            int len = size[(size_t)n * L];
            eval_tree(value + (size_t)n * L, type + (size_t)n * L, len, variables + (size_t)n * V, multi, (int)O, outs,
                      stack, &root);
            if (multi)
                for (unsigned o = 0; o < O; o++) results[(size_t)n * O + o] = outs[o];
            else
                results[n] = root;
        }
        free(stack);
        free(outs);
        _mm_setcsr(csr);
    }
}

/* Forest.batch_forward (tree/forest.py:143-176): every tree on every row -> [P,N,O]. */
void oracle_batch_forward(unsigned P, unsigned N, unsigned L, unsigned V, unsigned O, const float *value,
                          const int16_t *type, const int16_t *size, const float *variables, float *results,
                          int nthreads) {
    const int multi = O > 1;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        unsigned csr = ftz_on();
        float *stack = (float *)malloc(sizeof(float) * (ORC_MAX_STACK + 8));
        float *outs = (float *)malloc(sizeof(float) * (O + 1));
#pragma omp for schedule(dynamic, 16)
        for (long n = 0; n < (long)P; n++) {
            int len = size[(size_t)n * L];
            for (unsigned d = 0; d < N; d++) {
                float root;
                eval_tree(value + (size_t)n * L, type + (size_t)n * L, len, variables + (size_t)d * V, multi, (int)O,
                          outs, stack, &root);
                float *dst = results + ((size_t)n * N + d) * O;
                if (multi)
                    for (unsigned o = 0; o < O; o++) dst[o] = outs[o];
                else
                    dst[0] = root;
            }
        }
        free(stack);
        free(outs);
        _mm_setcsr(csr);
    }
}

/*
 * forward.cu:375-479: fitness[i] = (1/N) * sum_n sum_o loss(label[n,o] - out_o)
 * (calculate_fit :375-400, block reduction + atomicAdd :456-471,
 * averageFitnessValueKernel :474-479).  Not divided by O.  The reference sums
 * in a block tree; here the sum is sequential in double then rounded, which
 * is the more accurate of the two — comparisons use a relative tolerance.
 */
void oracle_sr_fitness(unsigned P, unsigned N, unsigned L, unsigned V, unsigned O, int useMSE, const float *value,
                       const int16_t *type, const int16_t *size, const float *variables, const float *labels,
                       float *fitness, int nthreads) {
    const int multi = O > 1;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        unsigned csr = ftz_on();
        float *stack = (float *)malloc(sizeof(float) * (ORC_MAX_STACK + 8));
        float *outs = (float *)malloc(sizeof(float) * (O + 1));
#pragma omp for schedule(dynamic, 16)
        for (long n = 0; n < (long)P; n++) {
            int len = size[(size_t)n * L];
            const float *tv = value + (size_t)n * L;
            const int16_t *tt = type + (size_t)n * L;
            double acc = 0.0;
            for (unsigned d = 0; d < N; d++) {
                float root;
                eval_tree(tv, tt, len, variables + (size_t)d * V, multi, (int)O, outs, stack, &root);
                const float *lab = labels + (size_t)d * O;
                if (multi) {
                    float fit = 0.0f;
                    for (unsigned o = 0; o < O; o++) {
                        float diff = lab[o] - outs[o];
                        fit += useMSE ? diff * diff : fabsf(diff);
                    }
                    acc += (double)fit;
                } else {
                    float diff = lab[0] - root;
                    acc += (double)(useMSE ? diff * diff : fabsf(diff));
                }
            }
            fitness[n] = (float)(acc / (double)N);
        }
        free(stack);
        free(outs);
        _mm_setcsr(csr);
    }
}

/* ------------------------------------------------------------------ */
/* RNG: kernel.h:157-180 hash(), thrust::random::taus88                */
/* (CUDA toolkit CCCL, thrust/random.h:90-97; not vendored by the      */
/* reference), thrust::uniform_real_distribution<float>(0,1).          */
/* ------------------------------------------------------------------ */

/* kernel.h:160-172: 64-bit FNV-1a over the 12 bytes of {n,k1,k2}, low 32 bits. */
uint32_t oracle_hash(uint32_t n, uint32_t k1, uint32_t k2) {
    const uint32_t a[3] = {n, k1, k2};
    const unsigned char *b = (const unsigned char *)a;
    uint64_t h = 14695981039346656037ULL;
    for (int i = 0; i < 12; i++) {
        h ^= (uint64_t)b[i];
        h *= 1099511628211ULL;
    }
    return (uint32_t)h;
}

typedef struct {
    uint32_t z1, z2, z3;
} taus88_t;

/* thrust/random/detail/linear_feedback_shift_engine.inl: b=(((z<<q)^z)>>(k-s));
 * z=((z & (~0u<<(32-k)))<<s)^b with (k,q,s) = (31,13,12),(29,2,4),(28,3,17);
 * xor_combine_engine seeds all three with the same word and XORs the outputs. */
static inline void taus88_seed(taus88_t *g, uint32_t s) { g->z1 = g->z2 = g->z3 = s; }
static inline uint32_t taus88_next(taus88_t *g) {
    uint32_t b;
    b = ((g->z1 << 13) ^ g->z1) >> 19;
    g->z1 = ((g->z1 & 0xFFFFFFFEu) << 12) ^ b;
    b = ((g->z2 << 2) ^ g->z2) >> 25;
    g->z2 = ((g->z2 & 0xFFFFFFF8u) << 4) ^ b;
    b = ((g->z3 << 3) ^ g->z3) >> 11;
    g->z3 = ((g->z3 & 0xFFFFFFF0u) << 17) ^ b;
    return g->z1 ^ g->z2 ^ g->z3;
}
/* uniform_real_distribution.inl:61-75: float(u32) / (1.0f + float(0xFFFFFFFF)). */
static inline float taus88_uniform(taus88_t *g) { return (float)taus88_next(g) / 4294967296.0f; }

/* test hooks */
void oracle_taus88_draws(uint32_t seed, int n, uint32_t *out) {
    taus88_t g;
    taus88_seed(&g, seed);
    for (int i = 0; i < n; i++) out[i] = taus88_next(&g);
}
uint32_t oracle_taus88_nth(uint32_t seed, int n) {
    taus88_t g;
    taus88_seed(&g, seed);
    uint32_t v = 0;
    for (int i = 0; i < n; i++) v = taus88_next(&g);
    return v;
}

/* Philox4x32-10 (Salmon et al., SC'11), counter (c0, c1, 0, 0), key (k0, k1) - the counter-based generator of this
 * library's own additions (evogp_b200/csrc/gen_tree.cuh): evogp_generate_philox, evogp_next_generation,
 * evogp_tournament_select.  Restated here so that those kernels have bit-exact parity tests. */
static void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    uint32_t c2 = 0, c3 = 0;
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
void oracle_philox(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t *out) { philox4x32_10(c0, c1, k0, k1, out); }

/* the per-tree random stream: thrust taus88 (the reference's), or Philox blocks (n, 0x10000 + j / 4) word j % 4 */
#define ORC_PHILOX_GEN_STREAM 0x10000u
typedef struct {
    int philox;
    taus88_t t;
    uint32_t n, k0, k1, blk, buf[4];
    int have;
} rng_t;
static void rng_seed_taus(rng_t *g, uint32_t seed) { g->philox = 0; taus88_seed(&g->t, seed); }
static void rng_seed_philox(rng_t *g, uint32_t n, uint32_t k0, uint32_t k1) {
    g->philox = 1; g->n = n; g->k0 = k0; g->k1 = k1; g->blk = 0; g->have = 0;
}
static uint32_t rng_next(rng_t *g) {
    if (!g->philox) return taus88_next(&g->t);
    if (g->have == 0) {
        philox4x32_10(g->n, ORC_PHILOX_GEN_STREAM + g->blk, g->k0, g->k1, g->buf);
        g->blk++;
        g->have = 4;
    }
    return g->buf[4 - g->have--];
}
static float rng_uniform(rng_t *g) { return (float)rng_next(g) / 4294967296.0f; }

/*
 * generate.cu:16-173 (treeGPGenerate).  Draw order per node:
 *   u (leaf test vs depth2leaf[depth])                        :71
 *   function: r (roulette, downward scan :74-84); multi-output
 *             only: u (vs outProb :88) and, if taken, one raw u32 (:93)
 *   leaf:     u (vs constProb :109), then one raw u32 (:112 / :118)
 * Frames {children_left, depth}: pop, decrement, emit node, re-push the
 * parent frame if it still has children, then push the new function's frame
 * (:60-128).  Subtree sizes by reverse scan (:130-158); only the valid prefix
 * is defined by the reference — this restatement zero-fills the tail.
 * Deviation (documented): depth >= MAX_FULL_DEPTH reads leafProbs out of
 * bounds in the reference; here it is treated as probability 1 (leaf).
 * Returns the node count; gv / gt / gs receive the tree (ORC_MAX_STACK entries each).
 */
static int grow_one(rng_t *g, int cap, unsigned V, unsigned O, unsigned S, float outProb, float constProb,
                    const float *depth2leaf, const float *roulette, const float *constSamples, float *gv, int16_t *gt,
                    int16_t *gs) {
    const int multi = O > 1;
    int16_t fr_childs[ORC_MAX_STACK], fr_depth[ORC_MAX_STACK];
    int nodeSize[ORC_MAX_STACK];
    fr_childs[0] = 1;
    fr_depth[0] = 0;
    int topGP = 0, top = 1;
    while (top > 0 && topGP < cap) {
        --top;
        int16_t cd_childs = (int16_t)(fr_childs[top] - 1), cd_depth = fr_depth[top];
        int16_t new_childs = 0, new_depth = 0;
        float nv;
        int16_t nt;
        float leafp = cd_depth < ORC_MAX_FULL_DEPTH ? depth2leaf[cd_depth] : 2.0f;
        if (rng_uniform(g) >= leafp) {
            float r = rng_uniform(g);
            int k = 0;
            for (int i = F_END - 1; i >= 0; i--)
                if (r >= roulette[i]) {
                    k = i + 1;
                    break;
                }
            int16_t t = k <= F_IF ? NT_TFUNC : (k <= F_GE ? NT_BFUNC : NT_UFUNC);
            nv = (float)k;
            nt = t;
            if (multi && rng_uniform(g) <= outProb) {
                uint32_t bits = ((uint32_t)(uint16_t)(int16_t)k) | ((uint32_t)(uint16_t)(int16_t)(rng_next(g) % O) << 16);
                memcpy(&nv, &bits, 4);
                nt = (int16_t)(t + NT_OUT);
            }
            new_childs = (int16_t)(t - 1);
            new_depth = (int16_t)(cd_depth + 1);
        } else {
            if (rng_uniform(g) <= constProb) {
                nv = constSamples[rng_next(g) % S];
                nt = NT_CONST;
            } else {
                nv = (float)(rng_next(g) % V);
                nt = NT_VAR;
            }
        }
        gv[topGP] = nv;
        gt[topGP] = nt;
        topGP++;
        if (cd_childs > 0) {
            fr_childs[top] = cd_childs;
            fr_depth[top] = cd_depth;
            top++;
        }
        if (new_childs > 0) {
            fr_childs[top] = new_childs;
            fr_depth[top] = new_depth;
            top++;
        }
    }
    top = 0;
    for (int i = topGP - 1; i >= 0; i--) {
        int t = gt[i] & NT_MASK;
        int sz = 1;
        if (t == NT_UFUNC) {
            sz += nodeSize[--top];
        } else if (t == NT_BFUNC) {
            sz += nodeSize[--top];
            sz += nodeSize[--top];
        } else if (t >= NT_TFUNC) {
            sz += nodeSize[--top];
            sz += nodeSize[--top];
            sz += nodeSize[--top];
        }
        nodeSize[top++] = sz;
        gs[i] = (int16_t)sz;
    }
    return topGP;
}

static void generate_any(int philox, unsigned P, unsigned L, unsigned V, unsigned O, unsigned S, float outProb, float constProb,
                         const uint32_t *keys, const float *depth2leaf, const float *roulette, const float *constSamples,
                         float *value_res, int16_t *type_res, int16_t *size_res, int nthreads) {
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads > 0 ? nthreads : 1)
    for (long n = 0; n < (long)P; n++) {
        float gv[ORC_MAX_STACK];
        int16_t gt[ORC_MAX_STACK], gs[ORC_MAX_STACK];
        rng_t g;
        if (philox) rng_seed_philox(&g, (uint32_t)n, keys[0], keys[1]);
        else rng_seed_taus(&g, oracle_hash((uint32_t)n, keys[0], keys[1]));
        grow_one(&g, ORC_MAX_STACK, V, O, S, outProb, constProb, depth2leaf, roulette, constSamples, gv, gt, gs);
        int len = gs[0];
        float *ov = value_res + (size_t)n * L;
        int16_t *ot = type_res + (size_t)n * L;
        int16_t *os = size_res + (size_t)n * L;
        for (unsigned i = 0; i < L; i++) {
            int in = (int)i < len;
            ov[i] = in ? gv[i] : 0.0f;
            ot[i] = in ? gt[i] : 0;
            os[i] = in ? gs[i] : 0;
        }
    }
}

void oracle_generate(unsigned P, unsigned L, unsigned V, unsigned O, unsigned S, float outProb, float constProb,
                     const uint32_t *keys, const float *depth2leaf, const float *roulette, const float *constSamples,
                     float *value_res, int16_t *type_res, int16_t *size_res, int nthreads) {
    generate_any(0, P, L, V, O, S, outProb, constProb, keys, depth2leaf, roulette, constSamples, value_res, type_res, size_res, nthreads);
}

/* evogp_generate_philox (evogp_b200/csrc/generate.cu): the same growth, draws from the Philox stream above */
void oracle_generate_philox(unsigned P, unsigned L, unsigned V, unsigned O, unsigned S, float outProb, float constProb,
                            const uint32_t *keys, const float *depth2leaf, const float *roulette, const float *constSamples,
                            float *value_res, int16_t *type_res, int16_t *size_res, int nthreads) {
    generate_any(1, P, L, V, O, S, outProb, constProb, keys, depth2leaf, roulette, constSamples, value_res, type_res, size_res, nthreads);
}

/* ------------------------------------------------------------------ */
/* splice: mutation.cu:5-115 (_gpTreeReplace)                          */
/* ------------------------------------------------------------------ */
static void copy_row(unsigned L, int len, const float *v, const int16_t *t, const int16_t *s, float *ov, int16_t *ot,
                     int16_t *os) {
    for (unsigned i = 0; i < L; i++) {
        int in = (int)i < len;
        ov[i] = in ? v[i] : 0.0f;
        ot[i] = in ? t[i] : 0;
        os[i] = in ? s[i] : 0;
    }
}

/*
 * recipient [0,pos) ++ donor [dpos, dpos+dsize) ++ recipient [pos+old_sub, old_size).
 * Ancestors of pos get size_diff added; the reference finds them by walking
 * root->pos choosing the child whose extent contains pos (mutation.cu:38-88).
 * That walk is restated literally here (the CUDA kernels use the equivalent
 * closed form  j < pos < j + size[j]).
 */
static void tree_replace(unsigned L, int pos, int dpos, int dsize, int old_offset, int old_size, int diff,
                         const float *v_old, const int16_t *t_old, const int16_t *s_old, const float *v_new,
                         const int16_t *t_new, const int16_t *s_new, float *ov, int16_t *ot, int16_t *os) {
    float sv[ORC_MAX_STACK * 2];
    int16_t st[ORC_MAX_STACK * 2], ss[ORC_MAX_STACK * 2];
    memset(ss, 0, sizeof ss);
    for (int i = 0; i < pos; i++) {
        sv[i] = v_old[i];
        st[i] = t_old[i];
        ss[i] = s_old[i];
    }
    int cur = 0;
    while (cur < pos) {
        ss[cur] = (int16_t)(ss[cur] + diff);
        int t = st[cur] & NT_MASK;
        cur++;
        if (cur >= pos) break;
        if (t == NT_BFUNC) {
            int right = cur + s_old[cur];
            if (!(pos < right)) cur = right;
        } else if (t == NT_TFUNC) {
            int mid = cur + s_old[cur];
            if (pos < mid) continue;
            int right = mid + s_old[mid]; /* reference reads an uninitialised slot when mid == pos; s_old is the intent */
            cur = pos < right ? mid : right;
        }
    }
    for (int i = 0; i < dsize; i++) {
        sv[pos + i] = v_new[dpos + i];
        st[pos + i] = t_new[dpos + i];
        ss[pos + i] = s_new[dpos + i];
    }
    for (int i = old_offset; i < old_size; i++) {
        sv[i + diff] = v_old[i];
        st[i + diff] = t_old[i];
        ss[i + diff] = s_old[i];
    }
    copy_row(L, ss[0], sv, st, ss, ov, ot, os);
}

/* mutation.cu:224-309 (treeGPCrossoverKernel) */
void oracle_crossover(int P_ori, int P_new, int L, const float *value, const int16_t *type, const int16_t *size,
                      const int *left_idx, const int *right_idx, const int *left_node, const int *right_node,
                      float *ov, int16_t *ot, int16_t *os, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (long n = 0; n < (long)P_new; n++) {
        size_t lo = (size_t)left_idx[n] * L, oo = (size_t)n * L;
        const float *lv = value + lo;
        const int16_t *lt = type + lo, *ls = size + lo;
        int left_size = ls[0];
        if (right_idx[n] < 0 || right_idx[n] >= P_ori) {
            copy_row(L, left_size, lv, lt, ls, ov + oo, ot + oo, os + oo);
            continue;
        }
        size_t ro = (size_t)right_idx[n] * L;
        int lsub = ls[left_node[n]], rsub = size[ro + right_node[n]];
        int diff = rsub - lsub;
        if (left_size + diff > L) {
            copy_row(L, left_size, lv, lt, ls, ov + oo, ot + oo, os + oo);
            continue;
        }
        tree_replace(L, left_node[n], right_node[n], rsub, left_node[n] + lsub, left_size, diff, lv, lt, ls, value + ro,
                     type + ro, size + ro, ov + oo, ot + oo, os + oo);
    }
}

/* mutation.cu:118-184 (treeGPMutationKernel) */
void oracle_mutate(int P, int L, const float *value, const int16_t *type, const int16_t *size, const int *mut_idx,
                   const float *nvalue, const int16_t *ntype, const int16_t *nsize, float *ov, int16_t *ot,
                   int16_t *os, int nthreads) {
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (long n = 0; n < (long)P; n++) {
        size_t o = (size_t)n * L;
        const float *lv = value + o;
        const int16_t *lt = type + o, *ls = size + o;
        int old_size = ls[0], pos = mut_idx[n];
        if (pos < 0 || pos >= old_size) {
            copy_row(L, old_size, lv, lt, ls, ov + o, ot + o, os + o);
            continue;
        }
        int osub = ls[pos], nsub = nsize[o];
        int diff = nsub - osub;
        if (old_size + diff > L) {
            copy_row(L, old_size, lv, lt, ls, ov + o, ot + o, os + o);
            continue;
        }
        tree_replace(L, pos, 0, nsub, pos + osub, old_size, diff, lv, lt, ls, nvalue + o, ntype + o, nsize + o, ov + o,
                     ot + o, os + o);
    }
}

/* ------------------------------------------------------------------ */
/* This library's own operators (no counterpart in the reference's      */
/* kernel.h): restated so that their kernels have bit-exact parity.     */
/* ------------------------------------------------------------------ */

/* evogp_extract_subtree (csrc/select.cu) = vmap_subtree / subtensor of
 * src/evogp/algorithm/mutation/mutation_utils.py:6-48: row n becomes the subtree rooted at pos[n], tail zero. */
void oracle_extract_subtree(int P, int L, const float *value, const int16_t *type, const int16_t *size, const int *pos,
                            float *ov, int16_t *ot, int16_t *os) {
    for (long n = 0; n < (long)P; n++) {
        size_t o = (size_t)n * L;
        int p = pos[n];
        int ok = p >= 0 && p < L;
        int len = ok ? size[o + p] : 0;
        for (int j = 0; j < L; j++) {
            int in = ok && j < len && p + j < L;
            ov[o + j] = in ? value[o + p + j] : 0.0f;
            ot[o + j] = in ? type[o + p + j] : 0;
            os[o + j] = in ? size[o + p + j] : 0;
        }
    }
}

/* keyed bijection of [0, n): 4-round Feistel network on 2 * half_bits bits, cycle-walked (csrc/select.cu) */
static uint32_t feistel_perm(uint32_t x, uint32_t n, int half_bits, const uint32_t rk[4]) {
    const uint32_t mask = (1u << half_bits) - 1u;
    do {
        uint32_t l = x >> half_bits, r = x & mask;
        for (int i = 0; i < 4; i++) {
            uint32_t f = (r ^ rk[i]) * 0x9E3779B1u;
            f ^= f >> 15;
            f *= 0x85EBCA77u;
            f ^= f >> 13;
            uint32_t nl = r;
            r = (l ^ f) & mask;
            l = nl;
        }
        x = (l << half_bits) | r;
    } while (x >= n);
    return x;
}
void oracle_feistel_perm(uint32_t n, uint32_t round, const uint32_t *keys, uint32_t *out) {
    int half_bits = 1;
    while ((1ull << (2 * half_bits)) < (unsigned long long)n) half_bits++;
    uint32_t rk[4];
    philox4x32_10(round, 0x30000u, keys[0], keys[1], rk);
    for (uint32_t x = 0; x < n; x++) out[x] = feistel_perm(x, n, half_bits, rk);
}

/* evogp_tournament_select (csrc/select.cu): the semantics of TournamentSelection
 * (src/evogp/algorithm/selection/tournament.py:59-133) with counter-based draws.  Tournament j: contenders i = 0..T-1
 * are Philox words (with replacement) or slice j % (P / T) of permutation j / (P / T) (without); the winner is the nth
 * best, nth geometric in best_p (tournament.py:97-101), ties broken by draw order. */
void oracle_tournament(int P, const float *fitness, int T, float best_p, int replace, int count, const uint32_t *keys,
                       int *winners) {
    const uint32_t k0 = keys[0], k1 = keys[1];
    int half_bits = 1;
    while ((1ull << (2 * half_bits)) < (unsigned long long)P) half_bits++;
    const int per_round = P / T;
    uint32_t *cs = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)T);
    float *fs = (float *)malloc(sizeof(float) * (size_t)T);
    for (int j = 0; j < count; j++) {
        uint32_t d[4];
        philox4x32_10((uint32_t)j, 0x20000u, k0, k1, d);
        int nth = 0;
        if (best_p < 1.0f) {
            float u = (float)d[0] / 4294967296.0f, q = 1.0f - best_p, thr = q;
            while (nth < T && u <= thr) {
                nth++;
                thr = thr * q;
            }
            if (nth >= T) nth = 0;
        }
        int round = j / per_round, slot = j - round * per_round;
        uint32_t rk[4];
        philox4x32_10((uint32_t)round, 0x30000u, k0, k1, rk);
        for (int i = 0; i < T; i++) {
            if (replace) {
                uint32_t w[4];
                philox4x32_10((uint32_t)j, 0x20000u + 1u + (uint32_t)(i >> 2), k0, k1, w);
                cs[i] = w[i & 3] % (uint32_t)P;
            } else {
                cs[i] = feistel_perm((uint32_t)(slot * T + i), (uint32_t)P, half_bits, rk);
            }
            float f = fitness[cs[i]];
            fs[i] = f == f ? f : -INFINITY;
        }
        int win = 0;
        for (int i = 0; i < T; i++) {
            int better = 0;
            for (int m = 0; m < T; m++)
                if (m != i && (fs[m] > fs[i] || (fs[m] == fs[i] && m < i))) better++;
            if (better == nth) win = (int)cs[i];
        }
        winners[j] = win;
    }
    free(cs);
    free(fs);
}

/* evogp_next_generation (csrc/nextgen.cu): a whole generation step — elitism, DefaultCrossover among the survivors,
 * DefaultMutation with freshly grown donors (reference: algorithm/genetic_programming.py:110-118,
 * crossover/default.py, mutation/default.py; splice rules mutation.cu:5-115 incl. the fallbacks :150,163,256,279) —
 * with the kernel's counter-based draws:  r0 = philox(n, 0), r1 = philox(n, 1) under the generation's keys;
 *   left parent  = order[r0.x % survivors], right parent = order[r0.y % survivors],
 *   left pos     = r0.z % len(left),        right pos    = r0.w % len(right),
 *   mutate iff float(r1.x) * 2^-32 < rate,  mutation pos = r1.y % len(child),
 *   donor        = grow (taus88 seeded with hash(n, k0 ^ 0x5bd1e995, k1)), rows capped at L nodes.
 * Built from tree_replace() above (the reference's root->pos walk), not from the kernel's closed form. */
void oracle_next_generation(int P, int L, const float *value, const int16_t *type, const int16_t *size,
                            const long long *order, int elite, int survivors, float rate, unsigned V, unsigned O, unsigned S,
                            float outProb, float constProb, const float *depth2leaf, const float *roulette,
                            const float *constSamples, const uint32_t *keys, float *ov, int16_t *ot, int16_t *os,
                            int nthreads) {
    const uint32_t k0 = keys[0], k1 = keys[1];
#pragma omp parallel for schedule(dynamic, 256) num_threads(nthreads > 0 ? nthreads : 1)
    for (long n = 0; n < (long)P; n++) {
        size_t oo = (size_t)n * L;
        if (n < elite) {
            size_t src = (size_t)order[n] * L;
            memcpy(ov + oo, value + src, sizeof(float) * (size_t)L);
            memcpy(ot + oo, type + src, sizeof(int16_t) * (size_t)L);
            memcpy(os + oo, size + src, sizeof(int16_t) * (size_t)L);
            continue;
        }
        uint32_t r0[4], r1[4];
        philox4x32_10((uint32_t)n, 0u, k0, k1, r0);
        philox4x32_10((uint32_t)n, 1u, k0, k1, r1);
        size_t lrow = (size_t)order[r0[0] % (uint32_t)survivors] * L, rrow = (size_t)order[r0[1] % (uint32_t)survivors] * L;
        int llen = size[lrow], rlen = size[rrow];
        int lpos = (int)(r0[2] % (uint32_t)(llen > 1 ? llen : 1)), rpos = (int)(r0[3] % (uint32_t)(rlen > 1 ? rlen : 1));
        int rows_ok = llen >= 1 && llen <= L && rlen >= 1 && rlen <= L;
        /* crossover into a temporary child */
        float cv[ORC_MAX_STACK];
        int16_t ct[ORC_MAX_STACK], cs[ORC_MAX_STACK];
        int lsub = rows_ok ? size[lrow + lpos] : 0, rsub = rows_ok ? size[rrow + rpos] : 0;
        if (rows_ok && rsub >= 1 && llen + rsub - lsub <= L)
            tree_replace((unsigned)L, lpos, rpos, rsub, lpos + lsub, llen, rsub - lsub, value + lrow, type + lrow, size + lrow,
                         value + rrow, type + rrow, size + rrow, cv, ct, cs);
        else
            copy_row((unsigned)L, llen >= 0 && llen <= L ? llen : 0, value + lrow, type + lrow, size + lrow, cv, ct, cs);
        int clen = cs[0];
        float u = (float)r1[0] / 4294967296.0f;
        if (u < rate) {
            float dv[ORC_MAX_STACK];
            int16_t dt[ORC_MAX_STACK], ds[ORC_MAX_STACK];
            rng_t g;
            rng_seed_taus(&g, oracle_hash((uint32_t)n, k0 ^ 0x5bd1e995u, k1));
            int cnt = grow_one(&g, L, V, O, S, outProb, constProb, depth2leaf, roulette, constSamples, dv, dt, ds);
            int dlen = cnt > 0 ? ds[0] : 0;
            int mpos = (int)(r1[1] % (uint32_t)(clen > 1 ? clen : 1));
            int msub = cs[mpos];
            if (dlen >= 1 && clen + dlen - msub <= L) {
                tree_replace((unsigned)L, mpos, 0, dlen, mpos + msub, clen, dlen - msub, cv, ct, cs, dv, dt, ds, ov + oo, ot + oo,
                             os + oo);
                continue;
            }
        }
        copy_row((unsigned)L, clen, cv, ct, cs, ov + oo, ot + oo, os + oo);
    }
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* scalar operator hooks for tests/host_lower_harness.cu (replays lowered programs on the CPU) */
float oracle_apply_unary(unsigned f, float a) { return op_unary(f, a); }
float oracle_apply_binary(unsigned f, float a, float b) { return op_binary(f, a, b); }
