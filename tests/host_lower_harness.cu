// host_lower_harness.cu — test infrastructure.  Runs the product's lowering pass
// (evogp_b200/csrc/lower.cuh, the same __host__ __device__ code the GPU runs) on the CPU
// and replays the emitted program with a scalar interpreter whose operator bodies are the
// oracle's.  tests/test_lowering.py compares the result with the oracle's direct
// stack evaluation: this pins the instruction set semantics and the lowering logic
// without a GPU.  The scalar interpreter below is the executable specification the
// replay kernel (eval.cu) follows.
#include <cmath>
#include <cstdlib>
#include <vector>
#include <xmmintrin.h>
#include <pmmintrin.h>
#include "../evogp_b200/csrc/lower.cuh"

extern "C" float oracle_apply_unary(unsigned f, float a);
extern "C" float oracle_apply_binary(unsigned f, float a, float b);

using namespace evogp;

template <bool MULTI>
static int run_row(const float *val, const int16_t *typ, int len, int L, int V, int O, const float *X, int N,
                   float *out, int *need_out, int *ninstr_out, int *maxsp_out) {
    const int Lp = (L + 1) & ~1;
    std::vector<uint2> prog(Lp);
    std::vector<uint32_t> SA(L + 1), SB(L + 1);
    const int budget = stack_depth_bound(L);
    const int need = lower_tree<MULTI>(val, typ, len, L, Lp, V, O, budget, prog.data(), SA.data(), SB.data(), 1);
    *need_out = need;
    int ninstr = 0;
    while (ninstr < Lp && (prog[ninstr].x & 0xFF) != C_END) ++ninstr;
    *ninstr_out = ninstr;
    int maxsp = 0;
    std::vector<float> stack(L + 8), outs(O > 0 ? O : 1);
    for (int d = 0; d < N; ++d) {
        const float *x = X + (size_t)d * V;
        float acc = 0.0f;
        int sp = 0;
        for (int o = 0; o < O; ++o) outs[o] = 0.0f;
        for (int pc = 0; pc < Lp; ++pc) {
            const uint32_t w = prog[pc].x;
            float cst;
            std::memcpy(&cst, &prog[pc].y, 4);
            const int code = w & 0xFF;
            if (code == C_END) break;
            if (!MULTI && (w & I_PUSH)) { stack[sp++] = acc; if (sp > maxsp) maxsp = sp; }
            const uint32_t ia = (w >> I_IDXA_SHIFT) & I_IDX_MASK, ib = w >> I_IDXB_SHIFT;
            const float la = (w & I_ACONST) ? cst : x[ia < (uint32_t)V ? ia : 0];
            float r = 0.0f;
            if (code == C_LOAD) { acc = la; continue; }
            if (code == C_NAN) { acc = NAN; for (int o = 0; o < O; ++o) outs[o] = NAN; continue; }
            if (code == C_IF3) {
                if (!MULTI || pc + 1 >= Lp) return -5;
                const uint2 ext = prog[++pc];
                auto leaf = [&](bool is_c, uint32_t word) { float f; std::memcpy(&f, &word, 4); return is_c ? f : x[(word & I_IDX_MASK) < (uint32_t)V ? (word & I_IDX_MASK) : 0]; };
                const float a = la, b = leaf(w & I_IF3_BCONST, ext.x), c = leaf(w & I_IF3_CCONST, ext.y);
                r = a > 0.0f ? b : c;
                if (ib != I_IDX_MASK) outs[ib] += r;
                acc = r;
                continue;
            }
            if (code == C_IF) {
                sp -= 2;
                if (sp < 0) return -2;
                const float t1 = stack[sp + 1], t2 = stack[sp];
                auto pick = [&](uint32_t s) { return s == 0 ? acc : (s == 1 ? t1 : t2); };
                const float a = pick(ia & 3), b = pick((ia >> 2) & 3), c = pick((ia >> 4) & 3);
                r = a > 0.0f ? b : c;
            } else if (code >= C_UA && code < C_UL) {
                r = oracle_apply_unary(code - C_UA + F_SIN, acc);
            } else if (code >= C_UL && code < C_AL) {
                r = oracle_apply_unary(code - C_UL + F_SIN, la);
            } else if (code >= C_AL && code < C_LA) {
                r = oracle_apply_binary(code - C_AL + F_ADD, acc, la);
            } else if (code >= C_LA && code < C_LL) {
                r = oracle_apply_binary(code - C_LA + F_ADD, la, acc);
            } else if (code >= C_LL && code < C_SA) {
                const float lb = (w & I_BCONST) ? cst : x[ib < (uint32_t)V ? ib : 0];
                r = oracle_apply_binary(code - C_LL + F_ADD, la, lb);
            } else if (code >= C_SA && code < C_AS) {
                if (--sp < 0) return -2;
                r = oracle_apply_binary(code - C_SA + F_ADD, stack[sp], acc);
            } else if (code >= C_AS && code < C_COUNT) {
                if (--sp < 0) return -2;
                const float s = stack[sp];
                r = oracle_apply_binary(code - C_AS + F_ADD, acc, s);
            } else {
                return -3;
            }
            if (MULTI && (w & I_OUT) && ib != I_IDX_MASK) outs[ib] += r;
            acc = r;
        }
        if (sp != 0) return -4;
        if (MULTI) for (int o = 0; o < O; ++o) out[(size_t)d * O + o] = outs[o];
        else out[d] = acc;
    }
    *maxsp_out = maxsp;
    return 0;
}

extern "C" int harness_batch_forward(int P, int N, int L, int V, int O, const float *value, const int16_t *type,
                                     const int16_t *size, const float *X, float *out, int *need, int *ninstr,
                                     int *maxsp) {
    // same floating-point environment as the oracle (and the GPU): flush-to-zero
    const unsigned saved_csr = _mm_getcsr();
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
    struct Restore { unsigned v; ~Restore() { _mm_setcsr(v); } } restore{saved_csr};
    for (int n = 0; n < P; ++n) {
        const int len = size[(size_t)n * L];
        int rc;
        if (O > 1) rc = run_row<true>(value + (size_t)n * L, type + (size_t)n * L, len, L, V, O, X, N,
                                      out + (size_t)n * N * O, need + n, ninstr + n, maxsp + n);
        else rc = run_row<false>(value + (size_t)n * L, type + (size_t)n * L, len, L, V, O, X, N,
                                 out + (size_t)n * N * O, need + n, ninstr + n, maxsp + n);
        if (rc) return rc * 1000000 - n;
    }
    return 0;
}

extern "C" int harness_depth_bound(int len) { return stack_depth_bound(len); }
