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
// constant folding in the lowering pass uses the interpreter's own operators (on the GPU: program.cuh's; here: the oracle's)
extern "C" float evogp_host_fold_unary(int u, float a) { return u < 15 ? oracle_apply_unary((unsigned)u + 14u, a) : 0.0f; }
extern "C" float evogp_host_fold_binary(int b, float x, float y) { return b < 13 ? oracle_apply_binary((unsigned)b + 1u, x, y) : 0.0f; }

using namespace evogp;

template <bool MULTI>
static int run_row(const float *val, const int16_t *typ, const int16_t *size, int len, int L, int V, int O, const float *X,
                   int N, float *out, int *need_out, int *ninstr_out, int *maxsp_out, bool split = false,
                   int deep_from = kNoDeepSlots, bool fold = true) {
    const int Lp = (L + 2) & ~1;   // prog_pitch(): one spare slot so C_END always fits
    std::vector<uint2> prog(Lp);
    std::vector<unsigned char> mem(lower_scratch_bytes(L) + 64);
    const LowerScratch scratch = carve_scratch(mem.data(), L);
    const int budget = stack_depth_bound(L);
    const int need = split ? lower_tree<MULTI, true>(Lanes{0, 1}, val, typ, size, len, L, Lp, V, O, budget, prog.data(), scratch, true, deep_from, fold)
                           : lower_tree<MULTI, false>(Lanes{0, 1}, val, typ, size, len, L, Lp, V, O, budget, prog.data(), scratch, true, deep_from, fold);
    *need_out = need;
    int ninstr = 0;
    while (ninstr < Lp && (prog[ninstr].x & I_CODE_MASK) != C_END) ++ninstr;
    *ninstr_out = ninstr;
    int maxsp = 0;
    std::vector<float> stack(L + 8), outs(O > 0 ? O : 1);
    for (int d = 0; d < N; ++d) {
        const float *x = X + (size_t)d * V;
        float acc = 0.0f;
        // operand stack with STATIC slot numbers; `filled` tracks which slots hold a live value so the
        // harness can verify the lowering pass's bookkeeping (push into a free slot, pop the top one)
        std::vector<char> filled(L + 8, 0);
        int height = 0;
        for (int o = 0; o < O; ++o) outs[o] = 0.0f;
        for (int pc = 0; pc < Lp; ++pc) {
            const uint32_t w = prog[pc].x;
            float cst;
            std::memcpy(&cst, &prog[pc].y, 4);
            const int code = w & I_CODE_MASK, form = code >> 4, op = code & 15;
            if (code == C_END) break;
            const uint32_t ia = (w >> I_IDXA_SHIFT) & I_IDXA_MASK, ib = (w >> I_IDXB_SHIFT) & I_IDXB_MASK;
            auto var = [&](uint32_t i) { return x[i < (uint32_t)V ? i : 0]; };
            auto pop = [&](int slot, float &v) {     // must be the top of the stack
                if (slot != height - 1 || !filled[slot]) return false;
                v = stack[slot]; filled[slot] = 0; --height; return true;
            };
            if (code == C_IF3) {
                if (!MULTI || pc + 1 >= Lp) return -5;
                const uint2 ext = prog[++pc];
                auto leaf = [&](bool is_c, uint32_t word) { float f; std::memcpy(&f, &word, 4); return is_c ? f : var(word & I_IDXA_MASK); };
                const float a = (w & I_IF3_ACONST) ? cst : var(ia);
                const float b = leaf(w & I_IF3_BCONST, ext.x), c = leaf(w & I_IF3_CCONST, ext.y);
                const float r = a > 0.0f ? b : c;
                if (ib != I_IDXB_MASK) outs[ib] += r;
                acc = r;
                continue;
            }
            if (!MULTI) {
                const int push = (w & I_PUSH_MASK) >> I_PUSH_SHIFT;
                if (push) {
                    const bool deep_load = code == C_LOAD_V_DEEP || code == C_LOAD_K_DEEP;
                    if (!(code == C_LOAD_V || code == C_LOAD_K || deep_load || form == FM_UV || form == FM_UK || form == FM_VV ||
                          form == FM_VK || form == FM_KV)) return -6;   // PUSH only on fresh-value instructions
                    if ((code <= C_LOAD_K_DEEP) && deep_load != (push - 1 >= deep_from)) return -8;   // LOADs into deep slots are marked in the opcode
                    if (push - 1 != height || filled[push - 1]) return -7;   // static slot == dynamic height
                    stack[push - 1] = acc; filled[push - 1] = 1; ++height;
                    if (height > maxsp) maxsp = height;
                }
            }
            if ((code == C_LOAD_V_DEEP || code == C_LOAD_K_DEEP) && !((w & I_PUSH_MASK) >> I_PUSH_SHIFT)) return -8;
            if (code == C_LOAD_V || code == C_LOAD_V_DEEP) { acc = var(ia); continue; }
            if (code == C_LOAD_K || code == C_LOAD_K_DEEP) { acc = cst; continue; }
            if (code == C_NAN) { acc = NAN; for (int o = 0; o < O; ++o) outs[o] = NAN; continue; }
            float r;
            if (code == C_IF) {
                float t1, t2;
                if (!pop((int)ib + 1, t1) || !pop((int)ib, t2)) return -2;
                auto pick = [&](uint32_t s) { return s == 0 ? acc : (s == 1 ? t1 : t2); };
                const float a = pick(ia & 3), b = pick((ia >> 2) & 3), c = pick((ia >> 4) & 3);
                r = a > 0.0f ? b : c;
            } else if (form >= FM_UA && form <= FM_UK) {
                const float a = form == FM_UA ? acc : (form == FM_UV ? var(ia) : cst);
                r = oracle_apply_unary(op + F_SIN, a);
            } else if (form >= FM_AV && form <= FM_AD) {
                float a, b, s;
                switch (form) {
                case FM_AV: a = acc; b = var(ia); break;
                case FM_AK: a = acc; b = cst; break;
                case FM_VA: a = var(ia); b = acc; break;
                case FM_KA: a = cst; b = acc; break;
                case FM_VV: a = var(ia); b = var(ib); break;
                case FM_VK: a = var(ia); b = cst; break;
                case FM_KV: a = cst; b = var(ia); break;
                case FM_SA: if ((int)ia >= deep_from || !pop((int)ia, s)) return -2; a = s; b = acc; break;
                case FM_AS: if ((int)ia >= deep_from || !pop((int)ia, s)) return -2; a = acc; b = s; break;
                case FM_DA: if ((int)ia < deep_from || !pop((int)ia, s)) return -2; a = s; b = acc; break;
                default: if ((int)ia < deep_from || !pop((int)ia, s)) return -2; a = acc; b = s; break;   // FM_AD
                }
                r = oracle_apply_binary(op + F_ADD, a, b);
            } else {
                return -3;
            }
            if (MULTI && (w & I_OUT) && ib != I_IDXB_MASK) outs[ib] += r;
            acc = r;
        }
        if (height != 0) return -4;
        if (MULTI) for (int o = 0; o < O; ++o) out[(size_t)d * O + o] = outs[o];
        else out[d] = acc;
    }
    *maxsp_out = maxsp;
    return 0;
}

extern "C" int harness_batch_forward(int P, int N, int L, int V, int O, const float *value, const int16_t *type,
                                     const int16_t *size, const float *X, float *out, int *need, int *ninstr,
                                     int *maxsp, int use_sizes) {
    // same floating-point environment as the oracle (and the GPU): flush-to-zero
    const unsigned saved_csr = _mm_getcsr();
    _MM_SET_FLUSH_ZERO_MODE(_MM_FLUSH_ZERO_ON);
    _MM_SET_DENORMALS_ZERO_MODE(_MM_DENORMALS_ZERO_ON);
    struct Restore { unsigned v; ~Restore() { _mm_setcsr(v); } } restore{saved_csr};
    for (int n = 0; n < P; ++n) {
        const int len = size[(size_t)n * L];
        int rc;
        // use_sizes bit 0: 1 = pass the subtree_size row (verified / trusted), 0 = pass none (recomputed from arities);
        // bit 1: lower in split mode (LOAD + acc-form instead of the fresh-value forms)
        const int16_t *srow = (use_sizes & 1) ? size + (size_t)n * L : nullptr;
        const bool split = (use_sizes & 2) != 0;
        const int deep_from = (use_sizes & 4) ? 1 : kNoDeepSlots;   // bit 2: slots >= 1 are deep (exercises the deep opcodes)
        const bool fold = !(use_sizes & 8);                         // bit 3: lower without constant folding
        if (O > 1) rc = run_row<true>(value + (size_t)n * L, type + (size_t)n * L, srow, len, L, V, O, X, N,
                                      out + (size_t)n * N * O, need + n, ninstr + n, maxsp + n, split, deep_from, fold);
        else rc = run_row<false>(value + (size_t)n * L, type + (size_t)n * L, srow, len, L, V, O, X, N,
                                 out + (size_t)n * N * O, need + n, ninstr + n, maxsp + n, split, deep_from, fold);
        if (rc) return rc * 1000000 - n;
    }
    return 0;
}

extern "C" int harness_depth_bound(int len) { return stack_depth_bound(len); }
