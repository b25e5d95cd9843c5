// lower.cuh — lowering pass: packed prefix row -> accumulator-machine program.
// One thread per tree (the pass is two O(len) dependent scans); per-node scratch
// lives in shared memory laid out [node][thread] so any mix of node indices across
// a warp is bank-conflict free.
//
// Structure is derived exactly the way the reference's evaluator derives it
// (forward.cu:277-296): from node types and subtree_size[0] only; interior
// subtree_size entries are recomputed, not trusted.
//
// lower_tree() is __host__ __device__ so that tests/host_lower_harness.cu can run the
// very same code on the CPU and replay its output against the oracle.
#pragma once
#include "program.cuh"

#include <cstring>
#ifdef __CUDA_ARCH__
#define EVOGP_LDG(p) __ldg(p)
#else
#define EVOGP_LDG(p) (*(p))
#endif

namespace evogp {

__host__ __device__ __forceinline__ uint32_t f32_bits(float x) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(x);
#else
    uint32_t u;
    std::memcpy(&u, &x, 4);
    return u;
#endif
}
// cvt.rzi.u32.f32 / cvt.rzi.s32.f32 semantics (saturating, NaN -> 0) on both sides
__host__ __device__ __forceinline__ unsigned f32_to_u32(float v) {
#ifdef __CUDA_ARCH__
    return __float2uint_rz(v);
#else
    if (!(v > 0.0f)) return 0u;
    return v >= 4294967296.0f ? 0xFFFFFFFFu : (unsigned)v;
#endif
}
__host__ __device__ __forceinline__ int f32_to_i32(float v) {
#ifdef __CUDA_ARCH__
    return __float2int_rz(v);
#else
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return -2147483647 - 1;
    return (int)v;
#endif
}
__host__ __device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }

// scratch word A: subtree size [0:11) | instruction slots [11:22) | stack need [22:30) | complex [31]
__host__ __device__ __forceinline__ uint32_t packA(int sz, int ni, int need, int cplx) {
    return (uint32_t)sz | ((uint32_t)ni << 11) | ((uint32_t)need << 22) | ((uint32_t)cplx << 31);
}
__host__ __device__ __forceinline__ int a_sz(uint32_t a) { return a & 0x7FF; }
__host__ __device__ __forceinline__ int a_ni(uint32_t a) { return (a >> 11) & 0x7FF; }
__host__ __device__ __forceinline__ int a_need(uint32_t a) { return (a >> 22) & 0xFF; }
__host__ __device__ __forceinline__ int a_cplx(uint32_t a) { return a >> 31; }

template <bool MULTI>
__host__ __device__ __forceinline__ int node_arity(int t) {
    if (MULTI) t &= NT_MASK;   // single-output mode does not mask (forward.cu:91-94)
    return (t == NT_VAR || t == NT_CONST) ? 0 : (t == NT_UFUNC ? 1 : (t == NT_BFUNC ? 2 : 3));
}

// a leaf operand: constant (bits) or variable (clamped index)
struct Leaf {
    bool is_const;
    uint32_t bits;   // constant bits, or variable index
};
__host__ __device__ __forceinline__ Leaf leaf_of(int t, float v, int V) {
    Leaf l;
    if ((t & NT_MASK) == NT_CONST) {
        l.is_const = true;
        l.bits = f32_bits(v);
    } else {
        int idx = f32_to_i32(v);               // forward.cu:100 `(int)node_value`
        idx = idx < 0 ? 0 : (idx >= V ? V - 1 : idx);   // reference reads out of bounds here; clamp
        l.is_const = false;
        l.bits = (uint32_t)idx;
    }
    return l;
}
// single-leaf instruction of form `fv` (variable) / `fk` (constant): LOAD, unary, AV/AK, VA/KA
__host__ __device__ __forceinline__ uint2 leaf_instr(int fv_code, int fk_code, Leaf l, uint32_t flags) {
    uint2 r;
    if (l.is_const) {
        r.x = (uint32_t)fk_code | flags;
        r.y = l.bits;
    } else {
        r.x = (uint32_t)fv_code | flags | (l.bits << I_IDXA_SHIFT);
        r.y = 0;
    }
    return r;
}

__host__ __device__ __forceinline__ uint2 mk2(uint32_t a, uint32_t b) {
    uint2 r;
    r.x = a;
    r.y = b;
    return r;
}

// Lowers one row.  SA/SB: two scratch arrays of `len` words, element i at [i * stride].
// Returns the operand-stack need of the program, or -1 for a malformed row (the program
// is then {C_NAN, C_END}).
__host__ __device__ inline int lower_tree_single(const float *val, const int16_t *typ, int len, int L, int Lp, int V, int O,
                                          int depth_budget, uint2 *out, uint32_t *SA, uint16_t *SB, int stride) {
    bool bad = len < 1 || len > L;
    if (bad) len = 0;

    // ---- pass A: leaves -> root.  size, slot count, Sethi-Ullman need per subtree ----
    for (int i = len - 1; i >= 0 && !bad; --i) {
        const int t = EVOGP_LDG(typ + i);
        const int ar = node_arity<false>(t);
        if (ar == 0) {
            SA[i * stride] = packA(1, 0, 0, 0);
            continue;
        }
        int c = i + 1, sz = 1;
        uint32_t ch[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < ar && !bad) {
                if (c >= len) {
                    bad = true;
                } else {
                    ch[k] = SA[c * stride];
                    c += a_sz(ch[k]);
                    sz += a_sz(ch[k]);
                }
            }
        }
        if (bad) break;
        int ni, need;
        if (ar == 1) {
            ni = 1 + (a_cplx(ch[0]) ? a_ni(ch[0]) : 0);
            need = a_cplx(ch[0]) ? a_need(ch[0]) : 0;
        } else if (ar == 2) {
            const int cx = a_cplx(ch[0]), cy = a_cplx(ch[1]);
            if (!cx && !cy) {
                const bool both_const = (EVOGP_LDG(typ + i + 1) & NT_MASK) == NT_CONST &&
                                        (EVOGP_LDG(typ + i + 2) & NT_MASK) == NT_CONST;
                ni = both_const ? 2 : 1;
                need = 0;
            } else if (cx && cy) {
                const int nx = a_need(ch[0]), ny = a_need(ch[1]);
                ni = a_ni(ch[0]) + a_ni(ch[1]) + 1;
                need = imax(imax(nx, ny), imin(nx, ny) + 1);
            } else {
                const uint32_t cc = cx ? ch[0] : ch[1];
                ni = a_ni(cc) + 1;
                need = a_need(cc);
            }
        } else {
            // ternary: every child (leaf or not) is produced as a value: leaf = one C_LOAD slot, need 0
            int nd[3], tot = 1;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                nd[k] = a_cplx(ch[k]) ? a_need(ch[k]) : 0;
                tot += a_cplx(ch[k]) ? a_ni(ch[k]) : 1;
            }
            const int hi = imax(nd[0], imax(nd[1], nd[2])), lo = imin(nd[0], imin(nd[1], nd[2]));
            const int mid = nd[0] + nd[1] + nd[2] - hi - lo;
            ni = tot;
            need = imax(hi, imax(mid + 1, lo + 2));
        }
        SA[i * stride] = packA(sz, ni, need, 1);
    }
    int root_need = 0;
    if (!bad && len > 0) {
        const uint32_t r = SA[0];
        root_need = a_cplx(r) ? a_need(r) : 0;
        if (a_sz(r) != len) bad = true;                  // prefix does not close at len
        else if (root_need > depth_budget) bad = true;   // cannot happen (stack_depth_bound)
    }
    if (bad || len == 0) {
        out[0] = mk2(C_NAN, 0);
        if (Lp > 1) out[1] = mk2(C_END, 0);
        return -1;
    }

    // ---- pass B: root -> leaves.  place each subtree's slot range, emit instructions ----
    {
        const uint32_t r = SA[0];
        if (!a_cplx(r)) {   // the tree is a single leaf
            out[0] = leaf_instr(C_LOAD_V, C_LOAD_K, leaf_of(EVOGP_LDG(typ), EVOGP_LDG(val), V), 0);
            if (Lp > 1) out[1] = mk2(C_END, 0);
            return 0;
        }
        SB[0] = 0;   // root: start 0, acc not live, stack empty  (start [0:11) | live [11] | height [12:16))
        if (a_ni(r) < Lp) out[a_ni(r)] = mk2(C_END, 0);
    }
    for (int i = 0; i < len; ++i) {
        const uint32_t me = SA[i * stride];
        if (!a_cplx(me)) continue;
        const uint32_t sb = SB[i * stride];   // start [0:11) | live [11] | stack height [12:16)
        const int st = sb & 0x7FF;
        const uint32_t live = (sb >> 11) & 1, height = (sb >> 12) & 0xF;   // stack height when this subtree starts
        const uint32_t live_push = live ? ((height + 1) << I_PUSH_SHIFT) : 0;   // a fresh value saves acc into slot `height`
        const uint32_t height_in = height + live;                            // height once that save has happened
        const int own = st + a_ni(me) - 1;
        const int t = EVOGP_LDG(typ + i);
        const float v = EVOGP_LDG(val + i);
        const int ar = node_arity<false>(t);
        const unsigned func = f32_to_u32(v);         // forward.cu:108 `(unsigned int)node_value`
        if (ar == 1) {
            const int u = unary_slot(func);
            const int c = i + 1;
            const uint32_t ci = SA[c * stride];
            if (a_cplx(ci)) {
                SB[c * stride] = (uint16_t)sb;    // same start, same liveness
                out[own] = mk2((uint32_t)opcode(FM_UA, u), 0);
            } else {
                out[own] = leaf_instr(opcode(FM_UV, u), opcode(FM_UK, u),
                                      leaf_of(EVOGP_LDG(typ + c), EVOGP_LDG(val + c), V), live_push);
            }
        } else if (ar == 2) {
            const int b = binary_slot(func);
            const int x = i + 1;
            const uint32_t xi = SA[x * stride];
            const int y = x + a_sz(xi);
            const uint32_t yi = SA[y * stride];
            const int cx = a_cplx(xi), cy = a_cplx(yi);
            if (!cx && !cy) {
                const int tx = EVOGP_LDG(typ + x), ty = EVOGP_LDG(typ + y);
                const float vx = EVOGP_LDG(val + x), vy = EVOGP_LDG(val + y);
                const Leaf lx = leaf_of(tx, vx, V), ly = leaf_of(ty, vy, V);
                if (a_ni(me) == 2) {          // two constants: load the first, then acc (op) const
                    out[st] = leaf_instr(C_LOAD_V, C_LOAD_K, lx, live_push);
                    out[st + 1] = leaf_instr(opcode(FM_AV, b), opcode(FM_AK, b), ly, 0);
                } else if (!lx.is_const && !ly.is_const) {
                    out[own] = mk2((uint32_t)opcode(FM_VV, b) | live_push | (lx.bits << I_IDXA_SHIFT) | (ly.bits << I_IDXB_SHIFT), 0);
                } else if (!lx.is_const) {
                    out[own] = mk2((uint32_t)opcode(FM_VK, b) | live_push | (lx.bits << I_IDXA_SHIFT), ly.bits);
                } else {
                    out[own] = mk2((uint32_t)opcode(FM_KV, b) | live_push | (ly.bits << I_IDXA_SHIFT), lx.bits);
                }
            } else if (cx && cy) {
                const bool x_first = a_need(xi) > a_need(yi);   // ties: right child first, as the reference does
                const int first = x_first ? x : y, second = x_first ? y : x;
                const int ni_first = x_first ? a_ni(xi) : a_ni(yi);
                SB[first * stride] = (uint16_t)sb;
                SB[second * stride] = (uint16_t)((uint32_t)(st + ni_first) | (1u << 11) | (height_in << 12));
                // the first child's value was saved into slot `height_in` by the second child's first instruction
                int form;
                uint32_t slot_arg = 0;
                if (height_in == 0 && kRegSlots > 0) form = x_first ? FM_BA : FM_AB;
                else if (height_in == 1 && kRegSlots > 1) form = x_first ? FM_CA : FM_AC;
                else { form = x_first ? FM_SA : FM_AS; slot_arg = (height_in - kRegSlots) << I_IDXA_SHIFT; }
                out[own] = mk2((uint32_t)opcode(form, b) | slot_arg, 0);
            } else {
                const int cc = cx ? x : y, lf = cx ? y : x;
                SB[cc * stride] = (uint16_t)sb;
                const Leaf l = leaf_of(EVOGP_LDG(typ + lf), EVOGP_LDG(val + lf), V);
                out[own] = cx ? leaf_instr(opcode(FM_AV, b), opcode(FM_AK, b), l, 0)
                              : leaf_instr(opcode(FM_VA, b), opcode(FM_KA, b), l, 0);
            }
        } else {
            // IF(a, b, c): produce the three values in descending-need order (ties: c, b, a — the
            // reference's order); the last produced sits in acc, the one before on the stack top.
            int pos[3];
            uint32_t inf[3];
            pos[0] = i + 1;
            inf[0] = SA[pos[0] * stride];
            pos[1] = pos[0] + a_sz(inf[0]);
            inf[1] = SA[pos[1] * stride];
            pos[2] = pos[1] + a_sz(inf[1]);
            inf[2] = SA[pos[2] * stride];
            int nd[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) nd[k] = a_cplx(inf[k]) ? a_need(inf[k]) : 0;
            int o0 = 2, o1 = 1, o2 = 0;   // c, b, a; stable bubble sort by need, descending
            if (nd[o1] > nd[o0]) { const int s = o0; o0 = o1; o1 = s; }
            if (nd[o2] > nd[o1]) { const int s = o1; o1 = o2; o2 = s; }
            if (nd[o1] > nd[o0]) { const int s = o0; o0 = o1; o1 = s; }
            int cur = st;
            uint32_t perm = 0;   // 2 bits per operand a,b,c: 0 acc, 1 stack top, 2 stack top-1
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int k = j == 0 ? o0 : (j == 1 ? o1 : o2);
                const uint32_t lv = j == 0 ? live : 1u;
                const uint32_t hj = j == 0 ? height : height_in + (uint32_t)(j - 1);   // stack height when entity j starts
                const uint32_t ik = k == 0 ? inf[0] : (k == 1 ? inf[1] : inf[2]);
                const int pk = k == 0 ? pos[0] : (k == 1 ? pos[1] : pos[2]);
                if (a_cplx(ik)) {
                    SB[pk * stride] = (uint16_t)((uint32_t)cur | (lv << 11) | (hj << 12));
                    cur += a_ni(ik);
                } else {
                    out[cur] = leaf_instr(C_LOAD_V, C_LOAD_K, leaf_of(EVOGP_LDG(typ + pk), EVOGP_LDG(val + pk), V),
                                          lv ? ((hj + 1) << I_PUSH_SHIFT) : 0);
                    cur += 1;
                }
                perm |= (uint32_t)(2 - j) << (2 * k);
            }
            // the first two produced values sit in slots height_in (older) and height_in + 1
            out[own] = mk2((uint32_t)C_IF | (perm << I_IDXA_SHIFT) | (height_in << I_IDXB_SHIFT), 0);
        }
    }
    return root_need;
}


// Multi-output rows (out_len > 1).  The reference's multiOutput branch makes EVERY function
// node hand its right-most child's value to its father (forward.cu:236-242: `top_val =
// right_node` is unconditional), so a subtree's value is simply its right-most leaf — the last
// node of its prefix span — and the only arithmetic with an effect is each OUT node applying its
// function to those leaves and adding the result to outs[outIndex].  The program is therefore a
// flat list of leaf-operand instructions, one per OUT node, emitted in the reference's
// processing order (last node first) so the float sums into outs[] associate identically.
// No operand stack.  Returns 0, or -1 for a malformed row.
__host__ __device__ inline int lower_tree_multi(const float *val, const int16_t *typ, int len, int L, int Lp, int V,
                                                int O, uint2 *out, uint32_t *SA, int stride) {
    bool bad = len < 1 || len > L;
    if (bad) len = 0;
    int slot = 0;
    for (int i = len - 1; i >= 0 && !bad; --i) {
        const int t = EVOGP_LDG(typ + i);
        const int ar = node_arity<true>(t);
        if (ar == 0) {
            SA[i * stride] = 1;
            continue;
        }
        int c = i + 1, sz = 1, last[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < ar && !bad) {
                if (c >= len) {
                    bad = true;
                } else {
                    const int cs = (int)SA[c * stride];
                    c += cs;
                    sz += cs;
                    last[k] = c - 1;   // right-most leaf of child k
                }
            }
        }
        if (bad) break;
        SA[i * stride] = (uint32_t)sz;
        if (!(t & NT_OUT)) continue;
        const uint32_t bits = f32_bits(EVOGP_LDG(val + i));           // kernel.h:105-113
        const unsigned func = (unsigned)(int)(int16_t)(bits & 0xFFFF);
        const unsigned oi = (unsigned)(int)(int16_t)(bits >> 16);
        const uint32_t outbits = I_OUT | ((oi < (unsigned)O ? oi : I_IDXB_MASK) << I_IDXB_SHIFT);
        if (slot + 2 > Lp) { bad = true; break; }                     // cannot happen: slots <= nodes
        if (ar == 1) {
            const int u = unary_slot(func);
            out[slot++] = leaf_instr(opcode(FM_UV, u), opcode(FM_UK, u),
                                     leaf_of(EVOGP_LDG(typ + last[0]), EVOGP_LDG(val + last[0]), V), outbits);
        } else if (ar == 2) {
            const int b = binary_slot(func);
            out[slot++] = leaf_instr(C_LOAD_V, C_LOAD_K, leaf_of(EVOGP_LDG(typ + last[0]), EVOGP_LDG(val + last[0]), V), 0);
            out[slot++] = leaf_instr(opcode(FM_AV, b), opcode(FM_AK, b),
                                     leaf_of(EVOGP_LDG(typ + last[1]), EVOGP_LDG(val + last[1]), V), outbits);
        } else {
            // C_IF3, two slots: {hdr, a} {b, c}; b/c words hold constant bits or a variable index
            const Leaf la = leaf_of(EVOGP_LDG(typ + last[0]), EVOGP_LDG(val + last[0]), V);
            const Leaf lb = leaf_of(EVOGP_LDG(typ + last[1]), EVOGP_LDG(val + last[1]), V);
            const Leaf lc = leaf_of(EVOGP_LDG(typ + last[2]), EVOGP_LDG(val + last[2]), V);
            const uint32_t hdr = (uint32_t)C_IF3 | outbits | (la.is_const ? I_IF3_ACONST : (la.bits << I_IDXA_SHIFT)) |
                                 (lb.is_const ? I_IF3_BCONST : 0u) | (lc.is_const ? I_IF3_CCONST : 0u);
            out[slot++] = mk2(hdr, la.is_const ? la.bits : 0u);
            out[slot++] = mk2(lb.bits, lc.bits);
        }
    }
    if (!bad && len > 0 && (int)SA[0] != len) bad = true;
    if (bad || len == 0) {
        out[0] = mk2(C_NAN, 0);
        if (Lp > 1) out[1] = mk2(C_END, 0);
        return -1;
    }
    if (slot < Lp) out[slot] = mk2(C_END, 0);
    return 0;
}

template <bool MULTI>
__host__ __device__ inline int lower_tree(const float *val, const int16_t *typ, int len, int L, int Lp, int V, int O,
                                          int depth_budget, uint2 *out, uint32_t *SA, uint16_t *SB, int stride) {
    if (MULTI) return lower_tree_multi(val, typ, len, L, Lp, V, O, out, SA, stride);
    return lower_tree_single(val, typ, len, L, Lp, V, O, depth_budget, out, SA, SB, stride);
}

#ifdef __CUDACC__
struct LowerArgs {
    const float *value;
    const int16_t *type;
    const int16_t *size;   // tree lengths: size[n * len_stride]  (len_stride = L for a packed subtree_size array)
    uint2 *prog;        // [P][Lp]
    unsigned *sched;    // 64 scheduler words, zeroed here (the replay kernel runs after this one)
    int P, L, Lp, V, O, depth_budget, len_stride;
};

// Shared-memory plan of one CTA (T trees, row width L): SA u32 [L][T], SB u16 [L][T] — per-node scratch laid
// out [node][thread], conflict-free for any mix of node indices.  Rows are read straight from global memory
// (L2-resident): staging them in shared memory with cp.async was measured SLOWER (158 us vs 84 us at config 2)
// because the extra 392 B/thread halves the resident warps of this latency-bound, divergent kernel.
__host__ __device__ inline size_t lower_smem_bytes(int L, int T) { return (size_t)L * T * 6; }

template <bool MULTI>
__global__ void __launch_bounds__(128) lower_kernel(LowerArgs g) {
    extern __shared__ __align__(16) uint32_t scratch[];
    const int T = blockDim.x, tid = threadIdx.x;
    uint32_t *SA = scratch;
    uint16_t *SB = reinterpret_cast<uint16_t *>(SA + (size_t)g.L * T);
    if (blockIdx.x == 0 && tid < 64) g.sched[tid] = 0;     // ticket counters of the replay kernel(s) that follow
    const int n = blockIdx.x * T + tid;
    if (n >= g.P) return;
    const int len = g.size[(size_t)n * g.len_stride];
    lower_tree<MULTI>(g.value + (size_t)n * g.L, g.type + (size_t)n * g.L, len, g.L, g.Lp, g.V, g.O, g.depth_budget,
                      g.prog + (size_t)n * g.Lp, SA + tid, SB + tid, T);
}
#endif

}  // namespace evogp
