// lower.cuh — lowering pass: packed prefix row -> accumulator-machine program.
// One thread per tree (the pass is two O(len) dependent scans); per-node scratch
// lives in shared memory laid out [node][thread] so any mix of node indices across
// a warp is bank-conflict free.
//
// Structure is derived exactly the way the reference's evaluator derives it
// (forward.cu:277-296): from node types and subtree_size[0] only; interior
// subtree_size entries are recomputed, not trusted.
#pragma once
#include "program.cuh"

namespace evogp {

// scratch word A: subtree size [0:11) | instruction slots [11:22) | stack need [22:30) | complex [31]
__device__ __forceinline__ uint32_t packA(int sz, int ni, int need, int cplx) {
    return (uint32_t)sz | ((uint32_t)ni << 11) | ((uint32_t)need << 22) | ((uint32_t)cplx << 31);
}
__device__ __forceinline__ int a_sz(uint32_t a) { return a & 0x7FF; }
__device__ __forceinline__ int a_ni(uint32_t a) { return (a >> 11) & 0x7FF; }
__device__ __forceinline__ int a_need(uint32_t a) { return (a >> 22) & 0xFF; }
__device__ __forceinline__ int a_cplx(uint32_t a) { return a >> 31; }

template <bool MULTI>
__device__ __forceinline__ int node_arity(int t) {
    if (MULTI) t &= NT_MASK;   // single-output mode does not mask (forward.cu:91-94)
    return (t == NT_VAR || t == NT_CONST) ? 0 : (t == NT_UFUNC ? 1 : (t == NT_BFUNC ? 2 : 3));
}

// leaf operand descriptor for slot A (shift 12 / flag A_CONST) or B (shift 22 / B_CONST)
__device__ __forceinline__ void leaf_desc(int t, float v, int V, bool slotB, uint32_t &hdr, uint32_t &cst) {
    if ((t & NT_MASK) == NT_CONST) {
        hdr |= slotB ? I_BCONST : I_ACONST;
        cst = __float_as_uint(v);
    } else {
        int idx = (int)v;                      // forward.cu:100 `(int)node_value`
        idx = idx < 0 ? 0 : (idx >= V ? V - 1 : idx);   // reference reads out of bounds here; clamp
        hdr |= (uint32_t)idx << (slotB ? I_IDXB_SHIFT : I_IDXA_SHIFT);
    }
}

struct LowerArgs {
    const float *value;
    const int16_t *type;
    const int16_t *size;
    uint2 *prog;        // [P][Lp]
    unsigned *sched;    // scheduler words (zeroed here for the replay kernel)
    unsigned *flags;    // [0]: count of malformed rows, [1]: max stack need seen
    int P, L, Lp, V, O, depth_budget;
};

template <bool MULTI>
__global__ void __launch_bounds__(128) lower_kernel(LowerArgs g) {
    extern __shared__ uint32_t scratch[];
    const int T = blockDim.x, tid = threadIdx.x;
    uint32_t *SA = scratch;               // [L][T]
    uint32_t *SB = scratch + g.L * T;     // [L][T]: start [0:11) | live [11]
    const int n = blockIdx.x * T + tid;
    if (blockIdx.x == 0 && tid < 4) g.sched[tid] = 0;
    if (n >= g.P) return;

    const float *val = g.value + (size_t)n * g.L;
    const int16_t *typ = g.type + (size_t)n * g.L;
    uint2 *out = g.prog + (size_t)n * g.Lp;
    int len = g.size[(size_t)n * g.L];
    bool bad = len < 1 || len > g.L;
    if (bad) len = 0;

    // ---- pass A: leaves -> root.  size, slot count, Sethi-Ullman need per subtree ----
    for (int i = len - 1; i >= 0 && !bad; --i) {
        const int t = __ldg(typ + i);
        const int ar = node_arity<MULTI>(t);
        if (ar == 0) {
            SA[i * T + tid] = packA(1, 0, 0, 0);
            continue;
        }
        int c = i + 1, sz = 1;
        uint32_t ch[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (k < ar) {
                if (c >= len) { bad = true; break; }
                ch[k] = SA[c * T + tid];
                c += a_sz(ch[k]);
                sz += a_sz(ch[k]);
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
                const bool both_const = (__ldg(typ + i + 1) & NT_MASK) == NT_CONST && (__ldg(typ + i + 2) & NT_MASK) == NT_CONST;
                const bool is_out = MULTI && (t & NT_OUT);
                ni = (both_const || is_out) ? 2 : 1;
                need = 0;
            } else if (cx && cy) {
                const int nx = a_need(ch[0]), ny = a_need(ch[1]);
                ni = a_ni(ch[0]) + a_ni(ch[1]) + 1;
                need = max(max(nx, ny), min(nx, ny) + 1);
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
            int hi = max(nd[0], max(nd[1], nd[2])), lo = min(nd[0], min(nd[1], nd[2]));
            int mid = nd[0] + nd[1] + nd[2] - hi - lo;
            ni = tot;
            need = max(hi, max(mid + 1, lo + 2));
        }
        SA[i * T + tid] = packA(sz, ni, need, 1);
    }
    if (!bad && len > 0) {
        const uint32_t r = SA[tid];
        if (a_sz(r) != len) bad = true;                          // prefix does not close at len
        else if (a_need(r) > g.depth_budget) bad = true;         // cannot happen (stack_depth_bound)
        else atomicMax(g.flags + 1, (unsigned)a_need(r));
    }
    if (bad) {
        out[0] = make_uint2(C_NAN, 0);
        if (g.Lp > 1) out[1] = make_uint2(C_END, 0);
        atomicAdd(g.flags, 1u);
        return;
    }

    // ---- pass B: root -> leaves.  place each subtree's slot range, emit instructions ----
    {
        const uint32_t r = SA[tid];
        if (!a_cplx(r)) {   // the tree is a single leaf
            uint32_t hdr = C_LOAD, cst = 0;
            leaf_desc(__ldg(typ), __ldg(val), g.V, false, hdr, cst);
            out[0] = make_uint2(hdr, cst);
            if (g.Lp > 1) out[1] = make_uint2(C_END, 0);
            return;
        }
        SB[tid] = 0;   // root: start 0, acc not live
        if (a_ni(r) < g.Lp) out[a_ni(r)] = make_uint2(C_END, 0);
    }
    for (int i = 0; i < len; ++i) {
        const uint32_t me = SA[i * T + tid];
        if (!a_cplx(me)) continue;
        const uint32_t sb = SB[i * T + tid];
        const int st = sb & 0x7FF;
        const uint32_t live_push = (sb >> 11) & 1 ? I_PUSH : 0;
        const int own = st + a_ni(me) - 1;
        const int t = __ldg(typ + i);
        const float v = __ldg(val + i);
        const int ar = node_arity<MULTI>(t);
        const bool is_out = MULTI && (t & NT_OUT);
        unsigned func = (unsigned)v;                 // forward.cu:108 `(unsigned int)node_value`
        uint32_t outbits = 0;
        if (is_out) {                                // kernel.h:105-113: {i16 function, i16 outIndex}
            const uint32_t bits = __float_as_uint(v);
            func = (unsigned)(int)(int16_t)(bits & 0xFFFF);
            const unsigned oi = (unsigned)(int)(int16_t)(bits >> 16);
            outbits = I_OUT | ((oi < (unsigned)g.O ? oi : I_IDX_MASK) << I_IDXB_SHIFT);
        }
        if (ar == 1) {
            const int u = unary_slot(func);
            const int c = i + 1;
            const uint32_t ci = SA[c * T + tid];
            if (a_cplx(ci)) {
                SB[c * T + tid] = sb;    // same start, same liveness
                out[own] = make_uint2((C_UA + u) | outbits, 0);
            } else {
                uint32_t hdr = (C_UL + u) | outbits | live_push, cst = 0;
                leaf_desc(__ldg(typ + c), __ldg(val + c), g.V, false, hdr, cst);
                out[own] = make_uint2(hdr, cst);
            }
        } else if (ar == 2) {
            const int b = binary_slot(func);
            const int x = i + 1;
            const uint32_t xi = SA[x * T + tid];
            const int y = x + a_sz(xi);
            const uint32_t yi = SA[y * T + tid];
            const int cx = a_cplx(xi), cy = a_cplx(yi);
            if (!cx && !cy) {
                const int tx = __ldg(typ + x), ty = __ldg(typ + y);
                const float vx = __ldg(val + x), vy = __ldg(val + y);
                if (a_ni(me) == 2) {
                    uint32_t h0 = C_LOAD | live_push, c0 = 0;
                    leaf_desc(tx, vx, g.V, false, h0, c0);
                    out[st] = make_uint2(h0, c0);
                    uint32_t h1 = (C_AL + b) | outbits, c1 = 0;
                    leaf_desc(ty, vy, g.V, false, h1, c1);
                    out[st + 1] = make_uint2(h1, c1);
                } else {
                    uint32_t hdr = (C_LL + b) | live_push, cst = 0;
                    leaf_desc(tx, vx, g.V, false, hdr, cst);
                    leaf_desc(ty, vy, g.V, true, hdr, cst);
                    out[own] = make_uint2(hdr, cst);
                }
            } else if (cx && cy) {
                const bool x_first = a_need(xi) > a_need(yi);   // ties: right child first, as the reference does
                const int first = x_first ? x : y, second = x_first ? y : x;
                const int ni_first = x_first ? a_ni(xi) : a_ni(yi);
                SB[first * T + tid] = sb;
                SB[second * T + tid] = (uint32_t)(st + ni_first) | (1u << 11);
                out[own] = make_uint2((x_first ? (C_SA + b) : (C_AS + b)) | outbits, 0);
            } else {
                const int cc = cx ? x : y, lf = cx ? y : x;
                SB[cc * T + tid] = sb;
                uint32_t hdr = (cx ? (C_AL + b) : (C_LA + b)) | outbits, cst = 0;
                leaf_desc(__ldg(typ + lf), __ldg(val + lf), g.V, false, hdr, cst);
                out[own] = make_uint2(hdr, cst);
            }
        } else {
            // IF(a, b, c): produce the three values in descending-need order (ties: c, b, a — the
            // reference's order); the last produced sits in acc, the one before on the stack top.
            int pos[3];
            uint32_t inf[3];
            pos[0] = i + 1;
            inf[0] = SA[pos[0] * T + tid];
            pos[1] = pos[0] + a_sz(inf[0]);
            inf[1] = SA[pos[1] * T + tid];
            pos[2] = pos[1] + a_sz(inf[1]);
            inf[2] = SA[pos[2] * T + tid];
            int nd[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) nd[k] = a_cplx(inf[k]) ? a_need(inf[k]) : 0;
            int ord[3] = {2, 1, 0};   // c, b, a
            // stable insertion sort by need, descending
            if (nd[ord[1]] > nd[ord[0]]) { int s = ord[0]; ord[0] = ord[1]; ord[1] = s; }
            if (nd[ord[2]] > nd[ord[1]]) { int s = ord[1]; ord[1] = ord[2]; ord[2] = s; }
            if (nd[ord[1]] > nd[ord[0]]) { int s = ord[0]; ord[0] = ord[1]; ord[1] = s; }
            int cur = st;
            uint32_t src[3] = {0, 0, 0};   // where IF finds a, b, c: 0 acc, 1 stack top, 2 stack top-1
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int k = ord[j];
                const uint32_t lv = j == 0 ? ((sb >> 11) & 1) : 1u;
                if (a_cplx(inf[k])) {
                    SB[pos[k] * T + tid] = (uint32_t)cur | (lv << 11);
                    cur += a_ni(inf[k]);
                } else {
                    uint32_t hdr = C_LOAD | (lv ? I_PUSH : 0), cst = 0;
                    leaf_desc(__ldg(typ + pos[k]), __ldg(val + pos[k]), g.V, false, hdr, cst);
                    out[cur] = make_uint2(hdr, cst);
                    cur += 1;
                }
                src[k] = 2 - j;
            }
            const uint32_t perm = src[0] | (src[1] << 2) | (src[2] << 4);
            out[own] = make_uint2(C_IF | outbits | (perm << I_IDXA_SHIFT), 0);
        }
    }
}

}  // namespace evogp
