// lower.cuh — lowering pass: packed prefix row -> accumulator-machine program.
//
// One WARP per tree, lanes over NODES.  The pass is a short sequence of data-parallel phases
// over the row (classify, prefix-sum the instruction slots, rank siblings, propagate slot
// offsets / stack heights root -> leaves, emit), so it has no divergence between trees and no
// per-thread serial scans.  (The first version ran one thread per tree with two dependent
// O(len) scans: latency-bound and divergent, ~140 us for 1e5 trees; see DESIGN.md.)
//
// Structure comes from node types and subtree_size; the sizes are verified against the arities
// and recomputed by one lane when they do not match (the reference's evaluator reads only
// subtree_size[0], forward.cu:283, so a row with stale interior sizes must still evaluate).
//
// Everything is written against a tiny lane abstraction (Lanes) so that
// tests/host_lower_harness.cu runs the very same source serially on the CPU and replays its
// output against the oracle.
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


// ---------------------------------------------------------------------------
// lane abstraction: device = a warp (lane, 32), host = one "lane" doing everything
// ---------------------------------------------------------------------------
struct Lanes {
    int lane, n;
};
__host__ __device__ __forceinline__ void lanes_sync() {
#ifdef __CUDA_ARCH__
    __syncwarp();
#endif
}
__host__ __device__ __forceinline__ bool lanes_any(bool p) {
#ifdef __CUDA_ARCH__
    return __any_sync(0xffffffffu, p);
#else
    return p;
#endif
}
__host__ __device__ __forceinline__ int lanes_max(int v) {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
    return v;
}
// in-place exclusive prefix sum of a[0..n) (u16), a[n] = total
__host__ __device__ inline void lanes_exclusive_scan(Lanes ln, uint16_t *a, int n) {
#ifdef __CUDA_ARCH__
    int carry = 0;
    for (int base = 0; base < n; base += 32) {
        const int j = base + ln.lane;
        const int x = j < n ? a[j] : 0;
        int incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (ln.lane >= o) incl += y;
        }
        if (j < n) a[j] = (uint16_t)(carry + incl - x);
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (ln.lane == 0) a[n] = (uint16_t)carry;
    __syncwarp();
#else
    int carry = 0;
    for (int j = 0; j < n; ++j) {
        const int x = a[j];
        a[j] = (uint16_t)carry;
        carry += x;
    }
    a[n] = (uint16_t)carry;
#endif
}

__host__ __device__ __forceinline__ void lanes_atomic_add(uint32_t *p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
// in-place inclusive prefix sum of a[0..n) (u32, wrap-around arithmetic)
__host__ __device__ inline void lanes_inclusive_scan32(Lanes ln, uint32_t *a, int n) {
#ifdef __CUDA_ARCH__
    uint32_t carry = 0;
    for (int base = 0; base < n; base += 32) {
        const int j = base + ln.lane;
        uint32_t incl = j < n ? a[j] : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (ln.lane >= o) incl += y;
        }
        if (j < n) a[j] = carry + incl;
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
#else
    uint32_t carry = 0;
    for (int j = 0; j < n; ++j) {
        carry += a[j];
        a[j] = carry;
    }
#endif
}

// per-tree scratch (one warp): arrays of L entries (M: L + 1).  In the fused kernel t/v/s are the TMA row buffer
// and the rest aliases the (idle) operand-stack area.
struct LowerScratch {
    int16_t *t;      // node types
    uint32_t *v;     // node value bits
    uint16_t *s;     // subtree sizes
    uint16_t *M;     // marks -> exclusive prefix sum of instruction slots
    uint32_t *D;     // path-sum workspace: [L + 1]
    uint16_t *q;     // value children: rank [0:2) | valid [15]
    uint16_t *st;    // spare (kept so the scratch layout is stable)
};
__host__ __device__ inline size_t lower_scratch_bytes(int L) { return (size_t)L * 18 + 32; }   // 4+2+2 rows, 2+4+2+2 aux
// the part that is not the staged rows: M, D, q, st
__host__ __device__ inline size_t lower_aux_bytes(int L) { return (size_t)L * 10 + 32; }
__host__ __device__ inline LowerScratch carve_aux(void *rows_v, void *rows_t, void *rows_s, void *aux, int L) {
    LowerScratch k;
    k.v = static_cast<uint32_t *>(rows_v);
    k.t = static_cast<int16_t *>(rows_t);
    k.s = static_cast<uint16_t *>(rows_s);
    unsigned char *b = static_cast<unsigned char *>(aux);
    k.D = reinterpret_cast<uint32_t *>(b);                 b += (size_t)(L + 1) * 4;
    k.M = reinterpret_cast<uint16_t *>(b);                 b += (size_t)(L + 1) * 2 + 2;
    k.q = reinterpret_cast<uint16_t *>(b);                 b += (size_t)L * 2;
    k.st = reinterpret_cast<uint16_t *>(b);
    return k;
}
__host__ __device__ inline LowerScratch carve_scratch(void *mem, int L) {
    LowerScratch k;
    unsigned char *b = static_cast<unsigned char *>(mem);
    k.v = reinterpret_cast<uint32_t *>(b);                 b += (size_t)L * 4;
    k.t = reinterpret_cast<int16_t *>(b);                  b += (size_t)L * 2;
    k.s = reinterpret_cast<uint16_t *>(b);                 b += (size_t)L * 2;
    k.D = reinterpret_cast<uint32_t *>(b);                 b += (size_t)(L + 1) * 4;
    k.M = reinterpret_cast<uint16_t *>(b);                 b += (size_t)(L + 1) * 2 + 2;   // keeps 4-byte multiples
    k.q = reinterpret_cast<uint16_t *>(b);                 b += (size_t)L * 2;
    k.st = reinterpret_cast<uint16_t *>(b);
    return k;
}

__host__ __device__ __forceinline__ int arity_of(int t, bool multi) {
    if (multi) t &= NT_MASK;   // single-output mode does not mask (forward.cu:91-94)
    return (t == NT_VAR || t == NT_CONST) ? 0 : (t == NT_UFUNC ? 1 : (t == NT_BFUNC ? 2 : 3));
}
__host__ __device__ __forceinline__ float bits_f32(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    std::memcpy(&f, &u, 4);
    return f;
#endif
}

// ---- constant folding (one level): a function whose operands are all constant LEAVES is itself a constant.  The
// value is computed by the very operator the replay kernels would run per datapoint (program.cuh, same translation
// unit, same -use_fast_math lowering), so every downstream value keeps its bits; the node then reads as a constant
// leaf - an operand of its father - and its instructions disappear from the program (BASELINE configs[1]: 13.9 -> 11.1
// instructions per tree).  Host builds (tests/host_lower_harness.cu) supply the oracle's operators instead.
#ifdef __CUDA_ARCH__
__device__ inline float fold_unary(int u, float a) {
#define EVOGP_FU(n) case n: return unary_op<n>(a);
    switch (u) {
        EVOGP_FU(0) EVOGP_FU(1) EVOGP_FU(2) EVOGP_FU(3) EVOGP_FU(4) EVOGP_FU(5) EVOGP_FU(6) EVOGP_FU(7)
        EVOGP_FU(8) EVOGP_FU(9) EVOGP_FU(10) EVOGP_FU(11) EVOGP_FU(12) EVOGP_FU(13) EVOGP_FU(14)
    default: return 0.0f;
    }
#undef EVOGP_FU
}
__device__ inline float fold_binary(int b, float x, float y) {
#define EVOGP_FB(n) case n: return binary_op<n>(x, y);
    switch (b) {
        EVOGP_FB(0) EVOGP_FB(1) EVOGP_FB(2) EVOGP_FB(3) EVOGP_FB(4) EVOGP_FB(5) EVOGP_FB(6)
        EVOGP_FB(7) EVOGP_FB(8) EVOGP_FB(9) EVOGP_FB(10) EVOGP_FB(11) EVOGP_FB(12)
    default: return 0.0f;
    }
#undef EVOGP_FB
}
#else
extern "C" float evogp_host_fold_unary(int u, float a);            // provided by the host harness (oracle operators)
extern "C" float evogp_host_fold_binary(int b, float x, float y);
inline float fold_unary(int u, float a) { return evogp_host_fold_unary(u, a); }
inline float fold_binary(int b, float x, float y) { return evogp_host_fold_binary(b, x, y); }
#endif

// Stage the row into k.t / k.v / k.s.  Returns false when the length is impossible.
// val == nullptr: the rows are already staged in k.t / k.v (/ k.s).
// size == nullptr with val != nullptr, or have_sizes == false: no size row; sizes are recomputed from the arities.
__host__ __device__ inline bool stage_rows(Lanes ln, const float *val, const int16_t *typ, const int16_t *size, int len,
                                           int L, LowerScratch k, bool have_sizes) {
    if (len < 1 || len > L) return false;
    if (val) {
        for (int j = ln.lane; j < len; j += ln.n) {
            k.t[j] = EVOGP_LDG(typ + j);
            k.v[j] = f32_bits(EVOGP_LDG(val + j));
            k.s[j] = (size && have_sizes) ? (uint16_t)EVOGP_LDG(size + j) : (uint16_t)0;
        }
    } else if (!have_sizes) {
        for (int j = ln.lane; j < len; j += ln.n) k.s[j] = 0;
    }
    lanes_sync();
    return true;
}

// Stale or inconsistent sizes: recompute them from the arities, leaves -> root (one lane; rare).
// Returns false when the prefix does not close at len (a malformed row).
template <bool MULTI>
__host__ __device__ inline bool fix_sizes(Lanes ln, int len, LowerScratch k) {
    bool ok = true;
    if (ln.lane == 0) {
        for (int i = len - 1; i >= 0 && ok; --i) {
            const int ar = arity_of(k.t[i], MULTI);
            int c = i + 1, tot = 1;
            for (int a = 0; a < ar; ++a) {
                if (c >= len) { ok = false; break; }
                tot += k.s[c];
                c += k.s[c];
            }
            k.s[i] = (uint16_t)tot;
        }
        if (ok && (int)k.s[0] != len) ok = false;
    }
    lanes_sync();
    return !lanes_any(!ok);
}

// Stage the row and make sure s[] holds arity-consistent subtree sizes.  Returns false for a malformed row.
template <bool MULTI>
__host__ __device__ inline bool lower_stage(Lanes ln, const float *val, const int16_t *typ, const int16_t *size, int len,
                                            int L, LowerScratch k, bool have_sizes) {
    if (!stage_rows(ln, val, typ, size, len, L, k, have_sizes)) return false;
    // verify: every function node's size is 1 + its children's, children stay inside the row, root spans it
    bool bad = false;
    for (int j = ln.lane; j < len; j += ln.n) {
        const int ar = arity_of(k.t[j], MULTI);
        int c = j + 1, tot = 1;
        for (int a = 0; a < ar && !bad; ++a) {
            if (c >= len) { bad = true; break; }
            const int cs = k.s[c];
            if (cs < 1) bad = true;
            tot += cs;
            c += cs;
        }
        if ((int)k.s[j] != tot || j + tot > len) bad = true;
    }
    if (ln.lane == 0 && (int)k.s[0] != len) bad = true;
    if (!lanes_any(bad)) return true;
    return fix_sizes<MULTI>(ln, len, k);
}

__host__ __device__ inline void emit_nan(Lanes ln, uint2 *out, int Lp) {
    if (ln.lane == 0) {
        out[0] = mk2(C_NAN, 0);
        if (Lp > 1) out[1] = mk2(C_END, 0);
    }
}

// Single-output rows.  Returns the operand-stack height the program needs, or -1 for a malformed row.
template <bool split>
__host__ __device__ inline int lower_tree_single(Lanes ln, const float *val, const int16_t *typ, const int16_t *size,
                                                 int len, int L, int Lp, int V, int depth_budget, uint2 *out,
                                                 LowerScratch k, bool have_sizes, int deep_from, bool fold) {
    // split: operators on leaves only are emitted as LOAD + acc-form instead of the fresh-value forms
    // (UV/UK/VV/VK/KV), so programs use 6 operand forms per operator instead of 9 - the K = 16 replay kernel
    // keeps only those laid out (its bodies are twice as long; the instruction cache is its limit)
    if (!stage_rows(ln, val, typ, size, len, L, k, have_sizes)) {
        emit_nan(ln, out, Lp);
        return -1;
    }
    auto is_func = [&](int j) { return arity_of(k.t[j], false) != 0; };
    auto is_const = [&](int j) { return (k.t[j] & NT_MASK) == NT_CONST; };
    // ---- one pass: verify the sizes (every function node's size is 1 + its children's, children stay inside
    //      the row, the root spans it) and mark the instruction slots each node contributes itself.  Stale or
    //      absent sizes are recomputed once by fix_sizes and the pass repeats. ----
    for (int attempt = 0;; ++attempt) {
        bool bad = false;
        for (int i = ln.lane; i < len; i += ln.n) {
            const int ar = arity_of(k.t[i], false);
            int tot = 1, pos = i + 1;
            bool mine = false;
            for (int a = 0; a < ar; ++a) {
                if (pos >= len) { mine = true; break; }
                const int cs = k.s[pos];
                if (cs < 1) { mine = true; break; }
                tot += cs;
                pos += cs;
            }
            if ((int)k.s[i] != tot || i + tot > len) mine = true;
            bad |= mine;
        }
        if (ln.lane == 0 && (int)k.s[0] != len) bad = true;
        if (!lanes_any(bad)) break;
        if (attempt || !fix_sizes<false>(ln, len, k)) {
            emit_nan(ln, out, Lp);
            return -1;
        }
    }
    // ---- constant folding, one level (decide from the untouched row, then apply): the folded node keeps its size, so
    //      every child position (i + 1, c + size[c]) stays valid; its former children are dead slots nobody visits ----
    if (fold) {
        for (int i = ln.lane; i < len; i += ln.n) {
            const int ar = arity_of(k.t[i], false);
            bool can = false;
            uint32_t fv = 0;
            if (ar == 1 || ar == 2) {
                const int x = i + 1, y = x + k.s[x];
                const unsigned func = f32_to_u32(bits_f32(k.v[i]));
                if (ar == 1 && k.t[x] == NT_CONST) {
                    can = true;
                    fv = f32_bits(fold_unary(unary_slot(func), bits_f32(k.v[x])));
                } else if (ar == 2 && k.t[x] == NT_CONST && k.t[y] == NT_CONST) {
                    can = true;
                    fv = f32_bits(fold_binary(binary_slot(func), bits_f32(k.v[x]), bits_f32(k.v[y])));
                }
            }
            k.q[i] = can ? 1 : 0;
            k.D[i] = fv;
        }
        lanes_sync();
        for (int i = ln.lane; i < len; i += ln.n)
            if (k.q[i]) { k.t[i] = NT_CONST; k.v[i] = k.D[i]; }
        lanes_sync();
    }
    // ---- mark the instruction slots each node contributes itself ----
    for (int i = ln.lane; i < len; i += ln.n) {
        const int ar = arity_of(k.t[i], false);
        int m = 0;
        if (ar == 1) {
            m = (split && !is_func(i + 1)) ? 2 : 1;
        } else if (ar == 2) {   // two constants (or split): LOAD + AK / AV
            const int x = i + 1, y = x + k.s[x];
            m = (!is_func(x) && !is_func(y) && (split || (is_const(x) && is_const(y)))) ? 2 : 1;
        } else if (ar == 3) {   // every leaf child is a LOAD
            const int a = i + 1, b = a + k.s[a], c = b + k.s[b];
            m = 1 + (!is_func(a)) + (!is_func(b)) + (!is_func(c));
        }
        k.M[i] = (uint16_t)m;
        k.q[i] = 0;
        k.D[i] = 0;
    }
    if (ln.lane == 0) k.D[len] = 0;
    if (!is_func(0)) {   // the tree is a single leaf
        if (ln.lane == 0) {
            out[0] = leaf_instr(C_LOAD_V, C_LOAD_K, leaf_of(k.t[0], bits_f32(k.v[0]), V), 0);
            if (Lp > 1) out[1] = mk2(C_END, 0);
        }
        return 0;
    }
    lanes_sync();
    lanes_exclusive_scan(ln, k.M, len);
    auto ni = [&](int j) { return (int)k.M[j + k.s[j]] - (int)k.M[j]; };   // slots of the whole subtree j
    // ---- rank the value-producing children of every function node: larger subtree first
    //      (ties: later child first, the reference's order), so the lighter sibling is the one evaluated
    //      with a value pending and the operand stack stays O(log len) deep (stack_depth_bound).
    //      In the same pass, root -> leaves as one prefix sum.  For a value child j (rank r, `before` slots
    //      emitted ahead of it among its siblings):  start(j) = sum of `before` over the path root..j;
    //      pending(j) = sum of r over that path = values alive (in acc or on the stack) when subtree j begins,
    //      so acc is live iff pending > 0 and the stack height is pending - 1.  A path sum in prefix order is a
    //      prefix sum of "add at j, subtract at j + size[j]" (j's contribution covers exactly its span). ----
    auto place = [&](int ch, uint32_t r, uint32_t before) {
        k.q[ch] = (uint16_t)(r | 0x8000);
        const uint32_t add = before | (r << 16);
        if (add) {
            lanes_atomic_add(&k.D[ch], add);
            lanes_atomic_add(&k.D[ch + k.s[ch]], 0u - add);
        }
    };
    for (int i = ln.lane; i < len; i += ln.n) {
        const int ar = arity_of(k.t[i], false);
        if (ar == 0) continue;
        const int c0 = i + 1;
        if (ar == 1) {
            if (is_func(c0)) k.q[c0] = 0x8000;
        } else if (ar == 2) {
            const int c1 = c0 + k.s[c0];
            const bool f0 = is_func(c0), f1 = is_func(c1);
            if (f0 && f1) {
                const bool x_first = k.s[c0] > k.s[c1];
                const int first = x_first ? c0 : c1, second = x_first ? c1 : c0;
                k.q[first] = 0x8000;
                place(second, 1, (uint32_t)ni(first));
            } else if (f0 || f1) {
                k.q[f0 ? c0 : c1] = 0x8000;
            }
        } else {
            const int c1 = c0 + k.s[c0], c2 = c1 + k.s[c1];
            const int c[3] = {c0, c1, c2};
            int o0 = 2, o1 = 1, o2 = 0;   // c, b, a; stable bubble sort by size, descending
            if (k.s[c[o1]] > k.s[c[o0]]) { const int sw = o0; o0 = o1; o1 = sw; }
            if (k.s[c[o2]] > k.s[c[o1]]) { const int sw = o1; o1 = o2; o2 = sw; }
            if (k.s[c[o1]] > k.s[c[o0]]) { const int sw = o0; o0 = o1; o1 = sw; }
            const int ord[3] = {o0, o1, o2};
            uint32_t before = 0;
            for (int r = 0; r < 3; ++r) {
                const int ch = ord[r] == 0 ? c0 : (ord[r] == 1 ? c1 : c2);
                place(ch, (uint32_t)r, before);
                before += is_func(ch) ? (uint32_t)ni(ch) : 1u;
            }
        }
    }
    lanes_sync();
    lanes_inclusive_scan32(ln, k.D, len);
    // ---- emit ----
    int my_max = 0;
    for (int i = ln.lane; i < len; i += ln.n) {
        const int ar = arity_of(k.t[i], false);
        const bool valued = i == 0 || (k.q[i] & 0x8000);
        if (!valued) continue;                    // folded leaf: an operand of its parent
        const uint32_t d = k.D[i], pending = d >> 16;
        const int st = d & 0xFFFF;
        const uint32_t live = pending ? 1u : 0u, height = pending ? pending - 1 : 0u;
        const uint32_t live_push = live ? ((height + 1) << I_PUSH_SHIFT) : 0;   // a fresh value saves acc into slot `height`
        const bool deep_push = live && (int)height >= deep_from;                // marks LOADs only; fresh forms carry the slot
        const int ld_v = deep_push ? C_LOAD_V_DEEP : C_LOAD_V, ld_k = deep_push ? C_LOAD_K_DEEP : C_LOAD_K;
        const uint32_t height_in = height + live;
        if (live) my_max = my_max > (int)height + 1 ? my_max : (int)height + 1;
        if (ar == 0) {                            // leaf child of a ternary node: produced by a LOAD
            out[st] = leaf_instr(ld_v, ld_k, leaf_of(k.t[i], bits_f32(k.v[i]), V), live_push);
            continue;
        }
        const int own = st + ni(i) - 1;
        const unsigned func = f32_to_u32(bits_f32(k.v[i]));   // forward.cu:108 `(unsigned int)node_value`
        if (ar == 1) {
            const int u = unary_slot(func), c = i + 1;
            if (is_func(c)) out[own] = mk2((uint32_t)opcode(FM_UA, u), 0);
            else if (split) {
                out[st] = leaf_instr(ld_v, ld_k, leaf_of(k.t[c], bits_f32(k.v[c]), V), live_push);
                out[own] = mk2((uint32_t)opcode(FM_UA, u), 0);
            } else out[own] = leaf_instr(opcode(FM_UV, u), opcode(FM_UK, u), leaf_of(k.t[c], bits_f32(k.v[c]), V), live_push);
        } else if (ar == 2) {
            const int b = binary_slot(func), x = i + 1, y = x + k.s[x];
            const bool cx = is_func(x), cy = is_func(y);
            if (!cx && !cy) {
                const Leaf lx = leaf_of(k.t[x], bits_f32(k.v[x]), V), ly = leaf_of(k.t[y], bits_f32(k.v[y]), V);
                if (split || (lx.is_const && ly.is_const)) {   // (two constants:) load the first, then acc (op) second
                    out[st] = leaf_instr(ld_v, ld_k, lx, live_push);
                    out[st + 1] = leaf_instr(opcode(FM_AV, b), opcode(FM_AK, b), ly, 0);
                } else if (!lx.is_const && !ly.is_const) {
                    out[own] = mk2((uint32_t)opcode(FM_VV, b) | live_push | (lx.bits << I_IDXA_SHIFT) | (ly.bits << I_IDXB_SHIFT), 0);
                } else if (!lx.is_const) {
                    out[own] = mk2((uint32_t)opcode(FM_VK, b) | live_push | (lx.bits << I_IDXA_SHIFT), ly.bits);
                } else {
                    out[own] = mk2((uint32_t)opcode(FM_KV, b) | live_push | (ly.bits << I_IDXA_SHIFT), lx.bits);
                }
            } else if (cx && cy) {
                const bool x_first = (k.q[x] & 3) == 0;
                // the first child's value was saved into slot `height_in` by the second child's first instruction
                const bool deep = (int)height_in >= deep_from;
                const int form = x_first ? (deep ? FM_DA : FM_SA) : (deep ? FM_AD : FM_AS);
                out[own] = mk2((uint32_t)opcode(form, b) | (height_in << I_IDXA_SHIFT), 0);
                my_max = my_max > (int)height_in + 1 ? my_max : (int)height_in + 1;
            } else {
                const int lf = cx ? y : x;
                const Leaf l = leaf_of(k.t[lf], bits_f32(k.v[lf]), V);
                out[own] = cx ? leaf_instr(opcode(FM_AV, b), opcode(FM_AK, b), l, 0)
                              : leaf_instr(opcode(FM_VA, b), opcode(FM_KA, b), l, 0);
            }
        } else {
            // IF(a, b, c): the value produced last sits in acc, the one before in slot height_in + 1,
            // the first in slot height_in
            const int a = i + 1, b = a + k.s[a], c = b + k.s[b];
            const uint32_t perm = (2u - (k.q[a] & 3)) | ((2u - (k.q[b] & 3)) << 2) | ((2u - (k.q[c] & 3)) << 4);
            out[own] = mk2((uint32_t)C_IF | (perm << I_IDXA_SHIFT) | (height_in << I_IDXB_SHIFT), 0);
            my_max = my_max > (int)height_in + 2 ? my_max : (int)height_in + 2;
        }
    }
    const int total = k.M[len];
    if (ln.lane == 0 && total < Lp) out[total] = mk2(C_END, 0);
    const int need = lanes_max(my_max);
    if (need > depth_budget) {   // cannot happen for well-formed rows (stack_depth_bound); fail safe
        lanes_sync();
        emit_nan(ln, out, Lp);
        return -1;
    }
    return need;
}

// Multi-output rows (out_len > 1).  The reference's multiOutput branch makes EVERY function
// node hand its right-most child's value to its father (forward.cu:236-242: `top_val =
// right_node` is unconditional), so a subtree's value is simply its right-most leaf — the last
// node of its prefix span — and the only arithmetic with an effect is each OUT node applying its
// function to those leaves and adding the result to outs[outIndex].  The program is therefore a
// flat list of leaf-operand instructions, one group per OUT node, in the reference's processing
// order (last node first) so the float sums into outs[] associate identically.  No operand stack.
__host__ __device__ inline int lower_tree_multi(Lanes ln, const float *val, const int16_t *typ, const int16_t *size,
                                                int len, int L, int Lp, int V, int O, uint2 *out, LowerScratch k,
                                                bool have_sizes) {
    if (!lower_stage<true>(ln, val, typ, size, len, L, k, have_sizes)) {
        emit_nan(ln, out, Lp);
        return -1;
    }
    for (int i = ln.lane; i < len; i += ln.n) {
        const int t = k.t[i], ar = arity_of(t, true);
        k.M[i] = (uint16_t)((ar != 0 && (t & NT_OUT)) ? (ar == 1 ? 1 : 2) : 0);
    }
    lanes_sync();
    lanes_exclusive_scan(ln, k.M, len);
    const int total = k.M[len];
    if (total + 1 > Lp) {   // cannot happen: an OUT node's slots never exceed its own node count
        emit_nan(ln, out, Lp);
        return -1;
    }
    for (int i = ln.lane; i < len; i += ln.n) {
        const int t = k.t[i], ar = arity_of(t, true);
        if (ar == 0 || !(t & NT_OUT)) continue;
        const int mine = ar == 1 ? 1 : 2;
        int slot = total - (int)k.M[i] - mine;             // nodes are emitted last-first
        int last[3] = {0, 0, 0}, c = i + 1;
        for (int a = 0; a < ar; ++a) {
            c += k.s[c];
            last[a] = c - 1;                               // right-most leaf of child a
        }
        const uint32_t bits = k.v[i];                      // kernel.h:105-113
        const unsigned func = (unsigned)(int)(int16_t)(bits & 0xFFFF);
        const unsigned oi = (unsigned)(int)(int16_t)(bits >> 16);
        const uint32_t outbits = I_OUT | ((oi < (unsigned)O ? oi : I_IDXB_MASK) << I_IDXB_SHIFT);
        const Leaf la = leaf_of(k.t[last[0]], bits_f32(k.v[last[0]]), V);
        if (ar == 1) {
            const int u = unary_slot(func);
            out[slot] = leaf_instr(opcode(FM_UV, u), opcode(FM_UK, u), la, outbits);
        } else if (ar == 2) {
            const int b = binary_slot(func);
            out[slot] = leaf_instr(C_LOAD_V, C_LOAD_K, la, 0);
            out[slot + 1] = leaf_instr(opcode(FM_AV, b), opcode(FM_AK, b), leaf_of(k.t[last[1]], bits_f32(k.v[last[1]]), V), outbits);
        } else {   // C_IF3, two slots: {hdr, a} {b, c}; b/c words hold constant bits or a variable index
            const Leaf lb = leaf_of(k.t[last[1]], bits_f32(k.v[last[1]]), V), lc = leaf_of(k.t[last[2]], bits_f32(k.v[last[2]]), V);
            const uint32_t hdr = (uint32_t)C_IF3 | outbits | (la.is_const ? I_IF3_ACONST : (la.bits << I_IDXA_SHIFT)) |
                                 (lb.is_const ? I_IF3_BCONST : 0u) | (lc.is_const ? I_IF3_CCONST : 0u);
            out[slot] = mk2(hdr, la.is_const ? la.bits : 0u);
            out[slot + 1] = mk2(lb.bits, lc.bits);
        }
    }
    if (ln.lane == 0 && total < Lp) out[total] = mk2(C_END, 0);
    return 0;
}

template <bool MULTI, bool SPLIT = false>
__host__ __device__ inline int lower_tree(Lanes ln, const float *val, const int16_t *typ, const int16_t *size, int len,
                                          int L, int Lp, int V, int O, int depth_budget, uint2 *out, LowerScratch k,
                                          bool have_sizes = true, int deep_from = kNoDeepSlots, bool fold = true) {
    if (MULTI) return lower_tree_multi(ln, val, typ, size, len, L, Lp, V, O, out, k, have_sizes);
    return lower_tree_single<SPLIT>(ln, val, typ, size, len, L, Lp, V, depth_budget, out, k, have_sizes,
                                    deep_from, fold);
}

#ifdef __CUDACC__
struct LowerArgs {
    const float *value;
    const int16_t *type;
    const int16_t *size;      // packed subtree_size rows [P][L], or (rows_have_sizes == 0) one length per tree
    const unsigned *offsets;  // non-null: value / type hold valid prefixes back to back, tree n at [offsets[n], offsets[n + 1])
    uint2 *prog;              // [P][Lp]
    unsigned *sched;          // 64 scheduler words, zeroed here (the replay kernel runs after this one)
    int P, L, Lp, V, O, depth_budget, rows_have_sizes;
    int deep_from;            // SPLIT: slots >= deep_from are addressed with the deep opcodes (program.cuh); else kNoDeepSlots
    int fold;                 // single-output: fold functions of constant leaves into constants
    unsigned *chunk_done;     // multi-GPU push counters of the replay kernel (eval.cu), zeroed here
    int nchunks;
};
__device__ __forceinline__ void lower_zero_scheduler_words(const LowerArgs &g) {
    if (blockIdx.x != 0) return;
    if (threadIdx.x < 64) g.sched[threadIdx.x] = 0;                          // ticket counters of the replay kernel
    for (int i = threadIdx.x; i < g.nchunks; i += blockDim.x) g.chunk_done[i] = 0;
}

// one warp per tree, grid-stride over the population
// SPLIT (single-output only): LOAD + acc-form for operators on leaves (see lower_tree_single)
template <bool MULTI, bool SPLIT = false>
__global__ void __launch_bounds__(256, 6) lower_kernel(LowerArgs g) {
    extern __shared__ __align__(16) unsigned char lower_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const size_t per_warp = (lower_scratch_bytes(g.L) + 15) & ~(size_t)15;
    const LowerScratch k = carve_scratch(lower_smem + warp * per_warp, g.L);
    lower_zero_scheduler_words(g);
    const Lanes ln{lane, 32};
    for (int n = blockIdx.x * nwarp + warp; n < g.P; n += gridDim.x * nwarp) {
        // rows_have_sizes == 0 (host path uploads one length per tree): sizes are recomputed from the arities
        const int16_t *srow = g.rows_have_sizes ? g.size + (size_t)n * g.L : nullptr;
        size_t at = (size_t)n * g.L;
        int len;
        if (g.offsets) {          // the host path's compact upload (host_api.cu): no size rows, no row padding
            const unsigned o0 = __ldg(g.offsets + n);
            at = o0;
            len = (int)(__ldg(g.offsets + n + 1) - o0);
        } else {
            len = g.rows_have_sizes ? (int)__ldg(srow) : (int)__ldg(g.size + n);
        }
        lower_tree<MULTI, SPLIT>(ln, g.value + at, g.type + at, srow, len, g.L, g.Lp, g.V, g.O,
                          g.depth_budget, g.prog + (size_t)n * g.Lp, k, g.rows_have_sizes != 0, g.deep_from, g.fold != 0);
        __syncwarp();
    }
}
#endif

}  // namespace evogp
