// lower_fast.cuh — the lowering pass for the common case, in registers.
//
// lower.cuh is the specification of the lowering pass: any row width, any arity, stale sizes, multi-output, and it
// compiles for the host (tests/host_lower_harness.cu).  It costs ~700 warp-instructions per tree - a third of an
// evaluation step at BASELINE configs[1] (profiles/r1_final_ncu.txt) - because every phase loops over the row
// through shared memory and the emitter branches per node kind.
//
// This file lowers the rows that matter for throughput with a fraction of that: single-output trees of unary and
// binary functions, max_tree_len <= 64, sizes consistent.  One warp per tree, lane i owns nodes i and i + 32 in
// REGISTERS; the two prefix sums (instruction slots; path sums of slot offsets and stack heights) are shuffle scans;
// shared memory only serves the data-dependent gathers (a child's type / size / value) and the scatter of the path-sum
// deltas; the emitter is branch-free (every function node computes its one instruction word with selects).  The
// programs are BIT-IDENTICAL to lower_tree_single's (tests/test_gpu_lowering.py compares them word for word), so
// everything proved about those programs (host harness, parity tests) carries over.  Any row outside the fast class
// (ternary node, inconsistent sizes, bad length) returns kLowerFallback and the caller runs the generic pass on it.
#pragma once
#include "lower.cuh"

namespace evogp {

constexpr int kLowerFallback = -2;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t x, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    return x;
}

// per-warp scratch: 4 arrays of NSETS * 32 + 1 words
template <int NSETS>
__host__ __device__ constexpr size_t lower_fast_scratch_bytes() { return (size_t)4 * (NSETS * 32 + 1) * 4; }

// Returns the operand-stack height the program needs (>= 0), -1 for a program that was replaced by C_NAN, or
// kLowerFallback when the row is outside the fast class (nothing was written).
template <int NSETS>
__device__ __forceinline__ int lower_fast_tree(const int lane, const float *__restrict__ val, const int16_t *__restrict__ typ,
                                               const int16_t *__restrict__ size, const int L, const int Lp, const int V,
                                               const int depth_budget, const int deep_from, const bool fold,
                                               uint2 *__restrict__ out, uint32_t *__restrict__ sm) {
    constexpr int CAP = NSETS * 32 + 1;
    uint32_t *TS = sm, *VB = sm + CAP, *M = sm + 2 * CAP, *D = sm + 3 * CAP;
    // ---- loads: issued before the length is known (one memory latency per tree) ----
    int t[NSETS], s[NSETS];
    uint32_t v[NSETS];
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        const int i = lane + 32 * k;
        const bool in = i < L;
        t[k] = in ? (int)(uint16_t)__ldg(typ + i) : 0;
        s[k] = in ? (int)(uint16_t)__ldg(size + i) : 0;
        v[k] = in ? __float_as_uint(__ldg(val + i)) : 0u;
    }
    const int len = __shfl_sync(0xffffffffu, s[0], 0);
    if (len < 1 || len > L) return kLowerFallback;
    const int nk = (NSETS > 1 && len > 32) ? NSETS : 1;       // warp-uniform: short trees skip the second node set
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        const int i = lane + 32 * k;
        if (i >= len) { t[k] = 0; s[k] = 0; v[k] = 0u; }
        TS[i] = (uint32_t)t[k] | ((uint32_t)s[k] << 16);
        VB[i] = v[k];
        D[i] = 0u;
    }
    if (lane == 0) { TS[CAP - 1] = 0u; D[CAP - 1] = 0u; }
    __syncwarp();

    // ---- children and the size check ----
    int ar[NSETS], c0s[NSETS], c1s[NSETS];
    uint32_t ts0[NSETS], ts1[NSETS], m[NSETS];
    bool bad = false;
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        ar[k] = 0; m[k] = 0u; ts0[k] = 0u; ts1[k] = 0u; c0s[k] = 0; c1s[k] = 0;
        if (k < nk) {
            const int i = lane + 32 * k;
            const int tt = t[k];
            const int a = (i < len) ? ((tt <= 1) ? 0 : (tt == 2 ? 1 : (tt == 3 ? 2 : 3))) : 0;   // arity_of(t, false): type NOT masked
            const int c0 = i + 1;
            const uint32_t w0 = TS[c0 < CAP ? c0 : CAP - 1];
            const int s0 = (int)(w0 >> 16);
            const int c1 = c0 + s0;
            const uint32_t w1 = TS[c1 < CAP ? c1 : CAP - 1];
            const int s1 = (int)(w1 >> 16);
            const int tot = 1 + (a >= 1 ? s0 : 0) + (a >= 2 ? s1 : 0);
            bad |= a == 3;
            bad |= a >= 1 && (c0 >= len || s0 < 1);
            bad |= a >= 2 && (c1 >= len || s1 < 1);
            bad |= i < len && (s[k] != tot || i + tot > len);
            ar[k] = a; c0s[k] = c0; c1s[k] = c1; ts0[k] = w0; ts1[k] = w1;
        }
    }
    if (__any_sync(0xffffffffu, bad)) return kLowerFallback;
    // ---- constant folding, one level (lower.cuh): a function of constant leaves becomes a constant leaf that keeps its
    //      size; decisions come from the children gathered above, i.e. from the untouched row ----
    if (fold) {
        bool any = false;
#pragma unroll
        for (int k = 0; k < NSETS; ++k) {
            if (k < nk) {
                const uint32_t t0 = ts0[k] & 0xFFFFu, t1 = ts1[k] & 0xFFFFu;
                const bool can = (ar[k] == 1 && t0 == 1u) || (ar[k] == 2 && t0 == 1u && t1 == 1u);
                if (can) {
                    const unsigned func = __float2uint_rz(__uint_as_float(v[k]));
                    const float x = __uint_as_float(VB[c0s[k]]);
                    const float r = ar[k] == 1 ? fold_unary(unary_slot(func), x)
                                               : fold_binary(binary_slot(func), x, __uint_as_float(VB[c1s[k]]));
                    t[k] = NT_CONST; v[k] = __float_as_uint(r); ar[k] = 0;
                    TS[lane + 32 * k] = (uint32_t)NT_CONST | ((uint32_t)s[k] << 16);
                    VB[lane + 32 * k] = v[k];
                }
                any |= can;
            }
        }
        if (__any_sync(0xffffffffu, any)) {   // fathers look at their children again
            __syncwarp();
#pragma unroll
            for (int k = 0; k < NSETS; ++k) {
                if (k < nk) {
                    ts0[k] = TS[c0s[k] < CAP ? c0s[k] : CAP - 1];
                    ts1[k] = TS[c1s[k] < CAP ? c1s[k] : CAP - 1];
                }
            }
        }
    }
    // ---- instruction slots each node contributes itself ----
#pragma unroll
    for (int k = 0; k < NSETS; ++k)
        if (k < nk) m[k] = ar[k] == 0 ? 0u : ((ar[k] == 2 && (ts0[k] & 0xFFFFu) == 1u && (ts1[k] & 0xFFFFu) == 1u) ? 2u : 1u);
    const int ar_root = __shfl_sync(0xffffffffu, ar[0], 0);
    if (ar_root == 0) {   // the tree is a single leaf
        if (lane == 0) {
            out[0] = leaf_instr(C_LOAD_V, C_LOAD_K, leaf_of(t[0], __uint_as_float(v[0]), V), 0);
            if (Lp > 1) out[1] = mk2(C_END, 0);
        }
        return 0;
    }
    // ---- exclusive prefix sum of the slots ----
    uint32_t mx[NSETS];
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        mx[k] = carry;
        if (k < nk) {
            const uint32_t incl = warp_incl_scan(m[k], lane);
            mx[k] = carry + incl - m[k];
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        M[lane + 32 * k] = mx[k];
    }
    const uint32_t total = carry;
    if (lane == 0) M[CAP - 1] = total;
    __syncwarp();
    // ---- sibling order: of two function children the larger subtree goes first (ties: the right one, the reference's
    //      order); the second one starts `slots of the first` later with one more value pending.  Path sums root -> node
    //      in prefix order = prefix sum of "add at j, subtract at j + size[j]". ----
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        if (k < nk && ar[k] == 2) {
            const uint32_t t0 = ts0[k] & 0xFFFFu, t1 = ts1[k] & 0xFFFFu;
            if (t0 > 1u && t1 > 1u) {
                const int s0 = (int)(ts0[k] >> 16), s1 = (int)(ts1[k] >> 16);
                const bool x_first = s0 > s1;
                const int first = x_first ? c0s[k] : c1s[k], second = x_first ? c1s[k] : c0s[k];
                const int s_first = x_first ? s0 : s1, s_second = x_first ? s1 : s0;
                const uint32_t add = (M[first + s_first] - M[first]) | (1u << 16);
                atomicAdd(&D[second], add);
                atomicAdd(&D[second + s_second], 0u - add);
            }
        }
    }
    __syncwarp();
    uint32_t d[NSETS];
    carry = 0;
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        d[k] = 0u;
        if (k < nk) {
            const uint32_t incl = warp_incl_scan(D[lane + 32 * k], lane);
            d[k] = carry + incl;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    // ---- emit: one instruction per function node (two for `const op const`), branch-free ----
    int my_max = 0;
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        if (k < nk && ar[k] != 0) {
            const int i = lane + 32 * k;
            const uint32_t pending = d[k] >> 16;
            const int st = (int)(d[k] & 0xFFFFu);
            const bool live = pending != 0u;
            const uint32_t height = live ? pending - 1u : 0u;
            const uint32_t live_push = live ? ((height + 1u) << I_PUSH_SHIFT) : 0u;   // a fresh value saves acc into slot `height`
            const bool deep_push = live && (int)height >= deep_from;                  // marks LOADs only; fresh forms carry the slot
            const uint32_t height_in = pending;
            if (live) my_max = max(my_max, (int)height + 1);
            const int own = st + (int)(M[i + s[k]] - mx[k]) - 1;
            const unsigned func = __float2uint_rz(__uint_as_float(v[k]));             // forward.cu:108 `(unsigned int)node_value`
            const uint32_t t0 = ts0[k] & 0xFFFFu, t1 = ts1[k] & 0xFFFFu;
            const bool f0 = t0 > 1u, f1 = t1 > 1u, k0 = t0 == 1u, k1 = t1 == 1u;
            // leaf operands (meaningful only where the child is a leaf): constant bits, or the clamped variable index
            const uint32_t v0 = VB[c0s[k]];
            int x0 = __float2int_rz(__uint_as_float(v0));
            x0 = x0 < 0 ? 0 : (x0 >= V ? V - 1 : x0);
            const uint32_t lx = k0 ? v0 : (uint32_t)x0;
            uint32_t w, c = 0u;
            if (ar[k] == 1) {
                const uint32_t u = (uint32_t)unary_slot(func);
                const int form = f0 ? FM_UA : (k0 ? FM_UK : FM_UV);
                w = (uint32_t)(form * 16) + u;
                if (!f0) {
                    w |= live_push;
                    if (k0) c = lx; else w |= lx << I_IDXA_SHIFT;
                }
            } else {
                const uint32_t b = (uint32_t)binary_slot(func);
                const uint32_t v1 = VB[c1s[k]];
                int x1 = __float2int_rz(__uint_as_float(v1));
                x1 = x1 < 0 ? 0 : (x1 >= V ? V - 1 : x1);
                const uint32_t ly = k1 ? v1 : (uint32_t)x1;
                if (f0 && f1) {
                    // the first child's value was saved into slot `height_in` by the second child's first instruction
                    const bool x_first = (ts0[k] >> 16) > (ts1[k] >> 16);
                    const int form = (x_first ? FM_SA : FM_AS) + (((int)height_in >= deep_from) ? 2 : 0);   // FM_DA = FM_SA + 2, FM_AD = FM_AS + 2
                    w = (uint32_t)(form * 16) + b + (height_in << I_IDXA_SHIFT);
                    my_max = max(my_max, (int)height_in + 1);
                } else if (f0) {            // acc (op) leaf y
                    w = (uint32_t)((k1 ? FM_AK : FM_AV) * 16) + b;
                    if (k1) c = ly; else w |= ly << I_IDXA_SHIFT;
                } else if (f1) {            // leaf x (op) acc
                    w = (uint32_t)((k0 ? FM_KA : FM_VA) * 16) + b;
                    if (k0) c = lx; else w |= lx << I_IDXA_SHIFT;
                } else if (k0 && k1) {      // load the first constant, then acc (op) second
                    out[st] = mk2((uint32_t)(deep_push ? C_LOAD_K_DEEP : C_LOAD_K) | live_push, lx);
                    w = (uint32_t)(FM_AK * 16) + b;
                    c = ly;
                } else if (!k0 && !k1) {
                    w = (uint32_t)(FM_VV * 16) + b + live_push + (lx << I_IDXA_SHIFT) + (ly << I_IDXB_SHIFT);
                } else if (!k0) {
                    w = (uint32_t)(FM_VK * 16) + b + live_push + (lx << I_IDXA_SHIFT);
                    c = ly;
                } else {
                    w = (uint32_t)(FM_KV * 16) + b + live_push + (ly << I_IDXA_SHIFT);
                    c = lx;
                }
            }
            out[own] = mk2(w, c);
        }
    }
    if (lane == 0 && (int)total < Lp) out[total] = mk2(C_END, 0);
    const int need = (int)__reduce_max_sync(0xffffffffu, (unsigned)my_max);
    if (need > depth_budget) {   // cannot happen for well-formed rows (stack_depth_bound); fail safe
        __syncwarp();
        if (lane == 0) {
            out[0] = mk2(C_NAN, 0);
            if (Lp > 1) out[1] = mk2(C_END, 0);
        }
        return -1;
    }
    return need;
}

static_assert(FM_DA == FM_SA + 2 && FM_AD == FM_AS + 2, "deep forms follow their shallow forms by two");

// the generic pass, out of line: rows outside the fast class are rare, their code stays out of the hot loop
__device__ __noinline__ void lower_generic_row(int lane, const float *val, const int16_t *typ, const int16_t *srow, int len, int L,
                                               int Lp, int V, int O, int depth_budget, uint2 *out, void *scratch, int deep_from, bool fold) {
    const LowerScratch k = carve_scratch(scratch, L);
    lower_tree<false, false>(Lanes{lane, 32}, val, typ, srow, len, L, Lp, V, O, depth_budget, out, k, true, deep_from, fold);
}

template <int NSETS>
__host__ __device__ inline size_t lower_fast_per_warp(int L) {
    const size_t a = lower_fast_scratch_bytes<NSETS>(), b = lower_scratch_bytes(L);
    return ((a > b ? a : b) + 15) & ~(size_t)15;
}

// one warp per tree, grid-stride over the population.  Requires packed subtree_size rows (rows_have_sizes).
template <int NSETS>
__global__ void __launch_bounds__(256) lower_fast_kernel(LowerArgs g) {
    extern __shared__ __align__(16) unsigned char lower_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    unsigned char *scratch = lower_smem + warp * lower_fast_per_warp<NSETS>(g.L);
    if (blockIdx.x == 0 && threadIdx.x < 64) g.sched[threadIdx.x] = 0;     // ticket counters of the replay kernel
    for (int n = blockIdx.x * nwarp + warp; n < g.P; n += gridDim.x * nwarp) {
        const float *val = g.value + (size_t)n * g.L;
        const int16_t *typ = g.type + (size_t)n * g.L, *srow = g.size + (size_t)n * g.L;
        uint2 *out = g.prog + (size_t)n * g.Lp;
        const int rc = lower_fast_tree<NSETS>(lane, val, typ, srow, g.L, g.Lp, g.V, g.depth_budget, g.deep_from, g.fold != 0, out,
                                              reinterpret_cast<uint32_t *>(scratch));
        if (rc == kLowerFallback) {
            __syncwarp();
            lower_generic_row(lane, val, typ, srow, (int)__ldg(srow), g.L, g.Lp, g.V, g.O, g.depth_budget, out, scratch, g.deep_from, g.fold != 0);
        }
        __syncwarp();
    }
}

}  // namespace evogp
