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

// folding anything but + - * / goes out of line: the operator bodies (sinh, pow, ...) are hundreds of instructions that
// would otherwise sit in the middle of the hot path's instruction stream
__device__ __noinline__ float fold_rare(unsigned arity, unsigned fid, float x, float y) {
    return arity == 1u ? fold_unary(unary_slot(fid), x) : fold_binary(binary_slot(fid), x, y);
}

// per-warp scratch: 4 arrays of NSETS * 32 + 1 words
template <int NSETS>
__host__ __device__ constexpr size_t lower_fast_scratch_bytes() { return (size_t)4 * (NSETS * 32 + 1) * 4; }

// Selector tables of the branch-free emitter, indexed by  idx = f0 << 3 | f1 << 2 | k0 << 1 | k1  (child 0 / child 1 is a
// function / a constant leaf).  Binary nodes:
//   idx  0 VV  1 VK  2 KV  3 KK (LOAD_K + AK)   4,5 VA  6,7 KA   8,10 AV  9,11 AK   12..15 SA / AS (by sibling order)
constexpr uint64_t kBinForm = 0x8ull | (0x9ull << 4) | (0xAull << 8) | (0x5ull << 12) | (0x6ull << 16) | (0x6ull << 20) |
                              (0x7ull << 24) | (0x7ull << 28) | (0x4ull << 32) | (0x5ull << 36) | (0x4ull << 40) | (0x5ull << 44) |
                              (0xCull << 48) | (0xCull << 52) | (0xCull << 56) | (0xCull << 60);
// what goes into the idxA field (0 nothing, 1 leaf x, 2 leaf y, 3 the stack slot) and into the constant word (0, 1 x, 2 y)
constexpr uint32_t kBinASel = 1u | (1u << 2) | (2u << 4) | (0u << 6) | (1u << 8) | (1u << 10) | (0u << 12) | (0u << 14) |
                              (2u << 16) | (0u << 18) | (2u << 20) | (0u << 22) | (3u << 24) | (3u << 26) | (3u << 28) | (3u << 30);
constexpr uint32_t kBinCSel = 0u | (2u << 2) | (1u << 4) | (2u << 6) | (0u << 8) | (0u << 10) | (1u << 12) | (1u << 14) |
                              (0u << 16) | (2u << 18) | (0u << 20) | (2u << 22);
static_assert(FM_VV == 8 && FM_VK == 9 && FM_KV == 10 && FM_AK == 5 && FM_AV == 4 && FM_VA == 6 && FM_KA == 7 && FM_SA == 11 &&
              FM_AS == 12 && FM_UA == 1 && FM_UV == 2 && FM_UK == 3, "emitter tables follow program.cuh's form numbers");

// Returns the operand-stack height the program needs (>= 0), -1 for a program that was replaced by C_NAN, or
// kLowerFallback when the row is outside the fast class (nothing was written).
template <int NSETS>
__device__ __forceinline__ int lower_fast_tree(const int lane, const float *val, const int16_t *typ, const int16_t *size,
                                               const int L, const int Lp, const int V, const int depth_budget,
                                               const int deep_from, const bool fold, uint2 *out, uint32_t *sm_) {
    constexpr int CAP = NSETS * 32 + 1;
    constexpr uint32_t POISON = 0x4000u;   // size of a slot beyond the row: any father reaching it fails its size check
    // volatile: other lanes write these words between this lane's accesses (a restrict-qualified or plain pointer lets
    // the compiler reuse a child's type word it loaded before the folding pass rewrote it)
    volatile uint32_t *sm = sm_;
    volatile uint32_t *TS = sm, *VB = sm + CAP, *M = sm + 2 * CAP, *D = sm + 3 * CAP;
    // ---- loads: issued before the length is known (one memory latency per tree) ----
    uint32_t t[NSETS], s[NSETS], v[NSETS];
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        const int i = lane + 32 * k;
        const bool in = i < L;
        t[k] = in ? (uint32_t)(uint16_t)__ldg(typ + i) : 0u;
        s[k] = in ? (uint32_t)(uint16_t)__ldg(size + i) : 0u;
        v[k] = in ? __float_as_uint(__ldg(val + i)) : 0u;
    }
    const int len = (int)__shfl_sync(0xffffffffu, s[0], 0);
    if (len < 1 || len > L) return kLowerFallback;
    const int nk = (NSETS > 1 && len > 32) ? NSETS : 1;       // warp-uniform: short trees skip the second node set
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        const int i = lane + 32 * k;
        const bool valid = i < len;
        if (!valid) { t[k] = 0u; s[k] = 1u; v[k] = 0u; }      // a lone leaf as far as this lane's own checks go
        TS[i] = valid ? (t[k] | (s[k] << 16)) : (POISON << 16);
        VB[i] = v[k];
        D[i] = 0u;
    }
    if (lane == 0) { TS[CAP - 1] = POISON << 16; D[CAP - 1] = 0u; }
    __syncwarp();

    // ---- children and the size check: size[i] == 1 + sizes of the children (with size[0] == len that also keeps
    //      every child inside the row); types beyond BFUNC leave the fast class ----
    uint32_t ar[NSETS], w0[NSETS], w1[NSETS], c1s[NSETS];
    bool bad = false;
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        ar[k] = 0u; w0[k] = 0u; w1[k] = 0u; c1s[k] = 0u;
        if (k < nk) {
            const int i = lane + 32 * k;
            const uint32_t a = t[k] > 1u ? t[k] - 1u : 0u;      // arity_of(t, false): the type is NOT masked in single-output mode
            const uint32_t x0 = TS[i + 1];
            const uint32_t c1 = min((uint32_t)(i + 1) + (x0 >> 16), (uint32_t)(CAP - 1));
            const uint32_t x1 = TS[c1];
            const uint32_t tot = 1u + (a >= 1u ? (x0 >> 16) : 0u) + (a >= 2u ? (x1 >> 16) : 0u);
            bad |= (a > 2u) | (s[k] != tot);
            ar[k] = a; w0[k] = x0; w1[k] = x1; c1s[k] = c1;
        }
    }
    if (__any_sync(0xffffffffu, bad)) return kLowerFallback;
    // ---- constant folding, one level (lower.cuh): a function of constant leaves becomes a constant leaf that keeps its
    //      size; decisions come from the children gathered above, i.e. from the untouched row ----
    unsigned fid[NSETS];
#pragma unroll
    for (int k = 0; k < NSETS; ++k) fid[k] = __float2uint_rz(__uint_as_float(v[k]));   // forward.cu:108 `(unsigned int)node_value`
    if (fold) {
        __syncwarp();          // every lane has read its children's words: the folding stores below may overwrite them
        bool any = false;
#pragma unroll
        for (int k = 0; k < NSETS; ++k) {
            if (k < nk) {
                const bool k0 = (w0[k] & 0xFFFFu) == 1u, k1 = (w1[k] & 0xFFFFu) == 1u;
                const bool can = (ar[k] == 2u && k0 && k1) || (ar[k] == 1u && k0);
                if (can) {
                    const float x = __uint_as_float(VB[lane + 32 * k + 1]);
                    const float y = ar[k] == 2u ? __uint_as_float(VB[c1s[k]]) : 0.0f;   // a unary node has no second child to read
                    const unsigned b = fid[k] - (unsigned)F_ADD;
                    float r;
                    if (ar[k] == 2u && b < 4u) {   // + - * / without a branch: the interpreter's own operator bodies, selected
                        const float r0 = binary_op<0>(x, y), r1 = binary_op<1>(x, y), r2 = binary_op<2>(x, y), r3 = binary_op<3>(x, y);
                        r = b == 0u ? r0 : (b == 1u ? r1 : (b == 2u ? r2 : r3));
                    } else {
                        r = fold_rare(ar[k], fid[k], x, y);
                    }
                    t[k] = NT_CONST; v[k] = __float_as_uint(r); ar[k] = 0u;
                    TS[lane + 32 * k] = (uint32_t)NT_CONST | (s[k] << 16);
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
                    w0[k] = TS[lane + 32 * k + 1];
                    w1[k] = TS[c1s[k]];
                }
            }
        }
    }
    const uint32_t ar_root = __shfl_sync(0xffffffffu, ar[0], 0);
    if (ar_root == 0u) {   // the tree is a single leaf
        if (lane == 0) {
            out[0] = leaf_instr(C_LOAD_V, C_LOAD_K, leaf_of((int)t[0], __uint_as_float(v[0]), V), 0);
            if (Lp > 1) out[1] = mk2(C_END, 0);
        }
        return 0;
    }
    // ---- instruction slots each node contributes itself (two for `const op const`), exclusive prefix sum ----
    uint32_t idx[NSETS], m[NSETS], mx[NSETS];
    uint32_t carry = 0;
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        idx[k] = 0u; m[k] = 0u; mx[k] = carry;
        if (k < nk) {
            const uint32_t t0 = w0[k] & 0xFFFFu, t1 = w1[k] & 0xFFFFu;
            idx[k] = (t0 > 1u ? 8u : 0u) | (t1 > 1u ? 4u : 0u) | (t0 == 1u ? 2u : 0u) | (t1 == 1u ? 1u : 0u);
            if (ar[k] == 1u) idx[k] &= 10u;                      // a unary node has no second child (w1 is whatever follows its subtree)
            m[k] = (ar[k] != 0u ? 1u : 0u) + ((ar[k] == 2u && idx[k] == 3u) ? 1u : 0u);
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
    uint32_t mend[NSETS];
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        mend[k] = 0u;
        if (k < nk) {
            const int i = lane + 32 * k;
            const uint32_t e = (uint32_t)i + s[k];                 // one past this subtree
            mend[k] = M[e];
            if (ar[k] == 2u && idx[k] >= 12u) {
                const bool x_first = (w0[k] >> 16) > (w1[k] >> 16);
                const uint32_t mc0 = mx[k] + m[k], mc1 = M[c1s[k]];  // slots before child 0 / before child 1
                const uint32_t add = (x_first ? mc1 - mc0 : mend[k] - mc1) | (1u << 16);
                atomicAdd(const_cast<uint32_t *>(&D[x_first ? c1s[k] : (uint32_t)(i + 1)]), add);
                atomicAdd(const_cast<uint32_t *>(&D[x_first ? e : c1s[k]]), 0u - add);
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
    // ---- emit: one instruction per function node (two for `const op const`); form and operand placement come from
    //      the selector tables, so lanes of every node kind run the same instructions ----
    uint32_t my_max = 0;
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
        if (k < nk && ar[k] != 0u) {
            const int i = lane + 32 * k;
            const uint32_t pending = d[k] >> 16, st = d[k] & 0xFFFFu;       // values alive when this subtree begins; its first slot
            const uint32_t live_push = pending << I_PUSH_SHIFT;             // a fresh value saves acc into slot pending - 1 (field = slot + 1)
            const uint32_t own = st + (mend[k] - mx[k]) - 1u;
            // leaf operands (meaningful only where the child is a leaf): constant bits, or the clamped variable index
            const uint32_t v0 = VB[i + 1], v1 = VB[c1s[k]];
            const int xi0 = min(max(__float2int_rz(__uint_as_float(v0)), 0), V - 1), xi1 = min(max(__float2int_rz(__uint_as_float(v1)), 0), V - 1);
            const uint32_t lx = (idx[k] & 2u) ? v0 : (uint32_t)xi0, ly = (idx[k] & 1u) ? v1 : (uint32_t)xi1;
            const bool unary = ar[k] == 1u;
            uint32_t form, asel, csel, op;
            if (unary) {      // UA / UV / UK by the child's kind
                const uint32_t u = fid[k] - (unsigned)F_SIN;
                op = u < 15u ? u : (uint32_t)U_ZERO;
                form = (idx[k] & 8u) ? (uint32_t)FM_UA : ((idx[k] & 2u) ? (uint32_t)FM_UK : (uint32_t)FM_UV);
                asel = (idx[k] & 10u) ? 0u : 1u;
                csel = (idx[k] & 2u) ? 1u : 0u;
            } else {
                const uint32_t b = fid[k] - (unsigned)F_ADD;
                op = b < 13u ? b : (uint32_t)B_ZERO;
                form = (uint32_t)(kBinForm >> (4u * idx[k])) & 15u;
                asel = (kBinASel >> (2u * idx[k])) & 3u;
                csel = (kBinCSel >> (2u * idx[k])) & 3u;
            }
            const bool ff = !unary && idx[k] >= 12u;
            if (ff) {
                // the first child's value was saved into slot `pending` by the second child's first instruction
                form -= ((w0[k] >> 16) > (w1[k] >> 16)) ? 1u : 0u;          // x first: SA, else AS
                form += ((int)pending >= deep_from) ? 2u : 0u;              // FM_DA = FM_SA + 2, FM_AD = FM_AS + 2
                my_max = max(my_max, pending + 1u);
            }
            my_max = max(my_max, pending);
            const bool fresh = (idx[k] & 12u) == 0u;                        // no function child: the instruction starts a value
            const uint32_t afield = asel == 1u ? lx : (asel == 2u ? ly : (asel == 3u ? pending : 0u));
            const uint32_t c = csel == 1u ? lx : (csel == 2u ? ly : 0u);
            uint32_t w = form * 16u + op + (afield << I_IDXA_SHIFT);
            if (!unary && idx[k] == 0u) w += ly << I_IDXB_SHIFT;
            if (!unary && idx[k] == 3u)       // const (op) const survives only without folding: load the first, then acc (op) second
                out[st] = mk2((uint32_t)(((int)pending > deep_from) ? C_LOAD_K_DEEP : C_LOAD_K) | live_push, lx);
            else if (fresh) w += live_push;
            out[own] = mk2(w, c);
        }
    }
    if (lane == 0 && (int)total < Lp) out[total] = mk2(C_END, 0);
    const int need = (int)__reduce_max_sync(0xffffffffu, my_max);
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
    lower_zero_scheduler_words(g);
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
