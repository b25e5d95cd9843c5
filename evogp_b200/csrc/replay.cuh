// replay.cuh — the tree-evaluation kernel (replay_kernel) and its launcher; see eval.cu for the design notes.
// Included by one translation unit per kernel flavour (FEAT): eval.cu (plain), eval_exchange.cu, eval_acc.cu.
#pragma once
#include <cstdlib>
#include "lower.cuh"
#include "fastpath_k8.inc"
#include "fastpath_k8_tmem.inc"
#include "fastpath_k16_tmem.inc"
#include "fastpath_k8_multi.inc"

namespace evogp {

// MODE_ACC: classification accuracy (problem/classification.py:54-67) - per datapoint the predicted class (arg-max of the
// outputs, or the rounded single output) is compared with the label inside the kernel; one float per tree leaves the SM.
// MODE_R2: squared Pearson correlation of the single output with the label (problem/transformation.py:36-43).
enum : int { MODE_MSE = 0, MODE_ABS = 1, MODE_ACC = 2, MODE_OUTPUT = 3, MODE_ROWWISE = 4 };
__host__ __device__ inline bool is_reduce_mode(int mode) { return mode <= MODE_ACC; }            // one float per tree
__host__ __device__ inline int label_columns(int mode, int O) { return mode <= MODE_ABS ? O : (mode == MODE_ACC ? 1 : 0); }

// Fitness exchange fused into the evaluation kernel (multi-GPU): every tree's fitness is stored straight into each
// rank's full-population buffer through peer-mapped memory (NVLink), at row_offset + tree.
struct Scatter {
    float *const *peers = nullptr;   // DEVICE array of `world` pointers, one full-population fitness buffer per rank
    int world = 0;
    unsigned row_offset = 0;
};



// Multi-GPU fitness exchange: trees are pushed to the peers in chunks of kPushChunk consecutive trees, by whichever warp
// finishes a chunk's last tree (tickets are handed out in order, so chunks complete - and travel - while later trees
// are still being evaluated): 128-byte coalesced stores over NVLink instead of one 4-byte store per tree per peer.
constexpr int kPushShift = 10, kPushChunk = 1 << kPushShift;

struct ReplayArgs {
    const uint2 *prog;      // [P][Lp]
    unsigned *sched;        // [0] ticket counter
    const float *X;         // MODE_ROWWISE: [P][V]; else [N][V]
    const float *labels;    // [N][O] (loss modes); [N] class ids (MODE_ACC)
    float acc_half, acc_max;   // MODE_ACC, single output: prediction = clamp(round(out + acc_half), 0, acc_max)
    float *out;             // fitness[P] | results[P][N][O] | results[P][O]
    int P, Lp, N, V, O;
    int NP;                 // N rounded up to a whole number of passes
    int npass, depth, mode;
    int smem_depth;         // operand-stack slots per warp kept in shared memory
    int tmem_slots;         // TSTK: slots [0, tmem_slots) live in tensor memory, the deeper ones in shared memory
    int tmem_cols;          // TSTK: tensor-memory columns the CTA allocates (power of two >= 32)
    // datapoint tiling (dataset larger than the shared-memory staging area): this launch covers datapoints
    // [d_base, d_base + N) of N_total; loss modes carry the running sum in out[] between launches
    int d_base, N_total, first_tile, last_tile;
    // multi-GPU: on the last tile also store the fitness into every rank's buffer (Scatter)
    float *const *peers;
    int world;
    unsigned row_offset;
    unsigned *chunk_done;   // [ceil(P / kPushChunk)] finished-tree counters (zeroed by the lowering kernel, self-resetting)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Tensor memory as a per-lane scratch: a warp owns TMEM lanes 32 * (warp % 4) .. + 31, thread i <-> lane i, and
// the .32x32b.x8 shapes move 8 consecutive 32-bit columns of every lane to / from 8 registers per thread.
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float *v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(v[0]),
                 "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld8(float *v, uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "r"(taddr)
                 : "memory");
}
// one operand-stack slot = K columns (K = 8 or 16)
template <int K>
__device__ __forceinline__ void tmem_store_slot(uint32_t taddr, const float (&v)[K]) {
#pragma unroll
    for (int j = 0; j < K / 8; ++j) tmem_st8(taddr + 8u * j, v + 8 * j);
}
template <int K>
__device__ __forceinline__ void tmem_load_slot(float (&v)[K], uint32_t taddr) {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < K / 8; ++j) tmem_ld8(v + 8 * j, taddr + 8u * j);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

#define FOR_K _Pragma("unroll") for (int k = 0; k < K; ++k)

// A lane's K values of one vector live at  base + lane*VW + j*32*VW + r  (VW = min(K,4)):
// float4 accesses for K >= 4, conflict-free for every K.
template <int K>
__device__ __forceinline__ void ld_vec(float (&r)[K], const float *p) {
    if constexpr (K >= 4) {
#pragma unroll
        for (int j = 0; j < K / 4; ++j) {
            const float4 v = *reinterpret_cast<const float4 *>(p + j * 128);
            r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
        }
    } else {
        FOR_K r[k] = p[k * 32];
    }
}
template <int K>
__device__ __forceinline__ void st_vec(float *p, const float (&r)[K]) {
    if constexpr (K >= 4) {
#pragma unroll
        for (int j = 0; j < K / 4; ++j)
            *reinterpret_cast<float4 *>(p + j * 128) = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
    } else {
        FOR_K p[k * 32] = r[k];
    }
}
// datapoint index (within a pass) of a lane's k-th value
template <int K>
__device__ __forceinline__ int dp_index(int lane, int k) {
    if constexpr (K >= 4) return (k >> 2) * 128 + lane * 4 + (k & 3);
    else return k * 32 + lane;
}

// TSTK: the operand stack lives in tensor memory instead of shared memory (K == 8, single-output only)
// FEAT: what besides the plain loss / output modes is compiled into the kernel.  The hot configuration (FEAT_PLAIN) carries
// neither the fitness-exchange protocol nor the classification epilogue: every extra basic block moves the PTX replay
// loop in the instruction cache, and the loop is sensitive to that (profiles/README.md: +10 % with both compiled in).
enum : int { FEAT_PLAIN = 0, FEAT_EXCHANGE = 1, FEAT_ACC = 2 };

template <int K, bool MULTI, bool ROWWISE, bool TSTK = false, int FEAT = FEAT_PLAIN>
__global__ void __launch_bounds__(K == 16 ? 1024 : 256, K == 16 ? 1 : ((TSTK && K == 8) ? 4 : 2)) replay_kernel(ReplayArgs g) {
    static_assert(!TSTK || ((K == 8 || K == 16) && !MULTI && !ROWWISE), "tensor-memory stack: K = 8 / 16 single-output only");
    static_assert(K != 16 || TSTK, "K = 16 exists only with the tensor-memory stack");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int VW = K >= 4 ? 4 : 1;
    constexpr int SLOT = K * 32;                 // floats per stack slot / per output accumulator
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int lane_off = lane * VW;

    // ---- shared memory carve-up ----
    float *Xs = reinterpret_cast<float *>(smem_raw);                    // [V][NP]   (not ROWWISE)
    float *Ys = Xs + (ROWWISE ? 0 : (size_t)g.V * g.NP);                // [O][NP]   (loss modes)
    const int LC = label_columns(g.mode, g.O);
    float *after = Ys + (size_t)LC * g.NP;
    uint2 *progs = reinterpret_cast<uint2 *>(after) + (size_t)warp * 2 * g.Lp;          // 2 rows / warp
    float *stacks = reinterpret_cast<float *>(reinterpret_cast<uint2 *>(after) + (size_t)nwarp * 2 * g.Lp);
    float *stack = stacks + (size_t)warp * g.smem_depth * SLOT;
    float *outs_all = stacks + (size_t)nwarp * g.smem_depth * SLOT;     // [O][SLOT] / warp (MULTI)
    float *outs = outs_all + (MULTI ? (size_t)warp * g.O * SLOT : 0);
    uint64_t *bars = reinterpret_cast<uint64_t *>(outs_all + (MULTI ? (size_t)nwarp * g.O * SLOT : 0)) + warp * 2;

    // ---- stage the dataset once per CTA: X[N][V] -> Xs[V][NP], labels[N][O] -> Ys[O][NP] ----
    if constexpr (!ROWWISE) {
        const int tot = g.NP * g.V;
        for (int idx = threadIdx.x; idx < tot; idx += blockDim.x) {
            const int d = idx / g.V, v = idx - d * g.V;
            Xs[v * g.NP + d] = d < g.N ? __ldg(g.X + idx) : 0.0f;
        }
        if (LC > 0) {
            const int tl = g.NP * LC;
            for (int idx = threadIdx.x; idx < tl; idx += blockDim.x) {
                const int d = idx / LC, o = idx - d * LC;
                Ys[o * g.NP + d] = d < g.N ? __ldg(g.labels + idx) : 0.0f;
            }
        }
    }
    if (lane == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // tensor-memory operand stack: warp 0 allocates the CTA's columns; warp w uses lanes 32 * (w % 4) and the
    // column block (w / 4) * depth * 8
    __shared__ uint32_t tmem_base_slot;
    uint32_t tstack = 0;
    if constexpr (TSTK) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                         "r"((uint32_t)g.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    __syncthreads();
    if constexpr (TSTK) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tstack = tmem_base_slot + ((uint32_t)(warp & 3) * 32u << 16) + (uint32_t)(warp >> 2) * (uint32_t)g.tmem_slots * (uint32_t)K;
    }

    // ---- fused all-gather over peer-mapped memory (NVLink): the warp that completes a chunk of kPushChunk consecutive
    //      trees pushes it to every rank with coalesced stores.  Completion is counted per chunk; the protocol is software-
    //      pipelined over this warp's trees so that no round trip is ever waited for: at the end of tree k lane 0
    //      (1) reads the count returned for tree k-2, (2) fences and counts tree k-1 (its fitness store was issued a whole
    //      tree ago, so the fence finds it already performed), (3) has just stored the fitness of tree k. ----
    int ex_stored = -1, ex_counted = -1;     // lane 0: tree whose fitness is stored but not counted / counted, result pending
    unsigned ex_done = 0;
    auto exchange_step = [&](int cur) {
        int push = -1;
        if (lane == 0) {
            if (ex_counted >= 0) {
                const int chunk = ex_counted >> kPushShift;
                if (ex_done == (unsigned)min(kPushChunk, g.P - (chunk << kPushShift))) push = chunk;
            }
            ex_counted = ex_stored;
            if (ex_stored >= 0) {
                __threadfence();                                               // that tree's fitness is visible before its count
                ex_done = atomicAdd(g.chunk_done + (ex_stored >> kPushShift), 1u) + 1u;
            }
            ex_stored = cur;
        }
        push = __shfl_sync(0xffffffffu, push, 0);
        if (push >= 0) {
            __threadfence();                                                   // every counted tree's fitness before the push
            const int first = push << kPushShift;
            const unsigned cnt = (unsigned)min(kPushChunk, g.P - first);
            const float *src = g.out + first;
            for (int r = 0; r < g.world; ++r) {
                float *dst = g.peers[r] + g.row_offset + (unsigned)first;
                for (unsigned i = lane; i < cnt; i += 32) dst[i] = __ldcg(src + i);
            }
            if (lane == 0) g.chunk_done[push] = 0u;                            // ready for the next launch
        }
    };

    const uint32_t row_bytes = (uint32_t)g.Lp * 8u;
    uint32_t phase0 = 0, phase1 = 0;
    int buf = 0;
    int tree = 0;
    if (lane == 0) tree = (int)atomicAdd(g.sched, 1u);
    tree = __shfl_sync(0xffffffffu, tree, 0);
    if (lane == 0 && tree < g.P) {
        mbar_expect_tx(bars, row_bytes);
        tma_load_1d(progs, g.prog + (size_t)tree * g.Lp, row_bytes, bars);
    }

    while (tree < g.P) {
        // ticket + prefetch for the tree after this one
        int next = 0;
        if (lane == 0) {
            next = (int)atomicAdd(g.sched, 1u);
            if (next < g.P) {
                mbar_expect_tx(bars + (buf ^ 1), row_bytes);
                tma_load_1d(progs + (size_t)(buf ^ 1) * g.Lp, g.prog + (size_t)next * g.Lp, row_bytes, bars + (buf ^ 1));
            }
        }
        next = __shfl_sync(0xffffffffu, next, 0);
        if (buf == 0) { mbar_wait(bars, phase0); phase0 ^= 1; }
        else          { mbar_wait(bars + 1, phase1); phase1 ^= 1; }
        const uint2 *prog = progs + (size_t)buf * g.Lp;

        float err = 0.0f;
        for (int pass = 0; pass < g.npass; ++pass) {
            const int pass_off = pass * SLOT;
            const float *xl;
            if constexpr (ROWWISE) xl = g.X + (size_t)tree * g.V;
            else xl = Xs + pass_off + lane_off;

            auto fetch_var = [&](float(&l)[K], uint32_t idx) {
                if constexpr (ROWWISE) {
                    const float x = __ldg(xl + idx);
                    FOR_K l[k] = x;
                } else {
                    ld_vec<K>(l, xl + (size_t)idx * g.NP);
                }
            };

            float acc[K];
            FOR_K acc[k] = 0.0f;
            if constexpr (MULTI) {
                for (int o = 0; o < g.O; ++o) st_vec<K>(outs + o * SLOT + lane_off, acc);
            }
            int pc = 0;
            // operand-stack slots are static (program.cuh).  TSTK: slots < tmem_slots in tensor memory, deeper ones
            // (reached only through the deep opcodes, i.e. through this generic path) in shared memory
            auto slot_store = [&](int slot) {
                if constexpr (TSTK) {
                    if (slot < g.tmem_slots) { tmem_store_slot<K>(tstack + (uint32_t)slot * (uint32_t)K, acc); return; }
                    slot -= g.tmem_slots;
                }
                st_vec<K>(stack + slot * SLOT + lane_off, acc);
            };
            auto slot_load = [&](float(&d)[K], int slot) {
                if constexpr (TSTK) {
                    if (slot < g.tmem_slots) { tmem_load_slot<K>(d, tstack + (uint32_t)slot * (uint32_t)K); return; }
                    slot -= g.tmem_slots;
                }
                ld_vec<K>(d, stack + slot * SLOT + lane_off);
            };

            // ---- generic interpreter: one instruction per call; two stages (operands by form,
            //      then ONE switch over the operator) keep it small enough to stay cache-resident ----
            auto step = [&]() -> bool {
                const uint2 ins = prog[pc];
                const uint32_t w = ins.x;
                const float cst = __uint_as_float(ins.y);
                const uint32_t code = w & I_CODE_MASK, form = code >> 4, op = code & 15u;
                const uint32_t ia = (w >> I_IDXA_SHIFT) & I_IDXA_MASK, ib = (w >> I_IDXB_SHIFT) & I_IDXB_MASK;
                ++pc;
                if (code == C_END) return true;
                float x[K], y[K], r[K];
                if (code == C_IF3) {
                    if constexpr (MULTI) {   // {hdr, a}{b, c}: three leaf operands
                        const uint2 ext = prog[pc];
                        ++pc;
                        float z[K];
                        if (w & I_IF3_ACONST) { FOR_K x[k] = cst; } else fetch_var(x, ia);
                        if (w & I_IF3_BCONST) { FOR_K y[k] = __uint_as_float(ext.x); } else fetch_var(y, ext.x & I_IDXA_MASK);
                        if (w & I_IF3_CCONST) { FOR_K z[k] = __uint_as_float(ext.y); } else fetch_var(z, ext.y & I_IDXA_MASK);
                        FOR_K r[k] = x[k] > 0.0f ? y[k] : z[k];
                    } else {
                        FOR_K r[k] = 0.0f;
                    }
                } else {
                    if constexpr (!MULTI) {   // multi-output programs have no operand stack
                        const uint32_t push = (w & I_PUSH_MASK) >> I_PUSH_SHIFT;
                        if (push) slot_store((int)push - 1);
                    }
                    switch (form) {
                    case FM_MISC:
                        if (code == C_LOAD_V || code == C_LOAD_V_DEEP) { fetch_var(acc, ia); return false; }
                        if (code == C_LOAD_K || code == C_LOAD_K_DEEP) { FOR_K acc[k] = cst; return false; }
                        if (code == C_IF) {   // forward.cu:223
                            float t1[K], t2[K];
                            slot_load(t1, (int)ib + 1);   // newer of the two saved values
                            slot_load(t2, (int)ib);
                            const uint32_t sa = ia & 3, sb = (ia >> 2) & 3, sc = (ia >> 4) & 3;
                            FOR_K {
                                const float a = sa == 0 ? acc[k] : (sa == 1 ? t1[k] : t2[k]);
                                const float b = sb == 0 ? acc[k] : (sb == 1 ? t1[k] : t2[k]);
                                const float c = sc == 0 ? acc[k] : (sc == 1 ? t1[k] : t2[k]);
                                acc[k] = a > 0.0f ? b : c;
                            }
                            return false;
                        }
                        // C_NAN (malformed row) and anything unknown
                        FOR_K acc[k] = __int_as_float(0x7fc00000);
                        if constexpr (MULTI)
                            for (int o = 0; o < g.O; ++o) st_vec<K>(outs + o * SLOT + lane_off, acc);
                        return false;
                    case FM_UA: FOR_K x[k] = acc[k]; break;
                    case FM_UV: fetch_var(x, ia); break;
                    case FM_UK: FOR_K x[k] = cst; break;
                    case FM_AV: FOR_K x[k] = acc[k]; fetch_var(y, ia); break;
                    case FM_AK: FOR_K { x[k] = acc[k]; y[k] = cst; } break;
                    case FM_VA: fetch_var(x, ia); FOR_K y[k] = acc[k]; break;
                    case FM_KA: FOR_K { x[k] = cst; y[k] = acc[k]; } break;
                    case FM_VV: fetch_var(x, ia); fetch_var(y, ib); break;
                    case FM_VK: fetch_var(x, ia); FOR_K y[k] = cst; break;
                    case FM_KV: FOR_K x[k] = cst; fetch_var(y, ia); break;
                    case FM_SA: case FM_DA: slot_load(x, (int)ia); FOR_K y[k] = acc[k]; break;
                    case FM_AS: case FM_AD: FOR_K x[k] = acc[k]; slot_load(y, (int)ia); break;
                    default: FOR_K { x[k] = 0.0f; y[k] = 0.0f; } break;
                    }
                    if (form <= FM_UK) {
#define U_CASE(u) case u: FOR_K r[k] = unary_op<u>(x[k]); break;
                        switch (op) {
                            U_CASE(0) U_CASE(1) U_CASE(2) U_CASE(3) U_CASE(4) U_CASE(5) U_CASE(6) U_CASE(7)
                            U_CASE(8) U_CASE(9) U_CASE(10) U_CASE(11) U_CASE(12) U_CASE(13) U_CASE(14)
                        default: FOR_K r[k] = 0.0f; break;
                        }
#undef U_CASE
                    } else {
#define B_CASE(b) case b: FOR_K r[k] = binary_op<b>(x[k], y[k]); break;
                        switch (op) {
                            B_CASE(0) B_CASE(1) B_CASE(2) B_CASE(3) B_CASE(4) B_CASE(5) B_CASE(6)
                            B_CASE(7) B_CASE(8) B_CASE(9) B_CASE(10) B_CASE(11) B_CASE(12)
                        default: FOR_K r[k] = 0.0f; break;
                        }
#undef B_CASE
                    }
                }
                if constexpr (MULTI) {   // every instruction of a multi-output program is an OUT node (or its LOAD)
                    if ((w & I_OUT) && ib != I_IDXB_MASK) {
                        float o_[K];
                        ld_vec<K>(o_, outs + ib * SLOT + lane_off);
                        FOR_K o_[k] += r[k];
                        st_vec<K>(outs + ib * SLOT + lane_off, o_);
                    }
                }
                FOR_K acc[k] = r[k];
                return false;
            };

            if constexpr ((K == 8 || K == 16) && !MULTI && !ROWWISE) {
                // PTX fast path (fastpath_k8.inc): brx.idx jump table, operands by opcode
                const uint32_t prog_base = smem_u32(prog);
                const uint32_t stack_base = TSTK ? tstack : smem_u32(stack + lane_off);
                uint32_t pc_addr = prog_base, status;
                const uint32_t xl_addr = smem_u32(xl), npb = (uint32_t)g.NP * 4u;
#if EVOGP_FASTPATH_K16_TMEM_ASM_FUSED_LOSS
                // K = 16: the PTX loop runs every pass of the tree and accumulates the loss itself when the passes are
                // whole and the mode is a plain loss (gen_fastpath.py FUSED_LOSS); it comes back here only for a slow-path
                // instruction (in whatever pass it is in) or when the last pass is done
                uint32_t xl_cur = xl_addr, yl_cur = smem_u32(Ys + pass_off + lane_off), pass_cur = (uint32_t)pass;
                const uint32_t loss_mode = (K == 16 && FEAT != FEAT_ACC && g.mode <= MODE_ABS && g.N % SLOT == 0)
                                               ? (g.mode == MODE_MSE ? 1u : 2u) : 0u;
#endif
                for (;;) {
                    if constexpr (K == 16) {
#if EVOGP_FASTPATH_K16_TMEM_ASM_FUSED_LOSS
                        asm volatile(EVOGP_FASTPATH_K16_TMEM_ASM
                                     : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]),
                                       "+f"(acc[6]), "+f"(acc[7]), "+f"(acc[8]), "+f"(acc[9]), "+f"(acc[10]), "+f"(acc[11]),
                                       "+f"(acc[12]), "+f"(acc[13]), "+f"(acc[14]), "+f"(acc[15]), "+r"(pc_addr), "=r"(status),
                                       "+r"(xl_cur), "+f"(err), "+r"(yl_cur), "+r"(pass_cur)
                                     : "r"(npb), "r"(stack_base), "r"(stack_base - 16u), "r"((uint32_t)g.npass), "r"(prog_base),
                                       "r"(loss_mode)
                                     : "memory");
                        if (loss_mode != 0u && status != 0u)      // the generic step below reads this pass's columns
                            xl = Xs + (size_t)pass_cur * SLOT + lane_off;
#else
                        asm volatile(EVOGP_FASTPATH_K16_TMEM_ASM
                                     : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]),
                                       "+f"(acc[6]), "+f"(acc[7]), "+f"(acc[8]), "+f"(acc[9]), "+f"(acc[10]), "+f"(acc[11]),
                                       "+f"(acc[12]), "+f"(acc[13]), "+f"(acc[14]), "+f"(acc[15]), "+r"(pc_addr), "=r"(status)
                                     : "r"(xl_addr), "r"(npb), "r"(stack_base), "r"(stack_base - 16u)
                                     : "memory");
#endif
                    } else if constexpr (TSTK) {
                        asm volatile(EVOGP_FASTPATH_K8_TMEM_ASM
                                     : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]),
                                       "+f"(acc[6]), "+f"(acc[7]), "+r"(pc_addr), "=r"(status)
                                     : "r"(xl_addr), "r"(npb), "r"(stack_base), "r"(stack_base - 8u)
                                     : "memory");
                    } else {
                        asm volatile(EVOGP_FASTPATH_K8_ASM
                                     : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]),
                                       "+f"(acc[6]), "+f"(acc[7]), "+r"(pc_addr), "=r"(status)
                                     : "r"(xl_addr), "r"(npb), "r"(stack_base)
                                     : "memory");
                    }
                    if (status == 0) break;
                    pc = (int)((pc_addr - prog_base) >> 3);
                    if (step()) break;
                    pc_addr = prog_base + ((uint32_t)pc << 3);
                }
#if EVOGP_FASTPATH_K16_TMEM_ASM_FUSED_LOSS
                if constexpr (K == 16) {
                    if (loss_mode != 0u) {     // every pass replayed and summed inside the loop
                        pass = g.npass;
                        continue;
                    }
                }
#endif
            } else if constexpr (K == 8 && MULTI && !ROWWISE) {
                // PTX loop for multi-output programs (fastpath_k8_multi.inc): leaf-operand forms, outs[] += in shared memory;
                // C_IF3 and the rare operators come back here one instruction at a time
                const uint32_t prog_base = smem_u32(prog), outs_base = smem_u32(outs + lane_off);
                uint32_t pc_addr = prog_base, status;
                const uint32_t xl_addr = smem_u32(xl), npb = (uint32_t)g.NP * 4u;
                for (;;) {
                    asm volatile(EVOGP_FASTPATH_K8_MULTI_ASM
                                 : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]),
                                   "+f"(acc[6]), "+f"(acc[7]), "+r"(pc_addr), "=r"(status)
                                 : "r"(xl_addr), "r"(npb), "r"(outs_base)
                                 : "memory");
                    if (status == 0) break;
                    pc = (int)((pc_addr - prog_base) >> 3);
                    if (step()) break;
                    pc_addr = prog_base + ((uint32_t)pc << 3);
                }
            } else {
                while (!step()) {
                }
            }

            // ---- per-pass epilogue ----
            if (FEAT == FEAT_ACC && g.mode == MODE_ACC) {
                float lab[K];
                ld_vec<K>(lab, Ys + pass_off + lane_off);
                if constexpr (MULTI) {
                    // arg-max of softmax(outputs) as torch computes it (classification.py:62-64): softmax is monotone, so
                    // the first maximal output wins; a NaN or an infinite maximum makes every probability NaN -> class 0
                    float best[K], cls[K];
                    bool poison[K];
                    ld_vec<K>(best, outs + lane_off);
                    FOR_K { cls[k] = 0.0f; poison[k] = !(best[k] == best[k]); }
                    for (int o = 1; o < g.O; ++o) {
                        float r[K];
                        ld_vec<K>(r, outs + o * SLOT + lane_off);
                        FOR_K {
                            poison[k] = poison[k] || !(r[k] == r[k]);
                            if (r[k] > best[k]) { best[k] = r[k]; cls[k] = (float)o; }
                        }
                    }
                    FOR_K {
                        const float pred = (poison[k] || fabsf(best[k]) == __int_as_float(0x7f800000)) ? 0.0f : cls[k];
                        if (pass_off + dp_index<K>(lane, k) < g.N && pred == lab[k]) err += 1.0f;
                    }
                } else {
                    FOR_K {   // transform(): clamp(round(out + max / 2), 0, max), round half to even (classification.py:51-52)
                        const float p = fminf(fmaxf(rintf(acc[k] + g.acc_half), 0.0f), g.acc_max);
                        if (pass_off + dp_index<K>(lane, k) < g.N && acc[k] == acc[k] && p == lab[k]) err += 1.0f;
                    }
                }
            } else if (g.mode <= MODE_ABS) {
                if constexpr (MULTI) {
                    for (int o = 0; o < g.O; ++o) {
                        float y[K], r[K];
                        ld_vec<K>(y, Ys + (size_t)o * g.NP + pass_off + lane_off);
                        ld_vec<K>(r, outs + o * SLOT + lane_off);
                        FOR_K {
                            const float diff = y[k] - r[k];
                            const float e = g.mode == MODE_MSE ? diff * diff : fabsf(diff);
                            if (pass_off + dp_index<K>(lane, k) < g.N) err += e;
                        }
                    }
                } else {
                    float y[K];
                    ld_vec<K>(y, Ys + pass_off + lane_off);
                    if (pass_off + SLOT <= g.N) {   // whole pass in range (warp-uniform): no per-datapoint bound checks
                        if (g.mode == MODE_MSE) { FOR_K { const float diff = y[k] - acc[k]; err += diff * diff; } }
                        else { FOR_K err += fabsf(y[k] - acc[k]); }
                    } else {
                        FOR_K {
                            const float diff = y[k] - acc[k];
                            const float e = g.mode == MODE_MSE ? diff * diff : fabsf(diff);
                            if (pass_off + dp_index<K>(lane, k) < g.N) err += e;
                        }
                    }
                }
            } else if (g.mode == MODE_OUTPUT) {
                FOR_K {
                    const int d = pass_off + dp_index<K>(lane, k);
                    if (d < g.N) {
                        float *dst = g.out + ((size_t)tree * g.N_total + g.d_base + d) * g.O;
                        if constexpr (MULTI) {
                            for (int o = 0; o < g.O; ++o) dst[o] = outs[o * SLOT + lane_off + (K >= 4 ? (k >> 2) * 128 + (k & 3) : k * 32)];
                        } else {
                            dst[0] = acc[k];
                        }
                    }
                }
            } else {   // MODE_ROWWISE: every lane computed the same value; lane 0 stores
                if (lane == 0) {
                    float *dst = g.out + (size_t)tree * g.O;
                    if constexpr (MULTI) {
                        for (int o = 0; o < g.O; ++o) dst[o] = outs[o * SLOT];
                    } else {
                        dst[0] = acc[0];
                    }
                }
            }
        }
        if (is_reduce_mode(g.mode)) {
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) err += __shfl_xor_sync(0xffffffffu, err, s);
            if (!g.first_tile) {                                              // running sum of the earlier tiles:
                float prev = lane == 0 ? g.out[tree] : 0.0f;                  // lane 0 reads (it is the one that writes)
                err += __shfl_sync(0xffffffffu, prev, 0);
            }
            // forward.cu:478 divides with div.approx (-use_fast_math), and so does this; the accuracy is count / N as torch
            // computes a tensor / python-scalar division (classification.py:66): count * (1 / N), both correctly rounded
            const float fit = !g.last_tile ? err : ((FEAT == FEAT_ACC && g.mode == MODE_ACC) ? __fmul_rn(err, __frcp_rn((float)(unsigned)g.N_total)) : err / (float)(unsigned)g.N_total);
            if (lane == 0) g.out[tree] = fit;
            if constexpr (FEAT == FEAT_EXCHANGE) {
                if (g.peers != nullptr && g.last_tile) exchange_step(tree);
            }
        }
        __syncwarp();   // every lane is done with prog[buf] before lane 0 re-targets it
        buf ^= 1;
        tree = next;
    }
    if constexpr (FEAT == FEAT_EXCHANGE) {
        if (g.peers != nullptr && g.last_tile && is_reduce_mode(g.mode)) {   // drain the exchange pipeline
            exchange_step(-1);
            exchange_step(-1);
        }
    }
    if constexpr (TSTK) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_slot), "r"((uint32_t)g.tmem_cols)
                         : "memory");
    }
}

// ---------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------
// device properties / measurement hooks of the calling thread's current device (owned by eval.cu)
struct ReplayEnv {
    int sm_count, max_smem, smem_per_sm;
    cudaEvent_t ev_begin, ev_end;
};
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

constexpr int kTmemSlots16 = 4;   // K = 16: operand-stack slots kept in tensor memory (deeper ones: shared memory)

// shared memory one warp of a replay CTA needs: two program rows, its shared-memory stack slots, multi-output accumulators
static size_t replay_per_warp(int K, bool multi, bool tmem, int depth, int Lp, int O) {
    const size_t SLOT = 32 * (size_t)K;
    int smem_depth = multi ? 1 : (depth > 0 ? depth : 1);
    if (tmem) smem_depth = K == 16 ? (depth > kTmemSlots16 ? depth - kTmemSlots16 : 0) : 0;
    return (size_t)2 * Lp * 8 + (size_t)smem_depth * SLOT * 4 + (multi ? (size_t)O * SLOT * 4 : 0) + 16;
}
// tensor-memory columns a CTA of `warps` warps needs for operand stacks of `depth` slots (8 columns per slot; the
// warps of one lane quarter share the columns), as the power of two >= 32 tcgen05.alloc accepts
static int tmem_stack_cols(int warps, int slots, int K = 8) {
    const int need = ((warps + 3) / 4) * slots * K;
    int cols = 32;
    while (cols < need) cols <<= 1;
    return cols;
}

template <int K, bool MULTI, bool ROWWISE, bool TSTK, int FEAT>
static int launch_replay_t(ReplayArgs a, int depth, cudaStream_t st, const ReplayEnv &env) {
    auto kern = replay_kernel<K, MULTI, ROWWISE, TSTK, FEAT>;
    const int g_max_smem = env.max_smem, g_sm_count = env.sm_count, g_smem_per_sm = env.smem_per_sm;
    const cudaEvent_t g_ev_replay_begin = env.ev_begin, g_ev_replay_end = env.ev_end;
    const int SLOT = K * 32;
    a.npass = ROWWISE ? 1 : (a.N + SLOT - 1) / SLOT;
    a.NP = a.npass * SLOT;
    a.depth = MULTI ? 1 : (depth > 0 ? depth : 1);   // multi-output programs use no operand stack
    depth = a.depth;
    // TSTK: K = 8 keeps the whole stack in tensor memory (the launcher only picks it for depth <= 8); K = 16 keeps
    // the first kTmemSlots16 slots there and the deeper, rarely reached ones in shared memory (deep opcodes)
    a.tmem_slots = TSTK ? (K == 16 ? (depth < kTmemSlots16 ? depth : kTmemSlots16) : depth) : 0;
    a.smem_depth = depth - a.tmem_slots;
    a.tmem_cols = 0;
    auto per_warp = [&]() { return (size_t)2 * a.Lp * 8 + (size_t)a.smem_depth * SLOT * 4 + (MULTI ? (size_t)a.O * SLOT * 4 : 0) + 16; };
    // the dataset slice a launch stages: all of it when it fits next to >= 4 warps, else whole passes of it
    const size_t per_dp = ROWWISE ? 0 : ((size_t)a.V + label_columns(a.mode, a.O)) * 4;   // bytes per datapoint
    const int N_total = a.N;
    int tile = a.NP;                                                                         // datapoints per launch
    // one launch when the dataset leaves room for two 8-warp CTAs per SM; otherwise tiles of <= 48 KB of dataset so
    // that occupancy survives (a 164 KB tile would leave one 4-warp CTA per SM: measured 6x slower on configs[3])
    if (per_dp && 2 * (per_dp * tile + 8 * per_warp() + 1024) > (size_t)g_max_smem) {
        size_t room = 48 * 1024;
        if (room + 4 * per_warp() > (size_t)g_max_smem) room = (size_t)g_max_smem > 4 * per_warp() ? (size_t)g_max_smem - 4 * per_warp() : 0;
        tile = (int)(room / per_dp / SLOT) * SLOT;
        if (tile < SLOT) tile = ((size_t)SLOT * per_dp + per_warp() <= (size_t)g_max_smem) ? SLOT : 0;   // one pass, as many warps as fit
        if (tile >= a.NP) tile = a.NP;
        if (tile < SLOT) {   // cannot happen: choose_replay picks a K whose single pass fits (replay_fits)
            set_error("one pass of %d datapoints x (%d inputs + %d labels) does not fit the %d B shared-memory staging area", SLOT, a.V, a.O, g_max_smem);
            return EVOGP_ERR_UNSUPPORTED;
        }
    }
    const float *X0 = a.X, *Y0 = a.labels;
    unsigned *sched0 = a.sched;
    if (g_ev_replay_begin) cudaEventRecord(g_ev_replay_begin, st);
    for (int d0 = 0, t = 0; d0 < N_total; d0 += tile, ++t) {
        a.N = N_total - d0 < tile ? N_total - d0 : tile;
        a.npass = ROWWISE ? 1 : (a.N + SLOT - 1) / SLOT;
        a.NP = a.npass * SLOT;
        a.d_base = d0; a.N_total = N_total; a.first_tile = d0 == 0; a.last_tile = d0 + tile >= N_total;
        if (!ROWWISE) a.X = X0 + (size_t)d0 * a.V;
        if (Y0) a.labels = Y0 + (size_t)d0 * label_columns(a.mode, a.O);
        // one ticket counter per launch: the 64 words lower_kernel zeroed, reused round-robin (launches are ordered
        // on the stream, so word t % 64 is idle again by the time launch t is enqueued)
        a.sched = sched0 + (t & 63);
        if (t >= 64) EVOGP_CUDA(cudaMemsetAsync(a.sched, 0, sizeof(unsigned), st));
        const size_t data = per_dp * a.NP;
        int warps = 8;
        if constexpr (K == 16) {
            // one CTA per SM of up to 32 warps (64 registers each; 4 slots x 16 columns x 8 warps per lane quarter =
            // the 512 columns), as many whole lane quarters as the shared memory left by the dataset holds
            const size_t room = (size_t)g_max_smem > data ? (size_t)g_max_smem - data : 0;
            warps = (int)(room / per_warp()) & ~3;
            warps = warps > 32 ? 32 : (warps < 4 ? 4 : warps);
        }
        while (warps > 1 && data + warps * per_warp() > (size_t)g_max_smem) warps >>= 1;
        size_t smem = data + warps * per_warp();
        EVOGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        EVOGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
        if (per_sm < 1) per_sm = 1;
        if constexpr (TSTK) {
            // The occupancy API answers 1 CTA/SM for a kernel that allocates tensor memory; the real limits are
            // registers, shared memory and the 512 columns (every resident CTA holds its columns until it exits,
            // and tcgen05.alloc blocks when they run out - CTAs beyond `fit` would only wait).
            a.tmem_cols = tmem_stack_cols(warps, a.tmem_slots, K);
            const int fit = 512 / a.tmem_cols;
            cudaFuncAttributes fa;
            EVOGP_CUDA(cudaFuncGetAttributes(&fa, kern));
            const int regs_per_warp = ((fa.numRegs * 32 + 255) / 256) * 256;
            const int by_regs = 65536 / (regs_per_warp * warps);
            const int by_smem = (int)((size_t)g_smem_per_sm / (smem + fa.sharedSizeBytes + 1024));
            per_sm = fit < by_regs ? fit : by_regs;
            if (by_smem < per_sm) per_sm = by_smem;
            if (per_sm < 1) per_sm = 1;
        }
        long long want = ((long long)a.P + warps - 1) / warps;
        int grid = (int)(want < (long long)per_sm * g_sm_count ? want : (long long)per_sm * g_sm_count);
        if (grid < 1) grid = 1;
        kern<<<grid, warps * 32, smem, st>>>(a);
        count_launch();
        const int rc = check_launch("replay_kernel");
        if (rc) return rc;
    }
    if (g_ev_replay_end) cudaEventRecord(g_ev_replay_end, st);
    return EVOGP_OK;
}

struct ReplayChoice {
    int K;
    bool tmem;
};

template <bool MULTI, int FEAT>
static int launch_replay(const ReplayArgs &a, int depth, ReplayChoice c, cudaStream_t st, const ReplayEnv &env) {
    if constexpr (FEAT == FEAT_PLAIN) {   // the row-wise `evaluate` mode exists only in the plain build
        if (a.mode == MODE_ROWWISE) return launch_replay_t<1, MULTI, true, false, FEAT>(a, depth, st, env);
    }
    if constexpr (!MULTI) {
        if (c.K == 16) return launch_replay_t<16, false, false, true, FEAT>(a, depth, st, env);
        if (c.K == 8 && c.tmem) return launch_replay_t<8, false, false, true, FEAT>(a, depth, st, env);
    }
    switch (c.K) {
    case 8: return launch_replay_t<8, MULTI, false, false, FEAT>(a, depth, st, env);
    case 4: return launch_replay_t<4, MULTI, false, false, FEAT>(a, depth, st, env);
    default: return launch_replay_t<1, MULTI, false, false, FEAT>(a, depth, st, env);
    }
}

// one translation unit per FEAT (eval.cu, eval_exchange.cu, eval_acc.cu) defines its dispatcher
#define EVOGP_DEFINE_REPLAY_DISPATCH(name, FEAT)                                                                          \
    int name(bool multi, const ReplayArgs &a, int depth, ReplayChoice c, void *stream, const ReplayEnv &env) {             \
        cudaStream_t st = static_cast<cudaStream_t>(stream);                                                               \
        return multi ? launch_replay<true, FEAT>(a, depth, c, st, env) : launch_replay<false, FEAT>(a, depth, c, st, env); \
    }
int launch_replay_plain(bool multi, const ReplayArgs &a, int depth, ReplayChoice c, void *stream, const ReplayEnv &env);
int launch_replay_exchange(bool multi, const ReplayArgs &a, int depth, ReplayChoice c, void *stream, const ReplayEnv &env);
int launch_replay_acc(bool multi, const ReplayArgs &a, int depth, ReplayChoice c, void *stream, const ReplayEnv &env);



}  // namespace evogp
