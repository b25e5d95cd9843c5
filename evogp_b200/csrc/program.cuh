// program.cuh — the accumulator-machine program a packed tree row is lowered to,
// and the scalar operator definitions shared by the lowering pass (which never
// evaluates them) and the replay kernels.
//
// Why a lowering pass at all: the reference interprets the prefix row once per
// (tree, datapoint) thread with a private operand stack in local memory
// (forward.cu:246-302).  Here one warp replays a tree over 32*K datapoints per
// pass, so control flow is warp-uniform and the only question is how many issue
// slots a node costs.  Lowering removes every leaf push (leaves become operands of
// their parent), orders sibling subtrees Sethi-Ullman style so the operand stack
// never holds more than ~log2(L) live vectors (it lives in shared memory, one
// float4 column per lane), and turns each function node into ONE instruction whose
// opcode already says where its operands are.
//
// Values computed per node are exactly the reference's: reordering sibling
// evaluation does not change any operand of any operator.
#pragma once
#include "common.cuh"

namespace evogp {

// ---------------------------------------------------------------------------
// instruction word (8 bytes): {header, constant}
//   header [8:0]   opcode = form * 16 + op      (form and op tables below)
//          [12:9]  PUSH    0: nothing; s+1: save acc into operand-stack slot s before executing.
//                          Only ever set on instructions that start a fresh value
//                          (LOAD_*, UV/UK, VV, VK, KV).  Multi-output programs never push; there
//                          bit 9 is OUT (add the result to outs[idxB]) and bits 10-12 are the
//                          constant flags of C_IF3.
//          [22:13] idxA    variable index of the (first) variable operand; SA/AS/DA/AD: stack slot;
//                          C_IF: operand permutation
//          [31:23] idxB    variable index of the second variable operand (VV); output index when
//                          OUT (0x1FF: out of range, result dropped); C_IF: lower of its two slots
//   constant       the constant operand of a *K* form, bit-cast
// Operand kinds are part of the opcode, so the replay loop never tests a flag to find an operand:
//   A accumulator, V variable (dataset column), K constant, S operand-stack slot, D "deep" operand-stack
//   slot.  Slot numbers are static: sibling order is fixed by the lowering pass, so the stack height at
//   every push and pop is known there and the replay loop keeps no stack pointer.
//   Deep slots: a kernel that keeps only the first T slots in its fast storage (tensor memory) asks the
//   lowering pass (deep_from = T) to mark every access to a slot >= T in the opcode - pushes as
//   C_LOAD_*_DEEP, pops as the DA / AD forms - so its hot bodies never test the slot number; those
//   opcodes take the generic path, which keeps slots >= T in shared memory.  Programs for kernels with
//   one stack medium (deep_from = none) never contain them.
// ---------------------------------------------------------------------------
constexpr uint32_t I_CODE_MASK = 0x1FFu;
constexpr int I_PUSH_SHIFT = 9;
constexpr uint32_t I_PUSH_MASK = 0xFu << I_PUSH_SHIFT;
constexpr uint32_t I_OUT = 1u << 9;                                                  // multi-output programs
constexpr uint32_t I_IF3_ACONST = 1u << 10, I_IF3_BCONST = 1u << 11, I_IF3_CCONST = 1u << 12;
#ifdef EVOGP_IDXA_TOP   // A/B layout: idxA in the top ten bits (one shift extracts it), idxB below it
constexpr int I_IDXA_SHIFT = 22, I_IDXB_SHIFT = 13;
#else
constexpr int I_IDXA_SHIFT = 13, I_IDXB_SHIFT = 23;
#endif
constexpr uint32_t I_IDXA_MASK = 0x3FFu, I_IDXB_MASK = 0x1FFu;
// (Operand-stack slots held in registers were measured and rejected: profiles/r1_replay_v3_regbanks.txt - 45 %
// fewer shared-memory wavefronts but 25 more registers and MOV work per save, 479 us vs 309 us.)
constexpr int kNoDeepSlots = 255;   // deep_from value meaning "every slot is an ordinary slot"

constexpr int NUM_U = 16;  // 15 unary functions (ids 14..28) + "unknown id -> 0"
constexpr int NUM_B = 14;  // 13 binary functions (ids 1..13) + "unknown id -> 0"
constexpr int U_ZERO = 15, B_ZERO = 13;

// forms (opcode >> 4)
enum : int {
    FM_MISC = 0,
    FM_UA = 1,   // acc = u(acc)
    FM_UV = 2,   // acc = u(var A)
    FM_UK = 3,   // acc = u(const)
    FM_AV = 4,   // acc = b(acc, var A)
    FM_AK = 5,   // acc = b(acc, const)
    FM_VA = 6,   // acc = b(var A, acc)
    FM_KA = 7,   // acc = b(const, acc)
    FM_VV = 8,   // acc = b(var A, var B)
    FM_VK = 9,   // acc = b(var A, const)
    FM_KV = 10,  // acc = b(const, var A)
    FM_SA = 11,  // acc = b(slot[idxA], acc)
    FM_AS = 12,  // acc = b(acc, slot[idxA])
    FM_DA = 13,  // acc = b(slot[idxA], acc)      idxA >= deep_from
    FM_AD = 14,  // acc = b(acc, slot[idxA])
    FM_COUNT = 15
};
// FM_MISC opcodes
enum : int {
    C_END = 0,     // stop
    C_LOAD_V = 1,  // acc = var A
    C_LOAD_K = 2,  // acc = const
    C_NAN = 3,     // malformed row: result NaN
    C_IF = 4,      // acc = a > 0 ? b : c; operands are acc / slot idxB+1 / slot idxB per idxA
    C_IF3 = 5,     // multi-output only, 2 slots {hdr, a}{b, c}: r = a > 0 ? b : c on three leaf operands
    C_LOAD_V_DEEP = 6,   // C_LOAD_V / C_LOAD_K whose PUSH targets a deep slot
    C_LOAD_K_DEEP = 7,
    C_COUNT = FM_COUNT * 16
};
__host__ __device__ inline int opcode(int form, int op) { return form * 16 + op; }
__host__ __device__ inline int unary_slot(unsigned f) { return (f >= (unsigned)F_SIN && f < (unsigned)F_END) ? (int)f - F_SIN : U_ZERO; }
__host__ __device__ inline int binary_slot(unsigned f) { return (f >= (unsigned)F_ADD && f <= (unsigned)F_GE) ? (int)f - F_ADD : B_ZERO; }

// Upper bound of the operand-stack depth any well-formed row of `len` nodes can need under
// the lowering pass's ordering (lower.cuh).  need() there is: unary = child; binary with two
// non-leaf children = max(max, min + 1); ternary = max(n0, n1 + 1, n2 + 2) over its three
// children sorted by need (a leaf child counts as need 0, so any ternary needs >= 2).
// M(d) = fewest nodes of a subtree needing >= d:  M(1) = M(2) = 4 (ternary of leaves),
// M(d) = min(1 + 2 M(d-1), 1 + 3 M(d-2)):  4, 4, 9, 13, 27, 40, 81, 121, 243, 364, 729, 1093.
__host__ __device__ inline int stack_depth_bound(int len) {
    if (len < 4) return 1;
    int d = 2, m_prev = 4, m_cur = 4;   // M(d-1), M(d)
    for (;;) {
        const int a = 1 + 2 * m_cur, b = 1 + 3 * m_prev;
        const int m_next = a < b ? a : b;
        if (m_next > len) return d;
        m_prev = m_cur;
        m_cur = m_next;
        ++d;
    }
}

// ---------------------------------------------------------------------------
// operator semantics — forward.cu:125-224, compiled like the reference with
// -use_fast_math (setup.py:55) so every intrinsic lowers identically
// (div.approx.ftz, sin/cos.approx, ex2/lg2.approx, tanh.approx, sqrt.approx).
// ---------------------------------------------------------------------------
template <int U>
__device__ __forceinline__ float unary_op(float a) {
    if constexpr (U == F_SIN - F_SIN) return sinf(a);
    else if constexpr (U == F_COS - F_SIN) return cosf(a);
    else if constexpr (U == F_TAN - F_SIN) return tanf(a);
    else if constexpr (U == F_SINH - F_SIN) return sinhf(a);
    else if constexpr (U == F_COSH - F_SIN) return coshf(a);
    else if constexpr (U == F_TANH - F_SIN) return tanhf(a);
    else if constexpr (U == F_LOG - F_SIN) return logf(a);
    else if constexpr (U == F_LOOSE_LOG - F_SIN) return a == 0.0f ? -kMaxVal : logf(fabsf(a));
    else if constexpr (U == F_EXP - F_SIN) return expf(a);
    else if constexpr (U == F_INV - F_SIN) return a == 0.0f ? __int_as_float(0x7fc00000) : 1.0f / a;
    else if constexpr (U == F_LOOSE_INV - F_SIN) {
        if (fabsf(a) <= kDelta) a = copysignf(kDelta, a);
        return 1.0f / a;
    } else if constexpr (U == F_NEG - F_SIN) return -a;
    else if constexpr (U == F_ABS - F_SIN) return fabsf(a);
    else if constexpr (U == F_SQRT - F_SIN) return sqrtf(a);
    else if constexpr (U == F_LOOSE_SQRT - F_SIN) {
        if (a <= 0.0f) a = fabsf(a);
        return sqrtf(a);
    } else return 0.0f;
}

template <int B>
__device__ __forceinline__ float binary_op(float a, float b) {
    if constexpr (B == F_ADD - F_ADD) return __fadd_rn(a, b);   // _rn: never contracted with a neighbour
    else if constexpr (B == F_SUB - F_ADD) return __fsub_rn(a, b);
    else if constexpr (B == F_MUL - F_ADD) return __fmul_rn(a, b);
    else if constexpr (B == F_DIV - F_ADD) return b == 0.0f ? __int_as_float(0x7fc00000) : a / b;
    else if constexpr (B == F_LOOSE_DIV - F_ADD) {
        if (fabsf(b) <= kDelta) b = copysignf(kDelta, b);
        return a / b;
    } else if constexpr (B == F_POW - F_ADD) return powf(a, b);
    else if constexpr (B == F_LOOSE_POW - F_ADD) return (a == 0.0f && b == 0.0f) ? 0.0f : powf(fabsf(a), b);
    else if constexpr (B == F_MAX - F_ADD) return a >= b ? a : b;
    else if constexpr (B == F_MIN - F_ADD) return a <= b ? a : b;
    else if constexpr (B == F_LT - F_ADD) return a < b ? 1.0f : -1.0f;
    else if constexpr (B == F_GT - F_ADD) return a > b ? 1.0f : -1.0f;
    else if constexpr (B == F_LE - F_ADD) return a <= b ? 1.0f : -1.0f;
    else if constexpr (B == F_GE - F_ADD) return a >= b ? 1.0f : -1.0f;
    else return 0.0f;
}

}  // namespace evogp
