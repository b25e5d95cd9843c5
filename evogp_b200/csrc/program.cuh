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
//   header [7:0]   opcode = form * 16 + op      (form and op tables below)
//          [8]     PUSH    spill acc to the operand stack before executing; only ever set on
//                          instructions that start a fresh value (LOAD_*, U?V/U?K, VV, VK, KV)
//          [11]    OUT     multi-output programs only: add the result to outs[idxB]
//                          (idxB == 0x3FF: output index out of range, result dropped)
//          [21:12] idxA    variable index of the (first) variable operand; C_IF: operand permutation
//          [31:22] idxB    variable index of the second variable operand (VV) / output index when OUT
//   constant       the constant operand of a *K* form, bit-cast
// Operand kinds are part of the opcode, so the replay loop never tests a flag to find an operand:
//   A = accumulator, V = variable (dataset column), K = constant, S = pop from the operand stack.
// ---------------------------------------------------------------------------
constexpr uint32_t I_PUSH = 1u << 8, I_OUT = 1u << 11;
// C_IF3 (multi-output, three leaf operands) flags which of a, b, c are constants
constexpr uint32_t I_IF3_BCONST = 1u << 8, I_IF3_ACONST = 1u << 9, I_IF3_CCONST = 1u << 10;
constexpr int I_IDXA_SHIFT = 12, I_IDXB_SHIFT = 22;
constexpr uint32_t I_IDX_MASK = 0x3FFu;

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
    FM_SA = 11,  // acc = b(pop, acc)
    FM_AS = 12,  // acc = b(acc, pop)
    FM_COUNT = 13
};
// FM_MISC opcodes
enum : int {
    C_END = 0,     // stop
    C_LOAD_V = 1,  // acc = var A
    C_LOAD_K = 2,  // acc = const
    C_NAN = 3,     // malformed row: result NaN
    C_IF = 4,      // acc = a > 0 ? b : c; operands are acc / stack top / stack top-1 per idxA
    C_IF3 = 5,     // multi-output only, 2 slots {hdr, a}{b, c}: r = a > 0 ? b : c on three leaf operands
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
