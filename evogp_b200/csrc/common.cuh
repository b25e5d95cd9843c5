// common.cuh — shared definitions of the sm_100a EvoGP kernels.
// Enumerations mirror the reference's numeric contract (src/evogp/cuda/defs.h:5-57);
// everything else is this library's own.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/evogp_b200.h"

namespace evogp {

constexpr int kMaxStack = EVOGP_MAX_STACK;
constexpr int kMaxFullDepth = EVOGP_MAX_FULL_DEPTH;
constexpr float kDelta = 1e-9f;   // defs.h:7
constexpr float kMaxVal = 1e9f;   // defs.h:8

// node_type low 7 bits (defs.h:10-22)
enum : int { NT_VAR = 0, NT_CONST = 1, NT_UFUNC = 2, NT_BFUNC = 3, NT_TFUNC = 4, NT_MASK = 0x7F, NT_OUT = 0x80 };

// function ids (defs.h:24-57)
enum : int {
    F_IF = 0,
    F_ADD = 1, F_SUB, F_MUL, F_DIV, F_LOOSE_DIV, F_POW, F_LOOSE_POW, F_MAX, F_MIN, F_LT, F_GT, F_LE, F_GE,
    F_SIN = 14, F_COS, F_TAN, F_SINH, F_COSH, F_TANH, F_LOG, F_LOOSE_LOG, F_EXP, F_INV, F_LOOSE_INV, F_NEG, F_ABS,
    F_SQRT, F_LOOSE_SQRT,
    F_END = 29
};

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int check_launch(const char *what);   // cudaGetLastError -> status
int ensure_device_ok();

}  // namespace evogp

#define EVOGP_REQUIRE(cond, ...)                      \
    do {                                              \
        if (!(cond)) {                                \
            evogp::set_error(__VA_ARGS__);            \
            return EVOGP_ERR_ARG;                     \
        }                                             \
    } while (0)

#define EVOGP_CUDA(call)                                                                       \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            evogp::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return EVOGP_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)
