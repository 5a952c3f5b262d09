// saf_isa.h -- reference ISA constants and the DEVICE program encoding (host + device).
//
// Reference ISA (input of the lowering): opcode numbering of
// src/evaluator/logic/bytecode/instruction.rs:295-395, operand order of
// .../compile/emit/assemble.rs:9-103, FnOp ids of .../bytecode/functions.rs:38-117.
//
// Device ISA (output of the lowering, input of the interpreter kernel): fixed-width 64-bit words
//     [ 7: 0] device opcode
//     [21: 8] d   destination field
//     [35:22] a   first source field
//     [49:36] b   second source field   (or FnOp id / second destination, see each opcode)
//     [63:50] c   third source field    (or FnOp id / second destination)
// A field is  kind(2) << 12 | index(12):
//     kind 0 = S[index]  live-value slot of this point in the on-chip register file
//     kind 1 = K[index]  constant table (constant bank / global memory, warp-uniform load)
//     kind 2 = ACC       the accumulator: result of the previous instruction, kept in real registers
//     kind 3 = none
// Every instruction leaves its result in ACC; if d is an S field it is also stored to that slot.
#pragma once
#include <stdint.h>

namespace saf {

// ---- reference opcodes ------------------------------------------------------------------------
enum RefOp : uint32_t {
    R_END = 0, R_COPY, R_NEG, R_SINCOS, R_ADD, R_ADD3, R_ADD4, R_ADDN, R_MUL, R_MUL3, R_MUL4, R_MULN,
    R_SUB, R_DIV, R_POW, R_MULADD, R_MULSUB, R_NEGMUL, R_NEGMULADD, R_NEGMULSUB, R_SQUARE, R_CUBE,
    R_POW4, R_POW3_2, R_INVPOW3_2, R_INVSQRT, R_INVSQUARE, R_INVCUBE, R_RECIP, R_POWI, R_SIN, R_COS,
    R_EXP, R_LN, R_SQRT, R_RECIPEXPM1, R_EXPSQR, R_EXPSQRNEG, R_BUILTIN1, R_BUILTIN2, R_BUILTIN3,
    R_BUILTIN4, R_OPCOUNT
};

// ---- FnOp ids ---------------------------------------------------------------------------------
enum FnOp : uint32_t {
    FN_SIN = 0, FN_COS, FN_TAN, FN_COT, FN_SEC, FN_CSC, FN_ASIN, FN_ACOS, FN_ATAN, FN_ACOT, FN_ASEC,
    FN_ACSC, FN_SINH, FN_COSH, FN_TANH, FN_COTH, FN_SECH, FN_CSCH, FN_ASINH, FN_ACOSH, FN_ATANH,
    FN_ACOTH, FN_ACSCH, FN_ASECH, FN_EXP, FN_EXPM1, FN_EXPNEG, FN_LN, FN_LOG1P, FN_SQRT, FN_CBRT,
    FN_ABS, FN_SIGNUM, FN_FLOOR, FN_CEIL, FN_ROUND, FN_ERF, FN_ERFC, FN_GAMMA, FN_LGAMMA, FN_DIGAMMA,
    FN_TRIGAMMA, FN_TETRAGAMMA, FN_SINC, FN_LAMBERTW, FN_ELLIPTICK, FN_ELLIPTICE, FN_ZETA,
    FN_EXPPOLAR, FN_ATAN2, FN_LOG, FN_BESSELJ, FN_BESSELY, FN_BESSELI, FN_BESSELK, FN_POLYGAMMA,
    FN_BETA, FN_ZETADERIV, FN_HERMITE, FN_ASSOCLEGENDRE, FN_SPHERICALHARMONIC, FN_COUNT
};

// ---- device opcodes ---------------------------------------------------------------------------
// Ordered by source arity so the interpreter fetches operands with two compares:
//   op <  D_FIRST_BINARY : one source  (a)
//   op <  D_FIRST_TERNARY: two sources (a, b)
//   else                 : three       (a, b, c)
enum DevOp : uint32_t {
    D_END = 0,
    // unary: d = f(a)
    D_COPY, D_NEG, D_SQUARE, D_CUBE, D_POW4, D_POW3_2, D_INVPOW3_2, D_INVSQRT, D_INVSQUARE,
    D_INVCUBE, D_RECIP, D_SIN, D_COS, D_EXP, D_LN, D_SQRT, D_RECIPEXPM1, D_EXPSQR, D_EXPSQRNEG,
    D_EXPNEG,      // Builtin1{ExpNeg}: exp(-a), the one builtin hot enough for its own opcode
    D_SINCOS,      // d = sin(a) (also ACC), field c = destination of cos(a) (S or none)
    D_BUILTIN1,    // d = fn(a), FnOp id in the index bits of field b
    D_BUILTIN3,    // d = fn(x0, x1, a), x0/x1 staged by the preceding D_ARG2, FnOp id in field b
    // binary: d = f(a, b)
    D_ADD, D_SUB, D_MUL, D_DIV, D_POW, D_NEGMUL,
    D_POWI,        // d = powi(a, (int)b)  -- b is a K field holding n as a double
    D_BUILTIN2,    // d = fn(a, b), FnOp id in the index bits of field c
    D_BUILTIN4,    // d = fn(x0, x1, a, b), FnOp id in field c
    D_ARG2,        // stage x0 = a, x1 = b for the next D_BUILTIN3/4; leaves ACC untouched
    // ternary: d = f(a, b, c)
    D_FMA,         // fma(a, b, c)
    D_FMS,         // fma(a, b, -c)
    D_FNMA,        // fma(-a, b, c)
    D_FNMS,        // fma(-a, b, -c)
    D_OPCOUNT
};
constexpr uint32_t D_FIRST_BINARY = D_ADD;
constexpr uint32_t D_FIRST_TERNARY = D_FMA;

constexpr uint32_t KIND_S = 0, KIND_K = 1, KIND_ACC = 2, KIND_NONE = 3;
constexpr uint32_t FIELD_BITS = 14, INDEX_BITS = 12, INDEX_MASK = 0xfffu, FIELD_MASK = 0x3fffu;
constexpr uint32_t FIELD_ACC = KIND_ACC << INDEX_BITS;
constexpr uint32_t FIELD_NONE = KIND_NONE << INDEX_BITS;
constexpr uint32_t MAX_INDEX = 1u << INDEX_BITS; // slots and constants per program

constexpr uint32_t make_field(uint32_t kind, uint32_t index) { return (kind << INDEX_BITS) | index; }
constexpr uint64_t make_word(uint32_t op, uint32_t d, uint32_t a, uint32_t b, uint32_t c) {
    return (uint64_t)op | ((uint64_t)d << 8) | ((uint64_t)a << 22) | ((uint64_t)b << 36) |
           ((uint64_t)c << 50);
}

// Launch-time limits of the inline descriptor tables (kernel parameter space).
constexpr int MAX_COLS = 32;
constexpr int MAX_OUTS = 32;

} // namespace saf
