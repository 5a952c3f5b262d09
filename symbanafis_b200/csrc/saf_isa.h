// saf_isa.h -- reference ISA constants and the DEVICE program encoding (host + device).
//
// Reference ISA (input of the lowering): opcode numbering of
// src/evaluator/logic/bytecode/instruction.rs:295-395, operand order of
// .../compile/emit/assemble.rs:9-103, FnOp ids of .../bytecode/functions.rs:38-117.
//
// Device ISA (output of the lowering, input of the interpreter kernel): a stream of 64-bit words.
// Every instruction has a `lo` word; opcodes that need a third field or an immediate (the FMA
// family, SinCos, Powi, the builtins -- see xop_has_hi) are followed by a `hi` word:
//     lo  [ 7: 0] device opcode         hi  [15: 0] c   third source field / second destination
//         [15: 8] prefix mode (heavy unary)  [63:32] imm FnOp id or i32 immediate
//         [31:16] d   destination field
//         [47:32] a   first source field
//         [63:48] b   second source field
// A field is  index(14) << 2 | kind(2):
//     kind 0 = S[index]  live-value slot of this point in the on-chip register file
//     kind 1 = K[index]  constant table (staged in shared memory, warp-uniform broadcast load)
//     kind 2 = ACC       the accumulator: result of the previous instruction, kept in real registers
//     kind 3 = none
// Every value-producing instruction leaves its result in ACC; if d is an S field it is also stored.
//
// Affine prefix (heavy unary opcodes only): the lowering folds a preceding single-use `Mul(x, K)` or
// `MulAdd(x, K1, K2)` into the transcendental that consumes it -- sin(a*x), sincos(f*x + p),
// exp(-u/w) are the dominant shapes of the reference's workloads -- saving one dispatch each:
//     mode 0: arg = a     mode 1: arg = a * K[b]     mode 2: arg = fma(a, K[b], K[b+1])
// with b the K field in lo[63:48].  The same correctly rounded operation runs, so results are
// unchanged bit for bit.
//
// The hot, cheap opcodes are SPECIALISED BY OPERAND KIND at lowering time (the opcode number says
// where each source comes from), so their handlers contain no operand-kind branches at all; the
// expensive opcodes (transcendentals, division, builtins, the FMA family) decode kinds at run time,
// where three compares vanish against the cost of the operation.
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

// ---- abstract device operations (what the lowering reasons about) ------------------------------
enum DevOp : uint32_t {
    D_END = 0,
    D_LOADK,       // ACC = K[a]                       (materialises a constant operand)
    // light unary: d = f(a)
    D_COPY, D_NEG, D_SQUARE, D_CUBE, D_POW4, D_POW3_2, D_INVPOW3_2, D_INVSQRT, D_INVSQUARE,
    D_INVCUBE, D_RECIP, D_SQRT,
    // heavy unary
    D_SIN, D_COS, D_EXP, D_LN, D_RECIPEXPM1, D_EXPSQR, D_EXPSQRNEG,
    D_EXPNEG,      // Builtin1{ExpNeg}: exp(-a)
    D_SINCOS,      // d = sin(a) (also ACC), field c = destination of cos(a) (S or none)
    D_BUILTIN1,    // d = fn(a), FnOp id in imm
    D_BUILTIN3,    // d = fn(x0, x1, a), x0/x1 staged by the preceding D_ARG2, FnOp id in imm
    // light binary: d = f(a, b)
    D_ADD, D_SUB, D_MUL, D_NEGMUL,
    // heavy binary
    D_DIV, D_POW,
    D_POWI,        // d = powi(a, imm)
    D_BUILTIN2,    // d = fn(a, b), FnOp id in imm
    D_BUILTIN4,    // d = fn(x0, x1, a, b), FnOp id in imm
    D_ARG2,        // stage x0 = a, x1 = b for the next D_BUILTIN3/4; leaves ACC untouched
    // ternary: d = f(a, b, c)
    D_FMA,         // fma(a, b, c)
    D_FMS,         // fma(a, b, -c)
    D_FNMA,        // fma(-a, b, c)
    D_FNMS,        // fma(-a, b, -c)
    D_OPCOUNT
};
constexpr uint32_t D_FIRST_LIGHT_UNARY = D_COPY, D_N_LIGHT_UNARY = D_SQRT - D_COPY + 1;   // 12
constexpr uint32_t D_FIRST_HEAVY_UNARY = D_SIN, D_N_HEAVY_UNARY = D_BUILTIN3 - D_SIN + 1; // 11
constexpr uint32_t D_FIRST_LIGHT_BINARY = D_ADD, D_N_LIGHT_BINARY = D_NEGMUL - D_ADD + 1; // 4
constexpr uint32_t D_FIRST_HEAVY_BINARY = D_DIV, D_N_HEAVY_BINARY = D_ARG2 - D_DIV + 1;   // 6
constexpr uint32_t D_FIRST_TERNARY = D_FMA, D_N_TERNARY = 4;

constexpr uint32_t KIND_S = 0, KIND_K = 1, KIND_ACC = 2, KIND_NONE = 3;
constexpr uint32_t KIND_BITS = 2, KIND_MASK = 3u, INDEX_BITS = 14;
constexpr uint32_t MAX_INDEX = 1u << INDEX_BITS; // slots and constants per program
constexpr uint32_t FIELD_ACC = KIND_ACC, FIELD_NONE = KIND_NONE;
constexpr uint32_t make_field(uint32_t kind, uint32_t index) { return (index << KIND_BITS) | kind; }
constexpr uint32_t field_kind(uint32_t f) { return f & KIND_MASK; }
constexpr uint32_t field_index(uint32_t f) { return (f >> KIND_BITS) & (MAX_INDEX - 1); }

// ---- encoded opcodes (the byte in lo[7:0]) -----------------------------------------------------
//   X_END, X_LOADK
//   X_ULIGHT + 2*u + (a is ACC)                 u = op - D_FIRST_LIGHT_UNARY      (a is S otherwise)
//   X_UHEAVY + h                                 h = op - D_FIRST_HEAVY_UNARY      (kinds at run time)
//   X_BLIGHT + 9*l + 3*kind(a) + kind(b)         l = op - D_FIRST_LIGHT_BINARY     (K,K never emitted)
//   X_BHEAVY + h                                 h = op - D_FIRST_HEAVY_BINARY     (kinds at run time)
//   X_TERN   + t                                 t = op - D_FIRST_TERNARY          (kinds at run time)
constexpr uint32_t X_END = 0, X_LOADK = 1;
constexpr uint32_t X_ULIGHT = 2;
constexpr uint32_t X_UHEAVY = X_ULIGHT + 2 * D_N_LIGHT_UNARY;  // 26
constexpr uint32_t X_BLIGHT = X_UHEAVY + D_N_HEAVY_UNARY;      // 37
constexpr uint32_t X_BHEAVY = X_BLIGHT + 9 * D_N_LIGHT_BINARY; // 73
constexpr uint32_t X_TERN = X_BHEAVY + D_N_HEAVY_BINARY;       // 79
constexpr uint32_t X_OPCOUNT = X_TERN + D_N_TERNARY;           // 83

constexpr uint32_t PREFIX_NONE = 0, PREFIX_MUL = 1, PREFIX_FMA = 2;
constexpr uint64_t make_lo(uint32_t xop, uint32_t d, uint32_t a, uint32_t b, uint32_t prefix = 0) {
    return (uint64_t)xop | ((uint64_t)prefix << 8) | ((uint64_t)d << 16) | ((uint64_t)a << 32) |
           ((uint64_t)b << 48);
}
constexpr uint64_t make_hi(uint32_t c, uint32_t imm) { return (uint64_t)c | ((uint64_t)imm << 32); }

// Decodes an encoded opcode back to (abstract op, kind(a), kind(b)); kinds are KIND_NONE when the
// opcode does not fix them.  Shared by the disassembler and the test emulator description.
inline void decode_xop(uint32_t x, uint32_t *op, uint32_t *ka, uint32_t *kb) {
    *ka = *kb = KIND_NONE;
    if (x == X_END) { *op = D_END; return; }
    if (x == X_LOADK) { *op = D_LOADK; *ka = KIND_K; return; }
    if (x < X_UHEAVY) { *op = D_FIRST_LIGHT_UNARY + (x - X_ULIGHT) / 2; *ka = ((x - X_ULIGHT) & 1) ? KIND_ACC : KIND_S; return; }
    if (x < X_BLIGHT) { *op = D_FIRST_HEAVY_UNARY + (x - X_UHEAVY); return; }
    if (x < X_BHEAVY) {
        uint32_t r = x - X_BLIGHT;
        *op = D_FIRST_LIGHT_BINARY + r / 9;
        *ka = (r % 9) / 3;
        *kb = r % 3;
        return;
    }
    if (x < X_TERN) { *op = D_FIRST_HEAVY_BINARY + (x - X_BHEAVY); return; }
    *op = D_FIRST_TERNARY + (x - X_TERN);
}

// Does the encoded opcode carry a second (hi) word?
inline bool xop_has_hi(uint32_t x) {
    if (x >= X_TERN) return true;
    uint32_t op, ka, kb;
    decode_xop(x, &op, &ka, &kb);
    return op == D_SINCOS || op == D_BUILTIN1 || op == D_BUILTIN2 || op == D_BUILTIN3 || op == D_BUILTIN4 ||
           op == D_POWI;
}

// Launch-time limits of the inline descriptor tables (kernel parameter space).
constexpr int MAX_COLS = 32;
constexpr int MAX_OUTS = 32;
// Threads per block of the interpreter kernel: fixed, so that the slot stride is a compile-time
// constant (one IMAD per shared-memory operand address).
constexpr int BLOCK_THREADS = 128;
// Widest block of the one-block-per-SM layout for programs with many live values (kernel.cu: WIDE).
constexpr int WIDE_MAX_THREADS = 512;

} // namespace saf
