// uop.h -- the interpreter kernel's PACKED instruction format (host + device).
//
// The device program of saf_isa.h (what saf_program_export returns and tests/dev_emulator.py runs) is
// layout-independent: fields are slot / constant INDICES.  Before a launch the host packs it for one
// register-file layout (points per thread P) into fixed-size 32-byte micro-instructions whose operands
// are ready-to-use BYTE OFFSETS, and whose opcode says where every operand comes from, so that a handler
// is: one integer add per shared-memory operand, one broadcast load per constant operand, the FP64
// instructions, and a sign-tested store.  No operand-kind decoding, shifting or masking happens in the
// hot handlers.
//
//   word 0: [31:0] uop of the NEXT micro-instruction to execute   [63:32] d   destination byte offset in the
//                  register file, < 0 = ACC only
//   word 1: [31:0] a              [63:32] b   S operands: byte offset of the slot in the register file
//   word 2: [31:0] c              [63:32] imm K operands: byte offset of the constant in the image
//   word 3: [31:0] kinds (generic ops: kind(a) | kind(b) << 2 | kind(c) << 4)   [63:32] uop of THIS instruction
// The dispatch loop carries the opcode from one instruction to the next (word 0), so the jump-table lookup of
// instruction i + 1 starts while the operands of instruction i + 1 are still being fetched; word 3 is read only
// when the loop is (re-)entered.  U_END: a = byte offset where the next iteration starts.
//
// Program counters are byte offsets into the CODE part of the image.  The kernel executes the code out of a
// shared-memory window of CODE_WINDOW_BYTES: the code is cut into windows of that size whose last slot holds
// U_NEXTWIN ("stage the next window, continue at its start"), so programs of any length run with on-chip
// instruction fetch and no per-instruction bounds check; programs that fit one window never restage.
//
// Operand kinds S (register-file slot), K (constant table), A (accumulator = previous result), see
// saf_isa.h.  Specialised micro-op families (operand kinds fixed by the opcode):
//   U1   light unary        x {S, A}
//   H1   heavy unary        S only (the packer gives an accumulator operand a slot: handlers that never read
//                           ACC let the compiler keep ACC in place across the dispatch loop)
//                           x {plain, affine prefix arg = fma(a, K[b], K[b+8])}
//   BC   commutative binary (add, mul, negmul), operands ordered S < K < A:   SS SK SA KA AA
//   BN   ordered binary (sub, div):                                 SS SK KS SA AS KA AK AA
//   T    fused multiply-add family, product operands ordered S < K < A, at most one A:
//                                                      SSS SSK SSA SKS SKK SKA SAS SAK KAS KAK
// Everything else (pow, powi, 1/expm1, builtins, operand staging, the rare kind combinations above) is COLD:
// generic micro-ops that decode `kinds` at run time inside one out-of-line function which works memory to
// memory (ACC is exchanged through a reserved register-file slot), so that their register needs and call
// sequences do not weigh on the hot dispatch loop.
#pragma once
#include <stdint.h>

#include "saf_isa.h"

namespace saf {

constexpr int UINSTR_BYTES = 32;
constexpr int UINSTR_WORDS = 4; // 64-bit words
constexpr uint32_t CODE_WINDOW_BYTES = 16 * 1024; // 512 micro-instructions, power of two
constexpr uint32_t CODE_WINDOW_INSTR = CODE_WINDOW_BYTES / UINSTR_BYTES;

constexpr uint32_t U_END = 0, U_LOADK = 1, U_NEXTWIN = 2;
constexpr uint32_t U_U1 = 3;                                  // + 2 * u + (a is A)
constexpr uint32_t U_H1 = U_U1 + 2 * D_N_LIGHT_UNARY;         // + 2 * h + prefixed
constexpr uint32_t H1_SIN = 0, H1_COS = 1, H1_EXP = 2, H1_LN = 3, H1_EXPSQR = 4, H1_EXPSQRNEG = 5, H1_EXPNEG = 6,
                   H1_SINCOS = 7, H1_COUNT = 8;
constexpr uint32_t U_BC = U_H1 + 2 * H1_COUNT;                // + 5 * i + variant
constexpr uint32_t BC_ADD = 0, BC_MUL = 1, BC_NEGMUL = 2, BC_COUNT = 3;
constexpr uint32_t BC_SS = 0, BC_SK = 1, BC_SA = 2, BC_KA = 3, BC_AA = 4;
constexpr uint32_t U_BN = U_BC + 5 * BC_COUNT;                // + 8 * i + variant
constexpr uint32_t BN_SUB = 0, BN_DIV = 1, BN_COUNT = 2;
constexpr uint32_t BN_SS = 0, BN_SK = 1, BN_KS = 2, BN_SA = 3, BN_AS = 4, BN_KA = 5, BN_AK = 6, BN_AA = 7;
constexpr uint32_t U_T = U_BN + 8 * BN_COUNT;                 // + 10 * t + variant
constexpr uint32_t T_FMA = 0, T_FMS = 1, T_FNMA = 2, T_FNMS = 3, T_COUNT = 4;
constexpr uint32_t T_SSS = 0, T_SSK = 1, T_SSA = 2, T_SKS = 3, T_SKK = 4, T_SKA = 5, T_SAS = 6, T_SAK = 7,
                   T_KAS = 8, T_KAK = 9;
// Fused chains (pack.cpp peephole): a multiply or a square whose result is only forwarded through ACC, followed by
// "ACC = K[c] * ACC" and/or "ACC = S[imm] + ACC" -- the shapes `(x * y) * K`, `x * y + S`, `(x * y) * K + S` of every
// term of a long sum.  The same correctly rounded operations run in the same order; one dispatch instead of two or
// three (the dispatch, not the arithmetic, is what a light micro-op costs).  Suffix operands: c = constant, imm = slot.
constexpr uint32_t U_FM = U_T + 10 * T_COUNT;                 // + 3 * BC variant of the multiply + suffix
constexpr uint32_t F_MULK = 0, F_ADDS = 1, F_MULK_ADDS = 2;
constexpr uint32_t U_FQ = U_FM + 3 * 5;                       // + 2 * (square operand is A) + suffix (F_MULK / F_ADDS)
// "t = S[a] + K[b]" whose result is also kept (slot offset in the `kinds` word, < 0 = not kept), followed by its square
// (optionally + S[imm]) or by two constant multiplies K[imm] * (K[c] * t): the (x - c)^2 and -2 (x - c) / w of a Gaussian
// term and of its gradient
constexpr uint32_t U_FS = U_FQ + 2 * 2;
constexpr uint32_t FS_SQ = 0, FS_SQ_ADDS = 1, FS_MULK2 = 2, FS_COUNT = 3;
// sin / cos (plain or with affine prefix) followed by "ACC = S[imm] + ACC" or "ACC = fma(S[imm], K[c], ACC)": the two
// output expressions of a map step such as Clifford's `sin(a y) + c cos(a x)`.  Four-points-per-thread layout only
// (its single dispatch does not replicate handler code; the threaded layouts lost more to the extra heavy handlers
// than they saved in dispatches).
constexpr uint32_t U_FT = U_FS + FS_COUNT;                    // + 4 * {sin, cos} + 2 * {adds, fma} + prefixed
constexpr uint32_t FT_ADDS = 0, FT_FMA = 1, FT_COUNT = 8;
// Two prefixed sin / cos of different slots combined by an add or an fma: `S = trig1(..)` kept only for the following
// U_FT that consumes and overwrites it -- one output expression of a map step (`sin(a y) + c cos(a x)`) in ONE dispatch.
// Fields: a, b = operand and prefix pair of trig1; imm, kinds word = operand and prefix pair of trig2; c = K of the fma.
constexpr uint32_t U_FD = U_FT + FT_COUNT;                    // + 4 * trig1 + 2 * trig2 + {adds, fma}
constexpr uint32_t FD_COUNT = 8;
// `S[d] = S[d] + K[kinds] * (S[imm] * (K[c] * (S[a] * K[b])))`: two fused multiply chains back to back -- one term of a
// gradient sum, `coefficient * (x - c) * exp(..)` accumulated into the running sum that is also its destination (the
// terrain gradient runs 218 of them per point; 1233 -> 1015 micro-ops)
constexpr uint32_t U_F5 = U_FD + FD_COUNT;
constexpr uint32_t U_G = U_F5 + 1;                            // generic micro-ops
constexpr uint32_t G_POW = 0, G_POWI = 1, G_BUILTIN1 = 2, G_BUILTIN2 = 3, G_BUILTIN3 = 4, G_BUILTIN4 = 5,
                   G_ARG2 = 6, G_FMA = 7, G_FMS = 8, G_FNMA = 9, G_FNMS = 10, G_DIV = 11, G_RECIPEXPM1 = 12, G_COUNT = 13;
constexpr uint32_t NO_PREFIX = 0xffffffffu; // field b of G_RECIPEXPM1 when there is no affine prefix
constexpr uint32_t U_COUNT = U_G + G_COUNT;

constexpr int32_t NO_DEST = -1;

} // namespace saf
