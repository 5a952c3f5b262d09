// ruop.h -- micro-ops of the REGISTER machine (host + device): the interpreter layout whose live values sit in the
// real register file of the SM and spill to shared memory.
//
// The shared-memory machine of uop.h pays a 16-byte shared-memory access per operand and pair of points; programs
// made of plain arithmetic (Euler/RK4 steps, polynomial terrain terms) are bound by that traffic, not by the FP64
// pipe.  An interpreter cannot index registers at run time -- so the register NUMBERS are part of the opcode: the
// dispatch loop keeps N_R named registers R0..R7 (4 points each) alive across all handlers, and there is one handler
// per (operation, destination register, operand register) combination.  Handlers are two-address:
//       Rd = Rd op X            X = R0..R7 | K (constant table) | S (shared-memory slot: spilled values, parameters)
// so a family costs 8 x 10 handlers instead of 8^3.  pack_r.cpp allocates registers over the device program
// (furthest-next-use eviction, in-place reuse of dying operands, MOVs otherwise).
//
// The micro-instruction format is uop.h's: 32 bytes,
//   word 0: [31:0] uop of the NEXT micro-instruction   [63:32] d  (byte offset of a slot: cold micro-ops only)
//   word 1: [31:0] a   [63:32] b      K / S operands as byte offsets (constant table / register file in smem)
//   word 2: [31:0] c   [63:32] imm
//   word 3: [31:0] kinds | cold micro-op << 8          [63:32] uop of THIS instruction
//
// Families (d = destination register, xk = operand kind 0..7 register, XK_K constant at a, XK_S slot at a):
//   RU_END              end of one iteration: a = code offset of the loop body; falls through after the last one
//   RU_EXIT             end of the program
//   RU_NEXTWIN          stage the next code window
//   RU_COLD             cold micro-op (pow, powi, builtins, ...): leaves the loop; operands and result in slots
//   RU_MOV  + d*10 + xk                 Rd = X
//   RU_ST   + d                         S[a] = Rd
//   RU_BIN  + (op*8 + d)*10 + xk        Rd = Rd op X       op: ADD MUL SUB RSUB(X - Rd) NEGMUL
//   RU_UN   + u*8 + d                   Rd = f(Rd)         u: NEG CUBE POW4            (square = MUL d, d)
//   RU_FRR  + s*36 + pair(a<=b)         R0 = +-(Ra * Rb) +- R0      s: FMA FMS FNMA FNMS (saf_isa.h)
//   RU_FKA  + s*8 + ra                  R0 = +-(Ra * K[a]) +- R0
//   RU_FKM  + s*10 + xk                 R0 = +-(R0 * K[a]) +- X[b]
//                                       (the fused family works on R0 only: per-register variants of it made the jump
//                                       table 9 KB, and its lookups missed the constant cache on every dispatch)
//   RU_H0   + 2*h + prefixed            R1 = f(R0)  (h: uop.h H1_*; sincos also writes R2 = cos);  prefix pair at b
//                                       (argument in R0, where the fused family and division leave their results)
//   RU_U0   + u                         R0 = f(R0)         u: POW3_2 INVPOW3_2 INVSQRT INVSQUARE INVCUBE RECIP SQRT
//   RU_DIV  + xk                        R0 = R0 / X
//   RU_RDIV + xk                        R0 = X / R0
#pragma once
#include <stdint.h>

#include "uop.h"

namespace saf {

constexpr int R_NREG = 8;
constexpr int R_POINTS = 4; // points per thread of the register machine
constexpr uint32_t XK_K = 8, XK_S = 9, XK_COUNT = 10;

constexpr uint32_t RU_END = 0, RU_EXIT = 1, RU_NEXTWIN = 2, RU_COLD = 3;
constexpr uint32_t RU_MOV = 4;                                    // + d*10 + xk
constexpr uint32_t RU_ST = RU_MOV + R_NREG * XK_COUNT;            // + d
constexpr uint32_t RB_ADD = 0, RB_MUL = 1, RB_SUB = 2, RB_RSUB = 3, RB_NEGMUL = 4, RB_COUNT = 5;
constexpr uint32_t RU_BIN = RU_ST + R_NREG;                       // + (op*8 + d)*10 + xk
constexpr uint32_t RUN_NEG = 0, RUN_CUBE = 1, RUN_POW4 = 2, RUN_COUNT = 3;
constexpr uint32_t RU_UN = RU_BIN + RB_COUNT * R_NREG * XK_COUNT; // + u*8 + d
constexpr uint32_t R_NPAIR = R_NREG * (R_NREG + 1) / 2;           // 36 unordered register pairs
constexpr uint32_t RU_FRR = RU_UN + RUN_COUNT * R_NREG;           // + s*36 + pair
constexpr uint32_t RU_FKA = RU_FRR + 4 * R_NPAIR;                 // + s*8 + ra
constexpr uint32_t RU_FKM = RU_FKA + 4 * R_NREG;                  // + s*10 + xk
constexpr uint32_t RU_H0 = RU_FKM + 4 * XK_COUNT;                 // + 2*h + prefixed
constexpr uint32_t RU0_POW3_2 = 0, RU0_INVPOW3_2 = 1, RU0_INVSQRT = 2, RU0_INVSQUARE = 3, RU0_INVCUBE = 4,
                   RU0_RECIP = 5, RU0_SQRT = 6, RU0_COUNT = 7;
constexpr uint32_t RU_U0 = RU_H0 + 2 * H1_COUNT;                  // + u
constexpr uint32_t RU_DIV = RU_U0 + RU0_COUNT;                    // + xk
constexpr uint32_t RU_RDIV = RU_DIV + XK_COUNT;                   // + xk
constexpr uint32_t RU_COUNT = RU_RDIV + XK_COUNT;

// index of the unordered pair (a <= b) in 0..35
constexpr uint32_t r_pair(uint32_t a, uint32_t b) { return a * R_NREG - a * (a - 1) / 2 + (b - a); }

} // namespace saf
