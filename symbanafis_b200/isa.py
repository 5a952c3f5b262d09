"""The SymbAnaFis evaluator ISA: opcodes, FnOp table, flat encoding, Program container.

Mirrors (not copies) the reference's wire format so that a program produced by the
reference compiler can be handed to this package unchanged:

* opcode numbering / operand order: ``src/evaluator/logic/bytecode/instruction.rs:295-395``
* flat ``u32`` encoding and trailing End sentinel: ``.../compile/emit/assemble.rs:9-103``
* FnOp numbering: ``.../bytecode/functions.rs:38-117``
* program payload (``flat_bytecode``, ``constants``, ``arg_pool``, ``workspace_size``,
  ``param_count``, ``result_reg``): ``src/evaluator/api.rs:125-142``

Register file layout per point (``instruction.rs:18-21``): ``R[0..param_count)`` inputs,
then the constant pool, then temporaries.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterable, List, Sequence, Tuple

import numpy as np

# name -> (opcode, operand kinds).  kinds: d = dest reg, r = source reg, s = pool start,
# c = pool count, i = i32 immediate, f = FnOp id.
_ISA: List[Tuple[str, str]] = [
    ("End", ""),
    ("Copy", "dr"),
    ("Neg", "dr"),
    ("SinCos", "ddr"),
    ("Add", "drr"),
    ("Add3", "drrr"),
    ("Add4", "drrrr"),
    ("AddN", "dsc"),
    ("Mul", "drr"),
    ("Mul3", "drrr"),
    ("Mul4", "drrrr"),
    ("MulN", "dsc"),
    ("Sub", "drr"),
    ("Div", "drr"),
    ("Pow", "drr"),
    ("MulAdd", "drrr"),
    ("MulSub", "drrr"),
    ("NegMul", "drr"),
    ("NegMulAdd", "drrr"),
    ("NegMulSub", "drrr"),
    ("Square", "dr"),
    ("Cube", "dr"),
    ("Pow4", "dr"),
    ("Pow3_2", "dr"),
    ("InvPow3_2", "dr"),
    ("InvSqrt", "dr"),
    ("InvSquare", "dr"),
    ("InvCube", "dr"),
    ("Recip", "dr"),
    ("Powi", "dri"),
    ("Sin", "dr"),
    ("Cos", "dr"),
    ("Exp", "dr"),
    ("Ln", "dr"),
    ("Sqrt", "dr"),
    ("RecipExpm1", "dr"),
    ("ExpSqr", "dr"),
    ("ExpSqrNeg", "dr"),
    ("Builtin1", "dfr"),
    ("Builtin2", "dfrr"),
    ("Builtin3", "dfrrr"),
    ("Builtin4", "dfrrrr"),
]

OPCODE = {name: i for i, (name, _) in enumerate(_ISA)}
OPNAME = [name for name, _ in _ISA]
OPERANDS = [kinds for _, kinds in _ISA]

# FnOp: (name, arity); index = numeric id
_FNOPS: List[Tuple[str, int]] = [
    ("sin", 1), ("cos", 1), ("tan", 1), ("cot", 1), ("sec", 1), ("csc", 1),
    ("asin", 1), ("acos", 1), ("atan", 1), ("acot", 1), ("asec", 1), ("acsc", 1),
    ("sinh", 1), ("cosh", 1), ("tanh", 1), ("coth", 1), ("sech", 1), ("csch", 1),
    ("asinh", 1), ("acosh", 1), ("atanh", 1), ("acoth", 1), ("acsch", 1), ("asech", 1),
    ("exp", 1), ("expm1", 1), ("exp_neg", 1), ("ln", 1), ("log1p", 1),
    ("sqrt", 1), ("cbrt", 1),
    ("abs", 1), ("signum", 1), ("floor", 1), ("ceil", 1), ("round", 1),
    ("erf", 1), ("erfc", 1), ("gamma", 1), ("lgamma", 1), ("digamma", 1), ("trigamma", 1),
    ("tetragamma", 1), ("sinc", 1), ("lambert_w", 1), ("elliptic_k", 1), ("elliptic_e", 1),
    ("zeta", 1), ("exp_polar", 1),
    ("atan2", 2), ("log", 2), ("bessel_j", 2), ("bessel_y", 2), ("bessel_i", 2), ("bessel_k", 2),
    ("polygamma", 2), ("beta", 2), ("zeta_deriv", 2), ("hermite", 2),
    ("assoc_legendre", 3), ("spherical_harmonic", 4),
]
FNOP = {name: i for i, (name, _) in enumerate(_FNOPS)}
FNOP_NAME = [name for name, _ in _FNOPS]
FNOP_ARITY = [a for _, a in _FNOPS]

Instr = Tuple  # ("Add", d, a, b) -- name followed by operand ints


class ProgramError(ValueError):
    """Malformed program (the reference's validate_program errors, helper.rs:98-196)."""


@dataclass
class Program:
    """Payload of a compiled evaluator (``api.rs:125-142``), generalised to k result registers."""

    flat: np.ndarray                 # uint32, End-terminated
    consts: np.ndarray               # float64
    arg_pool: np.ndarray             # uint32
    param_count: int
    workspace_size: int
    result_regs: List[int]
    param_names: List[str] = field(default_factory=list)

    def __post_init__(self) -> None:
        self.flat = np.ascontiguousarray(self.flat, dtype=np.uint32)
        self.consts = np.ascontiguousarray(self.consts, dtype=np.float64)
        self.arg_pool = np.ascontiguousarray(self.arg_pool, dtype=np.uint32)
        self.result_regs = [int(r) for r in self.result_regs]

    @property
    def result_reg(self) -> int:
        return self.result_regs[0]

    def instructions(self) -> List[Instr]:
        return decode(self.flat)

    def instruction_count(self) -> int:
        return len(self.instructions())

    def disassemble(self) -> str:
        """Readable listing in the spirit of ``CompiledEvaluator::disassemble`` (api.rs:190-267)."""
        out = ["=== Bytecode Disassembly ===",
               f"Parameters ({self.param_count}): [{', '.join(self.param_names)}]"]
        if len(self.consts):
            out.append(f"\n=== Constants ({len(self.consts)}) ===")
            for i, c in enumerate(self.consts):
                out.append(f"  R{self.param_count + i:<4} = {float(c)!r}")
        out.append("\n=== Memory Layout ===")
        out.append(f"Workspace Size: {self.workspace_size} registers")
        out.append(f"Arg Pool Size:  {len(self.arg_pool)} entries")
        out.append(f"Result registers: {self.result_regs}")
        instrs = self.instructions()
        out.append(f"\n=== Instructions ({len(instrs)}) ===")
        counts: dict = {}
        for i, ins in enumerate(instrs):
            counts[ins[0]] = counts.get(ins[0], 0) + 1
            out.append(f"  [{i:04}] {format_instr(ins, self.arg_pool)}")
        if instrs:
            out.append("\n=== Instruction Summary ===")
            for name, cnt in sorted(counts.items(), key=lambda kv: -kv[1]):
                out.append(f"  {name:<15}: {cnt:>6} ({100.0 * cnt / len(instrs):>5.1f}%)")
        return "\n".join(out) + "\n"


def assemble(instrs: Iterable[Instr]) -> np.ndarray:
    """``assemble_flat_bytecode`` (assemble.rs:9-103): dense u32 words + End sentinel."""
    words: List[int] = []
    for ins in instrs:
        name = ins[0]
        op = OPCODE[name]
        kinds = OPERANDS[op]
        if len(ins) - 1 != len(kinds):
            raise ProgramError(f"{name} takes {len(kinds)} operands, got {len(ins) - 1}")
        words.append(op)
        for kind, v in zip(kinds, ins[1:]):
            if kind == "f" and isinstance(v, str):
                v = FNOP[v]
            words.append(int(v) & 0xFFFFFFFF)
    words.append(0)
    return np.asarray(words, dtype=np.uint32)


def decode(flat: Sequence[int]) -> List[Instr]:
    """Inverse of :func:`assemble`; stops at the first End."""
    out: List[Instr] = []
    pc, n = 0, len(flat)
    while True:
        if pc >= n:
            raise ProgramError("bytecode stream not terminated by End")
        op = int(flat[pc])
        if op == 0:
            return out
        if op >= len(_ISA):
            raise ProgramError(f"unknown opcode {op} at word {pc}")
        kinds = OPERANDS[op]
        if pc + 1 + len(kinds) > n:
            raise ProgramError(f"truncated instruction at word {pc}")
        vals = []
        for kind, w in zip(kinds, flat[pc + 1: pc + 1 + len(kinds)]):
            w = int(w)
            if kind == "i" and w >= 1 << 31:
                w -= 1 << 32
            vals.append(w)
        out.append((OPNAME[op], *vals))
        pc += 1 + len(kinds)


def format_instr(ins: Instr, arg_pool: Sequence[int] = ()) -> str:
    name = ins[0]
    kinds = OPERANDS[OPCODE[name]]
    parts = []
    for kind, v in zip(kinds, ins[1:]):
        if kind in "dr":
            parts.append(f"R{v}")
        elif kind == "f":
            parts.append(FNOP_NAME[v & 0xFF])
        elif kind == "s":
            parts.append(f"pool[{v}..")
        elif kind == "c":
            parts[-1] += f"+{v})"
        else:
            parts.append(str(v))
    return f"{name} " + ", ".join(parts)


def reads_writes(ins: Instr, arg_pool: Sequence[int]) -> Tuple[List[int], List[int]]:
    """Registers read / written by an instruction (``for_each_read``/``for_each_write``)."""
    name = ins[0]
    kinds = OPERANDS[OPCODE[name]]
    reads: List[int] = []
    writes: List[int] = []
    start = None
    for kind, v in zip(kinds, ins[1:]):
        if kind == "d":
            writes.append(v)
        elif kind == "r":
            reads.append(v)
        elif kind == "s":
            start = v
        elif kind == "c":
            reads.extend(int(arg_pool[start + i]) for i in range(v))
    return reads, writes
