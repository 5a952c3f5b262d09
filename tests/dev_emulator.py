"""CPU emulator of the DEVICE program (test infrastructure).

Executes the 64-bit device words that ``lower.cpp`` produced (exported through
``saf_program_export``) with exactly the kernel's state model -- slots, constant table,
accumulator, ARG2 staging -- while every leaf operation is delegated to the oracle by running
a one-instruction reference program over the batch.  If the lowering (value numbering, chain
expansion, accumulator forwarding, slot allocation) is right, the emulator output is BIT-IDENTICAL
to the oracle running the original reference program.  This checks the host half of the product
on machines without a GPU.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from oracle import OracleProgram
from symbanafis_b200.isa import assemble

# device opcode numbering: symbanafis_b200/csrc/saf_isa.h (enum DevOp)
DEV_OPS = [
    "end", "copy", "neg", "square", "cube", "pow4", "pow3_2", "invpow3_2", "invsqrt", "invsquare",
    "invcube", "recip", "sin", "cos", "exp", "ln", "sqrt", "recipexpm1", "expsqr", "expsqrneg",
    "expneg", "sincos", "builtin1", "builtin3", "add", "sub", "mul", "div", "pow", "negmul", "powi",
    "builtin2", "builtin4", "arg2", "fma", "fms", "fnma", "fnms",
]
D = {n: i for i, n in enumerate(DEV_OPS)}
FIRST_BINARY = D["add"]
FIRST_TERNARY = D["fma"]
KIND_S, KIND_K, KIND_ACC, KIND_NONE = 0, 1, 2, 3

_UNARY_REF = {
    "neg": "Neg", "square": "Square", "cube": "Cube", "pow4": "Pow4", "pow3_2": "Pow3_2",
    "invpow3_2": "InvPow3_2", "invsqrt": "InvSqrt", "invsquare": "InvSquare", "invcube": "InvCube",
    "recip": "Recip", "sin": "Sin", "cos": "Cos", "exp": "Exp", "ln": "Ln", "sqrt": "Sqrt",
    "recipexpm1": "RecipExpm1", "expsqr": "ExpSqr", "expsqrneg": "ExpSqrNeg",
}
_BINARY_REF = {"add": "Add", "sub": "Sub", "mul": "Mul", "div": "Div", "pow": "Pow", "negmul": "NegMul"}
_TERNARY_REF = {"fma": "MulAdd", "fms": "MulSub", "fnma": "NegMulAdd", "fnms": "NegMulSub"}

_cache: Dict[tuple, OracleProgram] = {}


def _leaf(instr: tuple, n_in: int, n_out: int = 1) -> OracleProgram:
    key = (instr, n_in, n_out)
    p = _cache.get(key)
    if p is None:
        results = list(range(n_in, n_in + n_out))
        p = OracleProgram(assemble([instr]), [], [], n_in, n_in + n_out, results)
        _cache[key] = p
    return p


def run_device_program(code: Sequence[int], consts: np.ndarray, param_slot: Sequence[int],
                       result_field: Sequence[int], cols: Sequence[np.ndarray], n_points: int,
                       iters: int = 1) -> List[np.ndarray]:
    n_slots = 1 + max([int(s) for s in param_slot if s != 0xFFFF] +
                      [((int(w) >> sh) & 0xFFF) for w in code for sh in (8, 22, 36, 50)
                       if ((int(w) >> (sh + 12)) & 3) == KIND_S] + [0])
    S = [np.zeros(n_points) for _ in range(n_slots)]
    for j, s in enumerate(param_slot):
        if s == 0xFFFF:
            continue
        if j < len(cols):
            c = np.asarray(cols[j], dtype=np.float64)
            idx = np.minimum(np.arange(n_points), len(c) - 1)
            S[s] = c[idx].copy()
        else:
            S[s] = np.zeros(n_points)
    acc = np.zeros(n_points)
    x0 = x1 = np.zeros(n_points)

    def operand(f: int) -> np.ndarray:
        kind, idx = f >> 12, f & 0xFFF
        if kind == KIND_S:
            return S[idx]
        if kind == KIND_K:
            return np.full(n_points, consts[idx])
        assert kind == KIND_ACC, "operand field of kind none was read"
        return acc

    def result_value(f: int) -> np.ndarray:
        kind, idx = f >> 12, f & 0xFFF
        return S[idx].copy() if kind == KIND_S else np.full(n_points, consts[idx])

    for it in range(iters):
        for w in code:
            w = int(w)
            op = w & 0xFF
            if op == D["end"]:
                break
            fd, fa, fb, fc = (w >> 8) & 0x3FFF, (w >> 22) & 0x3FFF, (w >> 36) & 0x3FFF, (w >> 50) & 0x3FFF
            name = DEV_OPS[op]
            a = operand(fa)
            b = operand(fb) if op >= FIRST_BINARY else None
            c = operand(fc) if op >= FIRST_TERNARY else None
            if name == "copy":
                r = a.copy()
            elif name in _UNARY_REF:
                r = _leaf((_UNARY_REF[name], 1, 0), 1).eval_batch([a], n_points)[0]
            elif name == "expneg":
                r = _leaf(("Builtin1", 1, "exp_neg", 0), 1).eval_batch([a], n_points)[0]
            elif name == "sincos":
                s_, c_ = _leaf(("SinCos", 1, 2, 0), 1, 2).eval_batch([a], n_points)
                r = s_
                if (fc >> 12) == KIND_S:
                    S[fc & 0xFFF] = c_
            elif name == "builtin1":
                r = _leaf(("Builtin1", 1, fb & 0xFFF, 0), 1).eval_batch([a], n_points)[0]
            elif name == "builtin3":
                r = _leaf(("Builtin3", 3, fb & 0xFFF, 0, 1, 2), 3).eval_batch([x0, x1, a], n_points)[0]
            elif name in _BINARY_REF:
                r = _leaf((_BINARY_REF[name], 2, 0, 1), 2).eval_batch([a, b], n_points)[0]
            elif name == "powi":
                n = int(b[0])
                r = _leaf(("Powi", 1, 0, n), 1).eval_batch([a], n_points)[0]
            elif name == "builtin2":
                r = _leaf(("Builtin2", 2, fc & 0xFFF, 0, 1), 2).eval_batch([a, b], n_points)[0]
            elif name == "builtin4":
                r = _leaf(("Builtin4", 4, fc & 0xFFF, 0, 1, 2, 3), 4).eval_batch([x0, x1, a, b], n_points)[0]
            elif name == "arg2":
                x0, x1 = a.copy(), b.copy()
                continue
            elif name in _TERNARY_REF:
                r = _leaf((_TERNARY_REF[name], 3, 0, 1, 2), 3).eval_batch([a, b, c], n_points)[0]
            else:
                raise AssertionError(f"unknown device opcode {op}")
            acc = r
            if (fd >> 12) == KIND_S:
                S[fd & 0xFFF] = r.copy()
        if it + 1 < iters:
            for j, s in enumerate(param_slot):  # sequential, exactly like the kernel's feedback loop
                if s != 0xFFFF:
                    S[s] = result_value(int(result_field[j]))
    return [result_value(int(f)) for f in result_field]
