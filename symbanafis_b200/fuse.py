"""Bytecode-level stitching: build one multi-output program out of compiled programs + glue ops.

The reference evaluates ``dx, dy, dz`` as separate passes (``api.rs:365-375``) and leaves the Euler /
RK4 algebra between passes to NumPy (``examples/python/aizawa_flow.py:158-165``,
``double_pendulum_benchmark.py:317-365``).  :class:`Stitcher` writes that whole step as ONE program in
the reference ISA: compiled programs are inlined with their registers renamed, and the glue arithmetic
is appended as explicit ``Mul``/``Add`` instructions in the order NumPy would execute them -- so the
fused step performs exactly the same IEEE operations as the reference loop (no FMA is introduced), and
the device lowering (``lower.cpp``) then shares common subexpressions across the inlined copies.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np

from .isa import OPCODE, OPERANDS, Program, assemble

Reg = Tuple[str, int]  # ("p", i) | ("c", i) | ("t", i)


class Stitcher:
    def __init__(self, param_names: Sequence[str]):
        self.param_names = list(param_names)
        self.consts: List[float] = []
        self._cidx: Dict[bytes, int] = {}
        self.n_temps = 0
        self.instrs: List[tuple] = []  # (name, [operands: Reg | int | ("pool", [Reg...])])

    def param(self, i: int) -> Reg:
        return ("p", i)

    def const(self, v: float) -> Reg:
        key = np.float64(v).tobytes()
        i = self._cidx.get(key)
        if i is None:
            i = len(self.consts)
            self.consts.append(float(v))
            self._cidx[key] = i
        return ("c", i)

    def temp(self) -> Reg:
        self.n_temps += 1
        return ("t", self.n_temps - 1)

    def op(self, name: str, *srcs, imm=None) -> Reg:
        """Append ``name dest, srcs...`` with a fresh destination; returns the destination."""
        d = self.temp()
        kinds = OPERANDS[OPCODE[name]]
        ops: list = [d]
        it = iter(srcs)
        for k in kinds[1:]:
            if k == "r":
                ops.append(next(it))
            elif k in "if":
                ops.append(imm)
            else:
                raise ValueError(f"{name} cannot be appended with op()")
        self.instrs.append((name, ops))
        return d

    def call(self, prog: Program, args: Sequence[Reg]) -> List[Reg]:
        """Inline ``prog`` with parameter i bound to ``args[i]``; returns its result registers."""
        if len(args) != prog.param_count:
            raise ValueError("argument count does not match the program's parameter count")
        pc, nc = prog.param_count, len(prog.consts)
        tmap: Dict[int, Reg] = {}

        def src(r: int) -> Reg:
            if r < pc:
                return args[r]
            if r < pc + nc:
                return self.const(float(prog.consts[r - pc]))
            return tmap[r]

        def dst(r: int) -> Reg:
            if r < pc + nc:
                raise ValueError("inlined program writes a parameter/constant register")
            t = self.temp()  # every write gets a fresh register: in-place ops stay correct
            return t

        for ins in prog.instructions():
            name = ins[0]
            kinds = OPERANDS[OPCODE[name]]
            vals = list(ins[1:])
            # read sources first (they may name the register the instruction overwrites)
            ops: list = [None] * len(vals)
            pool_start = None
            for i, (k, v) in enumerate(zip(kinds, vals)):
                if k == "r":
                    ops[i] = src(v)
                elif k == "s":
                    pool_start = v
                elif k == "c":
                    ops[i - 1] = ("pool", [src(int(prog.arg_pool[pool_start + j])) for j in range(v)])
                    ops[i] = v
                elif k in "if":
                    ops[i] = v
            new_d = []
            for i, (k, v) in enumerate(zip(kinds, vals)):
                if k == "d":
                    ops[i] = dst(v)
                    new_d.append((v, ops[i]))
            for v, t in new_d:
                tmap[v] = t
            self.instrs.append((name, ops))
        return [src(r) for r in prog.result_regs]

    def finish(self, results: Sequence[Reg]) -> Program:
        npar, ncon = len(self.param_names), len(self.consts)

        def num(r: Reg) -> int:
            kind, i = r
            return i if kind == "p" else npar + i if kind == "c" else npar + ncon + i

        pool: List[int] = []
        out = []
        for name, ops in self.instrs:
            row = [name]
            for o in ops:
                if isinstance(o, tuple) and o and o[0] == "pool":
                    row.append(len(pool))
                    pool.extend(num(x) for x in o[1])
                elif isinstance(o, tuple):
                    row.append(num(o))
                else:
                    row.append(o)
            out.append(tuple(row))
        # a result that is a bare parameter/constant is legal: result_regs may name any register
        return Program(flat=assemble(out), consts=np.asarray(self.consts), arg_pool=np.asarray(pool, dtype=np.uint32),
                       param_count=npar, workspace_size=npar + ncon + self.n_temps,
                       result_regs=[num(r) for r in results], param_names=self.param_names)


def merge_programs(progs: Sequence[Program], param_names: Sequence[str]) -> Program:
    """k programs over the same parameters -> one k-output program (host-side twin of saf_program_fuse)."""
    st = Stitcher(param_names)
    args = [st.param(i) for i in range(len(param_names))]
    results: List[Reg] = []
    for p in progs:
        results.extend(st.call(p, args[: p.param_count]))
    return st.finish(results)


def euler_step(field: Program, dt: float, param_names: Sequence[str]) -> Program:
    """state + field(state)*dt per component, NumPy order: ``ds*dt`` then ``s + (...)``
    (aizawa_flow.py:163-165)."""
    st = Stitcher(param_names)
    s = [st.param(i) for i in range(len(param_names))]
    ds = st.call(field, s)
    dtc = st.const(dt)
    return st.finish([st.op("Add", s[i], st.op("Mul", ds[i], dtc)) for i in range(len(s))])


def rk4_pendulum_step(accel: Program, dt: float, damping: float) -> Program:
    """One damped RK4 step of the double pendulum, operation for operation the NumPy loop of
    ``double_pendulum_benchmark.py:317-365``.  ``accel`` maps (th1, th2, w1, w2) -> (a1, a2)."""
    names = ["th1", "th2", "w1", "w2"]
    st = Stitcher(names)
    s = [st.param(i) for i in range(4)]
    dt_half, dt_full, dt_sixth = st.const(dt * 0.5), st.const(dt), st.const(dt / 6.0)
    damp = st.const(damping)

    def deriv(state):  # k = [w1, w2, a1, a2]
        a = st.call(accel, state)
        return [state[2], state[3], a[0], a[1]]

    def advance(k, h):  # states + k*h  (np.multiply then np.add)
        return [st.op("Add", s[i], st.op("Mul", k[i], h)) for i in range(4)]

    k1 = deriv(s)
    k2 = deriv(advance(k1, dt_half))
    k3 = deriv(advance(k2, dt_half))
    k4 = deriv(advance(k3, dt_full))
    new = []
    for i in range(4):
        acc = st.op("Add", k1[i], k2[i])
        acc = st.op("Add", acc, k2[i])
        acc = st.op("Add", acc, k3[i])
        acc = st.op("Add", acc, k3[i])
        acc = st.op("Add", acc, k4[i])
        acc = st.op("Mul", acc, dt_sixth)
        new.append(st.op("Add", s[i], acc))
    new[2] = st.op("Mul", new[2], damp)
    new[3] = st.op("Mul", new[3], damp)
    return st.finish(new)
