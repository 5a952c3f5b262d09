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

# device encoding: symbanafis_b200/csrc/saf_isa.h (enum DevOp, X_* opcode layout, field = index << 2 | kind)
DEV_OPS = [
    "end", "loadk", "copy", "neg", "square", "cube", "pow4", "pow3_2", "invpow3_2", "invsqrt", "invsquare",
    "invcube", "recip", "sqrt", "sin", "cos", "exp", "ln", "recipexpm1", "expsqr", "expsqrneg", "expneg",
    "sincos", "builtin1", "builtin3", "add", "sub", "mul", "negmul", "div", "pow", "powi", "builtin2",
    "builtin4", "arg2", "fma", "fms", "fnma", "fnms",
]
D = {n: i for i, n in enumerate(DEV_OPS)}
KIND_S, KIND_K, KIND_ACC, KIND_NONE = 0, 1, 2, 3
N_LIGHT_UNARY, N_HEAVY_UNARY, N_LIGHT_BINARY, N_HEAVY_BINARY = 12, 11, 4, 6
X_END, X_LOADK, X_ULIGHT = 0, 1, 2
X_UHEAVY = X_ULIGHT + 2 * N_LIGHT_UNARY
X_BLIGHT = X_UHEAVY + N_HEAVY_UNARY
X_BHEAVY = X_BLIGHT + 9 * N_LIGHT_BINARY
X_TERN = X_BHEAVY + N_HEAVY_BINARY
HAS_HI = {"sincos", "builtin1", "builtin2", "builtin3", "builtin4", "powi", "fma", "fms", "fnma", "fnms"}


def decode_xop(x: int):
    """encoded opcode -> (name, kind(a) or None, kind(b) or None)"""
    if x == X_END:
        return "end", None, None
    if x == X_LOADK:
        return "loadk", KIND_K, None
    if x < X_UHEAVY:
        r = x - X_ULIGHT
        return DEV_OPS[D["copy"] + r // 2], (KIND_ACC if r & 1 else KIND_S), None
    if x < X_BLIGHT:
        return DEV_OPS[D["sin"] + x - X_UHEAVY], None, None
    if x < X_BHEAVY:
        r = x - X_BLIGHT
        return DEV_OPS[D["add"] + r // 9], (r % 9) // 3, r % 3
    if x < X_TERN:
        return DEV_OPS[D["div"] + x - X_BHEAVY], None, None
    return DEV_OPS[D["fma"] + x - X_TERN], None, None


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
    code = [int(w) for w in code]
    S: Dict[int, np.ndarray] = {}
    for j, s in enumerate(param_slot):
        if s == 0xFFFF:
            continue
        if j < len(cols):
            c = np.asarray(cols[j], dtype=np.float64)
            idx = np.minimum(np.arange(n_points), len(c) - 1)
            S[int(s)] = c[idx].copy()
        else:
            S[int(s)] = np.zeros(n_points)
    acc = np.zeros(n_points)
    x0 = x1 = np.zeros(n_points)

    def operand(f: int, forced_kind=None) -> np.ndarray:
        kind, idx = f & 3, f >> 2
        if forced_kind is not None:
            assert kind == forced_kind, "opcode specialisation disagrees with the operand field"
        if kind == KIND_S:
            return S[idx]  # KeyError = read of a slot that was never written
        if kind == KIND_K:
            return np.full(n_points, consts[idx])
        assert kind == KIND_ACC, "operand field of kind none was read"
        return acc

    def unary_arg(fa: int, fb: int, prefix: int) -> np.ndarray:
        """operand of a heavy unary opcode with its affine prefix (saf_isa.h): a, a*K[b] or fma(a,K[b],K[b+1])"""
        a = operand(fa)
        if prefix == 0:
            return a
        assert (fb & 3) == KIND_K
        k1 = np.full(n_points, consts[fb >> 2])
        if prefix == 1:
            return _leaf(("Mul", 2, 0, 1), 2).eval_batch([a, k1], n_points)[0]
        assert prefix == 2
        k2 = np.full(n_points, consts[(fb >> 2) + 1])
        return _leaf(("MulAdd", 3, 0, 1, 2), 3).eval_batch([a, k1, k2], n_points)[0]

    def result_value(f: int) -> np.ndarray:
        kind, idx = f & 3, f >> 2
        return S[idx].copy() if kind == KIND_S else np.full(n_points, consts[idx])

    for it in range(iters):
        pc = 0
        while True:
            lo = code[pc]
            pc += 1
            name, ka, kb = decode_xop(lo & 0xFF)
            if name == "end":
                break
            fd, fa, fb = (lo >> 16) & 0xFFFF, (lo >> 32) & 0xFFFF, (lo >> 48) & 0xFFFF
            prefix = (lo >> 8) & 0xFF
            fc, imm = KIND_NONE, 0
            if name in HAS_HI:
                hi = code[pc]
                pc += 1
                fc, imm = hi & 0xFFFF, (hi >> 32) & 0xFFFFFFFF
            if name == "loadk":
                r = operand(fa, KIND_K).copy()
            elif name == "copy":
                r = operand(fa, ka).copy()
            elif name in _UNARY_REF:
                arg = operand(fa, ka) if ka is not None else unary_arg(fa, fb, prefix)
                r = _leaf((_UNARY_REF[name], 1, 0), 1).eval_batch([arg], n_points)[0]
            elif name == "expneg":
                r = _leaf(("Builtin1", 1, "exp_neg", 0), 1).eval_batch([unary_arg(fa, fb, prefix)], n_points)[0]
            elif name == "sincos":
                s_, c_ = _leaf(("SinCos", 1, 2, 0), 1, 2).eval_batch([unary_arg(fa, fb, prefix)], n_points)
                r = s_
                if (fc & 3) == KIND_S:
                    S[fc >> 2] = c_
            elif name == "builtin1":
                r = _leaf(("Builtin1", 1, imm, 0), 1).eval_batch([operand(fa)], n_points)[0]
            elif name == "builtin3":
                r = _leaf(("Builtin3", 3, imm, 0, 1, 2), 3).eval_batch([x0, x1, operand(fa)], n_points)[0]
            elif name in _BINARY_REF:
                a, b = operand(fa, ka), operand(fb, kb)
                r = _leaf((_BINARY_REF[name], 2, 0, 1), 2).eval_batch([a, b], n_points)[0]
            elif name == "powi":
                n = imm - (1 << 32) if imm >= (1 << 31) else imm
                r = _leaf(("Powi", 1, 0, n), 1).eval_batch([operand(fa)], n_points)[0]
            elif name == "builtin2":
                r = _leaf(("Builtin2", 2, imm, 0, 1), 2).eval_batch([operand(fa), operand(fb)], n_points)[0]
            elif name == "builtin4":
                r = _leaf(("Builtin4", 4, imm, 0, 1, 2, 3), 4).eval_batch([x0, x1, operand(fa), operand(fb)],
                                                                          n_points)[0]
            elif name == "arg2":
                x0, x1 = operand(fa).copy(), operand(fb).copy()
                continue
            elif name in _TERNARY_REF:
                a, b, c = operand(fa), operand(fb), operand(fc)
                r = _leaf((_TERNARY_REF[name], 3, 0, 1, 2), 3).eval_batch([a, b, c], n_points)[0]
            else:
                raise AssertionError(f"unknown device opcode {name}")
            acc = r
            if (fd & 3) == KIND_S:
                S[fd >> 2] = r.copy()
        if it + 1 < iters:
            for j, s in enumerate(param_slot):  # sequential, exactly like the kernel's feedback loop
                if s != 0xFFFF:
                    S[int(s)] = result_value(int(result_field[j]))
    return [result_value(int(f)) for f in result_field]


# ---- packed micro-instruction image (symbanafis_b200/csrc/uop.h, produced by pack.cpp) -------------------
N_H1 = 8
U_END, U_LOADK, U_NEXTWIN, U_U1 = 0, 1, 2, 3
CODE_WINDOW_BYTES = 16 * 1024
U_H1 = U_U1 + 2 * N_LIGHT_UNARY
U_BC = U_H1 + 2 * N_H1
U_BN = U_BC + 5 * 3
U_T = U_BN + 8 * 2
U_FM = U_T + 10 * 4  # fused chains: multiply (5 operand-kind variants) + {mulK, addS, mulK then addS}
U_FQ = U_FM + 3 * 5  # square {S, A} + {mulK, addS}
U_FS = U_FQ + 2 * 2  # t = S + K (kept in the slot of the kinds word when >= 0), then t*t [+ S] or K * (K * t)
U_FT = U_FS + 3  # {sin, cos} x {S[imm] + ACC, fma(S[imm], K[c], ACC)} x {plain, affine prefix}
U_FD = U_FT + 8  # trig1(S[a] prefix b) combined with trig2(S[imm] prefix kinds-word) by add / fma(., K[c], .)
U_F5 = U_FD + 8  # S[d] += K[kinds] * (S[imm] * (K[c] * (S[a] * K[b])))
U_G = U_F5 + 1
H1_NAMES = ["sin", "cos", "exp", "ln", "expsqr", "expsqrneg", "expneg", "sincos"]
BC_NAMES = ["add", "mul", "negmul"]
BC_KINDS = ["SS", "SK", "SA", "KA", "AA"]
BN_NAMES = ["sub", "div"]
BN_KINDS = ["SS", "SK", "KS", "SA", "AS", "KA", "AK", "AA"]
T_NAMES = ["fma", "fms", "fnma", "fnms"]
T_KINDS = ["SSS", "SSK", "SSA", "SKS", "SKK", "SKA", "SAS", "SAK", "KAS", "KAK"]
G_NAMES = ["pow", "powi", "builtin1", "builtin2", "builtin3", "builtin4", "arg2", "fma", "fms", "fnma", "fnms", "div",
           "recipexpm1"]
NO_PREFIX = 0xFFFFFFFF


def run_packed_program(pk: dict, cols: Sequence[np.ndarray], n_points: int, iters: int = 1) -> List[np.ndarray]:
    """Executes the packed image exactly as kernel.cu does (register file addressed by byte offset, constants
    read from the image, accumulator, sign-tested destination store), leaf math delegated to the oracle."""
    image = np.asarray(pk["image"], dtype=np.uint64)
    kview = image.view(np.float64)
    P = int(pk["points_per_thread"])
    slot_bytes = int(pk.get("block_threads", 128)) * 8 * P
    rf_bytes = int(pk["n_slots"]) * slot_bytes
    RF: Dict[int, np.ndarray] = {}

    def K(off: int) -> np.ndarray:
        assert off % 8 == 0 and off < int(pk["code_off"]), "constant operand outside the constant table"
        return np.full(n_points, kview[off // 8])

    def S(off: int) -> np.ndarray:
        assert off % slot_bytes == 0 and 0 <= off < rf_bytes, "slot operand outside the register file"
        return RF[off]  # KeyError = slot read before it was written

    def store(off: int, v: np.ndarray):
        assert off % slot_bytes == 0 and 0 <= off < rf_bytes, "slot store outside the register file"
        RF[off] = v.copy()

    for j, off in enumerate(pk["param_off"]):
        off = int(off)
        if off < 0:
            continue
        if j < len(cols):
            c = np.asarray(cols[j], dtype=np.float64)
            store(off, c[np.minimum(np.arange(n_points), len(c) - 1)])
        else:
            store(off, np.zeros(n_points))
    acc = np.zeros(n_points)

    def by_letter(letter: str, off: int) -> np.ndarray:
        return S(off) if letter == "S" else K(off) if letter == "K" else acc

    def by_kind(kind: int, off: int) -> np.ndarray:
        return S(off) if kind == KIND_S else K(off) if kind == KIND_K else acc

    def ref1(name, x):
        return _leaf((name, 1, 0), 1).eval_batch([x], n_points)[0]

    def ref2(name, x, y):
        return _leaf((name, 2, 0, 1), 2).eval_batch([x, y], n_points)[0]

    def ref3(name, x, y, z):
        return _leaf((name, 3, 0, 1, 2), 3).eval_batch([x, y, z], n_points)[0]

    dbuf = bool(pk.get("double_buffered", False))
    final_off = pk["result_off_even"] if (dbuf and iters % 2 == 0) else pk["result_off"]

    def result_value(k: int, offs=None) -> np.ndarray:
        off = int((pk["result_off"] if offs is None else offs)[k])
        return K(off) if pk["result_is_k"][k] else S(off).copy()

    i32 = lambda v: v - (1 << 32) if v >= (1 << 31) else v  # noqa: E731
    code0 = int(pk["code_off"]) // 8  # program counters are byte offsets into the code part of the image
    pc = 0
    carried = None
    for it in range(iters):
        while True:
            w0, w1, w2, w3 = (int(x) for x in image[code0 + pc:code0 + pc + 4])
            pc += 4
            uop_next, fd = w0 & 0xFFFFFFFF, i32(w0 >> 32)
            fa, fb, fc, imm, kinds, uop = w1 & 0xFFFFFFFF, w1 >> 32, w2 & 0xFFFFFFFF, w2 >> 32, w3 & 0xFFFFFFFF, w3 >> 32
            # the kernel dispatches on the opcode carried by the PREVIOUS instruction: both must agree
            assert carried is None or carried == uop, "word 0 of the previous instruction does not carry this opcode"
            carried = uop_next
            if uop == U_END:
                assert fa % 32 == 0 and code0 + fa // 8 < len(image), "U_END must point at the next iteration's start"
                pc = fa // 8
                break
            if uop == U_NEXTWIN:  # last slot of a code window: execution continues in the next window
                assert (pc * 8) % CODE_WINDOW_BYTES == 0, "U_NEXTWIN must be the last slot of its window"
                continue
            assert ((pc - 4) * 8) % CODE_WINDOW_BYTES != CODE_WINDOW_BYTES - 32, "window's last slot must be U_NEXTWIN"
            keep_acc = False
            if uop == U_LOADK:
                r = K(fa)
            elif uop < U_H1:
                u, is_a = divmod(uop - U_U1, 2)
                name = DEV_OPS[D["copy"] + u]
                x = acc if is_a else S(fa)
                r = x.copy() if name == "copy" else ref1(_UNARY_REF[name], x)
            elif uop < U_BC:
                h, prefixed = divmod(uop - U_H1, 2)
                x = S(fa)  # the hot heavy handlers never read ACC
                if prefixed:
                    x = ref3("MulAdd", x, K(fb), K(fb + 8))
                name = H1_NAMES[h]
                if name == "expneg":
                    r = _leaf(("Builtin1", 1, "exp_neg", 0), 1).eval_batch([x], n_points)[0]
                elif name == "sincos":
                    r, c_ = _leaf(("SinCos", 1, 2, 0), 1, 2).eval_batch([x], n_points)
                    if i32(fc) >= 0:
                        store(fc, c_)
                else:
                    r = ref1(_UNARY_REF[name], x)
            elif uop < U_BN:
                i, v = divmod(uop - U_BC, 5)
                kk = BC_KINDS[v]
                r = ref2(_BINARY_REF[BC_NAMES[i]], by_letter(kk[0], fa), by_letter(kk[1], fb))
            elif uop < U_T:
                i, v = divmod(uop - U_BN, 8)
                kk = BN_KINDS[v]
                r = ref2(_BINARY_REF[BN_NAMES[i]], by_letter(kk[0], fa), by_letter(kk[1], fb))
            elif uop < U_FM:
                t, v = divmod(uop - U_T, 10)
                kk = T_KINDS[v]
                r = ref3(_TERNARY_REF[T_NAMES[t]], by_letter(kk[0], fa), by_letter(kk[1], fb), by_letter(kk[2], fc))
            elif uop < U_G:  # fused chains: the same operations, one dispatch (suffix operands: c = constant, imm = slot)
                if uop == U_F5:
                    t_ = ref2("Mul", K(fc), ref2("Mul", S(fa), K(fb)))
                    r = ref2("Add", S(fd), ref2("Mul", K(kinds), ref2("Mul", S(imm), t_)))
                    sfx = -1
                elif uop >= U_FD:
                    t1, rest = divmod(uop - U_FD, 4)
                    t2, tail = divmod(rest, 2)
                    u_ = ref1("Sin" if t1 == 0 else "Cos", ref3("MulAdd", S(fa), K(fb), K(fb + 8)))
                    v_ = ref1("Sin" if t2 == 0 else "Cos", ref3("MulAdd", S(imm), K(kinds), K(kinds + 8)))
                    r = ref2("Add", u_, v_) if tail == 0 else ref3("MulAdd", u_, K(fc), v_)
                    sfx = -1
                elif uop >= U_FT:
                    trig, rest = divmod(uop - U_FT, 4)
                    tail, prefixed = divmod(rest, 2)
                    x = S(fa)
                    if prefixed:
                        x = ref3("MulAdd", x, K(fb), K(fb + 8))
                    t_ = ref1("Sin" if trig == 0 else "Cos", x)
                    r = ref2("Add", S(imm), t_) if tail == 0 else ref3("MulAdd", S(imm), K(fc), t_)
                    sfx = -1
                elif uop >= U_FS:
                    t_ = ref2("Add", S(fa), K(fb))
                    if i32(kinds) >= 0:
                        store(i32(kinds), t_)
                    which = uop - U_FS
                    if which == 2:
                        r = ref2("Mul", K(imm), ref2("Mul", K(fc), t_))
                    else:
                        r = ref1("Square", t_)
                        if which == 1:
                            r = ref2("Add", S(imm), r)
                    sfx = -1
                elif uop < U_FQ:
                    v, sfx = divmod(uop - U_FM, 3)
                    kk = BC_KINDS[v]
                    r = ref2("Mul", by_letter(kk[0], fa), by_letter(kk[1], fb))
                else:
                    q, sfx = divmod(uop - U_FQ, 2)
                    r = ref1("Square", acc if q else S(fa))
                if sfx in (0, 2):
                    r = ref2("Mul", K(fc), r)
                if sfx in (1, 2):
                    r = ref2("Add", S(imm), r)
            else:
                name = G_NAMES[uop - U_G]
                ka, kb, kc = kinds & 3, (kinds >> 2) & 3, (kinds >> 4) & 3
                if name in ("pow", "div"):
                    r = ref2(_BINARY_REF[name], by_kind(ka, fa), by_kind(kb, fb))
                elif name == "recipexpm1":
                    x = by_kind(ka, fa)
                    if fb != NO_PREFIX:
                        x = ref3("MulAdd", x, K(fb), K(fb + 8))
                    r = ref1("RecipExpm1", x)
                elif name == "powi":
                    r = _leaf(("Powi", 1, 0, i32(imm)), 1).eval_batch([by_kind(ka, fa)], n_points)[0]
                elif name == "builtin1":
                    r = _leaf(("Builtin1", 1, imm, 0), 1).eval_batch([by_kind(ka, fa)], n_points)[0]
                elif name == "builtin2":
                    r = _leaf(("Builtin2", 2, imm, 0, 1), 2).eval_batch([by_kind(ka, fa), by_kind(kb, fb)], n_points)[0]
                elif name == "builtin3":
                    r = _leaf(("Builtin3", 3, imm, 0, 1, 2), 3).eval_batch(
                        [S(int(pk["x_off"][0])), S(int(pk["x_off"][1])), by_kind(ka, fa)], n_points)[0]
                elif name == "builtin4":
                    r = _leaf(("Builtin4", 4, imm, 0, 1, 2, 3), 4).eval_batch(
                        [S(int(pk["x_off"][0])), S(int(pk["x_off"][1])), by_kind(ka, fa), by_kind(kb, fb)], n_points)[0]
                elif name == "arg2":
                    x0, x1 = by_kind(ka, fa).copy(), by_kind(kb, fb).copy()
                    store(int(pk["x_off"][0]), x0)
                    store(int(pk["x_off"][1]), x1)
                    keep_acc = True
                    r = acc
                else:
                    r = ref3(_TERNARY_REF[name], by_kind(ka, fa), by_kind(kb, fb), by_kind(kc, fc))
            if not keep_acc:
                acc = r
            if fd >= 0:
                store(fd, acc)
        if it + 1 < iters and not dbuf:
            for j, off in enumerate(pk["param_off"]):  # sequential, exactly like the kernel's feedback loop
                if int(off) >= 0:
                    store(int(off), result_value(j))
    return [result_value(k, final_off) for k in range(len(pk["result_off"]))]


# ---- register-machine image (symbanafis_b200/csrc/ruop.h, produced by pack_r.cpp) ------------------------
R_NREG, XK_K, XK_S, XK_COUNT = 8, 8, 9, 10
RU_END, RU_EXIT, RU_NEXTWIN, RU_COLD, RU_MOV = 0, 1, 2, 3, 4
RU_ST = RU_MOV + R_NREG * XK_COUNT
RU_BIN = RU_ST + R_NREG
RB_NAMES = ["add", "mul", "sub", "rsub", "negmul"]
RU_UN = RU_BIN + len(RB_NAMES) * R_NREG * XK_COUNT
RUN_NAMES = ["neg", "cube", "pow4"]
R_NPAIR = R_NREG * (R_NREG + 1) // 2
RU_FRR = RU_UN + len(RUN_NAMES) * R_NREG
RU_FKA = RU_FRR + 4 * R_NPAIR  # the fused family works on R0 only
RU_FKM = RU_FKA + 4 * R_NREG
RU_H0 = RU_FKM + 4 * XK_COUNT
RU_U0 = RU_H0 + 2 * N_H1
RU0_NAMES = ["pow3_2", "invpow3_2", "invsqrt", "invsquare", "invcube", "recip", "sqrt"]
RU_DIV = RU_U0 + len(RU0_NAMES)
RU_RDIV = RU_DIV + XK_COUNT
RU_COUNT = RU_RDIV + XK_COUNT
R_PAIRS = [(a, b) for a in range(R_NREG) for b in range(a, R_NREG)]
R_LAYOUT = 104  # points_per_thread value that selects the register-machine image in saf_program_export_packed


def run_packed_r_program(pk: dict, cols: Sequence[np.ndarray], n_points: int, iters: int = 1) -> List[np.ndarray]:
    """Executes a register-machine image exactly as kernel_r.cuh does: eight registers that are LOST whenever the
    dispatch loop is left for a cold micro-op (they are poisoned with NaN here), slots addressed by byte offset,
    the opcode of every instruction carried by its predecessor.  Leaf math is delegated to the oracle."""
    image = np.asarray(pk["image"], dtype=np.uint64)
    kview = image.view(np.float64)
    P = 4
    slot_bytes = 128 * 8 * P
    rf_bytes = int(pk["n_slots"]) * slot_bytes
    RF: Dict[int, np.ndarray] = {}
    poison = np.full(n_points, np.nan)
    R = [poison.copy() for _ in range(R_NREG)]

    def K(off: int) -> np.ndarray:
        assert off % 8 == 0 and off < int(pk["code_off"]), "constant operand outside the constant table"
        return np.full(n_points, kview[off // 8])

    def S(off: int) -> np.ndarray:
        assert off % slot_bytes == 0 and 0 <= off < rf_bytes, "slot operand outside the register file"
        return RF[off]

    def store(off: int, v: np.ndarray):
        assert off % slot_bytes == 0 and 0 <= off < rf_bytes, "slot store outside the register file"
        RF[off] = v.copy()

    def X(xk: int, off: int) -> np.ndarray:
        return R[xk] if xk < R_NREG else K(off) if xk == XK_K else S(off)

    def ref1(name, x):
        return _leaf((name, 1, 0), 1).eval_batch([x], n_points)[0]

    def ref2(name, x, y):
        return _leaf((name, 2, 0, 1), 2).eval_batch([x, y], n_points)[0]

    def ref3(name, x, y, z):
        return _leaf((name, 3, 0, 1, 2), 3).eval_batch([x, y, z], n_points)[0]

    def by_kind(kind: int, off: int) -> np.ndarray:
        assert kind in (KIND_S, KIND_K), "cold micro-ops of the register machine read slots and constants only"
        return S(off) if kind == KIND_S else K(off)

    for j, off in enumerate(pk["param_off"]):
        off = int(off)
        if off < 0:
            continue
        if j < len(cols):
            c = np.asarray(cols[j], dtype=np.float64)
            store(off, c[np.minimum(np.arange(n_points), len(c) - 1)])
        else:
            store(off, np.zeros(n_points))

    i32 = lambda v: v - (1 << 32) if v >= (1 << 31) else v  # noqa: E731
    code0 = int(pk["code_off"]) // 8
    pc, it, carried, steps = 0, 0, None, 0
    while True:
        steps += 1
        assert steps < 50_000_000, "runaway register-machine program"
        w0, w1, w2, w3 = (int(x) for x in image[code0 + pc:code0 + pc + 4])
        here = pc
        pc += 4
        uop_next, fd = w0 & 0xFFFFFFFF, i32(w0 >> 32)
        fa, fb, imm, low3, uop = w1 & 0xFFFFFFFF, w1 >> 32, w2 >> 32, w3 & 0xFFFFFFFF, w3 >> 32
        assert carried is None or carried == uop, "word 0 of the previous instruction does not carry this opcode"
        carried = uop_next
        assert uop < RU_COUNT
        if uop == RU_NEXTWIN:
            assert (pc * 8) % CODE_WINDOW_BYTES == 0, "RU_NEXTWIN must be the last slot of its window"
            continue
        assert (here * 8) % CODE_WINDOW_BYTES != CODE_WINDOW_BYTES - 32, "window's last slot must be RU_NEXTWIN"
        if uop == RU_EXIT:
            break
        if uop == RU_END:
            it += 1
            if it < iters:
                assert fa % 32 == 0
                pc = fa // 8
            carried = None  # the kernel re-reads the opcode after RU_END
            continue
        if uop == RU_COLD:
            kinds, cold = low3 & 0xFF, low3 >> 8
            name = G_NAMES[cold - U_G]
            ka, kb = kinds & 3, (kinds >> 2) & 3
            r = None
            if name == "pow":
                r = ref2("Pow", by_kind(ka, fa), by_kind(kb, fb))
            elif name == "recipexpm1":
                x = by_kind(ka, fa)
                if fb != NO_PREFIX:
                    x = ref3("MulAdd", x, K(fb), K(fb + 8))
                r = ref1("RecipExpm1", x)
            elif name == "powi":
                r = _leaf(("Powi", 1, 0, i32(imm)), 1).eval_batch([by_kind(ka, fa)], n_points)[0]
            elif name == "builtin1":
                r = _leaf(("Builtin1", 1, imm, 0), 1).eval_batch([by_kind(ka, fa)], n_points)[0]
            elif name == "builtin2":
                r = _leaf(("Builtin2", 2, imm, 0, 1), 2).eval_batch([by_kind(ka, fa), by_kind(kb, fb)], n_points)[0]
            elif name == "builtin3":
                r = _leaf(("Builtin3", 3, imm, 0, 1, 2), 3).eval_batch(
                    [S(int(pk["x_off"][0])), S(int(pk["x_off"][1])), by_kind(ka, fa)], n_points)[0]
            elif name == "builtin4":
                r = _leaf(("Builtin4", 4, imm, 0, 1, 2, 3), 4).eval_batch(
                    [S(int(pk["x_off"][0])), S(int(pk["x_off"][1])), by_kind(ka, fa), by_kind(kb, fb)], n_points)[0]
            elif name == "arg2":
                x0, x1 = by_kind(ka, fa).copy(), by_kind(kb, fb).copy()
                store(int(pk["x_off"][0]), x0)
                store(int(pk["x_off"][1]), x1)
            else:
                raise AssertionError(f"cold micro-op {name} is not produced by pack_r.cpp")
            if r is not None:
                store(int(pk["x_off"][2]), r)
                if fd >= 0:
                    store(fd, r)
            R = [poison.copy() for _ in range(R_NREG)]  # the dispatch loop's registers do not survive
            carried = None
            continue
        if uop < RU_ST:
            d, xk = divmod(uop - RU_MOV, XK_COUNT)
            R[d] = X(xk, fa).copy()
        elif uop < RU_BIN:
            store(fa, R[uop - RU_ST])
        elif uop < RU_UN:
            od, xk = divmod(uop - RU_BIN, XK_COUNT)
            o, d = divmod(od, R_NREG)
            name, x = RB_NAMES[o], X(xk, fa)
            R[d] = ref2("Sub", x, R[d]) if name == "rsub" else ref2(_BINARY_REF[name], R[d], x)
        elif uop < RU_FRR:
            u, d = divmod(uop - RU_UN, R_NREG)
            R[d] = ref1(_UNARY_REF[RUN_NAMES[u]], R[d])
        elif uop < RU_FKA:
            s_, pair = divmod(uop - RU_FRR, R_NPAIR)
            a, b = R_PAIRS[pair]
            R[0] = ref3(_TERNARY_REF[T_NAMES[s_]], R[a], R[b], R[0])
        elif uop < RU_FKM:
            s_, a = divmod(uop - RU_FKA, R_NREG)
            R[0] = ref3(_TERNARY_REF[T_NAMES[s_]], R[a], K(fa), R[0])
        elif uop < RU_H0:
            s_, xk = divmod(uop - RU_FKM, XK_COUNT)
            R[0] = ref3(_TERNARY_REF[T_NAMES[s_]], R[0], K(fa), X(xk, fb))
        elif uop < RU_U0:
            h, prefixed = divmod(uop - RU_H0, 2)
            x = R[0]  # R1 = f(R0), sincos: R1 = sin, R2 = cos
            if prefixed:
                x = ref3("MulAdd", x, K(fb), K(fb + 8))
            name = H1_NAMES[h]
            if name == "expneg":
                R[1] = _leaf(("Builtin1", 1, "exp_neg", 0), 1).eval_batch([x], n_points)[0]
            elif name == "sincos":
                R[1], R[2] = _leaf(("SinCos", 1, 2, 0), 1, 2).eval_batch([x], n_points)
            else:
                R[1] = ref1(_UNARY_REF[name], x)
        elif uop < RU_DIV:
            R[0] = ref1(_UNARY_REF[RU0_NAMES[uop - RU_U0]], R[0])
        elif uop < RU_RDIV:
            R[0] = ref2("Div", R[0], X(uop - RU_DIV, fa))
        else:
            R[0] = ref2("Div", X(uop - RU_RDIV, fa), R[0])

    def result_value(k: int) -> np.ndarray:
        off = int(pk["result_off"][k])
        return K(off) if pk["result_is_k"][k] else S(off).copy()

    return [result_value(k) for k in range(len(pk["result_off"]))]
