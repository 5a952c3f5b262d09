#!/usr/bin/env python
"""Disassembles the register-machine image (csrc/ruop.h) of a named workload: python tools/rdis.py aizawa"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import dev_emulator as E
from symbanafis_b200 import workloads
from symbanafis_b200._lib import DeviceProgramHandle


def xname(xk, off, kview, sb):
    return f"R{xk}" if xk < 8 else f"K[{kview[off // 8]:.6g}]" if xk == E.XK_K else f"S{off // sb}"


def disassemble(pk):
    image = np.asarray(pk["image"], dtype=np.uint64); kview = image.view(np.float64)
    sb = 128 * 8 * 4; code0 = pk["code_off"] // 8; out = []
    for i in range((len(image) - code0) // 4):
        w0, w1, w2, w3 = (int(x) for x in image[code0 + 4 * i:code0 + 4 * i + 4])
        fa, fb, uop, low3 = w1 & 0xFFFFFFFF, w1 >> 32, w3 >> 32, w3 & 0xFFFFFFFF
        if uop == E.RU_END: t = f"end -> {fa // 32}"
        elif uop == E.RU_EXIT: t = "exit"
        elif uop == E.RU_NEXTWIN: t = "nextwin"
        elif uop == E.RU_COLD: t = f"cold {E.G_NAMES[(low3 >> 8) - E.U_G]}"
        elif uop < E.RU_ST: d, xk = divmod(uop - E.RU_MOV, 10); t = f"R{d} = {xname(xk, fa, kview, sb)}"
        elif uop < E.RU_BIN: t = f"S{fa // sb} = R{uop - E.RU_ST}"
        elif uop < E.RU_UN:
            od, xk = divmod(uop - E.RU_BIN, 10); o, d = divmod(od, 8); t = f"R{d} = {E.RB_NAMES[o]}(R{d}, {xname(xk, fa, kview, sb)})"
        elif uop < E.RU_FRR: u, d = divmod(uop - E.RU_UN, 8); t = f"R{d} = {E.RUN_NAMES[u]}(R{d})"
        elif uop < E.RU_FKA: s, p = divmod(uop - E.RU_FRR, 36); a, b = E.R_PAIRS[p]; t = f"R0 = {E.T_NAMES[s]}(R{a}, R{b}, R0)"
        elif uop < E.RU_FKM: s, a = divmod(uop - E.RU_FKA, 8); t = f"R0 = {E.T_NAMES[s]}(R{a}, K[{kview[fa // 8]:.6g}], R0)"
        elif uop < E.RU_H0: s, xk = divmod(uop - E.RU_FKM, 10); t = f"R0 = {E.T_NAMES[s]}(R0, K[{kview[fa // 8]:.6g}], {xname(xk, fb, kview, sb)})"
        elif uop < E.RU_U0: h, pre = divmod(uop - E.RU_H0, 2); t = f"R1{', R2' if E.H1_NAMES[h] == 'sincos' else ''} = {E.H1_NAMES[h]}(R0{' * K + K' if pre else ''})"
        elif uop < E.RU_DIV: t = f"R0 = {E.RU0_NAMES[uop - E.RU_U0]}(R0)"
        elif uop < E.RU_RDIV: t = f"R0 = div(R0, {xname(uop - E.RU_DIV, fa, kview, sb)})"
        else: t = f"R0 = div({xname(uop - E.RU_RDIV, fa, kview, sb)}, R0)"
        out.append(f"[{i}] {t}")
    return out


if __name__ == "__main__":
    w = workloads.get(sys.argv[1]); p = w.program
    h = DeviceProgramHandle.create(p.flat, p.consts, p.arg_pool, p.param_count, p.workspace_size, p.result_regs)
    print("\n".join(disassemble(h.export_packed(E.R_LAYOUT))))
