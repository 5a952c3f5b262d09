#!/usr/bin/env python
"""Disassembles the packed micro-op image (csrc/uop.h) of a named workload: python tools/udis.py aizawa [P]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import dev_emulator as E
from symbanafis_b200 import workloads
from symbanafis_b200._lib import DeviceProgramHandle

LIGHT_UNARY = ["copy", "neg", "square", "cube", "pow4", "pow3_2", "invpow3_2", "invsqrt", "invsquare", "invcube", "recip", "sqrt"]
BC_KINDS = ["SS", "SK", "SA", "KA", "AA"]
BN_KINDS = ["SS", "SK", "KS", "SA", "AS", "KA", "AK", "AA"]


def name_of(uop):
    if uop == E.U_END: return "end"
    if uop == E.U_LOADK: return "loadk"
    if uop == E.U_NEXTWIN: return "nextwin"
    if uop < E.U_H1: u, a = divmod(uop - E.U_U1, 2); return f"{LIGHT_UNARY[u]}:{'A' if a else 'S'}"
    if uop < E.U_BC: h, p = divmod(uop - E.U_H1, 2); return f"{E.H1_NAMES[h]}{'+prefix' if p else ''}"
    if uop < E.U_BN: i, v = divmod(uop - E.U_BC, 5); return f"{E.BC_NAMES[i]}:{BC_KINDS[v]}"
    if uop < E.U_T: i, v = divmod(uop - E.U_BN, 8); return f"{E.BN_NAMES[i]}:{BN_KINDS[v]}"
    if uop < E.U_FM: t, v = divmod(uop - E.U_T, 10); return f"{E.T_NAMES[t]}:{E.T_KINDS[v]}"
    if uop < E.U_FQ: v, sfx = divmod(uop - E.U_FM, 3); return f"FM mul:{BC_KINDS[v]} {['*K', '+S', '*K+S'][sfx]}"
    if uop < E.U_FS: q, sfx = divmod(uop - E.U_FQ, 2); return f"FQ square:{'A' if q else 'S'} {['*K', '+S'][sfx]}"
    if uop < E.U_FT: return "FS (S+K) " + ["^2", "^2+S", "*K*K"][uop - E.U_FS]
    if uop < E.U_FD: return f"FT {uop - E.U_FT}"
    if uop < E.U_F5: return f"FD {uop - E.U_FD}"
    if uop == E.U_F5: return "F5 S+=K*(S*(K*(S*K)))"
    return "cold " + E.G_NAMES[uop - E.U_G]


if __name__ == "__main__":
    w = workloads.get(sys.argv[1]); p = w.program
    P = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    h = DeviceProgramHandle.create(p.flat, p.consts, p.arg_pool, p.param_count, p.workspace_size, p.result_regs)
    pk = h.export_packed(P)
    image = np.asarray(pk["image"], dtype=np.uint64)
    code0 = pk["code_off"] // 8
    n = (len(image) - code0) // 4
    if pk["double_buffered"]: n //= 2
    hist = {}
    for i in range(n):
        w0, w1, w2, w3 = (int(x) for x in image[code0 + 4 * i:code0 + 4 * i + 4])
        uop = w3 >> 32
        d = w0 >> 32
        d = d - (1 << 32) if d >= (1 << 31) else d
        nm = name_of(uop)
        hist[nm] = hist.get(nm, 0) + 1
        print(f"[{i}] {nm}{'  -> slot' if d >= 0 else ''}")
    print("---", n, "micro-ops")
    for k, v in sorted(hist.items(), key=lambda kv: -kv[1]): print(f"{v:5d} {k}")
