#!/usr/bin/env python
"""Prints the dispatch-loop head of /tmp/probe.sass: from the prefetch LDC.64 pair to the first opcode compare."""
import re, sys
lines = open(sys.argv[1] if len(sys.argv) > 1 else "/tmp/probe.sass").read().split("\n")
# loop head = the region containing 'BSSY.RELIABLE' closest before first 'ISETP.GT.AND P0, PT, R., 0x' tree compare
idx = [i for i, l in enumerate(lines) if "BSSY.RELIABLE" in l]
for i in idx[:1]:
    j = i
    while j > 0 and "BRA" not in lines[j - 1] and i - j < 40:
        j -= 1
    k = i
    while k < len(lines) and not re.search(r"@!?P\d\s+BRA", lines[k]):
        k += 1
    print("\n".join(l.strip()[:90] for l in lines[j:k + 1]))
    print("head length", k + 1 - j)
