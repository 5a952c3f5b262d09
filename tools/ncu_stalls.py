#!/usr/bin/env python
"""Aggregates the sampled stall locations of an .ncu-rep by opcode: python tools/ncu_stalls.py <report.ncu-rep>"""
import csv, io, subprocess, sys, collections
rep=sys.argv[1]
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
h=rows[1]
iS,iE,iA,iN=h.index("Source"),h.index("Instructions Executed"),h.index("Address"),h.index("# Samples")
tot=0; byop=collections.Counter(); prev=None; after=collections.Counter()
for r in rows[2:]:
    try: e=int(r[iE]); s=int(r[iN])
    except (ValueError, IndexError): continue
    t=r[iS].split()
    if not t: continue
    o=(t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    tot+=s; byop[o]+=s
    if prev: after[prev]+=s
    prev=o
print("total samples", tot)
print("by opcode (samples at the instruction = stall waiting to issue it):", [(o, s, round(100*s/tot,1)) for o,s in byop.most_common(12)])
print("samples at instruction FOLLOWING:", [(o, s, round(100*s/tot,1)) for o,s in after.most_common(8)])
