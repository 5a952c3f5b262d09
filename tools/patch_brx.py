#!/usr/bin/env python
"""PTX peephole of the build: turns the compare tree nvcc emits for hot_run's `switch (uop)` into ONE brx.idx.

nvcc (NVVM) has no jump-table lowering for NVPTX: a C++ switch becomes a balanced tree of setp/bra whose leaves
ptxas folds into 3-entry BRX tables -- about 5 compare levels + a BRX sequence per dispatched micro-op.  The
interpreter's dispatch is the hottest path of the library, so the build rewrites it:

  kernel.cu --nvcc -ptx--> kernel.ptx --this script--> kernel.brx.ptx --ptxas--> kernel.cubin

kernel.cu marks the dispatch point and every case with PTX comments (SAF_DISPATCH <reg> / SAF_CASE <value>);
for every function that contains a dispatch marker this script
  * maps each case value to the label of the basic block its marker sits in,
  * declares a .branchtargets table over all micro-op values (values without a case -> the default block) and
  * inserts `brx.idx <reg>, <table>` right after the dispatch marker; the original tree becomes unreachable and
    ptxas drops it.
Nothing else in the PTX is touched.  The dispatch marker also carries the number of micro-ops (table size).
Usage: patch_brx.py in.ptx out.ptx
"""
import re
import sys


def patch(text: str) -> str:
    lines = text.split("\n")
    out, i, n_patched = [], 0, 0
    # function bodies: from a line that is exactly "{" to the matching "}" at column 0
    while i < len(lines):
        if lines[i].strip() != "{":
            out.append(lines[i])
            i += 1
            continue
        j = i + 1
        while j < len(lines) and lines[j] != "}":
            j += 1
        body = lines[i:j + 1]
        i = j + 1
        disp = [k for k, l in enumerate(body) if "// SAF_DISPATCH" in l]
        if not disp:
            out.extend(body)
            continue
        if len(disp) != 1:
            raise SystemExit("patch_brx: a function with more than one dispatch marker")
        m = re.search(r"SAF_DISPATCH\s+(%\w+)\s+(\d+);", body[disp[0]])
        reg, table_size = m.group(1), int(m.group(2))
        target = {}
        label = None
        for l in body:
            m = re.match(r"^(\$L__BB\w+):", l)
            if m:
                label = m.group(1)
            m = re.search(r"// SAF_CASE\s+(\w+);", l)
            if m:
                if label is None:
                    raise SystemExit("patch_brx: case marker before any label")
                key = m.group(1)
                if key in target and target[key] != label:
                    raise SystemExit(f"patch_brx: case {key} marked in two blocks")
                target[key] = label
        if "default" not in target:
            raise SystemExit("patch_brx: no default marker")
        n_patched += 1
        tbl = f"$SAF_BRX_{n_patched}"
        entries = [target.get(str(v), target["default"]) for v in range(table_size)]
        k = disp[0]
        while "// end inline asm" not in body[k]:
            k += 1
        ins = [f"{tbl}: .branchtargets " + ", ".join(entries) + ";", f"\tbrx.idx \t{reg}, {tbl};"]
        out.extend(body[:k + 1] + ins + body[k + 1:])
        sys.stderr.write(f"patch_brx: function {n_patched}: {len(target) - 1} cases + default, register {reg}\n")
    if not n_patched:
        raise SystemExit("patch_brx: no dispatch marker found")
    return "\n".join(out)


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    with open(src) as f:
        t = f.read()
    with open(dst, "w") as f:
        f.write(patch(t))
