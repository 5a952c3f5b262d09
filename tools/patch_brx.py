#!/usr/bin/env python
"""PTX peephole of the build: turns the compare tree nvcc emits for hot_run's `switch (uop)` into ONE brx.idx.

nvcc (NVVM) has no jump-table lowering for NVPTX: a C++ switch becomes a balanced tree of setp/bra whose leaves
ptxas folds into 3-entry BRX tables -- about 5 compare levels + a BRX sequence per dispatched micro-op.  The
interpreter's dispatch is the hottest path of the library, so the build rewrites it:

  kernel.cu --nvcc -ptx--> kernel.ptx --this script--> kernel.brx.ptx --ptxas--> kernel.cubin

kernel.cu marks the dispatch point and every case with PTX comments (SAF_DISPATCH <reg> / SAF_CASE <value>);
for every function that contains a dispatch marker this script
  * maps each case value to the label of the basic block its marker sits in,
  * declares a .branchtargets table over all micro-op values (values without a case -> the default block),
  * inserts `brx.idx <reg>, <table>` right after the dispatch marker; the original tree becomes unreachable and
    ptxas drops it, and
  * replaces the jump back to the loop head that follows every SAF_REDISPATCH marker (end of a handler) by a copy
    of what that jump would execute there -- the straight-line instructions the head block holds in front of the
    dispatch -- followed by its own `brx.idx`: threaded code, one indirect branch per micro-op.
Nothing else in the PTX is touched.  The dispatch marker also carries the number of micro-ops (table size).
Usage: patch_brx.py in.ptx out.ptx
"""
import os
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
        decl = f"{tbl}: .branchtargets " + ", ".join(entries) + ";"
        # the loop-head dispatch.  The head block may hold instructions in front of the marker (the compiler sinks
        # code common to all predecessors there, e.g. the program-counter increment): `prefix`
        k = disp[0]
        h = k
        while h > 0 and not re.match(r"^\$L__BB\w+:", body[h]):
            h -= 1
        head_label = body[h].split(":")[0]
        begin_asm = k
        while "// begin inline asm" not in body[begin_asm]:
            begin_asm -= 1
        prefix = [l for l in body[h + 1:begin_asm] if l.strip() and not l.strip().startswith(".loc")]
        threaded = any("// SAF_REDISPATCH" in l for l in body)
        for l in prefix:  # only copied (and therefore only constrained) when the function has threaded sites
            if threaded and (re.match(r"^\s*(@|bra|brx|ret|call|bar)", l) or "inline asm" in l):
                raise SystemExit("patch_brx: unexpected instruction in the loop head: " + l)
        while "// end inline asm" not in body[k]:
            k += 1
        body = body[:k + 1] + [f"\tbrx.idx \t{reg}, {tbl};"] + body[k + 1:]
        # threaded dispatch: after every SAF_REDISPATCH marker the block jumps back to the loop head; that jump
        # (and only that: register moves the compiler put before it stay) becomes another brx.idx
        n_re = 0
        i2 = 0
        while i2 < len(body):
            m = re.search(r"// SAF_REDISPATCH\s+(%\w+)\s+\d+;", body[i2])
            if not m:
                i2 += 1
                continue
            r2 = m.group(1)
            j2 = i2 + 1
            while j2 < len(body):
                t = body[j2].strip()
                if re.match(r"^\$L__BB\w+:", t):
                    raise SystemExit("patch_brx: no jump after a redispatch marker")
                mb = re.match(r"^bra(\.uni)?\s+(\$L__BB\w+);", t)
                if mb:
                    if mb.group(2) != head_label:
                        raise SystemExit("patch_brx: the jump after a redispatch marker does not go to the loop head")
                    break
                if t.startswith("@"):
                    raise SystemExit("patch_brx: conditional branch after a redispatch marker: " + t)
                j2 += 1
            # = what the jump to the loop head would have executed: the head's prefix, then the dispatch (on the
            # loop head's own opcode register, which the moves before the jump have just set)
            body[j2:j2 + 1] = prefix + [f"\tbrx.idx \t{reg}, {tbl};"]
            n_re += 1
            i2 = j2 + len(prefix) + 1
        # declare the table at the top of the function body (after the register declarations)
        d = 1
        while d < len(body) and (body[d].strip().startswith(".reg") or body[d].strip().startswith(".local") or
                                 body[d].strip().startswith(".loc") or body[d].strip().startswith(".shared") or
                                 body[d].strip() == "" or body[d].strip().startswith("//")):
            d += 1
        body = body[:d] + [decl] + body[d:]
        out.extend(body)
        sys.stderr.write(f"patch_brx: function {n_patched}: {len(target) - 1} cases + default, register {reg}, "
                         f"{n_re} threaded dispatch sites\n")
    if not n_patched:
        raise SystemExit("patch_brx: no dispatch marker found")
    return "\n".join(out)


if __name__ == "__main__":
    src, dst = sys.argv[1], sys.argv[2]
    with open(src) as f:
        t = f.read()
    patched = patch(t)  # raises before anything is written
    with open(dst + ".tmp", "w") as f:
        f.write(patched)
    os.replace(dst + ".tmp", dst)
