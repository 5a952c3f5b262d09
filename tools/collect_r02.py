#!/usr/bin/env python
"""Copies the judged artefacts of one tools/gpu_final.sh run from gpurun_out/<tag>/ into profiles/ under a prefix:
per-workload ncu summaries (tools/gpu_prof.sh), DRAM traffic per launch (-> profiles/r02_dram_traffic.json, read by
bench.py), the bench lines, the ncu launch list with the interpreter's share of a step.
Usage: python tools/collect_r02.py <tag> <prefix, e.g. r02_a>"""
import csv, json, os, shutil, sys

tag, prefix = sys.argv[1], sys.argv[2]
src, dst = os.path.join("gpurun_out", tag), "profiles"
traffic = {}
# captures of the interpreter (tools/gpu_prof.sh with MACHINE=interpreter) carry the suffix _interp in their names
tpath_spec = os.path.join(dst, "r02_dram_traffic_specialized.json")
tpath_interp = os.path.join(dst, "r02_dram_traffic.json")
traffic_spec = json.load(open(tpath_spec)) if os.path.exists(tpath_spec) else {}
tpath = tpath_interp
if os.path.exists(tpath):
    traffic = json.load(open(tpath))
for f in sorted(os.listdir(src)):
    if not f.startswith("ncu_summary_"):
        continue
    name = f[len("ncu_summary_"):-4]
    text = open(os.path.join(src, f)).read()
    plain = os.path.join(src, f"plain_{name}.log")
    line = ""
    if os.path.exists(plain):
        js = [l for l in open(plain) if l.startswith("{")]
        if js:
            d = json.loads(js[-1])
            line = (f"Plain run of the same command (exit 0, before the capture): {d['ms_per_step']:.4f} ms/step, "
                    f"{d['value']:.3e} point-evals/s, roofline {d['roofline']['bound']} {d['roofline']['frac']:.3f}.\n\n")
    with open(os.path.join(dst, f"{prefix}_{name}.md"), "w") as o:
        o.write(f"# Round 2, shipped build: {name} ({'interpreter kernel' if name.endswith('_interp') or '_interp_' in name else 'specialised kernel saf_spec'}) -- ncu --set full --clock-control none --import-source on, one launch "
                f"(tools/gpu_prof.sh, gpurun tag {tag})\n\n{line}```\n{text}```\n")
    for l in text.splitlines():
        if l.startswith("DRAM_TRAFFIC_JSON "):
            t = json.loads(l[len("DRAM_TRAFFIC_JSON "):])
            key = name if "_" not in name or not name.split("_")[-1].isdigit() else None
            tgt = traffic if key and key.endswith("_interp") else traffic_spec
            if key and key.endswith("_interp"):
                key = key[:-len("_interp")]
            if key:
                tgt[key] = dict(t, source=f"profiles/{prefix}_{name}.md (gpurun tag {tag})",
                                    note="one ncu --set full capture, per launch; output columns of a small launch are still in the "
                                         "126 MB L2 when the kernel ends, so their DRAM writes show up after it")
json.dump(traffic, open(tpath, "w"), indent=1)
json.dump(traffic_spec, open(tpath_spec, "w"), indent=1)
bl = os.path.join(src, "bench_full.log")
if os.path.exists(bl):
    js = [l for l in open(bl) if l.startswith("{")]
    ref = [l for l in open(os.path.join(src, "bench_reference.log")) if l.startswith("{")] if os.path.exists(os.path.join(src, "bench_reference.log")) else []
    with open(os.path.join(dst, f"{prefix}_bench_lines.md"), "w") as o:
        o.write(f"# Round 2 bench lines (one B200, tools/gpu_final.sh, gpurun tag {tag})\n\n## python bench.py --steps 20 --warmup 5\n"
                f"```json\n{js[-1] if js else 'no JSON line'}```\n## python bench.py --impl reference --steps 3 --warmup 1\n```json\n{ref[-1] if ref else 'no JSON line'}```\n")
        if js:
            d = json.loads(js[-1])
            o.write("\n| workload | machine | ms/step | point-evals/s | roofline | interpreter ms/step | e2e pinned | e2e pageable | CPU f64x4 arm |\n|---|---|---|---|---|---|---|---|---|\n")
            rows = [("clifford (headline)", d)] + list(d.get("workloads", {}).items())
            for k, v in rows:
                cb = v.get("cpu_baseline") or {}
                mi, it = v.get("machine") or {}, v.get("interpreter") or {}
                o.write(f"| {k} | {mi.get('name', 'interpreter')} P={mi.get('points_per_thread', '-')} x{mi.get('block_threads', '-')} | {v['ms_per_step']:.4f} | {v['value']:.3e} | {v['roofline']['bound']} {v['roofline']['frac']:.3f} | "
                        f"{it.get('ms_per_step', float('nan')):.4f} | {v['e2e']['value']:.3e} | {v['e2e_pageable']['value']:.3e} | {cb.get('value', 0):.3e} x{cb.get('cores', '?')} |\n")
lc = os.path.join(src, "launches.csv")
if os.path.exists(lc):
    shutil.copy(lc, os.path.join(dst, f"{prefix}_launches_clifford.csv"))
    rows = [r for r in csv.reader(l for l in open(lc) if l.startswith('"'))]
    h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    tot = {}
    for r in rows[1:]:
        tot.setdefault(r[ik], [0, 0.0]); tot[r[ik]][0] += 1; tot[r[ik]][1] += float(r[iv].replace(",", ""))
    with open(os.path.join(dst, f"{prefix}_launches_clifford.md"), "w") as o:
        o.write(f"# ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-workloads` ({tag}; gpu__time_duration.sum, --clock-control none)\n\n")
        o.write("| kernel | launches | total ms | mean ms |\n|---|---|---|---|\n")
        for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            o.write(f"| {k} | {n} | {ns / 1e6:.3f} | {ns / 1e6 / n:.3f} |\n")
        o.write("\nThe library's kernel (`saf_spec`, the specialised Clifford program; `saf_interp_p4_ksm` in the interpreter comparison) is the only kernel of the library inside a timed step (one launch per step = 1000 map "
                "iterations over 1 M particles); the rest are torch's L2-flush fill and state copies, the FP64-peak probe, the ideal-kernel probe and "
                "the host-path launches of the e2e legs, all outside the CUDA events of the device-timed steps.\n")
print("collected", tag, "->", prefix)
