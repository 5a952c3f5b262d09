#!/bin/bash
# Developer probe: compiles ONE instantiation of the interpreter kernel (P=4, small blob) and prints its register
# use, spills and the SASS of the dispatch-loop head.  No GPU needed.  Usage: tools/sass_probe.sh [extra nvcc flags]
cd "$(dirname "$0")/../symbanafis_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -DSAF_FAST_BUILD "$@" \
     -Xptxas -v -cubin kernel.cu -o /tmp/probe.cubin 2>&1 | grep -A2 "saf_interp" | grep -E "spill|Used"
cuobjdump -sass /tmp/probe.cubin | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s/ +;/;/' > /tmp/probe.sass
wc -l < /tmp/probe.sass
