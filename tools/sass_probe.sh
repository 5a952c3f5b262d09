#!/bin/bash
# Developer probe (no GPU needed): builds ONE interpreter entry point (P = 4, constants in shared memory) through the
# same pipeline as the Makefile (nvcc -ptx -> patch_brx.py -> ptxas) and leaves its SASS in /tmp/probe.sass.
# Usage: tools/sass_probe.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")/../symbanafis_b200/csrc"
nvcc -arch=compute_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -DSAF_FAST_BUILD "$@" -ptx kernel.cu -o /tmp/probe.ptx
python3 ../../tools/patch_brx.py /tmp/probe.ptx /tmp/probe.brx.ptx
ptxas -arch=sm_100a -lineinfo -v /tmp/probe.brx.ptx -o /tmp/probe.cubin 2>&1 | grep -E "Compiling entry|Used|hot_run" -A1 | grep -E "Used|spill|Compiling" | head -12
cuobjdump -sass -fun saf_interp_p4_ksm /tmp/probe.cubin | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s/ +;/;/' > /tmp/probe.sass
echo "$(wc -l < /tmp/probe.sass) SASS instructions in saf_interp_p4_ksm -> /tmp/probe.sass"
