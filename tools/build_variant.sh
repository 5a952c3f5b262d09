#!/bin/bash
# Builds a variant of the library next to the shipped one: tools/build_variant.sh <name> "<extra nvcc flags>"
#   -> build/variants/<name>.so   (load it with SAF_B200_LIB=$PWD/build/variants/<name>.so)
# e.g. tools/build_variant.sh checked -DSAF_CHECKED     (bounds-checked shared-memory / column accesses, kernel.cu)
set -e
NAME=$1; shift
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC=$ROOT/build/variants/${NAME}_src
mkdir -p $SRC/symbanafis_b200/csrc $SRC/tools $SRC/include
cp $ROOT/symbanafis_b200/csrc/*.cu $ROOT/symbanafis_b200/csrc/*.cuh $ROOT/symbanafis_b200/csrc/*.cpp $ROOT/symbanafis_b200/csrc/*.h \
   $ROOT/symbanafis_b200/csrc/Makefile $SRC/symbanafis_b200/csrc/
cp $ROOT/tools/patch_brx.py $SRC/tools/
cp $ROOT/include/*.h $SRC/include/
make -C $SRC/symbanafis_b200/csrc -j4 EXTRA="$*" OUT=$ROOT/build/variants/$NAME.so > $SRC/build.log 2>&1 || { grep -B2 -A6 "error" $SRC/build.log | head -40; exit 1; }
rm -rf $SRC  # the sources were a copy: the snapshot that travels to the GPU box stays small
ls -la $ROOT/build/variants/$NAME.so
