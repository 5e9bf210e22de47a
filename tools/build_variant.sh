#!/bin/bash
# Development aid: build a variant of the library with extra -D flags into build/lib_<tag>.so
# (one translation unit is recompiled, the other objects are reused).
# Usage: [SRC=fused] tools/build_variant.sh <tag> -DKI_EXP=1 ...
set -e
TAG=$1; shift
SRC=${SRC:-fused}
cd "$(dirname "$0")/.."
mkdir -p build/obj build/var
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -I include "$@" -c -o build/var/${SRC}_$TAG.o partner_b200/csrc/$SRC.cu
OBJS=$(ls build/obj/*.o | grep -v "/$SRC.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/lib_$TAG.so build/var/${SRC}_$TAG.o $OBJS
echo build/lib_$TAG.so
