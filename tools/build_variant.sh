#!/bin/bash
# Development aid: build a variant of the library with extra -D flags into build/lib_<tag>.so
# (fused.cu is recompiled, the other objects are reused).  Usage: tools/build_variant.sh <tag> -DKI_EXP=1 ...
set -e
TAG=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/obj
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -I include "$@" -c -o build/obj/fused_$TAG.o partner_b200/csrc/fused.cu
OBJS=$(ls build/obj/*.o | grep -v "/fused" )
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/lib_$TAG.so build/obj/fused_$TAG.o $OBJS
echo build/lib_$TAG.so
