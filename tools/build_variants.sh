#!/bin/bash
# experiment helper: tools/build_variants.sh <source basename, e.g. attn_tc> <DEFINE>... builds
# build_exp/lib_<DEFINE>.so = the normal library with that one source recompiled with -D<DEFINE> (build_exp/ is git-ignored)
set -e
cd "$(dirname "$0")/.."
src=$1; shift
mkdir -p build_exp
python gaot_b200/build.py > /dev/null
OBJS=$(ls gaot_b200/build/*.o | grep -v "/$src.o")
for v in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Wno-deprecated-gpu-targets -D$v -c gaot_b200/csrc/$src.cu -o build_exp/${src}_$v.o &
done
wait
for v in "$@"; do
  nvcc -shared -Wno-deprecated-gpu-targets -o build_exp/lib_$v.so $OBJS build_exp/${src}_$v.o
done
ls -la build_exp/*.so
