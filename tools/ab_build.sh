#!/bin/bash
# Build variant libraries for A/B runs on one GPU box: tools/ab_build.sh <tag> <extra nvcc flags...>  -> build/ab/lib<tag>.so
# (use with SELENITE_B200_LIB=build/ab/lib<tag>.so; build/ is git-ignored but travels with gpurun)
set -e
TAG=$1; shift
mkdir -p build/ab
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
  -o build/ab/lib${TAG}.so selenite_lite_b200/csrc/*.cu selenite_lite_b200/csrc/*.cpp 2>&1 | grep -E "error|spill" || true
ls -la build/ab/lib${TAG}.so
