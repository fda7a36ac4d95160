#!/bin/bash
# Builds a variant of libovb200.so with extra nvcc flags (kernel experiments), loaded with OVB_LIB_PATH.
#   scripts/build_variant_lib.sh <name> <flags...>   ->  scripts/microbench/build/libovb200_<name>.so
set -eu
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p scripts/microbench/build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
  --expt-relaxed-constexpr "$@" -shared omni_video_b200/csrc/*.cu -o scripts/microbench/build/libovb200_$name.so
echo built scripts/microbench/build/libovb200_$name.so
