#!/bin/bash
# Debug build of the attention kernel with pipeline time stamps (-DOVB_ATTN_TRACE) for scripts/attn_trace.py.
set -eu
cd "$(dirname "$0")/.."
mkdir -p scripts/microbench/build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
  --expt-relaxed-constexpr -DOVB_ATTN_TRACE ${OVB_TRACE_FLAGS:-} -shared \
  omni_video_b200/csrc/attn.cu omni_video_b200/csrc/host_util.cu -o scripts/microbench/build/${OVB_TRACE_OUT:-libovb_trace.so}
echo built scripts/microbench/build/${OVB_TRACE_OUT:-libovb_trace.so}
