#!/bin/bash
# Attention experiments: pipeline traces of the variant builds under scripts/microbench/build/libovb_trace*.so
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in scripts/microbench/build/libovb_trace*.so; do
  n=$(basename $lib .so)
  OVB_TRACE_LIB=$(basename $lib) timeout 120 python scripts/attn_trace.py > gpurun_out/trace_$n.log 2>&1
  grep -E "TF/s|cycles per KV" gpurun_out/trace_$n.log
done
