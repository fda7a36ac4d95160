#!/bin/bash
# Attention-only GPU session: parity tests of the attention kernel, A/B timing against the round-1 kernel
# (OVB_ATTN_V1=1), pipeline trace.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_properties.py -m gpu -x -q -k "attention" 2>&1 | tail -15 | tee gpurun_out/pytest_attn.log
timeout 300 python scripts/gpu_microbench.py --quick --attn-only 2>&1 | tee gpurun_out/attn_v2.log
OVB_ATTN_V1=1 timeout 300 python scripts/gpu_microbench.py --quick --attn-only 2>&1 | tee gpurun_out/attn_v1.log
timeout 120 python scripts/attn_trace.py 2>&1 | tee gpurun_out/attn_trace.log
