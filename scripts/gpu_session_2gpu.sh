#!/bin/bash
# 2-GPU session: parity of the parallel layouts on hardware (logs kept for profiles/), then bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 scripts/sp_check.py 2>&1 | grep -E "rank|Error|error" | tee gpurun_out/sp_check_2gpu.log
run 29512 scripts/vae_strip_check.py 2>&1 | grep -E "rank|Error|error" | tee gpurun_out/vae_strip_check_2gpu.log
timeout 300 python -m pytest tests/test_gpu_dit.py tests/test_gpu_vae.py -m gpu -q -k "two_gpus" 2>&1 | tail -3 | tee gpurun_out/pytest_2gpu.log
run 29513 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_2gpu.err | grep '^{' | tee gpurun_out/bench_2gpu_cfg2.json
run 29514 bench.py --gpus 2 --steps 5 --warmup 3 --no-cfg-split --no-vae 2>gpurun_out/bench_2gpu_u2.err | grep '^{' | tee gpurun_out/bench_2gpu_ulysses2.json
run 29515 scripts/vae_strip_check.py --config5 2>gpurun_out/vae_c5_2gpu.err | grep '^{' | tee gpurun_out/vae_config5_2gpu.json
if [ "${A14B:-0}" = "1" ]; then
  run 29516 bench.py --gpus 2 --model A14B --video 720p --steps 6 --warmup 2 --no-cfg-split --no-vae 2>gpurun_out/bench_2gpu_a14b_u2.err | grep '^{' | tee gpurun_out/bench_2gpu_A14B_720p_ulysses2.json
  run 29517 bench.py --gpus 2 --model A14B --video 720p --steps 6 --warmup 2 --no-vae 2>gpurun_out/bench_2gpu_a14b_cfg2.err | grep '^{' | tee gpurun_out/bench_2gpu_A14B_720p_cfg2.json
fi
tail -n 2 gpurun_out/*2gpu*.err gpurun_out/vae_c5_2gpu.err
