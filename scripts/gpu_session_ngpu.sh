#!/bin/bash
# usage: bash scripts/gpu_session_ngpu.sh N      (N = 4 or 8; one `gpurun --gpus N` call)
# 1.3B scaling point, A14B 720p with pure Ulysses-N and with cfg2 x Ulysses-N/2, VAE config 5, SP parity.
N=$1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# hard wall-clock caps: a hung multi-GPU job is charged N x its duration
run() { timeout $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 120 29520 scripts/sp_check.py 2>&1 | grep -E "rank 0|Error|error" | tee gpurun_out/sp_check_${N}gpu.log
run 240 29521 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_${N}gpu.err | grep '^{' | tee gpurun_out/bench_${N}gpu_1.3B_480p.json
run 300 29522 bench.py --gpus $N --model A14B --video 720p --steps 8 --warmup 2 --no-cfg-split --no-vae 2>gpurun_out/bench_${N}gpu_a14b_u.err | grep '^{' | tee gpurun_out/bench_${N}gpu_A14B_720p_ulysses${N}.json
run 300 29523 bench.py --gpus $N --model A14B --video 720p --steps 8 --warmup 2 --no-vae 2>gpurun_out/bench_${N}gpu_a14b_c.err | grep '^{' | tee gpurun_out/bench_${N}gpu_A14B_720p_cfg2.json
if [ "$N" = "8" ]; then
  run 240 29524 bench.py --gpus $N --model A14B --video 480p --steps 8 --warmup 2 --no-vae 2>gpurun_out/bench_${N}gpu_a14b_480.err | grep '^{' | tee gpurun_out/bench_${N}gpu_A14B_480p_cfg2.json
fi
run 240 29525 scripts/vae_strip_check.py --config5 2>gpurun_out/vae_c5_${N}gpu.err | grep '^{' | tee gpurun_out/vae_config5_${N}gpu.json
for f in gpurun_out/bench_${N}gpu*.err gpurun_out/vae_c5_${N}gpu.err; do echo "== $f"; tail -n 2 $f | cut -c1-300; done
