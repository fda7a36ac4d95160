#!/bin/bash
# One GPU-box session: parity tests, kernel microbench, bench line, ncu launch list + full capture.
# usage: scripts/gpu_session.sh [tests] [micro] [bench] [ref] [ncu]   (default: all)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what="${*:-tests micro bench ref ncu}"
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv | tee gpurun_out/gpu.txt
has() { [[ " $what " == *" $1 "* ]]; }

if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
fi
if has micro; then
  timeout 600 python scripts/gpu_microbench.py --json gpurun_out/microbench.json 2>&1 | tee gpurun_out/microbench.log
fi
if has bench; then
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "ref rc=$?"; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
fi
if has ncu; then
  CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
  $CMD > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1200 --csv \
      --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
  echo "ncu list rc=$?"
  $CMD > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'attn_fwd|gemm_bf16' -c 9 \
      -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
if has ncuk; then
  # single-kernel full captures (scripts/ncu_target.py): attention, GEMM (+fp32 residual epilogue), VAE conv, VAE norm
  for k in ${NCU_TARGETS:-attn_b2 gemm gemm_resid conv96 conv192 norm ln rope}; do
    python scripts/ncu_target.py $k > gpurun_out/plain_$k.log 2>&1 &&
    ncu --set full --clock-control none --import-source on --profile-from-start off -f \
        -o gpurun_out/prof_$k python scripts/ncu_target.py $k > gpurun_out/ncu_$k.log 2>&1
    echo "ncu $k rc=$?"
  done
fi
if has vae; then
  python scripts/gpu_vae_bench.py --video 480p --encode 2>&1 | tee gpurun_out/vae_bench_480p.log | head -30
fi
