cd /root/repo; mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 scripts/sp_check.py 2>&1 | grep -E "rank|Error|error" | tee gpurun_out/sp_check_2gpu.log
run 29512 scripts/vae_strip_check.py 2>&1 | grep -E "rank|Error|error" | tee gpurun_out/vae_strip_check_2gpu.log
timeout 300 python -m pytest tests/test_gpu_dit.py tests/test_gpu_vae.py -m gpu -q -k "two_gpus" 2>&1 | tail -3 | tee gpurun_out/pytest_2gpu.log
run 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_2gpu.err | grep '^{' | tee gpurun_out/bench_2gpu_cfg2.json
run 29515 scripts/vae_strip_check.py --config5 2>gpurun_out/vae_c5_2gpu.err | grep '^{' | tee gpurun_out/vae_config5_2gpu.json
tail -n 2 gpurun_out/bench_2gpu.err gpurun_out/vae_c5_2gpu.err
