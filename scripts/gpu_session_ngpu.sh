# usage: bash scripts/gpu_session_ngpu.sh N [extra bench flags]
N=$1; shift
# hard wall-clock cap: a hung multi-GPU job is charged N x its duration
cd /root/repo
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 "$@" 2>gpurun_out/bench_${N}gpu.err | tee gpurun_out/bench_${N}gpu.json
tail -n 3 gpurun_out/bench_${N}gpu.err
