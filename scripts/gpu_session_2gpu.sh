cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/sp_check.py 2>&1 | grep rank
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --no-vae 2>gpurun_out/bench2.err | tee gpurun_out/bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 3 --cfg-split --no-vae 2>gpurun_out/bench2c.err | tee gpurun_out/bench_2gpu_cfgsplit.json
tail -3 gpurun_out/bench2.err gpurun_out/bench2c.err
CUDA_VISIBLE_DEVICES=0 python bench.py --model A14B --steps 1 --warmup 1 --no-cpu-baseline 2>gpurun_out/bench_a14b.err | tee gpurun_out/bench_a14b_480p.json
tail -3 gpurun_out/bench_a14b.err
