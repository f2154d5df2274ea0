mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "=== ours 2 gpus"; timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; grep -E '^\{' gpurun_out/bench_2gpu.json | cut -c1-2600; tail -3 gpurun_out/bench_2gpu.err
