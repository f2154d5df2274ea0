mkdir -p gpurun_out
nvidia-smi -L
echo "=== reference arm"; timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-rays 1024 2>&1 | grep -E '^\{' | cut -c1-400
echo "=== ours 2 gpus"; timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; grep -E '^\{' gpurun_out/bench_2gpu.json | cut -c1-1200; tail -5 gpurun_out/bench_2gpu.err
