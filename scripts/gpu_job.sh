mkdir -p gpurun_out
echo "=== smoke"; timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | grep -E "smoke|rror" | tail -16
echo "=== bench"; timeout -s KILL 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench.err; cut -c1-160 gpurun_out/bench_r01.json; tail -2 gpurun_out/bench.err
echo "=== train steps"; for p in bf16 tf32; do CFN_TRAIN_PRECISION=$p timeout -s KILL 90 python scripts/train_step_bench.py 2>&1 | tail -1; done > gpurun_out/train_steps.json; cat gpurun_out/train_steps.json | cut -c1-200
