mkdir -p gpurun_out
echo "=== all gpu tests"; timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "=== kernel rooflines"; timeout -s KILL 600 python scripts/kernel_rooflines.py gpurun_out/kernel_rooflines.json
echo "=== bench"; timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench.err; cat gpurun_out/bench_r01.json; tail -3 gpurun_out/bench.err
