mkdir -p gpurun_out
echo "=== smoke"; timeout -s KILL 600 python __graft_entry__.py smoke 2>&1 | tail -16
echo "=== bench"; timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_tmp.json 2> gpurun_out/bench.err; cut -c1-2500 gpurun_out/bench_tmp.json; tail -3 gpurun_out/bench.err
echo "=== gpu tests"; timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
