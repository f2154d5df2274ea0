mkdir -p gpurun_out
echo "=== gpu tests"; timeout -s KILL 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
echo "=== bench"; timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac']); print(d['train_step'])"; tail -2 gpurun_out/bench.err
