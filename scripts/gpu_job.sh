mkdir -p gpurun_out
echo "=== train step bench tf32"; CFN_TRAIN_PRECISION=tf32 timeout -s KILL 90 python scripts/train_step_bench.py 2>&1 | tail -1
echo "=== gpu tests"; timeout -s KILL 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "=== bench"; timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['train_step'])"
