mkdir -p gpurun_out
echo "=== train step bench bf16"; CFN_TRAIN_PRECISION=bf16 timeout -s KILL 90 python scripts/train_step_bench.py 2>&1 | tail -1
echo "=== gpu tests"; timeout -s KILL 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
