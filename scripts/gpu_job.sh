mkdir -p gpurun_out
echo "=== all gpu tests"; timeout -s KILL 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "passed|failed|FAILED|stressed|Error" | head -20
echo "=== train step"; timeout -s KILL 600 python scripts/train_step_bench.py 2>&1 | tail -3
