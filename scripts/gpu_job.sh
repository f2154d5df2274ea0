mkdir -p gpurun_out
echo "=== all gpu tests"; timeout -s KILL 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | head -20
