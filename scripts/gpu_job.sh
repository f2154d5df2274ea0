mkdir -p gpurun_out
echo "=== gpu tests"; timeout -s KILL 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
