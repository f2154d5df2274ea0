mkdir -p gpurun_out
echo "=== tf32 network/training check"; timeout -s KILL 600 python scripts/tf32_train_check.py 2>&1 | tail -40
echo "=== train step bench tf32"; CFN_TRAIN_PRECISION=tf32 timeout -s KILL 300 python scripts/train_step_bench.py 2>&1 | tail -3
echo "=== train step bench fp32"; CFN_TRAIN_PRECISION=fp32 timeout -s KILL 300 python scripts/train_step_bench.py 2>&1 | tail -3
echo "=== gpu tests"; timeout -s KILL 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
