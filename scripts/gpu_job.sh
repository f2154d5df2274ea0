mkdir -p gpurun_out
echo "=== bf16 gemm check"; timeout -s KILL 100 python scripts/tgemm_bf16_check.py 2>&1 | grep -E "BAD|ALL OK|SOME|^bf16|torch|rror"
echo "=== train step bench bf16"; CFN_TRAIN_PRECISION=bf16 timeout -s KILL 90 python scripts/train_step_bench.py 2>&1 | tail -1
echo "=== train step bench tf32"; CFN_TRAIN_PRECISION=tf32 timeout -s KILL 90 python scripts/train_step_bench.py 2>&1 | tail -1
echo "=== gpu tests"; timeout -s KILL 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
