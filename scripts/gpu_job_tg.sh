mkdir -p gpurun_out
T="timeout -s KILL"
$T 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -m gpu -x -q -k "gemm or training or train or fused" 2>&1 | tail -3
$T 200 python scripts/tgemm_bf16_check.py 2>&1 | tail -12
$T 300 python scripts/r2_train_bench.py 2>&1 | tail -15
