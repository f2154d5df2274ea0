mkdir -p gpurun_out
echo "=== bf16 gemm check"; timeout -s KILL 100 python scripts/tgemm_bf16_check.py 2>&1 | tail -60
