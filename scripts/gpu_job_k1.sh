mkdir -p gpurun_out
T="timeout -s KILL"
$T 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
$T 200 python scripts/r2_k1_ab.py 2>&1 | tail -1
CFN_TC_BIAS_SIDE=0 $T 200 python scripts/r2_k1_ab.py 2>&1 | tail -1
CFN_TC_PROFILE=1 CFN_PRECISION=fp16 $T 120 python scripts/k1_timeline.py gpurun_out/k1_timeline_x.json 2>&1 | tail -16
