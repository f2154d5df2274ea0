# K2 training forward at small batches: four warps per ray (default) vs one (CFN_K2_SEG4=0)
T="timeout -s KILL"
$T 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 1 0 1 0; do
  echo "CFN_K2_SEG4=$v"
  CFN_K2_SEG4=$v $T 300 python scripts/r2_train_bench.py 2>&1 | grep -E 'k2_train_fwd_ms_512|"rays": 512, "graph": true|depth'
done
