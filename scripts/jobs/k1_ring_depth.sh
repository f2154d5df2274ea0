mkdir -p gpurun_out
T="timeout -s KILL 120"
: > gpurun_out/r2_k1_ring.txt
for cfgs in "512 4" "512 3" "512 2" "256 8" "256 6" "256 5" "256 4" "256 3"; do
  set -- $cfgs
  CFN_W=$1 CFN_TC_STAGES=$2 CFN_TC_PROFILE=1 $T python scripts/r2_k1_ring.py 2>&1 | tail -1 | tee -a gpurun_out/r2_k1_ring.txt
done
