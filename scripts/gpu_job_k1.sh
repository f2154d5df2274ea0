mkdir -p gpurun_out
T="timeout -s KILL"
for i in 1 2; do
CFN_AB_LIB=scripts/probe/libcfn_base.so $T 200 python scripts/r2_k1_ab.py 2>&1 | tail -1 | cut -c1-150
$T 200 python scripts/r2_k1_ab.py 2>&1 | tail -1 | cut -c1-150
CFN_TC_BIAS_SIDE=0 $T 200 python scripts/r2_k1_ab.py 2>&1 | tail -1 | cut -c1-150
done
