# same-box A/B of two builds of the library: scripts/probe/libcfn_base.so (built from another revision) vs the tree's
mkdir -p gpurun_out
T="timeout -s KILL"
for i in 1 2; do
CFN_AB_LIB=scripts/probe/libcfn_base.so $T 200 python scripts/r2_k1_ab.py 2>&1 | tail -1 | cut -c1-130
$T 200 python scripts/r2_k1_ab.py 2>&1 | tail -1 | cut -c1-130
done
CFN_TC_PROFILE=1 CFN_PRECISION=fp16 $T 120 python scripts/k1_timeline.py gpurun_out/k1_timeline_x.json 2>&1 | tail -3
