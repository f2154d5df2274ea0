# same-box A/B of the training step: scripts/probe/libcfn_base.so (another revision) vs the tree's build
T="timeout -s KILL"
for i in 1 2; do
echo base; CFN_AB_LIB=scripts/probe/libcfn_base.so $T 300 python scripts/r2_train_bench.py 2>&1 | grep -E '"rays": (512|4096), "graph": true'
echo tree; $T 300 python scripts/r2_train_bench.py 2>&1 | grep -E '"rays": (512|4096), "graph": true'
done
