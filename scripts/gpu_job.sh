mkdir -p gpurun_out
echo "=== bench"; timeout -s KILL 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_r01.json; tail -2 gpurun_out/bench.err
echo "=== ncu launch list (render)"
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log | cut -c1-200
echo "=== ncu full K1"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 2 -c 1 -f -o gpurun_out/r01_prof_k1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
echo "=== ncu full K2 + raw2outputs"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"flow_composite_fwd|raw2outputs" -s 4 -c 2 -f -o gpurun_out/r01_prof_k2 python scripts/kernel_rooflines.py gpurun_out/kr_tmp.json > gpurun_out/ncu_full2.log 2>&1
tail -1 gpurun_out/ncu_full2.log | cut -c1-200
echo "=== ncu full tgemm fwd (train step 3, trunk layer 3)"
CFN_TRAIN_PRECISION=tf32 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tgemm_kernel -s 83 -c 1 -f -o gpurun_out/r01_prof_tgemm_fwd python scripts/train_step_bench.py > gpurun_out/ncu_full3.log 2>&1
tail -1 gpurun_out/ncu_full3.log | cut -c1-200
echo "=== ncu full tgemm dgrad + wgrad"
CFN_TRAIN_PRECISION=tf32 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:tgemm_kernel -s 106 -c 2 -f -o gpurun_out/r01_prof_tgemm_bwd python scripts/train_step_bench.py > gpurun_out/ncu_full4.log 2>&1
tail -1 gpurun_out/ncu_full4.log | cut -c1-200
echo "=== kernel rooflines"; timeout -s KILL 300 python scripts/kernel_rooflines.py gpurun_out/kernel_rooflines.json 2>&1 | tail -3
echo "=== k1 timeline"; CFN_TC_PROFILE=1 timeout -s KILL 200 python scripts/k1_timeline.py > gpurun_out/k1_timeline.txt 2>&1; tail -5 gpurun_out/k1_timeline.txt
echo "=== tgemm timing"; timeout -s KILL 90 python scripts/tgemm_check.py 2>&1 | grep -E "ALL OK|SOME|^tf32|^fp32|torch" > gpurun_out/tgemm_timing.txt; cat gpurun_out/tgemm_timing.txt
echo "=== train step"; CFN_TRAIN_PRECISION=tf32 timeout -s KILL 90 python scripts/train_step_bench.py 2>&1 | tail -1 > gpurun_out/train_step_tf32.json; cat gpurun_out/train_step_tf32.json
CFN_TRAIN_PRECISION=fp32 timeout -s KILL 90 python scripts/train_step_bench.py 2>&1 | tail -1 > gpurun_out/train_step_fp32.json; cat gpurun_out/train_step_fp32.json
