mkdir -p gpurun_out
echo "=== bench"; timeout -s KILL 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_r01.json; tail -2 gpurun_out/bench.err
echo "=== ncu full bf16 tgemm fwd"
CFN_TRAIN_PRECISION=bf16 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:tgemm_kernel -s 83 -c 1 -f -o gpurun_out/r01_prof_tgemm_bf16_fwd python scripts/train_step_bench.py > gpurun_out/ncu_full3.log 2>&1
tail -1 gpurun_out/ncu_full3.log | cut -c1-200
echo "=== ncu full bf16 tgemm dgrad + wgrad"
CFN_TRAIN_PRECISION=bf16 timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:tgemm_kernel -s 106 -c 2 -f -o gpurun_out/r01_prof_tgemm_bf16_bwd python scripts/train_step_bench.py > gpurun_out/ncu_full4.log 2>&1
tail -1 gpurun_out/ncu_full4.log | cut -c1-200
echo "=== ncu launch list of the bf16 train step"
CFN_TRAIN_PRECISION=bf16 timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/train_launches_bf16.csv python scripts/train_step_bench.py > gpurun_out/train_ncu.log 2>&1
tail -1 gpurun_out/train_ncu.log | cut -c1-120
echo "=== train steps"; for p in bf16 tf32; do CFN_TRAIN_PRECISION=$p timeout -s KILL 90 python scripts/train_step_bench.py 2>&1 | tail -1; done > gpurun_out/train_steps.json; cat gpurun_out/train_steps.json
echo "=== bf16 gemm timing"; timeout -s KILL 100 python scripts/tgemm_bf16_check.py 2>&1 | grep -E "ALL OK|SOME|^bf16|torch" > gpurun_out/tgemm_bf16_timing.txt; cat gpurun_out/tgemm_bf16_timing.txt
