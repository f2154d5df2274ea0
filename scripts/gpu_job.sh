mkdir -p gpurun_out
echo "=== ncu launch list"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log | cut -c1-300
echo "=== ncu full on K1"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 2 -c 1 -o gpurun_out/r01_prof_k1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
echo "=== ncu full on K2 + raw2outputs"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"flow_composite_fwd|raw2outputs" -s 4 -c 2 -o gpurun_out/r01_prof_k2 python scripts/kernel_rooflines.py gpurun_out/kr_tmp.json > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log | cut -c1-200
ls -la gpurun_out | head -30
