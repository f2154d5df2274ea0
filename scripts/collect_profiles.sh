#!/bin/bash
# Regenerates the round-2 evidence under profiles/ on a B200 box (one GPU; ~8 minutes).  Run from the repo root:
#   gpurun --timeout 2400 -- 'bash scripts/collect_profiles.sh'
# then, back in the build container, `bash scripts/collect_profiles.sh summarize` turns gpurun_out/ into profiles/.
set -u
G=gpurun_out; P=profiles; R=r02
if [ "${1:-}" = "summarize" ]; then
  cp $G/${R}_bench_1gpu.json $P/${R}_bench_1gpu.json
  cp $G/${R}_bench_fern_1gpu.json $P/${R}_bench_fern_1gpu.json
  cp $G/${R}_bench_lego_1gpu.json $P/${R}_bench_lego_1gpu.json
  cp $G/${R}_launches.csv $P/${R}_launches.csv
  python scripts/ncu_summary.py list $G/${R}_launches.csv $P/${R}_launch_summary.csv
  python scripts/ncu_summary.py rep $G/${R}_prof_k1.ncu-rep $P/${R}_prof_k1_summary.csv
  python scripts/ncu_summary.py rep $G/${R}_prof_k2.ncu-rep $P/${R}_prof_k2_summary.csv
  python scripts/ncu_summary.py rep $G/${R}_prof_k2k4_train.ncu-rep $P/${R}_prof_k2k4_train_summary.csv
  python scripts/ncu_summary.py rep $G/${R}_prof_raw2outputs.ncu-rep $P/${R}_prof_raw2outputs_summary.csv
  python scripts/ncu_summary.py rep $G/${R}_prof_tgemm_train.ncu-rep $P/${R}_prof_tgemm_train_summary.csv
  cp $G/${R}_train4096_launches.csv $P/${R}_train_launches.csv
  python scripts/ncu_summary.py list $G/${R}_train4096_launches.csv $P/${R}_train_launch_summary.csv
  python scripts/ncu_summary.py list $G/${R}_train512_launches.csv $P/${R}_train512_launch_summary.csv
  cp $G/r2_train_bench.json $P/${R}_train_bench.json
  grep -q "tile total" $G/${R}_k1_timeline.txt 2>/dev/null && cp $G/${R}_k1_timeline.txt $P/${R}_k1_timeline.txt
  exit 0
fi
mkdir -p $G
T="timeout -s KILL"
$T 600 python bench.py --steps 5 --warmup 3 > $G/${R}_bench_1gpu.json 2> $G/bench.err
$T 600 python bench.py --config fern --steps 3 --warmup 3 --cpu-rays 128 > $G/${R}_bench_fern_1gpu.json 2> $G/bench_fern.err
$T 600 python bench.py --config lego --steps 3 --warmup 3 --cpu-rays 512 > $G/${R}_bench_lego_1gpu.json 2> $G/bench_lego.err
# launch list of the render leg (kernel shares must agree with the bench's CUDA-event attribution)
$T 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $G/${R}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-kernels > $G/ncu_list.log 2>&1
# full captures: K1 and K2 of the render leg, raw2outputs, K2 (train flavour) + K4 of a 4096-ray training step
$T 400 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 2 -c 1 -f -o $G/${R}_prof_k1 \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-kernels > $G/ncu_k1.log 2>&1
$T 400 ncu --set full --clock-control none --import-source on -k regex:flow_composite_fwd -s 2 -c 1 -f -o $G/${R}_prof_k2 \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-kernels > $G/ncu_k2.log 2>&1
$T 400 ncu --set full --clock-control none --import-source on -k regex:raw2outputs -s 2 -c 1 -f -o $G/${R}_prof_raw2outputs \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > $G/ncu_r2o.log 2>&1
CFN_RAYS=4096 CFN_STEPS=2 $T 300 ncu --set full --clock-control none --import-source on -k regex:flow_composite -s 2 -c 2 -f \
   -o $G/${R}_prof_k2k4_train python scripts/r2_step512.py > $G/ncu_k4.log 2>&1
# forward / dgrad / wgrad GEMMs of one trunk layer inside the bf16 training step
CFN_RAYS=4096 CFN_STEPS=2 $T 300 ncu --set full --clock-control none -k regex:tgemm_kernel -s 75 -c 8 -f \
   -o $G/${R}_prof_tgemm_train python scripts/r2_step512.py > $G/ncu_tg.log 2>&1
CFN_RAYS=4096 CFN_STEPS=3 $T 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $G/${R}_train4096_launches.csv python scripts/r2_step512.py > $G/ncu_t4096.log 2>&1
CFN_RAYS=512 CFN_STEPS=3 $T 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file $G/${R}_train512_launches.csv python scripts/r2_step512.py > $G/ncu_t512.log 2>&1
$T 400 python scripts/r2_train_bench.py > $G/r2_train_bench.log 2>&1
# (needs the diagnostic build of the library, -DCFN_TC_TIMELINE=1; with the shipped one it only prints how to build it)
CFN_TC_PROFILE=1 CFN_PRECISION=fp16 $T 200 python scripts/k1_timeline.py $G/k1_timeline.json > $G/${R}_k1_timeline_new.txt 2>&1
grep -q "tile total" $G/${R}_k1_timeline_new.txt && mv $G/${R}_k1_timeline_new.txt $G/${R}_k1_timeline.txt
for f in $G/bench.err $G/bench_fern.err $G/bench_lego.err; do tail -n 2 $f; done
cut -c1-400 $G/${R}_bench_fern_1gpu.json; echo; cut -c1-400 $G/${R}_bench_lego_1gpu.json
