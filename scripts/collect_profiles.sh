#!/bin/bash
# Regenerates the evidence under profiles/ on a B200 box (one GPU; ~6 minutes).  Run from the repo root:
#   gpurun --timeout 1800 -- 'bash scripts/collect_profiles.sh'
# then, back in the build container, `bash scripts/collect_profiles.sh summarize` turns gpurun_out/ into profiles/.
set -u
G=gpurun_out; P=profiles; R=r01
if [ "${1:-}" = "summarize" ]; then
  cp $G/bench_$R.json $P/${R}_bench_1gpu.json
  cp $G/${R}_launches.csv $P/${R}_launches.csv
  python scripts/ncu_summary.py list $G/${R}_launches.csv $P/${R}_launch_summary.csv
  python scripts/ncu_summary.py rep $G/${R}_prof_k1.ncu-rep $P/${R}_prof_k1_summary.csv
  python scripts/ncu_summary.py rep $G/${R}_prof_k2.ncu-rep $P/${R}_prof_k2_summary.csv
  python scripts/ncu_summary.py rep $G/${R}_prof_tgemm_bf16_fwd.ncu-rep $P/${R}_prof_tgemm_bf16_summary.csv
  python scripts/ncu_summary.py rep $G/${R}_prof_tgemm_bf16_bwd.ncu-rep /tmp/_bwd.csv && cat /tmp/_bwd.csv >> $P/${R}_prof_tgemm_bf16_summary.csv
  cp $G/kernel_rooflines.json $P/${R}_kernel_rooflines.json
  cp $G/k1_timeline.txt $P/${R}_k1_timeline.txt
  cp $G/train_launches_bf16.csv $P/${R}_train_launches.csv
  cat $G/train_steps.json > $P/${R}_train_step_1gpu.json
  exit 0
fi
mkdir -p $G
T="timeout -s KILL"
$T 400 python bench.py --steps 5 --warmup 3 > $G/bench_$R.json 2> $G/bench.err
# launch list of the render leg (kernel shares must agree with the bench's CUDA-event attribution)
$T 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $G/${R}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > $G/ncu_list.log 2>&1
# full captures: K1, the streaming kernels, and the three GEMM flavours of one bf16 training step
$T 400 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 2 -c 1 -f -o $G/${R}_prof_k1 \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > $G/ncu_k1.log 2>&1
$T 400 ncu --set full --clock-control none --import-source on -k regex:"flow_composite_fwd|raw2outputs" -s 4 -c 2 -f -o $G/${R}_prof_k2 \
   python scripts/kernel_rooflines.py $G/kr_tmp.json > $G/ncu_k2.log 2>&1
# tgemm launches of a training step: 40 per step; #83 = forward of trunk layer 3 in step 3, #106/#107 = dgrad / wgrad
CFN_TRAIN_PRECISION=bf16 $T 200 ncu --set full --clock-control none --import-source on -k regex:tgemm_kernel -s 83 -c 1 -f \
   -o $G/${R}_prof_tgemm_bf16_fwd python scripts/train_step_bench.py > $G/ncu_g1.log 2>&1
CFN_TRAIN_PRECISION=bf16 $T 200 ncu --set full --clock-control none --import-source on -k regex:tgemm_kernel -s 106 -c 2 -f \
   -o $G/${R}_prof_tgemm_bf16_bwd python scripts/train_step_bench.py > $G/ncu_g2.log 2>&1
CFN_TRAIN_PRECISION=bf16 $T 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
   --log-file $G/train_launches_bf16.csv python scripts/train_step_bench.py > $G/ncu_train.log 2>&1
$T 300 python scripts/kernel_rooflines.py $G/kernel_rooflines.json
CFN_TC_PROFILE=1 $T 200 python scripts/k1_timeline.py > $G/k1_timeline.txt 2>&1
for p in bf16 tf32 fp32; do CFN_TRAIN_PRECISION=$p $T 120 python scripts/train_step_bench.py 2>&1 | tail -1; done > $G/train_steps.json
$T 100 python scripts/tgemm_bf16_check.py 2>&1 | grep -E "ALL OK|SOME|^bf16|torch" > $G/tgemm_bf16_timing.txt
$T 100 python scripts/tgemm_check.py 2>&1 | grep -E "ALL OK|SOME|^tf32|^fp32|torch" > $G/tgemm_timing.txt
