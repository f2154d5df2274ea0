mkdir -p gpurun_out
echo "=== TC tests cg2"; CFN_TC_CTA_GROUP=2 timeout -s KILL 300 python -m pytest tests -m gpu -q -k "tensor_core" 2>&1 | tail -5
CFN_TC_PROFILE=1 CFN_TC_CTA_GROUP=2 timeout -s KILL 300 python scripts/k1_timeline.py gpurun_out/k1_timeline_cg2.json
python scripts/k1_mma_timeline.py gpurun_out/k1_timeline_cg2.json 4,18,18,18,18,20,18,18,8,18,9,4
for cg in 2; do
echo "=== bench cta_group $cg"
CFN_TC_CTA_GROUP=$cg timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cg$cg.json 2> gpurun_out/bench_cg$cg.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_cg$cg.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"roofline",d["roofline"]["achieved"],d["roofline"]["frac"],"k1share",d["roofline"]["k1_share_of_step"],d["clocks"])
PY
tail -3 gpurun_out/bench_cg$cg.err
done
