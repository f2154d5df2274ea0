mkdir -p gpurun_out
echo "=== TC tests"; timeout -s KILL 300 python -m pytest tests -m gpu -q -k "tensor_core or config" 2>&1 | tail -4
echo "=== TC tests cg1"; CFN_TC_CTA_GROUP=1 timeout -s KILL 300 python -m pytest tests -m gpu -q -k "tensor_core" 2>&1 | tail -3
CFN_TC_PROFILE=1 timeout -s KILL 300 python scripts/k1_timeline.py gpurun_out/k1_timeline_cg2.json | tail -16
python scripts/k1_mma_timeline.py gpurun_out/k1_timeline_cg2.json 4,18,18,18,18,20,18,18,8,18,9,4 | sed -n 2,4p
echo "=== bench"; timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tmp.json 2> gpurun_out/bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_tmp.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"roofline",d["roofline"]["achieved"],d["roofline"]["frac"],"k1share",d["roofline"]["k1_share_of_step"],d["clocks"])
PY
tail -3 gpurun_out/bench.err
