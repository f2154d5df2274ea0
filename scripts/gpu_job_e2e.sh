mkdir -p gpurun_out
T="timeout -s KILL"
$T 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "host_pipeline or sharding" 2>&1 | tail -3
$T 600 python bench.py --steps 5 --warmup 3 --no-train --no-kernels --no-cpu-baseline > gpurun_out/e2e_africa.json 2> gpurun_out/e2e_a.err; tail -2 gpurun_out/e2e_a.err
$T 600 python bench.py --config lego --steps 3 --warmup 3 --no-cpu-baseline --no-kernels > gpurun_out/e2e_lego.json 2> gpurun_out/e2e_l.err; tail -2 gpurun_out/e2e_l.err
$T 600 python bench.py --config fern --steps 3 --warmup 3 --no-cpu-baseline --no-kernels > gpurun_out/e2e_fern.json 2> gpurun_out/e2e_f.err; tail -2 gpurun_out/e2e_f.err
python - <<'PY'
import json
def load(p):
    for l in open(p):
        if l.startswith('{'): return json.loads(l)
for c in ('africa','lego','fern'):
    d=load(f'gpurun_out/e2e_{c}.json'); print(c, round(d['value']), round(d['e2e']['value']), d['e2e']['d2h_bytes_per_step'])
PY
