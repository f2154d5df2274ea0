# bench.py (africa) under torchrun on N GPUs:  gpurun --gpus N -- 'bash scripts/jobs/bench_ngpu.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/bench${N}.err
tail -n 2 gpurun_out/bench${N}.err
python - <<PY
import json
def load(p):
    for l in open(p):
        if l.startswith('{'): return json.loads(l)
d=load('gpurun_out/r02_bench_${N}gpu.json')
print(d['n_gpus'], round(d['value']), round(d['e2e']['value']), d['roofline']['frac'])
for k in ('train_step','train_step_strong','train_step_512'):
    t=d[k]; print(k, round(t['value']), round(t['ms_per_step'],3), t['rays_per_gpu'], t['ranks_hold_identical_weights'], t['scaling'])
PY
