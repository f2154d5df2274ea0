mkdir -p gpurun_out
nvidia-smi -L | wc -l
T="timeout -s KILL"
$T 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/bench8.err
tail -3 gpurun_out/bench8.err
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --config fern --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --no-kernels > gpurun_out/r02_bench_fern_8gpu.json 2> gpurun_out/bench8f.err
$T 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --config lego --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline --no-kernels > gpurun_out/r02_bench_lego_8gpu.json 2> gpurun_out/bench8l.err
python - <<'PY'
import json
def load(p):
    for l in open(p):
        if l.startswith('{'): return json.loads(l)
d=load('gpurun_out/r02_bench_8gpu.json')
print(d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['frac'])
for k in ('train_step','train_step_strong','train_step_512'):
    t=d[k]; print(k, round(t['value']), round(t['ms_per_step'],3), t['rays_per_gpu'], t['ranks_hold_identical_weights'], t['scaling'])
for c in ('fern','lego'):
    f=load(f'gpurun_out/r02_bench_{c}_8gpu.json'); print(c, f['n_gpus'], f['value'], f['e2e']['value'], f['config']['rays_per_step_per_gpu'])
PY
