mkdir -p gpurun_out
T="timeout -s KILL"
run() {
  $T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-kernels > gpurun_out/r02_bench_8gpu_$1.json 2> gpurun_out/bench8_$1.err
  tail -1 gpurun_out/bench8_$1.err
}
CFN_OVERLAP_ALLREDUCE=1 run overlap 29541
CFN_OVERLAP_ALLREDUCE=0 run nooverlap 29542
python - <<'PY'
import json
def load(p):
    for l in open(p):
        if l.startswith('{'): return json.loads(l)
for tag in ('overlap','nooverlap'):
    d=load(f'gpurun_out/r02_bench_8gpu_{tag}.json')
    print(tag, d['n_gpus'], round(d['value']), round(d['e2e']['value']))
    for k in ('train_step','train_step_strong'):
        t=d[k]; print('  ',k, round(t['value']), round(t['ms_per_step'],3), t['rays_per_gpu'], t['ranks_hold_identical_weights'])
PY
