mkdir -p gpurun_out
T="timeout -s KILL"
run() { # $1 = tag, env passed through
  $T 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-kernels > gpurun_out/r02_bench_2gpu_$1.json 2> gpurun_out/bench2_$1.err
  tail -2 gpurun_out/bench2_$1.err
}
CFN_OVERLAP_ALLREDUCE=0 run nooverlap 29531
CFN_OVERLAP_ALLREDUCE=1 run overlap 29532
python - <<'PY'
import json
def load(p):
    for l in open(p):
        if l.startswith('{'): return json.loads(l)
for tag in ('nooverlap','overlap'):
    d=load(f'gpurun_out/r02_bench_2gpu_{tag}.json')
    print(tag, d['n_gpus'], round(d['value']), round(d['e2e']['value']))
    for k in ('train_step','train_step_strong','train_step_512'):
        t=d[k]; print('  ',k, round(t['value']), round(t['ms_per_step'],3), t['rays_per_gpu'], t['ranks_hold_identical_weights'])
PY
