"""Achieved bandwidth / rate of the per-ray streaming kernels against the measured HBM peak (run on the GPU box).
Writes one JSON object per kernel; algorithmic bytes are the SURVEY §8(d) figures (inputs read once + outputs)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import cfnerf_b200 as cf
from cfnerf_b200.engine import _ptr, _stream
from oracle import cfnerf_oracle as O

dev = torch.device("cuda:0")
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
HBM = peaks["hbm_gbs"]
lib = cf._lib.load()


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


out = []
N, K = 128, 32
# ---- K2': raw2outputs, inputs (4.3 GB at 65536 rays) far larger than L2 ----
B = 65536
raw = torch.randn(B, N, K, 4, device=dev)
z = (1.2 + torch.cumsum(torch.rand(B, N, device=dev) * 0.05 + 0.01, -1)).contiguous()
d = torch.randn(B, 3, device=dev)
f32 = dict(dtype=torch.float32, device=dev)
rgb, disp, depth, w = torch.empty(B, 3, K, **f32), torch.empty(B, K, **f32), torch.empty(B, K, **f32), torch.empty(B, N, K, **f32)
for with_w in (False, True):
    def run():
        cf._lib.check(lib.cfn_raw2outputs_f32(_ptr(raw), _ptr(z), _ptr(d), 3, 0, _ptr(rgb), _ptr(disp), _ptr(w if with_w else None),
                                              _ptr(depth), B, N, K, _stream()))
    t = timeit(run)
    bytes_ray = 16 * N * K + 4 * N + 12 + 20 * K + (4 * N * K if with_w else 0)
    gbs = B * bytes_ray / t / 1e9
    out.append({"kernel": "raw2outputs_kernel" + ("<weights>" if with_w else ""), "rays": B, "ms": t * 1e3, "bytes_per_ray": bytes_ray,
                "achieved_gbs": gbs, "hbm_peak_gbs": HBM, "frac": gbs / HBM, "rays_per_s": B / t})
del raw, w
# ---- K3: sample_pdf ----
B3, M, Nf = 1 << 20, 63, 128
bins = torch.sort(torch.rand(B3, M, device=dev) * 5 + 1, -1).values
ww = torch.rand(B3, M - 1, device=dev)
u = torch.rand(B3, Nf, device=dev)
smp = torch.empty(B3, Nf, **f32)
t = timeit(lambda: cf._lib.check(lib.cfn_sample_pdf_f32(_ptr(bins), _ptr(ww), _ptr(u), _ptr(smp), None, B3, M, Nf, _stream())))
bytes_ray = 4 * (M + (M - 1) + Nf) + 4 * Nf
out.append({"kernel": "sample_pdf_kernel", "rays": B3, "ms": t * 1e3, "bytes_per_ray": bytes_ray, "achieved_gbs": B3 * bytes_ray / t / 1e9,
            "hbm_peak_gbs": HBM, "frac": B3 * bytes_ray / t / 1e9 / HBM, "rays_per_s": B3 / t})
del bins, ww, u, smp
# ---- K2: flows + compositing fed by the network stage's parameter records ----
cfg = O.CfnConfig()
net = cf.NeRFFlowsParams.from_oracle_params(cfg, O.make_params(cfg, 0, "lively"), *O.make_latents(cfg, 0)).to(dev)
eng = cf.engine_for(net, dev, "bf16")
B2 = 32768
rays = O.synthetic_rays(B2, 1).to(dev)
zz = eng.zvals(rays, cf.reference_t_schedule(N, dev), None, False)
fp = eng.network(B2, N, rays=rays, z_vals=zz)
ea, er = cf.test_latents(net, dev)
t = timeit(lambda: eng.flow_composite(fp, zz, rays[:, 3:6], 11, ea, er, False, want_kstats=True))
bytes_ray = 72 * 4 * N + 4 * N + 44 + 20 * K + 32
out.append({"kernel": "flow_composite_fwd_kernel", "rays": B2, "ms": t * 1e3, "bytes_per_ray": bytes_ray,
            "achieved_gbs": B2 * bytes_ray / t / 1e9, "hbm_peak_gbs": HBM, "frac": B2 * bytes_ray / t / 1e9 / HBM,
            "rays_per_s": B2 / t, "note": "transcendental-bound (16 tanh + 3 sigmoid + softplus + exp per (point,k)), not HBM-bound"})
t = timeit(lambda: eng.network(B2, N, rays=rays, z_vals=zz), iters=5)
out.append({"kernel": "mlp_tc_kernel", "rays": B2, "ms": t * 1e3, "tflops": B2 * N * 4708864 / t / 1e12, "rays_per_s": B2 / t})
for o in out:
    print(json.dumps(o))
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/kernel_rooflines.json"
os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
json.dump(out, open(path, "w"), indent=1)
