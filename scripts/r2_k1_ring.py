"""Diagnostic: K1's per-K-chunk MMA issue cadence (clock64 stamps of the issuer warp of CTA 0) as a function of the
weight-ring depth.  CFN_W = netwidth, CFN_TC_STAGES caps the ring.  Needs the diagnostic build of the library
(CFN_NVCC_EXTRA=-DCFN_TC_TIMELINE=1 python -m cfnerf_b200.build --force); run on the GPU box with CFN_TC_PROFILE=1."""
import ctypes as C
import json
import os
import statistics
import sys

os.environ.setdefault("CFN_TC_PROFILE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import cfnerf_b200 as cf
from oracle import cfnerf_oracle as O

dev = torch.device("cuda:0")
cfg = O.CfnConfig(W=int(os.environ.get("CFN_W", "512")))
net = cf.NeRFFlowsParams.from_oracle_params(cfg, O.make_params(cfg, 0), *O.make_latents(cfg, 0)).to(dev)
eng = cf.engine_for(net, dev, os.environ.get("CFN_PRECISION", "fp16"))
B = 148 * 4
rays = O.synthetic_rays(B, 1).to(dev)
z = eng.zvals(rays, cf.reference_t_schedule(128, dev), None, False)
for _ in range(2):
    eng.network(B, 128, rays=rays, z_vals=z)
torch.cuda.synchronize()
N = 4096
buf = (C.c_uint64 * (4 * N))()
cf._lib.check(eng.lib.cfn_debug_profile(eng.h, buf, 4 * N))
mma = [x for x in buf[N:2 * N] if x]
tma = [x for x in buf[2 * N:3 * N] if x]
land = [x for x in buf[3 * N:4 * N] if x]
d = [b - a for a, b in zip(mma, mma[1:])]
d = [x for x in d[len(d) // 2:] if x < 3000]
lat = [b - a for a, b in zip(tma, land)] if len(tma) == len(land) else [0]
# timed run of a big launch
Bb = 32768
raysb = O.synthetic_rays(Bb, 1).to(dev)
zb = eng.zvals(raysb, cf.reference_t_schedule(128, dev), None, False)
os.environ.pop("CFN_TC_PROFILE", None)
for _ in range(3):
    eng.network(Bb, 128, rays=raysb, z_vals=zb)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8):
    eng.network(Bb, 128, rays=raysb, z_vals=zb)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"W": cfg.W, "stages_env": os.environ.get("CFN_TC_STAGES"), "chunk_cadence_median": statistics.median(d),
                  "chunk_cadence_mean": sum(d) / len(d), "tma_issue_to_land_median": statistics.median(lat),
                  "ms_per_32768_rays": e0.elapsed_time(e1) / 8}))
