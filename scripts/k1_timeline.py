"""Diagnostic: per-step timeline of the tensor-core network kernel (CTA 0) from its clock64() stamps.
The stamps are compiled OUT of the shipped library (even predicated off they cost 1.8 % of the kernel); build the
diagnostic library first:   CFN_NVCC_EXTRA=-DCFN_TC_TIMELINE=1 python -m cfnerf_b200.build --force
then on the GPU box:         CFN_TC_PROFILE=1 python scripts/k1_timeline.py [out.json]"""
import ctypes as C
import json
import os
import sys

os.environ.setdefault("CFN_TC_PROFILE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import cfnerf_b200 as cf
from oracle import cfnerf_oracle as O

dev = torch.device("cuda:0")
cfg = O.CfnConfig()
net = cf.NeRFFlowsParams.from_oracle_params(cfg, O.make_params(cfg, 0), *O.make_latents(cfg, 0)).to(dev)
eng = cf.engine_for(net, dev, os.environ.get("CFN_PRECISION", "bf16"))
n_tiles = 4
rays = O.synthetic_rays(148 * n_tiles, 1).to(dev)   # 148 CTAs x n_tiles tiles of 128 points (N=128: one ray per tile)
t = cf.reference_t_schedule(128, dev)
z = eng.zvals(rays, t, None, False)
for _ in range(2):
    eng.network(rays.shape[0], 128, rays=rays, z_vals=z)
torch.cuda.synchronize()
N = 4096
buf = (C.c_uint64 * (4 * N))()
cf._lib.check(eng.lib.cfn_debug_profile(eng.h, buf, 4 * N))
roles = {"epilogue": list(buf[0:N]), "mma": list(buf[N:2 * N]), "tma": list(buf[2 * N:3 * N]), "detail": list(buf[3 * N:4 * N])}
roles = {k: [x for x in v if x] for k, v in roles.items()}
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/k1_timeline.json"
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump(roles, open(out, "w"))
ep = roles["epilogue"]
if not ep:
    sys.exit("no stamps: this library was built without -DCFN_TC_TIMELINE=1 (see the docstring)")
# epilogue stamps per tile: tile start, after encode, then per step (after acc_full wait, after epilogue) -> 2 + 2*steps
n_steps = 12
per_tile = 2 + 2 * n_steps
t0 = ep[0]
print("epilogue warp timeline (cycles relative to kernel start), last profiled tile:")
k = (len(ep) // per_tile - 1) * per_tile
base = ep[k]
print(f" tile start {ep[k]-t0}, encode {ep[k+1]-ep[k]}")
prev = ep[k + 1]
for g in range(n_steps):
    a, b = ep[k + 2 + 2 * g], ep[k + 3 + 2 * g]
    print(f" step {g:2d}: waited acc_full {a-prev:7d}  epilogue {b-a:7d}   (t={a-base})")
    prev = b
print(f" tile total {prev-base}")

det = roles["detail"]
if det:
    print("kind-2 epilogue detail (last 11 stamps, deltas):", [det[i + 1] - det[i] for i in range(len(det) - 11, len(det) - 1)])
