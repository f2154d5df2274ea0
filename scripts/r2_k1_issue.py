"""Diagnostic: where the MMA issuer of K1 spends its time per ring slot (three stamps per slot: loop top, after the
waits, after the issue).  Needs the diagnostic build of the library:
    CFN_NVCC_EXTRA='-DCFN_TC_TIMELINE=1 -DCFN_TC_ISSUE_STAMPS=1' python -m cfnerf_b200.build --force      (here; the .so travels with gpurun)
then on the GPU box:  CFN_TC_PROFILE=1 python scripts/r2_k1_issue.py"""
import ctypes as C, json, os, sys
os.environ.setdefault("CFN_TC_PROFILE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import cfnerf_b200 as cf
from oracle import cfnerf_oracle as O
dev = torch.device("cuda:0")
cfg = O.CfnConfig()
net = cf.NeRFFlowsParams.from_oracle_params(cfg, O.make_params(cfg, 0), *O.make_latents(cfg, 0)).to(dev)
eng = cf.engine_for(net, dev, "fp16")
B = 148 * 3
rays = O.synthetic_rays(B, 1).to(dev)
z = eng.zvals(rays, cf.reference_t_schedule(128, dev), None, False)
eng.network(B, 128, rays=rays, z_vals=z)
torch.cuda.synchronize()
N = 4096
buf = (C.c_uint64 * (4 * N))()
cf._lib.check(eng.lib.cfn_debug_profile(eng.h, buf, 4 * N))
mma = [x for x in buf[N:2 * N] if x]
json.dump({"mma": mma}, open("gpurun_out/k1_issue.json", "w"))
steps = [2] + [18] * 7 + [8, 18, 9, 4]
per_tile = sum(3 * s + 1 for s in steps)
print(len(mma), per_tile)
o = 2 * per_tile   # third tile
i = o
for si, sn in enumerate(steps):
    for k in range(sn):
        a0, b0, c0 = mma[i], mma[i + 1], mma[i + 2]
        nxt = mma[i + 3]
        if si in (1, 2, 8, 9, 10, 11):
            print(f"step {si:2d} slot {k:2d}: wait {b0-a0:5d} issue {c0-b0:5d} tail {nxt-c0:5d}   t={a0-mma[o]}")
        i += 3
    i += 1
