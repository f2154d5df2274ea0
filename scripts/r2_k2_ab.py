"""K2 (flow stacks + compositing) timing, render and training flavours, K = 32 and 128, with a parity check against the
fp32 check mode.  CFN_AB_LIB=<other .so> times another build of the library on the same box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import cfnerf_b200 as cf
if os.environ.get("CFN_AB_LIB"):
    import cfnerf_b200._lib as _L
    _L.LIB_PATH = os.environ["CFN_AB_LIB"]
from oracle import cfnerf_oracle as O

dev = torch.device("cuda:0")
res = {"lib": os.environ.get("CFN_AB_LIB", "tree")}
for K in (32, 128):
    cfg = O.CfnConfig(K=K)
    sa, sr = O.make_latents(cfg, 0)
    net = cf.NeRFFlowsParams.from_oracle_params(cfg, O.make_params(cfg, 0, "stressed"), sa, sr).to(dev)
    eng = cf.engine_for(net, dev, "fp16")
    B, N = 16384, 128
    rays = O.synthetic_rays(B, 1).to(dev)
    z = eng.zvals(rays, cf.reference_t_schedule(N, dev), None, False)
    fp = eng.network(B, N, rays=rays, z_vals=z)
    ea, er = cf.test_latents(net, dev)
    ea, er = ea.reshape(-1).contiguous(), er.contiguous()
    def run(train):
        return eng.flow_composite(fp, z, rays[:, 3:6], 11, ea, er, False, train=train, want_trans=train)
    for train in (False, True):
        for _ in range(3): out = run(train)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): out = run(train)
        e1.record(); torch.cuda.synchronize()
        res[f"K{K}_{'train' if train else 'render'}_ms"] = e0.elapsed_time(e1) / 10
    ref = cf.engine_for(net, dev, "fp32").flow_composite(fp[:256 * N], z[:256], rays[:256, 3:6], 11, ea, er, False)
    out = run(False)
    res[f"K{K}_max_err_vs_accurate"] = max(float((out[k][:256] - ref[k]).abs().max()) for k in ("rgb_map", "depth_map"))
print(json.dumps(res))
