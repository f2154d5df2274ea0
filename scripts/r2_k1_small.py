import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cfnerf_b200 as cf
from oracle import cfnerf_oracle as O
dev = torch.device("cuda:0")
cfg = O.CfnConfig()
B = int(os.environ.get("CFN_RAYS", "64"))
rays = O.synthetic_rays(B, 1).to(dev)
net = cf.NeRFFlowsParams.from_oracle_params(cfg, O.make_params(cfg, 0), *O.make_latents(cfg, 0)).to(dev)
eng = cf.engine_for(net, dev, "fp16")
z = eng.zvals(rays, cf.reference_t_schedule(128, dev), None, False)
fp = eng.network(B, 128, rays=rays, z_vals=z)
torch.cuda.synchronize()
ref = cf.engine_for(net, dev, "fp32").network(B, 128, rays=rays, z_vals=z)
print("max err", float((fp - ref).abs().max()))
