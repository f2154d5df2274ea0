"""One FusedTrainStep shape, a few eager steps (launch list / ncu captures of the training kernels).
  CFN_RAYS (512), CFN_TRAIN_PRECISION (bf16), CFN_STEPS (3)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import cfnerf_b200 as cf
from cfnerf_b200 import dist as D
from oracle import cfnerf_oracle as O

dev = torch.device("cuda:0")
cfg = O.CfnConfig()
B = int(os.environ.get("CFN_RAYS", "512"))
net = cf.NeRFFlowsParams.from_oracle_params(cfg, O.make_params(cfg, 0), *O.make_latents(cfg, 0)).to(dev)
tr = D.FusedTrainStep(net, lr=5e-4, precision=os.environ.get("CFN_TRAIN_PRECISION", "bf16"), use_graph=False)
rays = O.synthetic_rays(B, 1).to(dev)
target = torch.rand(B, 3, generator=torch.Generator().manual_seed(2)).to(dev)
for _ in range(int(os.environ.get("CFN_STEPS", "3"))):
    tr.step(rays, target, want_loss=False)
torch.cuda.synchronize()
print("done")
