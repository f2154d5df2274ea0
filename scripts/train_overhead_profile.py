"""Host-side overhead of one training step: a tiny batch (8 rays) makes the device work negligible, cProfile shows where
the per-step CPU time goes."""
import cProfile, os, pstats, sys, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cfnerf_b200 as cf
from cfnerf_b200 import dist as D
from cfnerf_b200.optim import FusedAdam
from oracle import cfnerf_oracle as O

dev = torch.device("cuda:0")
cfg = O.CfnConfig()
net = cf.NeRFFlowsParams.from_oracle_params(cfg, O.make_params(cfg, 0), *O.make_latents(cfg, 0)).to(dev)
params = [q for n, q in net.named_parameters() if not n.startswith("alpha_linear") and not n.startswith("alpha_std_linear")]
opt = FusedAdam(params, lr=5e-4)
R = int(os.environ.get("CFN_TRAIN_RAYS", "8"))
rays = O.synthetic_rays(R, 1).to(dev)
target = torch.rand(R, 3).to(dev)
prec = os.environ.get("CFN_TRAIN_PRECISION", "bf16")
for _ in range(5):
    D.train_step(net, opt, rays, target, None, precision=prec)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(50):
    D.train_step(net, opt, rays, target, None, precision=prec)
torch.cuda.synchronize()
print("ms per step at %d rays: %.3f" % (R, (time.perf_counter() - t0) / 50 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    D.train_step(net, opt, rays, target, None, precision=prec)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
