"""BASELINE.json configs[2]: one training step (render -> K-mean -> KDE-NLL + 0.01*entropy -> backward ->
gradient all-reduce -> Adam) on a 4096-ray global batch, data-parallel over the ranks of torchrun.
CFN_TRAIN_PRECISION selects the GEMM engine of the step: bf16 (bf16 storage, tcgen05 kind::f16; default), tf32 (fp32
storage, tcgen05 kind::tf32) or fp32 (CUDA-core FMA check engine)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import cfnerf_b200 as cf
from cfnerf_b200 import dist as D
from oracle import cfnerf_oracle as O

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
GLOBAL = int(os.environ.get("CFN_TRAIN_RAYS", "4096"))
PREC = os.environ.get("CFN_TRAIN_PRECISION", "bf16")
steps, warm = 5, 2
cfg = O.CfnConfig()
net = cf.NeRFFlowsParams.from_oracle_params(cfg, O.make_params(cfg, 0), *O.make_latents(cfg, 0)).to(dev)
params = [q for n, q in net.named_parameters() if not n.startswith("alpha_linear") and not n.startswith("alpha_std_linear")]
from cfnerf_b200.optim import FusedAdam
opt = FusedAdam(params, lr=5e-4, betas=(0.9, 0.999)) if os.environ.get("CFN_TORCH_ADAM") != "1" else torch.optim.Adam(params, lr=5e-4)
bucket = D.GradBucket(params)
rays = D.shard_rays(O.synthetic_rays(GLOBAL, 1), rank, world).to(dev)
g = torch.Generator().manual_seed(2)
target = D.shard_rays(torch.rand(GLOBAL, 3, generator=g), rank, world).to(dev)
torch.manual_seed(100 + rank)       # each rank draws its own latent noise, like each DataParallel replica (models.py:233-235)
FUSED = os.environ.get("CFN_TRAIN_AUTOGRAD") != "1"      # default: the autograd-free FusedTrainStep
trainer = D.FusedTrainStep(net, lr=5e-4, precision=PREC) if FUSED else None
one_step = (lambda: trainer.step(rays, target)) if FUSED else (lambda: D.train_step(net, opt, rays, target, bucket, precision=PREC))
for _ in range(warm):
    out = one_step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    out = one_step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    t = float(ms) / steps * 1e-3
    print(json.dumps({"metric": "rays/sec (train step)", "value": GLOBAL / t, "unit": "rays/s", "n_gpus": world,
                      "ms_per_step": t * 1e3, "global_batch_rays": GLOBAL, "dtype": PREC,
                      "loss": float(out["loss"]), "psnr": float(out["psnr"]),
                      "gflops_per_step": GLOBAL * 128 * 4708864 * 3 / 1e9, "achieved_tflops": GLOBAL * 128 * 4708864 * 3 / t / 1e12}))
if world > 1:
    dist.destroy_process_group()
