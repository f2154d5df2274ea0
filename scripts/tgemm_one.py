"""One GEMM flavour of the training shapes, for ncu captures: python scripts/tgemm_one.py fwd|dgrad|wgrad [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from cfnerf_b200.api import gemm
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
Mp = 4096 * 128
X = torch.randn(Mp, 512, device=dev); Wt = torch.randn(512, 512, device=dev) * 0.05
G = torch.randn(Mp, 512, device=dev); bias = torch.randn(512, device=dev)
Y = torch.empty(Mp, 512, device=dev); dW = torch.zeros(512, 512, device=dev)
for _ in range(reps):
    if which == "fwd": gemm(X, Wt.t(), engine="tf32", bias=bias, epilogue="relu", out=Y, round_out=True)
    elif which == "dgrad": gemm(G, Wt, engine="tf32", epilogue="relu_mask_mul", aux=X, out=Y, round_out=True)
    else: gemm(G.t(), X, engine="tf32", out=dW, split_k=18)
torch.cuda.synchronize()
print("done", which)
