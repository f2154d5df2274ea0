"""Experiment: does running the layer-by-layer training chain in point chunks that fit the 126 MB L2 pay?
8 forward layers (X -> relu(X W^T + b) with bit masks, bf16 storage) + 8 dgrad + 8 wgrad over 524 288 points, either
whole-batch per layer (what mlp_chain.cu does) or chunk by chunk (each chunk walks all layers before the next starts).
All activations are saved (distinct buffers per layer), as in training.  CUDA events, CUDA-graph replay to hide launches."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cfnerf_b200.api import gemm_bf16

dev = torch.device("cuda:0")
torch.manual_seed(0)
bf = torch.bfloat16
M, W, L = 524288, 512, 8
Ws = [(torch.randn(W, W, device=dev) * 0.06).to(bf) for _ in range(L)]
bs = [torch.randn(W, device=dev) * 0.1 for _ in range(L)]
acts = [torch.empty(M, W, device=dev, dtype=bf) for _ in range(L + 1)]
acts[0].copy_(torch.randn(M, W, device=dev).to(bf))
masks = [torch.empty(M, W // 32, device=dev, dtype=torch.int32) for _ in range(L)]
grads = [torch.empty(M, W, device=dev, dtype=bf) for _ in range(2)]
dWs = [torch.zeros(W, W, device=dev) for _ in range(L)]
rs = [torch.zeros(W, device=dev) for _ in range(L)]

def fwd(c0, c1):
    for l in range(L):
        gemm_bf16(acts[l][c0:c1], Ws[l].t(), bias=bs[l], epilogue="relu", mask_out=masks[l][c0:c1], out=acts[l + 1][c0:c1])

def bwd(c0, c1, split):
    g = grads[0][c0:c1]
    for l in reversed(range(L)):
        # wgrad: dW += dY^T X (split-K atomics), dgrad: dX = (dY W) * relu'(h_{l-1})
        gemm_bf16(g.t(), acts[l][c0:c1], out_dtype=torch.float32, split_k=split, out=dWs[l])
        if l > 0:
            gn = grads[(L - l) & 1][c0:c1]
            gemm_bf16(g, Ws[l], epilogue="relu_mask_mul", aux_bits=masks[l - 1][c0:c1], out=gn)
            g = gn

def run(chunk, what):
    split_full = 74
    def body():
        for c0 in range(0, M, chunk):
            if what in ("fwd", "both"): fwd(c0, c0 + chunk)
            if what in ("bwd", "both"): bwd(c0, c0 + chunk, max(1, split_full * chunk // M) if chunk < M else split_full)
    grads[0].copy_(torch.randn(M, W, device=dev).to(bf))
    body(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        body()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            body()
    torch.cuda.synchronize()
    for _ in range(2): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5

res = {}
for what in ("fwd", "bwd"):
    for chunk in (M, 131072, 65536, 32768):
        try:
            res[f"{what}_chunk{chunk}"] = run(chunk, what)
        except Exception as e:
            res[f"{what}_chunk{chunk}"] = f"failed: {str(e)[:200]}"
        print(what, chunk, res[f"{what}_chunk{chunk}"], flush=True)
json.dump(res, open("gpurun_out/r2_l2_chunk.json", "w"), indent=1)
