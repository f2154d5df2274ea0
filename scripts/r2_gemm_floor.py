"""Per-launch floor of the tcgen05 GEMM: forward flavour (relu + mask, bf16) at N = K = 512 for M from 256 to 524288,
CUDA-graph replay of 20 back-to-back launches (what a training step looks like to the GPU)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cfnerf_b200.api import gemm_bf16
dev = torch.device("cuda:0"); bf = torch.bfloat16
W = (torch.randn(512, 512, device=dev) * 0.05).to(bf); b = torch.randn(512, device=dev)
res = {}
for M in (256, 2048, 16384, 65536, 131072, 524288):
    X = torch.randn(M, 512, device=dev).to(bf); Y = torch.empty(M, 512, device=dev, dtype=bf)
    bits = torch.empty(M, 16, dtype=torch.int32, device=dev)
    def body():
        for _ in range(20): gemm_bf16(X, W.t(), bias=b, epilogue="relu", mask_out=bits, out=Y)
    body(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        body(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s): body()
    torch.cuda.synchronize()
    for _ in range(3): g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    res[M] = e0.elapsed_time(e1) / 200 * 1e3
    print(M, f"{res[M]:.1f} us per launch", flush=True)
json.dump(res, open("gpurun_out/r2_gemm_floor.json", "w"))
