"""GPU check + timing of the bf16-storage tcgen05 GEMM (cfn_gemm_bf16) against torch fp64 on the same bf16 operands."""
import sys, json
import torch
sys.path.insert(0, ".")
from cfnerf_b200.api import gemm_bf16

dev = torch.device("cuda:0")
torch.manual_seed(0)
bf = torch.bfloat16
ok = True

def report(name, C, ref, tol_rel):
    global ok
    err = (C.double() - ref).abs().max().item(); scale = max(1.0, ref.abs().max().item())
    good = err <= tol_rel * scale and bool(torch.isfinite(C.float()).all())
    ok &= good
    print(("OK " if good else "BAD"), name, f"err={err:.3e} scale={scale:.2e}", flush=True)

for (M, N, K) in [(128, 16, 64), (300, 104, 72), (5000, 512, 576), (1000, 256, 544)]:
    A = torch.randn(M, K, device=dev).to(bf); W = torch.randn(N, K, device=dev).to(bf)
    ref = A.double() @ W.double().t()
    report(f"KK plain M={M} N={N} K={K}", gemm_bf16(A, W.t()), ref, 6e-3)
    bias = torch.randn(N, device=dev)
    bits = torch.zeros(M, (N + 31) // 32, dtype=torch.int32, device=dev)
    Y = gemm_bf16(A, W.t(), bias=bias, epilogue="relu", mask_out=bits)
    refy = (ref + bias.double()).clamp_min(0)
    report(f"KK relu  M={M} N={N} K={K}", Y, refy, 6e-3)
    # bit mask == (Y > 0)
    want = (Y.float() > 0)
    got = torch.zeros_like(want)
    for j in range(N): got[:, j] = ((bits[:, j // 32] >> (8 * (j % 4) + (j % 32) // 4)) & 1).bool()
    same = bool((want == got).all()); ok &= same
    print("OK " if same else "BAD", "  relu bit mask", flush=True)
    flags = (torch.rand(N, device=dev) > 0.5).float()
    T = gemm_bf16(A, W.t(), bias=bias, epilogue="tanh_mask", aux=flags, out_dtype=torch.float32)
    reft = torch.where(flags.bool()[None, :], torch.tanh(ref + bias.double()), ref + bias.double())
    report(f"KK tanh  M={M} N={N} K={K} (fp32 out)", T, reft, 2e-5)
    # dgrad flavours: G (M x N) . W (N x K) -> (M x K), masked by the bits of an (M x K) activation
    G = torch.randn(M, N, device=dev).to(bf)
    refd = G.double() @ W.double()
    report(f"KMN plain M={M} N={K} K={N}", gemm_bf16(G, W), refd, 6e-3)
    act = torch.randn(M, K, device=dev)
    abits = torch.zeros(M, (K + 31) // 32, dtype=torch.int32, device=dev)
    for j in range(K): abits[:, j // 32] |= ((act[:, j] > 0).int() << (8 * (j % 4) + (j % 32) // 4))
    D = gemm_bf16(G, W, epilogue="relu_mask_mul", aux_bits=abits)
    report(f"KMN mask  M={M} N={K} K={N}", D, torch.where(act > 0, refd, torch.zeros_like(refd)), 6e-3)

for (O_, I_, P, split) in [(512, 512, 40000, 18), (512, 576, 40000, 12), (64, 512, 30000, 74), (16, 64, 3000, 5), (256, 544, 9999, 9)]:
    G = torch.randn(P, O_, device=dev).to(bf); X = torch.randn(P, I_, device=dev).to(bf)
    ref = G.double().t() @ X.double()
    rs = torch.zeros(O_, device=dev)
    dW = gemm_bf16(G.t(), X, out_dtype=torch.float32, split_k=split, rowsum=rs)
    report(f"MNMN wgrad out={O_} in={I_} pts={P} split={split}", dW, ref, 2e-5)
    report("  rowsum", rs, G.double().sum(0), 2e-5)
print("ALL OK" if ok else "SOME BAD", flush=True)

def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

Mp = 4096 * 128
X = torch.randn(Mp, 512, device=dev).to(bf); Wt = (torch.randn(512, 512, device=dev) * 0.05).to(bf)
G = torch.randn(Mp, 512, device=dev).to(bf); bias = torch.randn(512, device=dev)
Y = torch.empty(Mp, 512, device=dev, dtype=bf); bits = torch.zeros(Mp, 16, dtype=torch.int32, device=dev)
dW = torch.zeros(512, 512, device=dev); rs = torch.zeros(512, device=dev)
t_f = timeit(lambda: gemm_bf16(X, Wt.t(), bias=bias, epilogue="relu", mask_out=bits, out=Y))
t_d = timeit(lambda: gemm_bf16(G, Wt, epilogue="relu_mask_mul", aux_bits=bits, out=Y))
t_w = timeit(lambda: gemm_bf16(G.t(), X, out=dW, split_k=18, rowsum=rs), n=5)
fl = 2.0 * Mp * 512 * 512
print("bf16", json.dumps({"fwd_ms": t_f, "dgrad_ms": t_d, "wgrad_ms": t_w, "fwd_tflops": fl / t_f / 1e9, "dgrad_tflops": fl / t_d / 1e9, "wgrad_tflops": fl / t_w / 1e9}))
t_t = timeit(lambda: torch.matmul(X, Wt.t()))
print("torch bf16 matmul: %.3f ms = %.1f TFLOP/s" % (t_t, fl / t_t / 1e9))
