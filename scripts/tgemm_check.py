"""GPU check + timing of the tcgen05 TF32 GEMM (cfn_gemm_f32 engine 1) against torch fp64 on tf32-rounded operands."""
import sys, time, json
import torch
sys.path.insert(0, ".")
from cfnerf_b200.api import gemm

dev = torch.device("cuda:0")
torch.manual_seed(0)

def tf32_round(x):
    # round-to-nearest-even onto 10 mantissa bits (what a well-behaved producer stores)
    i = x.view(torch.int32)
    r = ((i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF)
    return r.view(torch.float32)

def run_case(M, N, K, a_mn, b_mn, epi="none", bias=False, acc=False, split=1, lda_pad=0, name=""):
    # logical A (M,K), B (K,N); storage selects the major
    if a_mn:
        As = tf32_round(torch.randn(K, M + lda_pad, device=dev)); A = As[:, :M].t()
    else:
        As = tf32_round(torch.randn(M, K + lda_pad, device=dev)); A = As[:, :K]
    if b_mn:
        Bs = tf32_round(torch.randn(K, N + lda_pad, device=dev)); B = Bs[:, :N]
    else:
        Bs = tf32_round(torch.randn(N, K + lda_pad, device=dev)); B = Bs[:, :K].t()
    bias_t = torch.randn(N, device=dev) if bias else None
    aux = None
    if epi == "relu_mask_mul": aux = torch.randn(M, N, device=dev)
    if epi == "tanh_mask": aux = (torch.rand(N, device=dev) > 0.5).float()
    out = None
    ref = A.double() @ B.double()
    if bias: ref = ref + bias_t.double()
    if acc:
        out = torch.randn(M, N, device=dev); ref = ref + out.double()
    if epi == "relu": ref = ref.clamp_min(0)
    if epi == "tanh_mask": ref = torch.where(aux.bool()[None, :], torch.tanh(ref), ref)
    if epi == "relu_mask_mul": ref = torch.where(aux > 0, ref, torch.zeros_like(ref))
    C = gemm(A, B, engine="tf32", bias=bias_t, epilogue=epi, aux=aux, out=out, accumulate=acc, split_k=split)
    torch.cuda.synchronize()
    err = (C.double() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-30
    ok = err <= 2e-5 * max(1.0, scale) and bool(torch.isfinite(C).all())
    print(f"{'OK ' if ok else 'BAD'} {name or ''} M={M} N={N} K={K} a_mn={int(a_mn)} b_mn={int(b_mn)} epi={epi} bias={int(bias)} acc={int(acc)} split={split} pad={lda_pad} err={err:.3e} scale={scale:.2e}", flush=True)
    return ok

ok = True
quick = [
    (128, 16, 32, False, False), (128, 256, 32, False, False), (128, 256, 64, False, False), (256, 256, 512, False, False),
    (300, 100, 72, False, False), (1000, 512, 512, False, False), (5000, 512, 576, False, False),
    (128, 256, 64, False, True), (1000, 512, 512, False, True), (700, 64, 12, False, True), (700, 60, 64, False, True),
    (128, 256, 64, True, True), (512, 512, 4096, True, True), (512, 576, 5000, True, True), (64, 512, 3000, True, True),
    (12, 64, 3000, True, True), (256, 540, 1000, True, True), (512, 64, 999, True, True),
    (256, 128, 64, True, False),
]
for (M, N, K, a, b) in quick:
    ok &= run_case(M, N, K, a, b)
ok &= run_case(3000, 512, 512, False, False, epi="relu", bias=True, name="fwd")
ok &= run_case(3000, 72, 64, False, False, epi="tanh_mask", bias=True, name="heads", lda_pad=8)
ok &= run_case(3000, 512, 512, False, True, epi="relu_mask_mul", name="dgrad")
ok &= run_case(3000, 512, 512, False, True, epi="relu_mask_mul", acc=True, name="dgrad+acc")
ok &= run_case(512, 512, 40000, True, True, split=37, name="wgrad")
ok &= run_case(512, 576, 40000, True, True, split=24, name="wgrad skip")
ok &= run_case(64, 512, 40000, True, True, split=148, name="wgrad head")
ok &= run_case(300, 100, 72, False, False, lda_pad=4, name="padded ld")
print("ALL OK" if ok else "SOME BAD", flush=True)

# ---- timing on the training shapes (4096 rays x 128 samples) ----
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

Mp = 4096 * 128
X = tf32_round(torch.randn(Mp, 512, device=dev)); Wt = tf32_round(torch.randn(512, 512, device=dev) * 0.05)
G = tf32_round(torch.randn(Mp, 512, device=dev)); bias = torch.randn(512, device=dev)
Y = torch.empty(Mp, 512, device=dev)
res = {}
for eng in ("tf32", "fp32"):
    t_f = timeit(lambda: gemm(X, Wt.t(), engine=eng, bias=bias, epilogue="relu", out=Y))
    t_d = timeit(lambda: gemm(G, Wt, engine=eng, epilogue="relu_mask_mul", aux=X, out=Y))
    dW = torch.zeros(512, 512, device=dev)
    t_w = timeit(lambda: gemm(G.t(), X, engine=eng, out=dW, split_k=37), n=5)
    fl = 2.0 * Mp * 512 * 512
    res[eng] = {"fwd_ms": t_f, "dgrad_ms": t_d, "wgrad_ms": t_w, "fwd_tflops": fl / t_f / 1e9, "dgrad_tflops": fl / t_d / 1e9, "wgrad_tflops": fl / t_w / 1e9}
    print(eng, json.dumps(res[eng]), flush=True)
t_t = timeit(lambda: torch.matmul(X, Wt.t()))
print("torch fp32 matmul (allow_tf32=%s): %.3f ms" % (torch.backends.cuda.matmul.allow_tf32, t_t))
torch.backends.cuda.matmul.allow_tf32 = True
t_t = timeit(lambda: torch.matmul(X, Wt.t()))
print("torch tf32 matmul: %.3f ms = %.1f TFLOP/s" % (t_t, 2.0 * Mp * 512 * 512 / t_t / 1e9))
