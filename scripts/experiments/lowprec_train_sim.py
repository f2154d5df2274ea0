"""CPU experiment (oracle only): how do tf32 / bf16 STORAGE of activations and gradients between the layer GEMMs move
the gradients and the PSNR after a few Adam steps?  Decides whether a bf16-storage training chain can meet the 0.1 dB bar.
  python scripts/experiments/lowprec_train_sim.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from oracle import cfnerf_oracle as O


def rnd(x, fmt):
    if fmt == "fp32": return x
    if fmt == "bf16": return x.to(torch.bfloat16).to(torch.float32)
    i = x.contiguous().view(torch.int32)
    return ((i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)


class RoundFB(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fmt_f, fmt_b):
        ctx.fmt_b = fmt_b
        return rnd(x, fmt_f)

    @staticmethod
    def backward(ctx, g):
        return rnd(g, ctx.fmt_b), None, None


FMT = "fp32"


def lin(x, w, b):
    # operands rounded to the storage format; the output gradient is rounded on its way back (stored dY)
    y = F.linear(RoundFB.apply(x, FMT, "fp32"), RoundFB.apply(w, FMT, "fp32"), b)
    return RoundFB.apply(y, "fp32", FMT)


def mlp_encode(p, cfg, embedded):
    g_pos, g_dir = embedded[:, : cfg.in_pos], embedded[:, cfg.in_pos:]
    h = g_pos
    for i in range(cfg.D):
        h = F.relu(lin(h, p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"]))
        if i == cfg.skip:
            h = torch.cat([g_pos, h], -1)
    h_alpha = lin(h, p["h_alpha_linear.weight"], p["h_alpha_linear.bias"])
    feat = lin(h, p["feature_linear.weight"], p["feature_linear.bias"])
    v = F.relu(lin(torch.cat([feat, g_dir], -1), p["views_linears.0.weight"], p["views_linears.0.bias"]))
    h_rgb = lin(v, p["h_rgb_linear.weight"], p["h_rgb_linear.bias"])
    return h_alpha, h_rgb


O.mlp_encode = mlp_encode
cfg = O.CfnConfig(W=256, K=64, h_alpha=32)
B, steps = 16, 6
res = {}
for fmt in ("fp32", "tf32", "bf16"):
    FMT = fmt
    p0 = O.make_params(cfg, 5, "lively")
    p = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    live = [k for k in p if not k.startswith("alpha_linear") and not k.startswith("alpha_std_linear")]
    opt = torch.optim.Adam([p[k] for k in live], lr=5e-4, betas=(0.9, 0.999))
    g = torch.Generator().manual_seed(12)
    rays = O.synthetic_rays(B, 13)
    target = torch.rand(B, 3, generator=g)
    psnrs, grads0 = [], None
    for it in range(steps):
        t_rand = torch.rand(B, 128, generator=g)
        ea, er = torch.randn(cfg.K, 1, generator=g), torch.randn(cfg.K, 3, generator=g)
        out = O.render_rays(p, cfg, rays, ea, er, True, t_rand=t_rand, faithful=False)
        l = O.kde_nll_loss(out["rgb_map"], target, out["loss_entropy"], cfg.K, 0.01)
        opt.zero_grad()
        l["loss"].backward()
        if it == 0:
            grads0 = {k: p[k].grad.clone() for k in live if p[k].grad is not None}
        opt.step()
        psnrs.append(float(l["psnr"]))
    res[fmt] = (psnrs, grads0)
    print(fmt, ["%.4f" % x for x in psnrs], flush=True)
for fmt in ("tf32", "bf16"):
    worst = 0
    for k, g0 in res["fp32"][1].items():
        n = g0.norm().item()
        if n == 0: continue
        e = (g0 - res[fmt][1][k]).norm().item() / n
        worst = max(worst, e)
    d = max(abs(a - b) for a, b in zip(res["fp32"][0], res[fmt][0]))
    print(f"{fmt}: worst grad rel l2 err at step 0 = {worst:.3e}; max |dPSNR| over {steps} steps = {d:.4f} dB")
