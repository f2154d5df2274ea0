"""GPU check of the TF32 tensor-core network stage: forward vs the fp32 check mode, gradients vs the fp32 path."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cfnerf_b200 as cf
from oracle import cfnerf_oracle as O

dev = torch.device("cuda:0")
for kw, B in ((dict(), 64), (dict(W=256, K=64, h_alpha=32), 33), (dict(D=7, W=128, F=2), 17)):
    cfg = O.CfnConfig(**kw)
    p = O.make_params(cfg, 3, "lively")
    sa, sr = O.make_latents(cfg, 3)
    rays = O.synthetic_rays(B, 4).to(dev)
    g = torch.Generator().manual_seed(7)
    target = torch.rand(B, 3, generator=g).to(dev)
    t_rand = torch.rand(B, 128, generator=g).to(dev)
    ea, er = torch.randn(cfg.K, 1, generator=g).to(dev), torch.randn(cfg.K, 3, generator=g).to(dev)
    res = {}
    for prec in ("fp32", "tf32", "bf16"):
        net = cf.NeRFFlowsParams.from_oracle_params(cfg, p, sa, sr).to(dev)
        o_test = cf.render_rays(rays, net, None, 128, False, False, precision=prec)
        out = cf.render_rays(rays, net, None, 128, True, False, perturb=1., raw_noise_std=1., t_rand=t_rand, eps_alpha=ea,
                             eps_rgb=er, precision=prec)
        l = cf.kde_nll_loss(out["rgb_map"], target, out["loss_entropy"], cfg.K, 0.01)
        net.zero_grad()
        l["loss"].backward()
        res[prec] = dict(test=o_test, out=out, loss=float(l["loss"]), grads={n: q.grad.clone() for n, q in net.named_parameters() if q.grad is not None})
    for other in ("tf32", "bf16"):
        a, b = res["fp32"], res[other]
        print("config", kw, "B", B, "fp32 vs", other)
        for k in ("rgb_map", "depth_map"):
            print("  train %-10s max abs diff %.3e" % (k, (a["out"][k] - b["out"][k]).abs().max().item()))
        print("  loss fp32 %.6f other %.6f" % (a["loss"], b["loss"]))
        worst, wname = 0.0, ""
        for n in a["grads"]:
            ga, gb = a["grads"][n].double(), b["grads"][n].double()
            na = ga.norm().item()
            rel = (ga - gb).norm().item() / max(na, 1e-30)
            if na > 0 and rel > worst:
                worst, wname = rel, n
            if not torch.isfinite(gb).all():
                print("  NON-FINITE grad", n)
        print("  worst grad rel l2 err %.3e (%s)" % (worst, wname), flush=True)
