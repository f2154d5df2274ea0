"""CPU envelope of the low-precision training chains (no GPU needed): rounding every activation / weight operand and
every back-propagated gradient of the MLP chain to tf32 or bf16 STORAGE inside the oracle must keep the PSNR after a few
Adam steps within the north star's 0.1 dB of the fp32 run.  This is the experiment that decided the bf16-storage chain
(scripts/experiments/lowprec_train_sim.py is the long form); the GPU counterpart is
tests/test_gpu_parity.py::test_training_steps_track_the_oracle[tf32|bf16]."""
import torch
import torch.nn.functional as F

from oracle import cfnerf_oracle as O


def _rnd(x, fmt):
    if fmt == "fp32":
        return x
    if fmt == "bf16":
        return x.to(torch.bfloat16).to(torch.float32)
    i = x.contiguous().view(torch.int32)                      # tf32: round to nearest even on 10 mantissa bits
    return ((i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)


class _RoundFB(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, fmt_f, fmt_b):
        ctx.fmt_b = fmt_b
        return _rnd(x, fmt_f)

    @staticmethod
    def backward(ctx, g):
        return _rnd(g, ctx.fmt_b), None, None


def _encode_with_storage(fmt):
    def lin(x, w, b):
        y = F.linear(_RoundFB.apply(x, fmt, "fp32"), _RoundFB.apply(w, fmt, "fp32"), b)
        return _RoundFB.apply(y, "fp32", fmt)               # the stored dY

    def mlp_encode(p, cfg, embedded):                       # same graph as oracle.mlp_encode (models.py:165-186)
        g_pos, g_dir = embedded[:, : cfg.in_pos], embedded[:, cfg.in_pos:]
        h = g_pos
        for i in range(cfg.D):
            h = F.relu(lin(h, p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"]))
            if i == cfg.skip:
                h = torch.cat([g_pos, h], -1)
        h_alpha = lin(h, p["h_alpha_linear.weight"], p["h_alpha_linear.bias"])
        feat = lin(h, p["feature_linear.weight"], p["feature_linear.bias"])
        v = F.relu(lin(torch.cat([feat, g_dir], -1), p["views_linears.0.weight"], p["views_linears.0.bias"]))
        return h_alpha, lin(v, p["h_rgb_linear.weight"], p["h_rgb_linear.bias"])
    return mlp_encode


def _train(fmt, monkeypatch, steps=3, B=8):
    monkeypatch.setattr(O, "mlp_encode", _encode_with_storage(fmt))
    cfg = O.CfnConfig(W=128, K=16, h_alpha=32, h_rgb=32, D=4)
    p = {k: v.clone().requires_grad_(True) for k, v in O.make_params(cfg, 5, "lively").items()}
    live = [k for k in p if not k.startswith("alpha_linear") and not k.startswith("alpha_std_linear")]
    opt = torch.optim.Adam([p[k] for k in live], lr=5e-4, betas=(0.9, 0.999))
    g = torch.Generator().manual_seed(12)
    rays = O.synthetic_rays(B, 13)
    target = torch.rand(B, 3, generator=g)
    psnr, grads0 = [], None
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
        psnr.append(float(l["psnr"].detach()))
    return psnr, grads0


def test_tf32_and_bf16_storage_keep_the_training_psnr(monkeypatch):
    ref_psnr, ref_g = _train("fp32", monkeypatch)
    for fmt, tol_grad in (("tf32", 0.1), ("bf16", 0.3)):
        psnr, g = _train(fmt, monkeypatch)
        assert max(abs(a - b) for a, b in zip(ref_psnr, psnr)) <= 0.1, (fmt, ref_psnr, psnr)
        worst = max(((ref_g[k] - g[k]).norm() / ref_g[k].norm()).item() for k in ref_g if ref_g[k].norm() > 0)
        assert worst <= tol_grad, (fmt, worst)
