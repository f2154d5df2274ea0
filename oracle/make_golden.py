"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (the only place /root/reference exists):

    python -m oracle.make_golden

Every array written under an ``out_`` key was produced by the reference's own functions
(`render_rays`, `raw2outputs`, `NeRF_Flows.forward`, autograd through them), loaded through
oracle/refload.py.  Inputs (rays, noise, uniforms) are stored beside them; network parameters are
NOT stored (9.4 MB) — they are regenerated from ``oracle.cfnerf_oracle.make_params(cfg, seed,
variant)`` and the file carries ``params_checksum`` so RNG drift is detected instead of silently
comparing different networks.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import cfnerf_oracle as O
from oracle import refload

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _cfg_dict(cfg):
    return dict(D=cfg.D, W=cfg.W, L_pos=cfg.L_pos, L_dir=cfg.L_dir, h_alpha=cfg.h_alpha, h_rgb=cfg.h_rgb,
                F=cfg.F, K=cfg.K)


def _save(name, cfg, seed, variant, params, **arrays):
    out = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    out["cfg"] = np.array([cfg.D, cfg.W, cfg.L_pos, cfg.L_dir, cfg.h_alpha, cfg.h_rgb, cfg.F, cfg.K], np.int64)
    out["seed"] = np.int64(seed)
    out["variant"] = np.array(variant)
    out["params_checksum"] = np.float64(O.params_checksum(params))
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def golden_render_test(name, cfg, seed, variant, n_rays, **kw):
    p = O.make_params(cfg, seed, variant)
    sa, sr = O.make_latents(cfg, seed)
    main, model, nq = refload.build_reference_model(cfg, p, sa, sr)
    rays = O.synthetic_rays(n_rays, seed + 1)
    with torch.no_grad():
        ref = main.render_rays(rays, model, nq, 128, False, False, K_samples=cfg.K, perturb=0.0,
                               raw_noise_std=0.0, **kw)
    _save(name, cfg, seed, variant, p, in_rays=rays, in_sample_alpha=sa, in_sample_rgb=sr,
          in_lindisp=np.bool_(kw.get("lindisp", False)), in_white_bkgd=np.bool_(kw.get("white_bkgd", False)),
          out_rgb_map=ref["rgb_map"], out_disp_map=ref["disp_map"], out_depth_map=ref["depth_map"])


def golden_render_train(name, cfg, seed, variant, n_rays):
    p = O.make_params(cfg, seed, variant)
    sa, sr = O.make_latents(cfg, seed)
    main, model, nq = refload.build_reference_model(cfg, p, sa, sr)
    rays = O.synthetic_rays(n_rays, seed + 1)
    rng_seed = 77 + seed
    torch.manual_seed(rng_seed)
    ref = main.render_rays(rays, model, nq, 128, True, False, K_samples=cfg.K, perturb=1.0, raw_noise_std=1.0)
    # replay the reference's RNG consumption order: t_rand -> eps_alpha -> eps_rgb
    torch.manual_seed(rng_seed)
    t_rand = torch.rand(n_rays, 128)
    eps_a = torch.empty(cfg.K, 1).normal_()
    eps_c = torch.empty(cfg.K, 3).normal_()
    g = torch.Generator().manual_seed(seed + 2)
    target = torch.rand(n_rays, 3, generator=g)
    # the trainer's loss, main:1027-1050 (inline code there; restated in the oracle)
    ent = ref["loss_entropy"].mean()
    losses = O.kde_nll_loss(ref["rgb_map"], target, ent, cfg.K, beta1=0.01)
    model.zero_grad()
    losses["loss"].backward()
    grads = {n: q.grad for n, q in model.named_parameters()}
    gsel = {}
    norms = []
    names = []
    for n, gr in sorted(grads.items()):
        names.append(n)
        norms.append(0.0 if gr is None else float(gr.double().pow(2).sum().sqrt()))
    for n in ("alpha_mean", "alpha_std", "rgb_mean", "rgb_std", "pts_linears.0.bias", "pts_linears.7.bias",
              "h_alpha_linear.bias", "h_rgb_linear.bias", "views_linears.0.bias", "feature_linear.bias",
              "flows_rgb.amor_d.bias", "flows_rgb.amor_diag1.0.bias", "flows_rgb.amor_diag2.0.bias",
              "flows_rgb.amor_b.bias", "flows_alpha.amor_diag1.0.bias", "flows_alpha.amor_diag2.0.bias",
              "flows_alpha.amor_b.bias", "flows_rgb.amor_d.weight", "flows_alpha.amor_b.weight"):
        gsel["grad__" + n] = grads[n]
    for n in ("pts_linears.0.weight", "pts_linears.5.weight", "views_linears.0.weight", "h_alpha_linear.weight"):
        gsel["gradrows__" + n] = grads[n][:4]
    _save(name, cfg, seed, variant, p, in_rays=rays, in_t_rand=t_rand, in_eps_alpha=eps_a, in_eps_rgb=eps_c,
          in_target=target, in_beta1=np.float64(0.01),
          out_rgb_map=ref["rgb_map"], out_disp_map=ref["disp_map"], out_depth_map=ref["depth_map"],
          out_raw_ray0=ref["raw"][0], out_loss_entropy=ent, out_loss=losses["loss"], out_loss_nll=losses["loss_nll"],
          out_psnr=losses["psnr"], out_grad_names=np.array(names), out_grad_norms=np.array(norms, np.float64), **gsel)


def golden_raw2outputs(name, seed):
    main, _ = refload.load_reference()
    g = torch.Generator().manual_seed(seed)
    B, N, K = 6, 128, 32
    raw = torch.randn(B, N, K, 4, generator=g) * 2.0
    z = torch.sort(torch.rand(B, N, generator=g) * 6.0 + 1.0, -1).values
    d = torch.randn(B, 3, generator=g)
    outs = {}
    for wb in (False, True):
        rgb_map, disp, w, depth = main.raw2outputs(raw, z, d, raw_noise_std=0.0, white_bkgd=wb)
        tag = "wb" if wb else "nb"
        outs.update({f"out_rgb_map_{tag}": rgb_map, f"out_disp_{tag}": disp, f"out_weights_{tag}": w,
                     f"out_depth_{tag}": depth})
    cfg = O.CfnConfig()
    _save(name, cfg, seed, "none", {}, in_raw=raw, in_z_vals=z, in_rays_d=d, **outs)


def golden_network(name, cfg, seed, variant, n_pts):
    p = O.make_params(cfg, seed, variant)
    sa, sr = O.make_latents(cfg, seed)
    main, model, nq = refload.build_reference_model(cfg, p, sa, sr)
    g = torch.Generator().manual_seed(seed + 9)
    pts = torch.randn(n_pts, 3, generator=g) * 2.0
    dirs = torch.randn(n_pts, 3, generator=g)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    emb_p, _ = main.get_embedder(cfg.L_pos, 0)
    emb_d, _ = main.get_embedder(cfg.L_dir, 0)
    x = torch.cat([emb_p(pts), emb_d(dirs)], -1)
    with torch.no_grad():
        h_alpha, h_rgb = model.encode(x)
        raw, zeros = model(x, is_val=False, is_test=True)
        r1, r2, b = model.flows_rgb.encode(h_rgb)
    _save(name, cfg, seed, variant, p, in_pts=pts, in_dirs=dirs, in_sample_alpha=sa, in_sample_rgb=sr,
          out_embedded=x, out_h_alpha=h_alpha, out_h_rgb=h_rgb, out_raw=raw, out_r1_rgb=r1, out_r2_rgb=r2,
          out_b_rgb=b)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    canon = O.CfnConfig()
    small = O.CfnConfig(W=256, K=64, h_alpha=32)       # parser-default variant
    golden_render_test("render_test_canonical", canon, 0, "lively", 24)
    golden_render_test("render_test_default_init", canon, 1, "default", 8)
    golden_render_test("render_test_small_wb_lindisp", small, 3, "lively", 16, lindisp=True, white_bkgd=True)
    golden_render_train("render_train_canonical", canon, 0, "lively", 16)
    golden_render_train("render_train_small", small, 3, "lively", 8)
    golden_raw2outputs("raw2outputs_random", 11)
    golden_network("network_canonical", canon, 0, "lively", 96)
    golden_network("network_stressed", canon, 2, "stressed", 64)


if __name__ == "__main__":
    main()
