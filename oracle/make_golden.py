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

import math
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


def golden_train_depth(name, cfg, seed, variant, n_rgb, n_depth, depth_lambda=0.01):
    """One iteration of the shipped recipe's trainer body (--colmap_depth, --depth_lambda; main:1009-1067) through the
    UNMODIFIED `render()` -> `batchify_rays` -> `render_rays` -> `run_network`/`batchify` -> `NeRF_Flows.forward`, with
    netchunk = n_rgb * 128 so that, as in the shipped recipe (512 * 128 = 65536), the colour rays fill the first network
    call and the depth rays the second one (each call draws its own noise).  The loss lines are inline code of `train()`
    (main:1018-1055) and are restated here verbatim in torch."""
    p = O.make_params(cfg, seed, variant)
    sa, sr = O.make_latents(cfg, seed)
    main, model, _ = refload.build_reference_model(cfg, p, sa, sr)
    embed_fn, _ = main.get_embedder(cfg.L_pos, 0)
    embeddirs_fn, _ = main.get_embedder(cfg.L_dir, 0)
    netchunk = n_rgb * 128

    def nq(inputs, viewdirs, network_fn, is_val, is_test):                                   # main:333-336
        return main.run_network(inputs, viewdirs, network_fn, is_val, is_test, embed_fn=embed_fn,
                                embeddirs_fn=embeddirs_fn, netchunk=netchunk)

    B = n_rgb + n_depth
    rays = O.synthetic_rays(B, seed + 1)
    batch_rays = torch.stack([rays[:, 0:3], rays[:, 3:6]], 0)                                # (2, B, 3), main:1011
    kw = dict(is_train=True, uniformsample=False, network_query_fn=nq, perturb=1.0, N_importance=0, N_samples=128,
              K_samples=cfg.K, network_fn=model, use_viewdirs=True, white_bkgd=False, raw_noise_std=1.0, ndc=False,
              lindisp=False, near=1.2, far=8.0)                                              # main:382-399, 822-827
    rng_seed = 91 + seed
    torch.manual_seed(rng_seed)
    rgbs, disp, depth, extras = main.render(8, 8, 10.0, chunk=1024 * 32, rays=batch_rays, verbose=False, retraw=False,
                                            **kw)                                            # main:1014-1016
    # replay the RNG order: t_rand, then (eps_alpha, eps_rgb) of call 1, then of call 2
    torch.manual_seed(rng_seed)
    t_rand = torch.rand(B, 128)
    eps = [(torch.empty(cfg.K, 1).normal_(), torch.empty(cfg.K, 3).normal_()) for _ in range(2)]
    g = torch.Generator().manual_seed(seed + 2)
    target_s = torch.rand(n_rgb, 3, generator=g)
    target_depth = 1.2 + 6.8 * torch.rand(n_depth, generator=g)
    # ---- main:1018-1055, verbatim
    N_batch = n_rgb
    K = cfg.K
    depth_m = torch.mean(depth, -1)
    rgbs_ = rgbs[:N_batch, :]
    depth_, depth_col = depth_m[:N_batch], depth_m[N_batch:]
    extras_ = {x: extras[x][:N_batch] for x in extras}
    rgb_mean = torch.mean(rgbs_, -1)
    mse_train = main.img2mse(rgb_mean, target_s)
    psnr_train = main.mse2psnr(mse_train)
    eps_ = 1e-05
    n = K
    rgb_std = torch.std(rgbs_, -1) * n / (n - 1)
    H_sqrt = rgb_std.detach() * torch.pow(0.8 / n, torch.tensor(-1 / 7)) + eps_
    H_sqrt = H_sqrt[..., None]
    r_P_C_1 = torch.exp(-((rgbs_ - target_s[..., None]) ** 2) / (2 * H_sqrt * H_sqrt))
    r_P_C_2 = torch.pow(torch.tensor(2 * math.pi), -1.5) / H_sqrt
    r_P_C_mean = (r_P_C_1 * r_P_C_2).mean(-1) + eps_
    loss_nll = -torch.log(r_P_C_mean).mean()
    loss_entropy = extras_["loss_entropy"].mean()
    loss = loss_nll + 0.01 * loss_entropy
    depth_loss = main.img2mse(depth_col, target_depth)
    loss = loss + depth_lambda * depth_loss
    model.zero_grad()
    loss.backward()
    grads = {n_: q.grad for n_, q in model.named_parameters()}
    names, norms, gsel = [], [], {}
    for n_, gr in sorted(grads.items()):
        names.append(n_)
        norms.append(0.0 if gr is None else float(gr.double().pow(2).sum().sqrt()))
    for n_ in ("alpha_mean", "alpha_std", "rgb_mean", "rgb_std", "pts_linears.0.bias", "pts_linears.7.bias",
               "h_alpha_linear.bias", "h_rgb_linear.bias", "views_linears.0.bias", "feature_linear.bias",
               "flows_rgb.amor_d.bias", "flows_rgb.amor_b.bias", "flows_alpha.amor_diag1.0.bias",
               "flows_alpha.amor_diag2.0.bias", "flows_alpha.amor_b.bias", "flows_alpha.amor_b.weight"):
        gsel["grad__" + n_] = grads[n_]
    for n_ in ("pts_linears.0.weight", "pts_linears.5.weight", "h_alpha_linear.weight"):
        gsel["gradrows__" + n_] = grads[n_][:4]
    _save(name, cfg, seed, variant, p, in_rays=rays, in_t_rand=t_rand,
          in_eps_alpha=torch.stack([e[0] for e in eps], 0), in_eps_rgb=torch.stack([e[1] for e in eps], 0),
          in_target=target_s, in_target_depth=target_depth, in_beta1=np.float64(0.01),
          in_depth_lambda=np.float64(depth_lambda), in_n_rgb=np.int64(n_rgb), in_netchunk=np.int64(netchunk),
          out_rgb_map=rgbs, out_depth_map=depth, out_loss_entropy=loss_entropy, out_loss_nll=loss_nll,
          out_depth_loss=depth_loss, out_loss=loss, out_psnr=psnr_train, out_grad_names=np.array(names),
          out_grad_norms=np.array(norms, np.float64), **gsel)


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
    golden_train_depth("train_depth_small", small, 4, "lively", 8, 4)
    golden_raw2outputs("raw2outputs_random", 11)
    golden_network("network_canonical", canon, 0, "lively", 96)
    golden_network("network_stressed", canon, 2, "stressed", 64)


if __name__ == "__main__":
    main()
