"""CPU: the oracle restatement reproduces every fixture the unmodified reference generated."""
import numpy as np
import pytest
import torch

from conftest import T, load_golden
from oracle import cfnerf_oracle as O

TOL = 2e-6  # fp32 CPU vs fp32 CPU, same op order; observed 0.0 in the build container


@pytest.mark.parametrize("name", ["render_test_canonical", "render_test_default_init", "render_test_small_wb_lindisp"])
@pytest.mark.parametrize("faithful", [True, False])
def test_render_test_mode(name, faithful):
    g, cfg, p = load_golden(name)
    ea, er = O.test_latents(T(g["in_sample_alpha"]), T(g["in_sample_rgb"]))
    with torch.no_grad():
        out = O.render_rays(p, cfg, T(g["in_rays"]), ea, er, False, lindisp=bool(g["in_lindisp"]),
                            white_bkgd=bool(g["in_white_bkgd"]), faithful=faithful)
    for k in ("rgb_map", "disp_map", "depth_map"):
        np.testing.assert_allclose(out[k].numpy(), g["out_" + k], rtol=0, atol=TOL)


@pytest.mark.parametrize("name", ["render_train_canonical", "render_train_small"])
def test_render_train_mode_and_loss(name):
    g, cfg, p = load_golden(name)
    out = O.render_rays(p, cfg, T(g["in_rays"]), T(g["in_eps_alpha"]), T(g["in_eps_rgb"]), True,
                        t_rand=T(g["in_t_rand"]))
    for k in ("rgb_map", "disp_map", "depth_map"):
        np.testing.assert_allclose(out[k].detach().numpy(), g["out_" + k], rtol=0, atol=TOL)
    np.testing.assert_allclose(out["raw"][0].detach().numpy(), g["out_raw_ray0"], rtol=0, atol=TOL)
    np.testing.assert_allclose(float(out["loss_entropy"]), float(g["out_loss_entropy"]), rtol=1e-6)
    losses = O.kde_nll_loss(out["rgb_map"], T(g["in_target"]), out["loss_entropy"], cfg.K, float(g["in_beta1"]))
    np.testing.assert_allclose(float(losses["loss"]), float(g["out_loss"]), rtol=1e-6)
    np.testing.assert_allclose(float(losses["psnr"]), float(g["out_psnr"]), rtol=1e-6)


def test_train_gradients_small():
    """autograd through the restatement == autograd through the reference (pins the backward oracle)."""
    g, cfg, p = load_golden("render_train_small")
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    out = O.render_rays(p, cfg, T(g["in_rays"]), T(g["in_eps_alpha"]), T(g["in_eps_rgb"]), True,
                        t_rand=T(g["in_t_rand"]))
    losses = O.kde_nll_loss(out["rgb_map"], T(g["in_target"]), out["loss_entropy"], cfg.K, float(g["in_beta1"]))
    losses["loss"].backward()
    names = [str(n) for n in g["out_grad_names"]]
    norms = dict(zip(names, g["out_grad_norms"]))
    for n, q in p.items():
        mine = 0.0 if q.grad is None else float(q.grad.double().pow(2).sum().sqrt())
        assert abs(mine - norms[n]) <= 1e-5 * max(1e-6, norms[n]) + 1e-9, n
    for k in g:
        if k.startswith("grad__"):
            np.testing.assert_allclose(p[k[6:]].grad.numpy(), g[k], rtol=1e-4, atol=1e-7, err_msg=k)
        if k.startswith("gradrows__"):
            np.testing.assert_allclose(p[k[10:]].grad[:4].numpy(), g[k], rtol=1e-4, atol=1e-7, err_msg=k)
    # dead parameters
    assert p["alpha_linear.weight"].grad is None and p["alpha_std_linear.weight"].grad is None
    assert float(p["flows_alpha.amor_d.weight"].grad.abs().max()) == 0.0


def test_raw2outputs_golden():
    g, _, _ = load_golden("raw2outputs_random")
    for wb, tag in ((False, "nb"), (True, "wb")):
        rgb, disp, w, depth = O.raw2outputs(T(g["in_raw"]), T(g["in_z_vals"]), T(g["in_rays_d"]), wb)
        np.testing.assert_allclose(rgb.numpy(), g[f"out_rgb_map_{tag}"], rtol=0, atol=TOL)
        np.testing.assert_allclose(disp.numpy(), g[f"out_disp_{tag}"], rtol=1e-6, atol=TOL)
        np.testing.assert_allclose(w.numpy(), g[f"out_weights_{tag}"], rtol=0, atol=TOL)
        np.testing.assert_allclose(depth.numpy(), g[f"out_depth_{tag}"], rtol=0, atol=TOL)


@pytest.mark.parametrize("name", ["network_canonical", "network_stressed"])
def test_network_golden(name):
    g, cfg, p = load_golden(name)
    x = torch.cat([O.positional_encoding(T(g["in_pts"]), cfg.L_pos),
                   O.positional_encoding(T(g["in_dirs"]), cfg.L_dir)], -1)
    np.testing.assert_allclose(x.numpy(), g["out_embedded"], rtol=0, atol=TOL)
    with torch.no_grad():
        ha, hr = O.mlp_encode(p, cfg, x)
        ea, er = O.test_latents(T(g["in_sample_alpha"]), T(g["in_sample_rgb"]))
        raw, _ = O.nerf_flows_forward(p, cfg, x, ea, er, False)
        r1, r2, b = O.flow_conditioning(p, "flows_rgb", hr, 3, cfg.F)
    np.testing.assert_allclose(ha.numpy(), g["out_h_alpha"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(hr.numpy(), g["out_h_rgb"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(raw.numpy(), g["out_raw"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(r1.numpy(), g["out_r1_rgb"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(r2.numpy(), g["out_r2_rgb"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(b.numpy(), g["out_b_rgb"], rtol=0, atol=1e-5)


def test_param_count_matches_reference():
    assert O.CfnConfig().n_params() == 2360546  # SURVEY §8 [probe]
    p = O.make_params(O.CfnConfig(), 0)
    assert sum(v.numel() for v in p.values()) == 2360546


def test_sample_pdf_sequential_oracle_agrees_with_upstream_formulation():
    """The bit-exact (sequential fp32) oracle and the torch-library formulation agree to rounding."""
    g = torch.Generator().manual_seed(4)
    B, M, Nf = 64, 63, 128
    bins = torch.sort(torch.rand(B, M, generator=g) * 5 + 1, -1).values
    w = torch.rand(B, M - 1, generator=g) ** 4
    w[:4] = 0.0  # empty-weight rays (pdf uniform through the +1e-5)
    for u in (torch.linspace(0, 1, Nf).expand(B, Nf).contiguous(), torch.rand(B, Nf, generator=g)):
        s, below = O.sample_pdf(bins.numpy(), w.numpy(), u.numpy())
        s_t = O.sample_pdf_upstream_torch(bins, w, u).numpy()
        # the algorithm itself is discontinuous where denom crosses its 1e-5 guard (tiny-pdf bins), so
        # a few samples may legitimately land elsewhere in their bin; everything else agrees to rounding
        diff = np.abs(s - s_t)
        assert np.median(diff) <= 1e-6 and (diff > 2e-5).mean() < 0.01
        assert below.min() >= 0 and below.max() <= M - 1
        assert np.all(s >= bins.numpy()[:, :1] - 1e-6) and np.all(s <= bins.numpy()[:, -1:] + 1e-6)
    # deterministic u is sorted -> samples sorted
    s, _ = O.sample_pdf(bins.numpy(), w.numpy(), torch.linspace(0, 1, Nf).expand(B, Nf).numpy())
    assert np.all(np.diff(s, axis=-1) >= 0)


def test_depth_supervised_trainer_body_golden():
    """The shipped recipe's loss (--colmap_depth, main:1009-1055) from the unmodified reference: colour rays in the first
    network call, depth rays in the second (own noise), entropy = the first call's scalar, depth MSE on mean_K depth."""
    g, cfg, p = load_golden("train_depth_small")
    n_rgb = int(g["in_n_rgb"])
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    out = O.render_rays(pr, cfg, T(g["in_rays"]), T(g["in_eps_alpha"]), T(g["in_eps_rgb"]), True,
                        t_rand=T(g["in_t_rand"]), faithful=True, netchunk=int(g["in_netchunk"]))
    np.testing.assert_allclose(out["rgb_map"].detach().numpy(), g["out_rgb_map"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["depth_map"].detach().numpy(), g["out_depth_map"], rtol=0, atol=2e-5)
    assert len(out["entropy_calls"]) == 2 and out["entropy_calls"][0][1] == n_rgb * 128
    res = O.trainer_loss(out, T(g["in_target"]), cfg.K, float(g["in_beta1"]), target_depth=T(g["in_target_depth"]),
                         depth_lambda=float(g["in_depth_lambda"]))
    for k in ("loss_entropy", "loss_nll", "depth_loss", "loss", "psnr"):
        np.testing.assert_allclose(float(res[k]), float(g["out_" + k]), rtol=2e-5, err_msg=k)
    res["loss"].backward()
    names = [str(n) for n in g["out_grad_names"]]
    for n, ref_norm in zip(names, g["out_grad_norms"]):
        gr = pr[n].grad
        mine = 0.0 if gr is None else float(gr.double().pow(2).sum().sqrt())
        assert abs(mine - ref_norm) <= 1e-3 * max(ref_norm, 1e-7) + 1e-9, (n, mine, ref_norm)
    for k in g:
        if k.startswith("grad__"):
            np.testing.assert_allclose(pr[k[6:]].grad.numpy(), g[k], rtol=1e-3, atol=1e-3 * np.abs(g[k]).max() + 1e-9,
                                       err_msg=k)
