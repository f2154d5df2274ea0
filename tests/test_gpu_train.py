"""GPU parity tests of the TRAINING half of the path and of the reference-facing boundary,
all through the C-ABI:

  * the shipped recipe's depth-supervised trainer body (--colmap_depth, --depth_lambda; main:1009-1067) against a fixture
    the UNMODIFIED reference generated (loss terms and autograd gradients), on the autograd path and on FusedTrainStep;
  * the K4 adjoints of depth_map / rgb_map against the oracle's autograd;
  * PSNR after a fixed number of optimisation steps (north star: within 0.1 dB of the reference path) in a run that can
    fail: 200 Adam steps with the reference's lr decay fitting a frozen teacher, required gain >= 3 dB;
  * train-mode coarse + fine (gradients into both networks);
  * `network_fn(embedded, is_val, is_test)`; `install()` behind the real `run_nerf_uncertainty_NF` module staged under
    oracle/_ref (render() and one iteration of the trainer body).
"""
import math
import os

import numpy as np
import pytest
import torch

from conftest import T, load_golden
from oracle import cfnerf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cf():
    import cfnerf_b200
    return cfnerf_b200


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def make_net(cf, cfg, params, sa, sr, dev):
    return cf.NeRFFlowsParams.from_oracle_params(cfg, params, sa, sr).to(dev)


def _live(named):
    return [(n, q) for n, q in named if not n.startswith("alpha_linear") and not n.startswith("alpha_std_linear")]


# ------------------------------------------------------------------------------------------------
# F1 / A12: depth-supervised trainer body vs the unmodified reference
# ------------------------------------------------------------------------------------------------
def _check_grads_vs_golden(g, grads, rtol):
    names = [str(n) for n in g["out_grad_names"]]
    for n, ref_norm in zip(names, g["out_grad_norms"]):
        gr = grads.get(n)
        mine = 0.0 if gr is None else float(gr.double().pow(2).sum().sqrt())
        assert abs(mine - ref_norm) <= rtol * max(ref_norm, 1e-7) + 1e-8, f"|grad {n}| = {mine} vs {ref_norm}"
    for k in g:
        if k.startswith("grad__"):
            ref = g[k]
            np.testing.assert_allclose(grads[k[6:]].cpu().numpy().reshape(ref.shape), ref, rtol=rtol,
                                       atol=rtol * np.abs(ref).max() + 1e-9, err_msg=k)
        if k.startswith("gradrows__"):
            ref = g[k]
            np.testing.assert_allclose(grads[k[10:]][:4].cpu().numpy(), ref, rtol=rtol,
                                       atol=rtol * np.abs(ref).max() + 1e-9, err_msg=k)


def test_depth_supervised_trainer_body_vs_reference_golden(cf, dev):
    """fp32 check mode: render_rays over [colour rays | depth rays] with one latent draw per network call (netchunk),
    trainer_loss (KDE-NLL + beta1 * first-call entropy + depth_lambda * depth MSE), autograd backward — against the
    reference's own numbers at 2e-3 (observed ~1e-5)."""
    g, cfg, p = load_golden("train_depth_small")
    sa, sr = O.make_latents(cfg, int(g["seed"]))
    net = make_net(cf, cfg, p, sa, sr, dev)
    n_rgb = int(g["in_n_rgb"])
    out = cf.render_rays(T(g["in_rays"]).to(dev), net, None, 128, True, False, K_samples=cfg.K, perturb=1.,
                         raw_noise_std=1., t_rand=T(g["in_t_rand"]).to(dev), eps_alpha=T(g["in_eps_alpha"]).to(dev),
                         eps_rgb=T(g["in_eps_rgb"]).to(dev), precision="fp32", netchunk=int(g["in_netchunk"]))
    np.testing.assert_allclose(out["rgb_map"].detach().cpu().numpy(), g["out_rgb_map"], rtol=0, atol=1e-5)
    np.testing.assert_allclose(out["depth_map"].detach().cpu().numpy(), g["out_depth_map"], rtol=0, atol=3e-5)
    res = cf.trainer_loss(out, T(g["in_target"]).to(dev), cfg.K, float(g["in_beta1"]),
                          target_depth=T(g["in_target_depth"]).to(dev), depth_lambda=float(g["in_depth_lambda"]))
    for k in ("loss_entropy", "loss_nll", "depth_loss", "loss", "psnr"):
        np.testing.assert_allclose(float(res[k]), float(g["out_" + k].reshape(-1)[0]), rtol=3e-5, err_msg=k)
    net.zero_grad()
    res["loss"].backward()
    _check_grads_vs_golden(g, {n: q.grad for n, q in net.named_parameters() if q.grad is not None}, 2e-3)
    assert n_rgb * 128 == int(g["in_netchunk"])


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_train_step_with_depth_rays_vs_reference_golden(cf, dev, use_graph):
    """FusedTrainStep (the autograd-free chain: fused trainer loss, per-ray log-det seeds, device-clock Adam) on the same
    fixture: loss terms and the raw gradient buffer against the reference's autograd, eager and replayed from CUDA
    graphs (step 1 runs eagerly and captures, step 2 replays: same inputs, second Adam step)."""
    from cfnerf_b200 import dist as D
    g, cfg, p = load_golden("train_depth_small")
    sa, sr = O.make_latents(cfg, int(g["seed"]))
    net = make_net(cf, cfg, p, sa, sr, dev)
    n_rgb = int(g["in_n_rgb"])
    rays = T(g["in_rays"]).to(dev)
    tr = D.FusedTrainStep(net, lr=0.0, precision="fp32", beta1=float(g["in_beta1"]),
                          depth_lambda=float(g["in_depth_lambda"]), netchunk=int(g["in_netchunk"]), use_graph=use_graph)
    for it in range(2 if use_graph else 1):       # lr = 0: the weights do not move, every step sees the fixture
        res = tr.step(rays[:n_rgb], T(g["in_target"]).to(dev), t_rand=T(g["in_t_rand"]).to(dev),
                      eps_alpha=T(g["in_eps_alpha"]).to(dev), eps_rgb=T(g["in_eps_rgb"]).to(dev),
                      depth_rays=rays[n_rgb:], target_depth=T(g["in_target_depth"]).to(dev))
        torch.cuda.synchronize()
        for k in ("loss_entropy", "loss_nll", "depth_loss", "loss", "psnr"):
            np.testing.assert_allclose(float(res[k]), float(g["out_" + k].reshape(-1)[0]), rtol=3e-5, err_msg=f"{k} step {it}")
        _check_grads_vs_golden(g, dict(zip(tr.eng.names, tr.grads)), 2e-3)


def test_k4_adjoints_of_depth_and_colour_maps_vs_oracle_autograd(cf, dev):
    """K4's g_depth_map / g_rgb_map adjoints in isolation: L = <A, rgb_map> + <Bm, depth_map> with random A, Bm, gradients
    of every parameter against autograd through the oracle (fp32 check mode, white background on: exercises gA)."""
    cfg = O.CfnConfig(W=128, D=4, K=16, h_alpha=32, h_rgb=32)
    p = O.make_params(cfg, 6, "lively")
    sa, sr = O.make_latents(cfg, 6)
    net = make_net(cf, cfg, p, sa, sr, dev)
    B = 6
    rays = O.synthetic_rays(B, 17)
    g = torch.Generator().manual_seed(4)
    t_rand = torch.rand(B, 128, generator=g)
    ea, er = torch.randn(cfg.K, 1, generator=g), torch.randn(cfg.K, 3, generator=g)
    A, Bm = torch.randn(B, 3, cfg.K, generator=g), torch.randn(B, cfg.K, generator=g) * 0.2
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    o_ref = O.render_rays(pr, cfg, rays, ea, er, True, t_rand=t_rand, white_bkgd=True, faithful=False)
    ((o_ref["rgb_map"] * A).sum() + (o_ref["depth_map"] * Bm).sum()).backward()
    out = cf.render_rays(rays.to(dev), net, None, 128, True, False, perturb=1., white_bkgd=True, t_rand=t_rand.to(dev),
                         eps_alpha=ea.to(dev), eps_rgb=er.to(dev), precision="fp32")
    net.zero_grad()
    ((out["rgb_map"] * A.to(dev)).sum() + (out["depth_map"] * Bm.to(dev)).sum()).backward()
    for n, q in _live(net.named_parameters()):
        ref = pr[n].grad
        if ref is None:
            continue
        err = (q.grad.cpu().double() - ref.double()).norm().item()
        assert err <= 2e-3 * ref.double().norm().item() + 1e-7, (n, err, ref.norm().item())


# ------------------------------------------------------------------------------------------------
# the 0.1 dB bar, in a run that can fail
# ------------------------------------------------------------------------------------------------
def _teacher_setup():
    cfg = O.CfnConfig(D=4, W=128, K=8, h_alpha=32, h_rgb=32)
    teacher = O.make_params(cfg, 0, "lively")
    p0 = O.make_params(cfg, 5, "default")
    sa, sr = O.make_latents(cfg, 5)
    B = 64
    rays = O.synthetic_rays(B, 13)
    ea_t, er_t = O.test_latents(*O.make_latents(cfg, 0))
    with torch.no_grad():
        target = O.render_rays(teacher, cfg, rays, ea_t, er_t, False, faithful=False)["rgb_map"].mean(-1)
    return cfg, p0, sa, sr, rays, target


_ORACLE_RUN = {}


def _oracle_training_run(steps, lr, decay_steps):
    """The reference trainer on the host: oracle forward (pinned to the reference, incl. its autograd gradients),
    torch.optim.Adam, lr schedule of main:1073-1077 (global_step lags the iteration by one)."""
    key = (steps, lr, decay_steps)
    if key in _ORACLE_RUN:
        return _ORACLE_RUN[key]
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    cfg, p0, sa, sr, rays, target = _teacher_setup()
    p = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    live = [k for k in p if not k.startswith("alpha_linear") and not k.startswith("alpha_std_linear")]
    opt = torch.optim.Adam([p[k] for k in live], lr=lr, betas=(0.9, 0.999))
    g = torch.Generator().manual_seed(12)
    draws = []
    for it in range(steps):
        for grp in opt.param_groups:
            grp["lr"] = lr * 0.1 ** (max(it - 1, 0) / decay_steps)
        t_rand = torch.rand(rays.shape[0], 128, generator=g)
        ea, er = torch.randn(cfg.K, 1, generator=g), torch.randn(cfg.K, 3, generator=g)
        draws.append((t_rand, ea, er))
        out = O.render_rays(p, cfg, rays, ea, er, True, t_rand=t_rand, faithful=False)
        l = O.kde_nll_loss(out["rgb_map"], target, out["loss_entropy"], cfg.K, 0.01)
        opt.zero_grad()
        l["loss"].backward()
        opt.step()
    _ORACLE_RUN[key] = ({k: v.detach() for k, v in p.items()}, draws)
    return _ORACLE_RUN[key]


def _eval_psnr_oracle(cfg, p, sa, sr, rays, target):
    ea, er = O.test_latents(sa, sr)
    with torch.no_grad():
        m = O.render_rays(p, cfg, rays, ea, er, False, faithful=False)["rgb_map"].mean(-1)
    return -10.0 * math.log10(float(((m - target) ** 2).mean()))


@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_psnr_after_200_steps_within_0p1_db_of_the_reference_trainer(cf, dev, precision):
    """Fit the mean colours a frozen teacher renders, 200 Adam steps, lr 1e-3 decayed 10x per 100 steps (the reference's
    schedule, which also makes the end point a property of the optimisation rather than of chaotic step-to-step noise:
    with a constant lr even two fp32 runs differing in the last bit part by ~1 dB).  Both sides see the same rays,
    targets, jitter and latent draws.  The GPU trainer (tensor-core chains) must end within 0.1 dB of the oracle
    trainer, and both must have gained >= 3 dB — a wrong gradient fails one or the other."""
    from cfnerf_b200 import dist as D
    steps, lr, decay = 200, 1e-3, 100.0
    cfg, p0, sa, sr, rays, target = _teacher_setup()
    p_ref, draws = _oracle_training_run(steps, lr, decay)
    psnr0 = _eval_psnr_oracle(cfg, p0, sa, sr, rays, target)
    psnr_ref = _eval_psnr_oracle(cfg, p_ref, sa, sr, rays, target)
    net = make_net(cf, cfg, p0, sa, sr, dev)
    tr = D.FusedTrainStep(net, lr=lr, precision=precision, lrate_decay=decay / 1000.0)
    rd, td = rays.to(dev), target.to(dev)
    for t_rand, ea, er in draws:
        tr.step(rd, td, t_rand=t_rand.to(dev), eps_alpha=ea.to(dev), eps_rgb=er.to(dev), want_loss=False)
    torch.cuda.synchronize()
    assert abs(float(tr.adam_state[0]) - steps) < 0.5
    np.testing.assert_allclose(float(tr.adam_state[1]), lr * 0.1 ** ((steps - 2) / decay), rtol=1e-4)
    # evaluate the GPU-trained weights with the same (host, fp32) evaluator as the reference run
    p_mine = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    psnr_mine = _eval_psnr_oracle(cfg, p_mine, sa, sr, rays, target)
    # ... and through the library's own render (fp32 check mode): the two evaluators agree
    out = cf.render_rays(rd, net, None, 128, False, False, precision="fp32", want_kstats=True)
    psnr_lib = -10.0 * math.log10(float(((out["kstats"][:, 0:3].cpu() - target) ** 2).mean()))
    print(f"PSNR start {psnr0:.3f} dB, oracle trainer {psnr_ref:.3f} dB, {precision} chain {psnr_mine:.3f} dB "
          f"(library evaluator {psnr_lib:.3f} dB)")
    assert psnr_ref - psnr0 >= 3.0 and psnr_mine - psnr0 >= 3.0
    assert abs(psnr_mine - psnr_ref) <= 0.1, (psnr_mine, psnr_ref)
    assert abs(psnr_lib - psnr_mine) <= 0.02


# ------------------------------------------------------------------------------------------------
# A10 in train mode: gradients into both networks
# ------------------------------------------------------------------------------------------------
def test_coarse_fine_train_mode_gradients_into_both_networks(cf, dev):
    """Extension spec A10: loss = KDE-NLL on the fine maps + KDE-NLL on the coarse maps (+ both entropy terms); the fine
    grid is detached (sample_pdf).  Gradients of the coarse and the fine network against autograd through the oracle
    composition, fp32 check mode."""
    cfg = O.CfnConfig(W=128, D=4, K=16, h_alpha=32, h_rgb=32)
    pc, pf = O.make_params(cfg, 0, "lively"), O.make_params(cfg, 1, "lively")
    sa, sr = O.make_latents(cfg, 0)
    net_c, net_f = make_net(cf, cfg, pc, sa, sr, dev), make_net(cf, cfg, pf, sa, sr, dev)
    B, Nc, Nf = 6, 64, 128
    rays = O.synthetic_rays(B, 8)
    g = torch.Generator().manual_seed(21)
    target = torch.rand(B, 3, generator=g)
    t_rand = torch.rand(B, Nc, generator=g)
    u = torch.rand(B, Nf, generator=g)
    ea, er = torch.randn(cfg.K, 1, generator=g), torch.randn(cfg.K, 3, generator=g)
    prc = {k: v.clone().requires_grad_(True) for k, v in pc.items()}
    prf = {k: v.clone().requires_grad_(True) for k, v in pf.items()}
    ref = O.render_rays_hier(prc, prf, cfg, rays, ea, er, True, Nc, Nf, t_rand=t_rand, u=u)
    l_ref = (O.kde_nll_loss(ref["rgb_map"], target, ref["loss_entropy"], cfg.K, 0.01)["loss"] +
             O.kde_nll_loss(ref["rgb0"], target, ref["loss_entropy0"], cfg.K, 0.01)["loss"])
    l_ref.backward()
    out = cf.render_rays(rays.to(dev), net_c, None, Nc, True, False, K_samples=cfg.K, perturb=1., N_importance=Nf,
                         network_fine=net_f, t_rand=t_rand.to(dev), u=u.to(dev), eps_alpha=ea.to(dev),
                         eps_rgb=er.to(dev), precision="fp32")
    assert (out["z_vals"].cpu() - ref["z_vals"]).abs().max().item() <= 1e-3
    l = (cf.kde_nll_loss(out["rgb_map"], target.to(dev), out["loss_entropy"], cfg.K, 0.01)["loss"] +
         cf.kde_nll_loss(out["rgb0"], target.to(dev), out["loss_entropy0"], cfg.K, 0.01)["loss"])
    net_c.zero_grad(); net_f.zero_grad()
    l.backward()
    torch.cuda.synchronize()
    assert abs(float(l) - float(l_ref)) <= 2e-3 * max(1.0, abs(float(l_ref)))
    worst = 0.0
    for net, pr, tag in ((net_c, prc, "coarse"), (net_f, prf, "fine")):
        for n, q in _live(net.named_parameters()):
            r = pr[n].grad
            if r is None or r.norm() == 0:
                continue
            assert q.grad is not None, (tag, n)
            rel = (q.grad.cpu().double() - r.double()).norm().item() / r.double().norm().item()
            worst = max(worst, rel)
            # a resampled depth that lands in another CDF bin moves one of 192 samples of one ray: loose per tensor
            assert rel <= 2e-2, (tag, n, rel)
    print(f"coarse+fine train mode: worst relative gradient distance {worst:.2e}")


# ------------------------------------------------------------------------------------------------
# boundary: network_fn(embedded, is_val, is_test) and install() behind the real module
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("fp16", 2e-3)])
def test_network_fn_call_surface_on_embedded_inputs(cf, dev, precision, tol):
    """`network_fn(embedded, is_val, is_test)` (models.py:188) — what `batchify` calls (main:55) — on the reference's own
    embedded input: raw (M,K,4) and the zeros second value against the golden the unmodified reference wrote."""
    g, cfg, p = load_golden("network_canonical")
    sa, sr = T(g["in_sample_alpha"]), T(g["in_sample_rgb"])
    net = make_net(cf, cfg, p, sa, sr, dev)
    emb = T(g["out_embedded"]).to(dev)
    raw, zeros = net(emb, is_val=False, is_test=True, precision=precision)
    assert raw.shape == (emb.shape[0], cfg.K, 4) and zeros.shape == raw.shape and float(zeros.abs().max()) == 0.0
    err = (raw.cpu() - T(g["out_raw"])).abs().max().item()
    assert err <= tol, err
    # train-mode call: entropy scalar broadcast to (M,K,1) (models.py:291), same draws -> oracle value
    gen = torch.Generator().manual_seed(2)
    ea, er = torch.randn(cfg.K, 1, generator=gen), torch.randn(cfg.K, 3, generator=gen)
    raw_t, ent = net(emb, is_val=False, is_test=False, eps_alpha=ea.to(dev), eps_rgb=er.to(dev), precision="fp32")
    with torch.no_grad():
        raw_o, ent_o = O.nerf_flows_forward(p, cfg, T(g["out_embedded"]), ea, er, True, faithful=False)
    assert ent.shape == (emb.shape[0], cfg.K, 1)
    assert (raw_t.cpu() - raw_o).abs().max().item() <= 2e-5
    np.testing.assert_allclose(float(ent[0, 0, 0]), float(ent_o), rtol=2e-5)
    with pytest.raises(ValueError):
        net(emb[:, :50], False, True)


def _real_reference():
    from oracle import refload
    if not refload.reference_available():
        pytest.skip("oracle/_ref not staged (python oracle/build_ref.py in the build container)")
    main, _ = refload.load_reference()
    return main


def test_install_behind_the_real_reference_module(cf, dev):
    """`cfnerf_b200.install(run_nerf_uncertainty_NF)` on the UNMODIFIED module object (staged under oracle/_ref): its own
    `render()` -> `batchify_rays` then reach the CUDA path through the rebound globals (main:93, 540).  Full-image branch
    (c2w=pose, render_kwargs_test incl. the reference's network_query_fn lambda) against the oracle, in the fp32 check
    mode and in the default (fp16) render mode."""
    R = _real_reference()
    saved = (R.render_rays, R.raw2outputs)
    try:
        cfg = O.CfnConfig()
        p = O.make_params(cfg, 0, "lively")
        sa, sr = O.make_latents(cfg, 0)
        net = torch.nn.DataParallel(make_net(cf, cfg, p, sa, sr, dev), device_ids=[0])              # main:330
        embed_fn, _ = R.get_embedder(cfg.L_pos, 0)
        embeddirs_fn, _ = R.get_embedder(cfg.L_dir, 0)
        nq = lambda inputs, viewdirs, network_fn, is_val, is_test: R.run_network(                    # noqa: E731 main:333-336
            inputs, viewdirs, network_fn, is_val, is_test, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=65536)
        kwargs_test = dict(is_train=False, uniformsample=False, network_query_fn=nq, perturb=False, N_importance=0,
                           N_samples=128, K_samples=cfg.K, network_fn=net, use_viewdirs=True, white_bkgd=False,
                           raw_noise_std=0., ndc=False, lindisp=False, retraw=True)                  # main:382-405
        H, W, focal = 6, 8, 7.0
        o, d = O.get_rays(H, W, focal, torch.eye(4)[:3])
        ea, er = O.test_latents(sa, sr)
        with torch.no_grad():
            ref = O.render_rays(p, cfg, O.pack_ray_batch(o, d, 1.2, 8.0), ea, er, False, faithful=False)
        for prec, tol in (("fp32", 1e-5), (None, 2e-3)):
            cf.install(R, precision=prec or "fp16")
            rgb, disp, depth, extras = R.render(H, W, focal, chunk=20, c2w=torch.eye(4)[:3].to(dev), near=1.2, far=8.0,
                                                **kwargs_test)
            assert rgb.shape == (H, W, 3, cfg.K) and disp.shape == (H, W, cfg.K) and extras == {}
            assert (rgb.reshape(-1, 3, cfg.K).cpu() - ref["rgb_map"]).abs().max().item() <= tol
            assert (depth.reshape(-1, cfg.K).cpu() - ref["depth_map"]).abs().max().item() <= tol
        # the untouched reference `run_network` -> `batchify` -> network_fn(embedded, ...) also lands on the CUDA path
        pts = torch.randn(5, 7, 3, generator=torch.Generator().manual_seed(1))
        vd = torch.nn.functional.normalize(torch.randn(5, 3, generator=torch.Generator().manual_seed(2)), dim=-1)
        cf.install(R, precision="fp32")
        raw, zeros = nq(pts.to(dev), vd.to(dev), net, False, True)     # DataParallel.__call__ -> NeRFFlowsParams.forward
        with torch.no_grad():
            raw_o, _ = O.run_network(p, cfg, pts, vd, ea, er, False, faithful=False)
        assert raw.shape == (5, 7, cfg.K, 4) and (raw.cpu() - raw_o).abs().max().item() <= 2e-5
    finally:
        R.render_rays, R.raw2outputs = saved
        cf.configure(precision="fp16", train_precision="tf32")   # restore the module defaults


def test_one_real_trainer_body_iteration_behind_install(cf, dev):
    """One iteration of the reference trainer body (main:1009-1067) executed against the REAL module after install():
    `R.render(rays=batch_rays_train, **render_kwargs_train)` with depth rays appended, the loss lines of main:1018-1055
    written with the reference's own helpers (`img2mse`, `mse2psnr`), `loss.backward()`, `torch.optim.Adam.step()` on the
    module's parameters (main:339) — loss terms and the updated weights against the oracle doing the same on the host."""
    R = _real_reference()
    saved = (R.render_rays, R.raw2outputs)
    try:
        cf.install(R, train_precision="fp32", netchunk=8 * 128)
        cfg = O.CfnConfig(W=256, K=64, h_alpha=32)
        p = O.make_params(cfg, 4, "lively")
        sa, sr = O.make_latents(cfg, 4)
        model = torch.nn.DataParallel(make_net(cf, cfg, p, sa, sr, dev), device_ids=[0])
        grad_vars = list(model.parameters())
        optimizer = torch.optim.Adam(params=grad_vars, lr=5e-4, betas=(0.9, 0.999))                  # main:339
        n_rgb, n_depth, K = 8, 4, cfg.K
        rays = O.synthetic_rays(n_rgb + n_depth, 5)
        g = torch.Generator().manual_seed(6)
        target_s = torch.rand(n_rgb, 3, generator=g)
        target_depth = 1.2 + 6.8 * torch.rand(n_depth, generator=g)
        t_rand = torch.rand(n_rgb + n_depth, 128, generator=g)
        eps_a = torch.randn(2, K, 1, generator=g)
        eps_c = torch.randn(2, K, 3, generator=g)
        batch_rays = torch.stack([rays[:, 0:3], rays[:, 3:6]], 0).to(dev)
        kwargs_train = dict(is_train=True, uniformsample=False, network_query_fn=None, perturb=1.0, N_importance=0,
                            N_samples=128, K_samples=K, network_fn=model, use_viewdirs=True, white_bkgd=False,
                            raw_noise_std=1.0, ndc=False, lindisp=False,
                            t_rand=t_rand.to(dev), eps_alpha=eps_a.to(dev), eps_rgb=eps_c.to(dev))    # injected draws
        rgbs, disp, depth, extras = R.render(8, 8, 10.0, chunk=1024 * 32, rays=batch_rays, near=1.2, far=8.0,
                                             verbose=False, retraw=False, **kwargs_train)             # main:1014-1016
        # ---- main:1018-1055 with the reference's helpers
        N_batch = n_rgb
        depth_m = torch.mean(depth, -1)
        rgbs_ = rgbs[:N_batch, :]
        depth_col = depth_m[N_batch:]
        extras_ = {x: extras[x][:N_batch] for x in extras}
        rgb_mean = torch.mean(rgbs_, -1)
        ts, td = target_s.to(dev), target_depth.to(dev)
        # mse2psnr (helpers:16) builds torch.Tensor([10.]) on the DEFAULT device (the host script sets it to CUDA,
        # main:1201): same expression with the constant on the device
        psnr_train = -10. * torch.log(R.img2mse(rgb_mean, ts)) / torch.log(torch.tensor([10.], device=dev))
        rgb_std = torch.std(rgbs_, -1) * K / (K - 1)
        H_sqrt = (rgb_std.detach() * (0.8 / K) ** (-1 / 7) + 1e-05)[..., None]
        r1 = torch.exp(-((rgbs_ - ts[..., None]) ** 2) / (2 * H_sqrt * H_sqrt))
        r2 = (2 * math.pi) ** (-1.5) / H_sqrt
        loss_nll = -torch.log((r1 * r2).mean(-1) + 1e-05).mean()
        loss_entropy = extras_["loss_entropy"].mean()
        depth_loss = R.img2mse(depth_col, td)
        loss = loss_nll + 0.01 * loss_entropy + 0.01 * depth_loss
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        torch.cuda.synchronize()
        # ---- the same on the host
        pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        live = [k for k in pr]
        opt_ref = torch.optim.Adam([pr[k] for k in live], lr=5e-4, betas=(0.9, 0.999))
        o_ref = O.render_rays(pr, cfg, rays, eps_a, eps_c, True, t_rand=t_rand, faithful=False, netchunk=8 * 128)
        res = O.trainer_loss(o_ref, target_s, K, 0.01, target_depth=target_depth, depth_lambda=0.01)
        opt_ref.zero_grad()
        res["loss"].backward()
        opt_ref.step()
        for mine, ref_k in ((loss, "loss"), (loss_nll, "loss_nll"), (loss_entropy, "loss_entropy"),
                            (depth_loss, "depth_loss"), (psnr_train, "psnr")):
            np.testing.assert_allclose(float(mine), float(res[ref_k]), rtol=5e-5, err_msg=ref_k)
        sd = model.module.state_dict()
        for k in live:
            if pr[k].grad is None:
                continue
            # Adam's first step moves every weight by lr * sign(grad): equal up to gradients that are numerically zero
            d = (sd[k].cpu() - pr[k].detach()).abs()
            assert (d > 1e-5).float().mean().item() <= 1e-2, (k, d.max().item())
    finally:
        R.render_rays, R.raw2outputs = saved
        cf.configure(precision="fp16", train_precision="tf32", netchunk=1024 * 64)


# ------------------------------------------------------------------------------------------------
# CUDA-graph replay, second device
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["bf16", "tf32"])
def test_fused_train_step_graph_replay_equals_eager(cf, dev, precision):
    """Three optimisation steps replayed from the captured CUDA graphs (device-resident optimiser clock) against the same
    three steps launched eagerly: same losses, same weights up to the fp32-atomics noise of the split-K wgrads."""
    from cfnerf_b200 import dist as D
    cfg = O.CfnConfig()
    p = O.make_params(cfg, 4, "lively")
    sa, sr = O.make_latents(cfg, 4)
    B = 40
    rays = O.synthetic_rays(B, 6).to(dev)
    g = torch.Generator().manual_seed(9)
    target = torch.rand(B, 3, generator=g).to(dev)
    draws = [(torch.rand(B, 128, generator=g).to(dev), torch.randn(cfg.K, 1, generator=g).to(dev),
              torch.randn(cfg.K, 3, generator=g).to(dev)) for _ in range(3)]
    nets = [make_net(cf, cfg, p, sa, sr, dev) for _ in range(2)]
    trainers = [D.FusedTrainStep(nets[0], lr=5e-4, precision=precision, lrate_decay=0.25, use_graph=False),
                D.FusedTrainStep(nets[1], lr=5e-4, precision=precision, lrate_decay=0.25, use_graph=True)]
    for it, (t_rand, ea, er) in enumerate(draws):
        la, lb = (tr.step(rays, target, t_rand=t_rand, eps_alpha=ea, eps_rgb=er) for tr in trainers)
        assert abs(float(la["loss"]) - float(lb["loss"])) <= 2e-4 * max(1.0, abs(float(la["loss"]))), it
    assert trainers[1]._shapes[(B, 0)]["graph_a"] is not None
    assert torch.equal(trainers[0].adam_state, trainers[1].adam_state)
    d = (trainers[0].flat_param - trainers[1].flat_param).abs()
    assert d.max().item() <= 3 * 2 * 5e-4 + 1e-6 and (d > 5e-5).float().mean().item() <= 1e-2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_second_device_in_the_same_process(cf):
    """Function attributes (dynamic shared memory opt-in) and the SM count are per device: a render and a training step on
    cuda:1 after cuda:0 was used must work and agree with cuda:0."""
    from cfnerf_b200 import dist as D
    cfg = O.CfnConfig()
    p = O.make_params(cfg, 0, "lively")
    sa, sr = O.make_latents(cfg, 0)
    rays = O.synthetic_rays(32, 3)
    g = torch.Generator().manual_seed(1)
    target, t_rand = torch.rand(32, 3, generator=g), torch.rand(32, 128, generator=g)
    ea, er = torch.randn(cfg.K, 1, generator=g), torch.randn(cfg.K, 3, generator=g)
    outs, losses = [], []
    for i in (0, 1):
        d = torch.device("cuda", i)
        net = make_net(cf, cfg, p, sa, sr, d)
        for prec in ("fp16", "tf32", "fp32"):
            outs.append(cf.render_rays(rays.to(d), net, None, 128, False, False, precision=prec)["rgb_map"].cpu())
        tr = D.FusedTrainStep(net, precision="bf16")
        losses.append(float(tr.step(rays.to(d), target.to(d), t_rand=t_rand.to(d), eps_alpha=ea.to(d),
                                    eps_rgb=er.to(d))["loss"]))
    for a, b in zip(outs[:3], outs[3:]):
        assert torch.equal(a, b)
    assert abs(losses[0] - losses[1]) <= 1e-4 * max(1.0, abs(losses[0]))


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_deterministic_mode_gives_bitwise_identical_gradients(cf, dev, precision):
    """cfn_set_deterministic: two-pass split-K weight gradients (no fp32 atomics).  The same training step twice gives
    bit-identical gradient buffers; the default (atomic) mode agrees with it to fp32 summation-order noise."""
    from cfnerf_b200 import dist as D
    cfg = O.CfnConfig()
    p = O.make_params(cfg, 4, "lively")
    sa, sr = O.make_latents(cfg, 4)
    B = 96
    rays = O.synthetic_rays(B, 6).to(dev)
    g = torch.Generator().manual_seed(9)
    target = torch.rand(B, 3, generator=g).to(dev)
    t_rand = torch.rand(B, 128, generator=g).to(dev)
    ea, er = torch.randn(cfg.K, 1, generator=g).to(dev), torch.randn(cfg.K, 3, generator=g).to(dev)
    grads = []
    for det in (True, True, False):
        net = make_net(cf, cfg, p, sa, sr, dev)
        tr = D.FusedTrainStep(net, lr=0.0, precision=precision, deterministic=det)
        tr.eng.set_deterministic(det)       # engines are cached per (module, precision): set explicitly both ways
        tr.step(rays, target, t_rand=t_rand, eps_alpha=ea, eps_rgb=er, want_loss=False)
        torch.cuda.synchronize()
        grads.append(tr.flat_grad.clone())
        tr.eng.set_deterministic(False)
    assert torch.equal(grads[0], grads[1]), "deterministic mode must be bitwise reproducible"
    ref = grads[0].double()
    assert (grads[2].double() - ref).norm().item() <= 1e-5 * ref.norm().item()
    assert float(ref.abs().max()) > 0


@pytest.mark.parametrize("precision,n_rays,tol", [("fp32", 1024, 2e-5), ("bf16", 4096, 2e-3)])
def test_global_batch_gradient_equals_mean_of_shard_gradients(cf, dev, precision, n_rays, tol):
    """Data parallelism on one GPU (N-rank training == 1-rank global-batch training): the gradient of the
    whole batch — 8 network calls of the reference, one latent draw each — equals the mean of the gradients of its 8
    equal shards, each computed alone with its own draw, which is what the all-reduce of `FusedTrainStep` averages.  The
    bf16 case runs at BASELINE configs[2]'s full size (4096-ray global batch, 512 rays per rank)."""
    from cfnerf_b200 import dist as D
    cfg = O.CfnConfig()
    p = O.make_params(cfg, 4, "lively")
    sa, sr = O.make_latents(cfg, 4)
    shards = 8
    per = n_rays // shards
    rays = O.synthetic_rays(n_rays, 6).to(dev)
    g = torch.Generator().manual_seed(9)
    target = torch.rand(n_rays, 3, generator=g).to(dev)
    t_rand = torch.rand(n_rays, 128, generator=g).to(dev)
    ea, er = torch.randn(shards, cfg.K, 1, generator=g).to(dev), torch.randn(shards, cfg.K, 3, generator=g).to(dev)
    net = make_net(cf, cfg, p, sa, sr, dev)
    full = D.FusedTrainStep(net, lr=0.0, precision=precision, netchunk=per * 128)
    lf = full.step(rays, target, t_rand=t_rand, eps_alpha=ea, eps_rgb=er)
    assert full._shapes[(n_rays, 0)]["G"] == shards
    g_full = full.flat_grad.clone()
    net2 = make_net(cf, cfg, p, sa, sr, dev)
    part = D.FusedTrainStep(net2, lr=0.0, precision=precision, netchunk=per * 128)
    acc, losses = torch.zeros_like(g_full, dtype=torch.float64), []
    for s in range(shards):
        sl = slice(s * per, (s + 1) * per)
        ls = part.step(rays[sl], target[sl], t_rand=t_rand[sl], eps_alpha=ea[s], eps_rgb=er[s])
        acc += part.flat_grad.double()
        losses.append(float(ls["loss"]))
    g_mean = acc / shards
    rel = (g_full.double() - g_mean).norm().item() / g_mean.norm().item()
    print(f"global batch vs mean of {shards} shards [{precision}, {n_rays} rays]: relative gradient distance {rel:.2e}")
    assert rel <= tol, rel
    assert abs(float(lf["loss"]) - sum(losses) / shards) <= 1e-4 * max(1.0, abs(float(lf["loss"])))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_part_backward_equals_the_whole_backward(cf, dev, precision):
    """cfn_network_bwd_part (part 1 down to the weight gradient of trunk layer D/2, part 2 the rest) is what lets the
    data-parallel trainer all-reduce the finished half of its gradient bucket under the rest of the backward: the two
    calls must write exactly what cfn_network_bwd writes (deterministic mode: bit for bit)."""
    from cfnerf_b200 import dist as D
    cfg = O.CfnConfig()
    p = O.make_params(cfg, 5, "lively")
    sa, sr = O.make_latents(cfg, 5)
    B = 96
    rays = O.synthetic_rays(B, 3).to(dev)
    g = torch.Generator().manual_seed(11)
    target = torch.rand(B, 3, generator=g).to(dev)
    t_rand = torch.rand(B, 128, generator=g).to(dev)
    ea, er = torch.randn(cfg.K, 1, generator=g).to(dev), torch.randn(cfg.K, 3, generator=g).to(dev)
    grads = []
    for two in (False, True):
        net = make_net(cf, cfg, p, sa, sr, dev)
        tr = D.FusedTrainStep(net, lr=0.0, precision=precision, deterministic=True)
        tr.two_part_backward = two
        tr.step(rays, target, t_rand=t_rand, eps_alpha=ea, eps_rgb=er)
        grads.append(tr.flat_grad.clone())
        assert 0 < tr._split_off < tr.flat_grad.numel()
    assert float(grads[0].abs().max()) > 0
    assert torch.equal(grads[0], grads[1])


@pytest.mark.parametrize("precision,K", [("fp32", 32), ("bf16", 32), ("fp32", 40)])
def test_small_batch_training_forward_equals_the_one_warp_per_ray_kernel(cf, dev, precision, K):
    """Batches of at most 1536 rays take the four-warps-per-ray training forward (each warp walks a quarter of the samples
    from a local transmittance of 1; the ranges are stitched afterwards).  It must return what the sequential kernel
    returns — maps, log-det sums, the transmittances and the range sums the backward reads — up to the re-association of
    the transmittance product.  A 1600-ray batch takes the sequential kernel; its first 700 rays alone take the new one."""
    cfg = O.CfnConfig(K=K)
    p = O.make_params(cfg, 7, "lively")
    sa, sr = O.make_latents(cfg, 7)
    net = make_net(cf, cfg, p, sa, sr, dev)
    eng = cf.engine_for(net, dev, precision)
    B, N, Bs = 1600, 128, 700
    rays = O.synthetic_rays(B, 9).to(dev)
    z = eng.zvals(rays, cf.reference_t_schedule(N, dev), torch.rand(B, N, generator=torch.Generator().manual_seed(3)).to(dev), False)
    g = torch.Generator().manual_seed(4)
    ea, er = torch.randn(K, generator=g).to(dev), torch.randn(K, 3, generator=g).to(dev)
    fp = eng.network(B, N, rays=rays, z_vals=z)
    big = eng.flow_composite(fp, z, rays[:, 3:6], 11, ea, er, True, want_raw=True, train=True, want_trans=True)
    small = eng.flow_composite(fp[:Bs * N], z[:Bs], rays[:Bs, 3:6], 11, ea, er, True, want_raw=True, train=True, want_trans=True)
    assert big["seg_sums"].shape[1] == 4 and small["seg_sums"].shape[1] == 4
    assert torch.equal(small["raw"], big["raw"][:Bs])                      # per-sample arithmetic is identical
    for k in ("rgb_map", "depth_map", "disp_map", "trans", "seg_sums", "logdet_sums"):
        a, b = small[k].double(), big[k][:Bs].double()
        err = ((a - b).abs() / (b.abs() + 1e-3)).max().item()
        assert err <= 2e-5, (k, err)       # 128-term sums re-associated as 4 x 32 (+ the white-background 1 - acc)
