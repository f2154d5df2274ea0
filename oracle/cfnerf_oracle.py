"""TEST INFRASTRUCTURE ONLY — CPU restatement of CF-NeRF's per-ray K-sample render/train hot path.

This file is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The shipped path
(``cfnerf_b200``) never routes through it and fails loudly when its CUDA library is missing.

Parity pin: every function below that has a counterpart in the reference is checked, in this
container, against the unmodified reference executed live through ``oracle/refload.py``
(``tests/test_oracle_vs_reference.py``) and against the committed fixtures that
``oracle/make_golden.py`` produced from that reference (``tests/golden/*.npz``).  The reference has
no tests, golden vectors or known-answer data of its own.

``sample_pdf`` / the coarse+fine pass are **parity unpinned by the reference**: they do not exist in
it (only the comment at run_nerf_helpers.py:9-11 survives).  Their semantics are those of the
upstream ancestor the reference names (README.md:93, yenchenlin/nerf-pytorch, un-vendored, no pinned
commit), restated here with an explicit, sequential fp32 operation order so that a GPU kernel can
be bit-exact against it.

All citations are relative to /root/reference: main = run_nerf_uncertainty_NF.py,
helpers = run_nerf_helpers.py, models = model/models.py, flows = model/flow/flows.py.

Everything is plain torch on CPU (the reference is plain torch), dtype-parametric so that an fp64
run can referee fp32 disagreements.  ``faithful=True`` keeps the reference's K-fold materialisation
of the flow conditioning (models:210-217, 255-257, 271-273) — that is the variant timed as the CPU
baseline ("port").  ``faithful=False`` amortises once per point (same numbers).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# configuration + deterministic parameters
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class CfnConfig:
    """Architecture of one NeRF_Flows network (models:20-36; canonical = africa.txt + train_NF.sh)."""

    D: int = 8          # --netdepth
    W: int = 512        # --netwidth
    L_pos: int = 10     # --multires
    L_dir: int = 4      # --multires_views
    h_alpha: int = 64   # --h_alpha_size
    h_rgb: int = 64     # --h_rgb_size
    F: int = 4          # --n_flows
    K: int = 32         # --K_samples

    @property
    def in_pos(self) -> int:
        return 3 + 6 * self.L_pos

    @property
    def in_dir(self) -> int:
        return 3 + 6 * self.L_dir

    @property
    def skip(self) -> int:
        # skips=[netdepth/2] with TRUE division (main:327): for an odd depth the list holds x.5 and `i in skips`
        # (models:39, 171) never matches -> no skip connection at all
        return self.D // 2 if self.D % 2 == 0 else -1

    def n_params(self) -> int:
        W, ip, idr = self.W, self.in_pos, self.in_dir
        n = ip * W + W + (self.D - 2) * (W * W + W) + ((W + (ip if self.skip >= 0 else 0)) * W + W)
        n += (W + idr) * (W // 2) + W // 2 + W * W + W + 2 * (W + 1)
        n += W * self.h_alpha + self.h_alpha + (W // 2) * self.h_rgb + self.h_rgb + 8
        for z, h in ((1, self.h_alpha), (3, self.h_rgb)):
            n += (h + 1) * self.F * (z * z + 3 * z)
        return n


def _linear_init(gen, out_f, in_f):
    # nn.Linear default (kaiming_uniform a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in)) for weight and bias
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound
    return w, b


def make_params(cfg: CfnConfig, seed: int = 0, variant: str = "default") -> dict:
    """Deterministic fp32 parameters keyed like ``NeRF_Flows.state_dict()`` (models:38-67, 339-350).

    variant "default": nn.Linear-style init, globals mean 0 / std 1 (models:44-48).
    variant "lively" : same, but non-trivial globals and x3 conditioning so the flows leave identity
                       (used by parity tests so that every term is exercised).
    variant "stressed": SURVEY §8(d) "stressed heads": amortisation weights x8, head weights x4,
                       biases ~ N(0,1) — exposes GEMM-precision loss.
    """
    g = torch.Generator().manual_seed(seed)
    W, D = cfg.W, cfg.D
    p = {}
    p["alpha_mean"] = torch.zeros(1)
    p["alpha_std"] = torch.ones(1)
    p["rgb_mean"] = torch.zeros(3)
    p["rgb_std"] = torch.ones(3)
    for i in range(D):
        if i == 0:
            fin = cfg.in_pos
        elif i == cfg.skip + 1:
            fin = W + cfg.in_pos
        else:
            fin = W
        p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"] = _linear_init(g, W, fin)
    p["views_linears.0.weight"], p["views_linears.0.bias"] = _linear_init(g, W // 2, W + cfg.in_dir)
    p["feature_linear.weight"], p["feature_linear.bias"] = _linear_init(g, W, W)
    p["alpha_linear.weight"], p["alpha_linear.bias"] = _linear_init(g, 1, W)           # dead (models:59)
    p["alpha_std_linear.weight"], p["alpha_std_linear.bias"] = _linear_init(g, 1, W)   # dead (models:60)
    p["h_alpha_linear.weight"], p["h_alpha_linear.bias"] = _linear_init(g, cfg.h_alpha, W)
    p["h_rgb_linear.weight"], p["h_rgb_linear.bias"] = _linear_init(g, cfg.h_rgb, W // 2)
    for name, z, h in (("flows_rgb", 3, cfg.h_rgb), ("flows_alpha", 1, cfg.h_alpha)):
        p[f"{name}.amor_d.weight"], p[f"{name}.amor_d.bias"] = _linear_init(g, cfg.F * z * z, h)
        p[f"{name}.amor_diag1.0.weight"], p[f"{name}.amor_diag1.0.bias"] = _linear_init(g, cfg.F * z, h)
        p[f"{name}.amor_diag2.0.weight"], p[f"{name}.amor_diag2.0.bias"] = _linear_init(g, cfg.F * z, h)
        p[f"{name}.amor_b.weight"], p[f"{name}.amor_b.bias"] = _linear_init(g, cfg.F * z, h)
    if variant in ("lively", "stressed"):
        p["alpha_mean"] = torch.tensor([0.35])
        p["alpha_std"] = torch.tensor([0.8])
        p["rgb_mean"] = torch.tensor([0.2, -0.1, 0.3])
        p["rgb_std"] = torch.tensor([0.9, 1.1, 0.7])
        amp_amor, amp_head = (3.0, 2.0) if variant == "lively" else (8.0, 4.0)
        for name in ("flows_rgb", "flows_alpha"):
            for sub in ("amor_d", "amor_diag1.0", "amor_diag2.0", "amor_b"):
                p[f"{name}.{sub}.weight"] = p[f"{name}.{sub}.weight"] * amp_amor
                p[f"{name}.{sub}.bias"] = torch.randn(p[f"{name}.{sub}.bias"].shape, generator=g) * (
                    1.0 if variant == "stressed" else 0.3)
        for head in ("h_alpha_linear", "h_rgb_linear"):
            p[f"{head}.weight"] = p[f"{head}.weight"] * amp_head
            if variant == "stressed":
                p[f"{head}.bias"] = torch.randn(p[f"{head}.bias"].shape, generator=g)
    elif variant != "default":
        raise ValueError(variant)
    return {k: v.contiguous() for k, v in p.items()}


def make_latents(cfg: CfnConfig, seed: int = 0):
    """K fixed base draws, (K,1) then (K,3) — stands in for models:53-55 (not in state_dict there)."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randn(cfg.K, 1, generator=g), torch.randn(cfg.K, 3, generator=g)


def params_checksum(params: dict) -> float:
    """Order-independent fp64 fingerprint, stored in golden files to detect RNG drift."""
    acc = 0.0
    for i, k in enumerate(sorted(params)):
        v = params[k].double()
        acc += float((v * torch.cos(torch.arange(v.numel(), dtype=torch.float64).reshape(v.shape) * 0.37 + i)).sum())
    return acc


def _cast(params: dict, dtype):
    return {k: v.to(dtype) for k, v in params.items()}


# --------------------------------------------------------------------------------------------
# A3 positional encoding (helpers:21-69)
# --------------------------------------------------------------------------------------------
def positional_encoding(x: torch.Tensor, L: int) -> torch.Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)], blocks of 3 (helpers:29-51)."""
    out = [x]
    freqs = 2.0 ** torch.linspace(0.0, L - 1, steps=L)  # helpers:38 (fp32 linspace, exact powers of two)
    for f in freqs:
        xf = x * f.to(x.dtype)
        out.append(torch.sin(xf))
        out.append(torch.cos(xf))
    return torch.cat(out, -1)


# --------------------------------------------------------------------------------------------
# A4 trunk + heads (models:165-186)
# --------------------------------------------------------------------------------------------
def mlp_encode(p: dict, cfg: CfnConfig, embedded: torch.Tensor):
    """(M, in_pos+in_dir) -> h_alpha (M,h_alpha), h_rgb (M,h_rgb)."""
    g_pos, g_dir = embedded[:, : cfg.in_pos], embedded[:, cfg.in_pos:]
    h = g_pos
    for i in range(cfg.D):
        h = F.relu(F.linear(h, p[f"pts_linears.{i}.weight"], p[f"pts_linears.{i}.bias"]))
        if i == cfg.skip:
            h = torch.cat([g_pos, h], -1)  # gamma(p) FIRST (models:171-172)
    h_alpha = F.linear(h, p["h_alpha_linear.weight"], p["h_alpha_linear.bias"])          # models:175
    feat = F.linear(h, p["feature_linear.weight"], p["feature_linear.bias"])             # models:176
    v = F.relu(F.linear(torch.cat([feat, g_dir], -1), p["views_linears.0.weight"],
                        p["views_linears.0.bias"]))                                      # models:177-181
    h_rgb = F.linear(v, p["h_rgb_linear.weight"], p["h_rgb_linear.bias"])                # models:182
    return h_alpha, h_rgb


# --------------------------------------------------------------------------------------------
# A5 amortised flow parameters (models:358-385)
# --------------------------------------------------------------------------------------------
def flow_conditioning(p: dict, name: str, h: torch.Tensor, z: int, nF: int):
    """h (R,hdim) -> r1, r2 (R,z,z,F) upper-triangular with tanh'd diagonals, b (R,1,z,F)."""
    R = h.shape[0]
    full_d = F.linear(h, p[f"{name}.amor_d.weight"], p[f"{name}.amor_d.bias"]).reshape(R, z, z, nF)
    d1 = torch.tanh(F.linear(h, p[f"{name}.amor_diag1.0.weight"], p[f"{name}.amor_diag1.0.bias"])).reshape(R, z, nF)
    d2 = torch.tanh(F.linear(h, p[f"{name}.amor_diag2.0.weight"], p[f"{name}.amor_diag2.0.bias"])).reshape(R, z, nF)
    b = F.linear(h, p[f"{name}.amor_b.weight"], p[f"{name}.amor_b.bias"]).reshape(R, 1, z, nF)
    mask = torch.triu(torch.ones(z, z, dtype=h.dtype), diagonal=1)[None, :, :, None]     # models:327-328
    r1 = full_d * mask                                                                    # models:374
    r2 = full_d.transpose(2, 1) * mask                                                    # models:375
    idx = torch.arange(z)
    r1[:, idx, idx, :] = d1                                                               # models:377
    r2[:, idx, idx, :] = d2                                                               # models:378
    return r1, r2, b


# --------------------------------------------------------------------------------------------
# A6 triangular Sylvester flow stack (models:387-416, flows:189-268)
# --------------------------------------------------------------------------------------------
def flow_stack(z0: torch.Tensor, r1, r2, b, want_logdet: bool):
    """z0 (R,z) -> z_F (R,z), sum of per-flow log|det J| (R,) (0 when not wanted, flows:223)."""
    R, zdim = z0.shape
    nF = r1.shape[-1]
    z = z0
    logdet = torch.zeros(R, dtype=z0.dtype) if want_logdet else 0
    flip = torch.arange(zdim - 1, -1, -1)                                                 # models:323
    diag = torch.arange(zdim)
    for f in range(nF):
        R1, R2, bf = r1[..., f], r2[..., f], b[..., f]
        zk = z.unsqueeze(1)                                                               # (R,1,z)
        zp = zk[:, :, flip] if f % 2 == 1 else zk                                         # models:404-408
        pre = torch.bmm(zp, R2.transpose(2, 1)) + bf                                      # flows:213/238
        t = torch.tanh(pre)
        step = torch.bmm(t, R1.transpose(2, 1))                                           # flows:214/239
        if f % 2 == 1:
            step = step[:, :, flip]
        z = (step + zk).squeeze(1)
        if want_logdet:
            dj = R1[:, diag, diag] * R2[:, diag, diag]                                    # flows:229-230, 251
            dj = (1 - t.squeeze(1) ** 2) * dj + 1.0                                       # flows:252-253
            logdet = logdet + torch.log(dj.abs() + 1e-8).sum(-1)                          # flows:255-262
    return z, logdet


# --------------------------------------------------------------------------------------------
# A7 NeRF_Flows.forward (models:188-291)
# --------------------------------------------------------------------------------------------
def test_latents(sample_alpha, sample_rgb):
    """Test-mode base noise: constructor-time draws with the LAST sample zeroed (models:198-205)."""
    ea, er = sample_alpha.clone(), sample_rgb.clone()
    ea[-1] = 0
    er[-1] = 0
    return ea, er


def nerf_flows_forward(p: dict, cfg: CfnConfig, embedded: torch.Tensor, eps_alpha: torch.Tensor,
                       eps_rgb: torch.Tensor, train: bool, faithful: bool = True):
    """embedded (M,90), eps_alpha (K,1), eps_rgb (K,3) -> raw (M,K,4) [rgb|sigma], entropy scalar.

    train=False: caller passes ``test_latents(...)``; second value is 0 (the reference returns
    zeros_like(raw), models:223).  train=True: eps are the fresh per-call draws (models:234, 246).
    """
    M, K = embedded.shape[0], cfg.K
    h_alpha, h_rgb = mlp_encode(p, cfg, embedded)
    a_mean, a_std = p["alpha_mean"], p["alpha_std"]
    c_mean, c_std = p["rgb_mean"], p["rgb_std"]
    a0 = (eps_alpha[None] * a_std[None, None, :] + a_mean[None, None, :]).expand(M, K, 1)   # models:200/239
    c0 = (eps_rgb[None] * c_std[None, None, :] + c_mean[None, None, :]).expand(M, K, 3)     # models:206/251
    if faithful:
        ha = h_alpha[:, None, :].expand(M, K, cfg.h_alpha).reshape(M * K, cfg.h_alpha)       # models:210-211
        hr = h_rgb[:, None, :].expand(M, K, cfg.h_rgb).reshape(M * K, cfg.h_rgb)             # models:215-216
        r1a, r2a, ba = flow_conditioning(p, "flows_alpha", ha, 1, cfg.F)
        r1c, r2c, bc = flow_conditioning(p, "flows_rgb", hr, 3, cfg.F)
    else:
        def rep(t):
            return t[:, None].expand(M, K, *t.shape[1:]).reshape(M * K, *t.shape[1:])
        r1a, r2a, ba = (rep(t) for t in flow_conditioning(p, "flows_alpha", h_alpha, 1, cfg.F))
        r1c, r2c, bc = (rep(t) for t in flow_conditioning(p, "flows_rgb", h_rgb, 3, cfg.F))
    za, ld_a = flow_stack(a0.reshape(M * K, 1), r1a, r2a, ba, train)
    zc, ld_c = flow_stack(c0.reshape(M * K, 3), r1c, r2c, bc, train)
    za = za.reshape(M, K, 1)
    zc = zc.reshape(M, K, 3)
    raw = torch.cat([zc, za], -1)                                                            # models:221/289
    if not train:
        return raw, raw.new_zeros(())
    ld_a = ld_a.reshape(M, K) + (za.sum(-1) - F.softplus(za).sum(-1))                        # models:263
    ld_c = ld_c.reshape(M, K) + (zc.sum(-1) - 2 * F.softplus(zc).sum(-1))                    # models:278
    base_a = -0.5 * (a_std.log() * 2 + (a0 - a_mean) * (a0 - a_mean) * (a_std ** 2).reciprocal())   # models:268
    base_c = -0.5 * (c_std.log() * 2 + (c0 - c_mean) * (c0 - c_mean) * (c_std ** 2).reciprocal())   # models:283
    entropy = base_a.mean() - ld_a.mean() + base_c.mean() - ld_c.mean()                      # models:286
    return raw, entropy


# --------------------------------------------------------------------------------------------
# A2 run_network (main:67-85), with the netchunk loop of batchify (main:47-64)
# --------------------------------------------------------------------------------------------
def run_network(p: dict, cfg: CfnConfig, pts: torch.Tensor, viewdirs: torch.Tensor, eps_alpha, eps_rgb,
                train: bool, netchunk: int = 1024 * 64, faithful: bool = True):
    """pts (B,N,3), viewdirs (B,3) -> raw (B,N,K,4), per-chunk entropies [(scalar, n_points)]."""
    B, N = pts.shape[:2]
    flat = pts.reshape(-1, 3)
    emb = positional_encoding(flat, cfg.L_pos)
    dirs = viewdirs[:, None].expand(B, N, 3).reshape(-1, 3)                                  # main:74-78
    emb = torch.cat([emb, positional_encoding(dirs, cfg.L_dir)], -1)                         # main:79-80
    raws, ents = [], []
    for ci, i in enumerate(range(0, emb.shape[0], netchunk)):
        # the reference draws fresh noise inside EVERY network call in train mode (models:233-251): a leading axis on
        # the eps arguments, (G,K,1) / (G,K,3), gives call ci its own draw; plain (K,1) / (K,3) are shared by all calls
        ea = eps_alpha[ci] if eps_alpha.dim() == 3 else eps_alpha
        er = eps_rgb[ci] if eps_rgb.dim() == 3 else eps_rgb
        r, e = nerf_flows_forward(p, cfg, emb[i:i + netchunk], ea, er, train, faithful)
        raws.append(r)
        ents.append((e, min(netchunk, emb.shape[0] - i)))
    return torch.cat(raws, 0).reshape(B, N, cfg.K, 4), ents


def entropy_mean(ents) -> torch.Tensor:
    """The trainer's ``extras['loss_entropy'].mean()`` (main:1045): point-weighted mean of per-call scalars."""
    tot = sum(n for _, n in ents)
    return sum(e * (n / tot) for e, n in ents)


# --------------------------------------------------------------------------------------------
# A8 raw2outputs (main:411-454)
# --------------------------------------------------------------------------------------------
def raw2outputs(raw: torch.Tensor, z_vals: torch.Tensor, rays_d: torch.Tensor, white_bkgd: bool = False):
    """raw (B,N,K,4), z_vals (B,N), rays_d (B,3) -> rgb_map (B,3,K), disp (B,K), weights (B,N,K), depth (B,K)."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.full_like(dists[..., :1], 1e1)], -1)                     # last = 10.0 (main:427)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)                                 # main:429
    rgb = torch.sigmoid(raw[..., :3])                                                        # main:431
    alpha = 1.0 - torch.exp(-F.softplus(raw[..., 3]) * dists[..., None])                     # main:424, 442 (noise never added)
    ones = torch.ones((alpha.shape[0], 1, alpha.shape[-1]), dtype=raw.dtype)
    trans = torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-10], -2), -2)[:, :-1, :]         # main:443
    weights = alpha * trans
    rgb_map = torch.sum(weights[..., None] * rgb, -3).transpose(-1, -2)                      # (B,3,K) main:444-445
    depth_map = torch.sum(weights * z_vals[..., None], -2)                                   # main:447
    acc_map = torch.sum(weights, -2)                                                         # main:449
    disp_map = 1.0 / torch.max(1e-10 * torch.ones_like(depth_map) + 1e-10,
                               depth_map / (acc_map + 1e-10) + 1e-10)                        # main:448
    if white_bkgd:
        rgb_map = rgb_map + (1.0 - acc_map[:, None, :])                                      # main:451-452
    return rgb_map, disp_map, weights, depth_map


# --------------------------------------------------------------------------------------------
# A1 render_rays (main:457-553)
# --------------------------------------------------------------------------------------------
def reference_t_schedule(dtype=torch.float32) -> torch.Tensor:
    """The hard-coded 96+32 sample schedule on [0,1] (main:510)."""
    return torch.cat([torch.linspace(0.0, 0.5, steps=97)[:-1], torch.linspace(0.5, 1.0, steps=32)], 0).to(dtype)


def z_from_t(t_vals, near, far, lindisp: bool, t_rand=None):
    """near/far (B,1); t_vals (N,) -> z_vals (B,N), optionally stratified-jittered (main:511-532)."""
    if not lindisp:
        z = near * (1.0 - t_vals) + far * t_vals
    else:
        z = 1.0 / (1.0 / near * (1.0 - t_vals) + 1.0 / far * t_vals)
    z = z.expand(near.shape[0], t_vals.shape[0])
    if t_rand is not None:
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    return z


def render_rays(p: dict, cfg: CfnConfig, ray_batch: torch.Tensor, eps_alpha, eps_rgb, train: bool,
                t_rand=None, lindisp: bool = False, white_bkgd: bool = False, faithful: bool = True,
                netchunk: int = 1024 * 64):
    """ray_batch (B,11)=[o d near far viewdir] -> dict like main:542-547 (loss_entropy as the scalar; the per-call
    scalars and their point counts, i.e. the rows of the reference's (B*N,K,1) tensor, under "entropy_calls")."""
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:]
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    z_vals = z_from_t(reference_t_schedule(ray_batch.dtype), near, far, lindisp, t_rand)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]                 # main:534
    raw, ents = run_network(p, cfg, pts, viewdirs, eps_alpha, eps_rgb, train, netchunk=netchunk, faithful=faithful)
    rgb_map, disp_map, weights, depth_map = raw2outputs(raw, z_vals, rays_d, white_bkgd)
    ret = {"rgb_map": rgb_map, "disp_map": disp_map, "depth_map": depth_map, "weights": weights,
           "z_vals": z_vals}
    if train:
        ret["raw"] = raw
        ret["loss_entropy"] = entropy_mean(ents)
        ret["entropy_calls"] = ents
        ret["pts"] = pts
    return ret


# --------------------------------------------------------------------------------------------
# A9 sample_pdf — EXTENSION, parity unpinned by the reference (see module docstring)
# --------------------------------------------------------------------------------------------
def sample_pdf(bins, weights, u):
    """Inverse-CDF resampling with an explicit sequential fp32 operation order.

    bins (B,M) fp32 = midpoints of the coarse z; weights (B,M-1) fp32; u (B,Nf) fp32 in [0,1].
    Returns samples (B,Nf) fp32 and the integer bracket ``below`` (B,Nf) int32.

    Order of operations (what the CUDA kernel reproduces bit-for-bit):
      w = weights + 1e-5f ; total = sequential left-to-right fp32 sum of w ; pdf = w / total ;
      cdf[0] = 0, cdf[j] = cdf[j-1] + pdf[j-1] (sequential fp32) ;
      i = #{j : cdf[j] <= u}  (searchsorted right=True) ; below = max(i-1,0) ; above = min(i, M-1) ;
      denom = cdf[above]-cdf[below], replaced by 1 where < 1e-5 ; t = (u-cdf[below])/denom ;
      sample = bins[below] + t*(bins[above]-bins[below])   (separate mul and add, no FMA).
    numpy float32 arithmetic is IEEE and never contracts, so this *is* that order.
    """
    bins = np.ascontiguousarray(np.asarray(bins, dtype=np.float32))
    w = np.asarray(weights, dtype=np.float32) + np.float32(1e-5)
    u = np.ascontiguousarray(np.asarray(u, dtype=np.float32))
    total = np.cumsum(w, axis=-1, dtype=np.float32)[..., -1:]          # np.cumsum is strictly sequential
    pdf = (w / total).astype(np.float32)
    cdf = np.concatenate([np.zeros_like(pdf[..., :1]), np.cumsum(pdf, axis=-1, dtype=np.float32)], -1)
    M = cdf.shape[-1]
    assert bins.shape[-1] == M
    inds = (cdf[:, None, :] <= u[:, :, None]).sum(-1).astype(np.int64)  # right=True
    below = np.maximum(inds - 1, 0)
    above = np.minimum(inds, M - 1)
    cdf_b = np.take_along_axis(cdf, below, -1)
    cdf_a = np.take_along_axis(cdf, above, -1)
    bin_b = np.take_along_axis(bins, below, -1)
    bin_a = np.take_along_axis(bins, above, -1)
    denom = (cdf_a - cdf_b).astype(np.float32)
    denom = np.where(denom < np.float32(1e-5), np.float32(1.0), denom)
    t = ((u - cdf_b) / denom).astype(np.float32)
    samples = (bin_b + (t * (bin_a - bin_b)).astype(np.float32)).astype(np.float32)
    return samples, below.astype(np.int32)


def sample_pdf_upstream_torch(bins: torch.Tensor, weights: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """The same resampling written with torch library calls (cumsum / searchsorted / gather), as the
    upstream ancestor does.  Only used to show that the sequential-order oracle above agrees with it
    to rounding (torch's CPU reductions associate differently, so it cannot be the bit-exact pin)."""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    inds = torch.searchsorted(cdf.contiguous(), u.contiguous(), right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, -1, below), torch.gather(cdf, -1, above)
    bin_b, bin_a = torch.gather(bins, -1, below), torch.gather(bins, -1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return bin_b + (u - cdf_b) / denom * (bin_a - bin_b)


def merge_sorted(z_coarse, z_fine):
    """sort(cat[z_coarse, z_fine]) along the last axis (values only; fp32 compare is exact)."""
    return np.sort(np.concatenate([np.asarray(z_coarse, np.float32), np.asarray(z_fine, np.float32)], -1), -1)


# --------------------------------------------------------------------------------------------
# A10 coarse + fine pass — EXTENSION, parity unpinned by the reference
# --------------------------------------------------------------------------------------------
def coarse_t_schedule(n_coarse: int, dtype=torch.float32) -> torch.Tensor:
    """Upstream linspace(0,1,Nc); the reference's 96+32 schedule bit-for-bit when Nc == 128."""
    if n_coarse == 128:
        return reference_t_schedule(dtype)
    return torch.linspace(0.0, 1.0, steps=n_coarse).to(dtype)


def render_rays_hier(p_coarse: dict, p_fine: dict, cfg: CfnConfig, ray_batch, eps_alpha, eps_rgb, train: bool,
                     n_coarse: int, n_fine: int, t_rand=None, u=None, lindisp=False, white_bkgd=False,
                     faithful: bool = False):
    """Coarse pass -> K-mean weights -> sample_pdf -> fine pass over the merged grid.

    CF-NeRF-specific decision: the coarse weights are (B,Nc,K); all K fields share ONE
    fine grid built from their mean over K, so the MLP is still evaluated once per point.
    ``u`` (B,Nf) are the resampling uniforms; None = deterministic linspace(0,1,Nf) (perturb == 0).
    """
    dtype = ray_batch.dtype
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:]
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    B = ray_batch.shape[0]
    z_c = z_from_t(coarse_t_schedule(n_coarse, dtype), near, far, lindisp, t_rand)
    pts_c = rays_o[..., None, :] + rays_d[..., None, :] * z_c[..., :, None]
    raw_c, ents_c = run_network(p_coarse, cfg, pts_c, viewdirs, eps_alpha, eps_rgb, train, faithful=faithful)
    rgb0, disp0, w_c, depth0 = raw2outputs(raw_c, z_c, rays_d, white_bkgd)
    w_mean = w_c.mean(-1)                                                 # (B,Nc)
    z_mid = 0.5 * (z_c[..., 1:] + z_c[..., :-1])
    if u is None:
        u = torch.linspace(0.0, 1.0, steps=n_fine).expand(B, n_fine)
    z_s, _ = sample_pdf(z_mid.detach().float().numpy(), w_mean[..., 1:-1].detach().float().numpy(),
                        u.float().numpy())
    z_all = torch.from_numpy(merge_sorted(z_c.detach().float().numpy(), z_s)).to(dtype)
    pts_f = rays_o[..., None, :] + rays_d[..., None, :] * z_all[..., :, None]
    raw_f, ents_f = run_network(p_fine, cfg, pts_f, viewdirs, eps_alpha, eps_rgb, train, faithful=faithful)
    rgb_map, disp_map, w_f, depth_map = raw2outputs(raw_f, z_all, rays_d, white_bkgd)
    ret = {"rgb_map": rgb_map, "disp_map": disp_map, "depth_map": depth_map, "rgb0": rgb0, "disp0": disp0,
           "depth0": depth0, "z_vals": z_all, "z_samples": torch.from_numpy(z_s), "weights0": w_c}
    if train:
        ret["loss_entropy"] = entropy_mean(ents_f)
        ret["loss_entropy0"] = entropy_mean(ents_c)
    return ret


# --------------------------------------------------------------------------------------------
# A11 K-reduction + KDE-NLL loss of the caller (main:1027-1050), TB variant (main:1122-1131)
# --------------------------------------------------------------------------------------------
def k_reduce(rgb_map: torch.Tensor, depth_map: torch.Tensor, K: int):
    """Predictive mean, 'uncertainty' std (unbiased std x K/(K-1), main:1034/1130) and mean depth."""
    rgb_mean = rgb_map.mean(-1)
    rgb_std = torch.std(rgb_map, -1) * K / (K - 1)
    return rgb_mean, rgb_std, depth_map.mean(-1)


def kde_nll_loss(rgb_map: torch.Tensor, target: torch.Tensor, loss_entropy: torch.Tensor, K: int,
                 beta1: float = 0.01):
    """rgb_map (B,3,K), target (B,3) -> dict(loss, loss_nll, mse, psnr) exactly as main:1027-1048."""
    eps = 1e-05
    rgb_mean = rgb_map.mean(-1)
    mse = torch.mean((rgb_mean - target) ** 2)                                               # helpers:15
    psnr = -10.0 * torch.log(mse) / math.log(10.0)                                           # helpers:16
    rgb_std = torch.std(rgb_map, -1) * K / (K - 1)                                           # main:1034
    h = rgb_std.detach() * (0.8 / K) ** (-1.0 / 7.0) + eps                                   # main:1036
    h = h[..., None]
    p1 = torch.exp(-((rgb_map - target[..., None]) ** 2) / (2 * h * h))                      # main:1038
    p2 = (2 * math.pi) ** (-1.5) / h                                                         # main:1039
    nll = -torch.log((p1 * p2).mean(-1) + eps).mean()                                        # main:1040-1042
    loss = nll + beta1 * loss_entropy if beta1 else nll                                      # main:1047-1050
    return {"loss": loss, "loss_nll": nll, "mse": mse, "psnr": psnr}


def trainer_loss(out: dict, target_s: torch.Tensor, K: int, beta1: float = 0.01, target_depth=None,
                 depth_lambda: float = 0.0):
    """The loss of the trainer body, main:1018-1055, on `render_rays(cat[colour rays, depth rays], train=True)`.

    With --colmap_depth the batch is [N_batch colour rays | depth rays] (main:1009-1011): `depth = mean_K depth_map`,
    colours keep the first N_batch rays (main:1020-1022), and `extras[x][:N_batch]` (main:1023) slices the first
    N_batch ROWS of the (B*N,K,1) entropy tensor — rows are points, so with N_batch <= netchunk that is the FIRST network
    call's scalar only.  `depth_loss = img2mse(depth_col, target_depth)` enters with weight depth_lambda (main:1053-1054).
    """
    B = out["rgb_map"].shape[0]
    B_depth = 0 if target_depth is None else int(target_depth.shape[0])
    N_batch = B - B_depth
    ents = out["entropy_calls"]
    if B_depth:
        assert N_batch <= ents[0][1], "the first N_batch rows must lie inside the first network call"
        ent = ents[0][0]
    else:
        ent = entropy_mean(ents)                                                             # main:1045
    res = kde_nll_loss(out["rgb_map"][:N_batch], target_s, ent, K, beta1)
    res["loss_entropy"] = ent
    if B_depth:
        depth_col = out["depth_map"].mean(-1)[N_batch:]                                      # main:1020, 1022
        res["depth_loss"] = torch.mean((depth_col - target_depth) ** 2)                      # main:1053
        res["loss"] = res["loss"] + depth_lambda * res["depth_loss"]                         # main:1054
    return res


# --------------------------------------------------------------------------------------------
# ray generation (helpers:288-297, 360-377) — adjacent to the path, used to build the bench configs
# --------------------------------------------------------------------------------------------
def get_rays(H: int, W: int, focal: float, c2w: torch.Tensor):
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing="ij")
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - W * 0.5) / focal, -(j - H * 0.5) / focal, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    o0 = -1.0 / (W / (2.0 * focal)) * rays_o[..., 0] / rays_o[..., 2]
    o1 = -1.0 / (H / (2.0 * focal)) * rays_o[..., 1] / rays_o[..., 2]
    o2 = 1.0 + 2.0 * near / rays_o[..., 2]
    d0 = -1.0 / (W / (2.0 * focal)) * (rays_d[..., 0] / rays_d[..., 2] - rays_o[..., 0] / rays_o[..., 2])
    d1 = -1.0 / (H / (2.0 * focal)) * (rays_d[..., 1] / rays_d[..., 2] - rays_o[..., 1] / rays_o[..., 2])
    d2 = -2.0 * near / rays_o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def pack_ray_batch(rays_o, rays_d, near: float, far: float):
    """What ``render`` assembles before ``batchify_rays`` (main:136-158): (B,11)."""
    rays_o = rays_o.reshape(-1, 3).float()
    rays_d = rays_d.reshape(-1, 3).float()
    viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    nf = torch.ones_like(rays_d[..., :1])
    return torch.cat([rays_o, rays_d, near * nf, far * nf, viewdirs], -1)


def synthetic_rays(n: int, seed: int = 1, near: float = 1.2, far: float = 8.0) -> torch.Tensor:
    """o ~ N(0,0.1^2), d = unit Gaussian direction with d_z <- -|d_z|."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g) * 0.1
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    d[:, 2] = -d[:, 2].abs()
    return pack_ray_batch(o, d, near, far)
