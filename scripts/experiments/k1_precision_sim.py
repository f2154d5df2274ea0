"""CPU simulation of K1's operand rounding on the 'stressed heads' weights: which GEMMs need more than
bf16 to keep predictive mean / std / depth inside 2e-3?  Test infrastructure (imports oracle/)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from oracle import cfnerf_oracle as O

torch.set_num_threads(8)

def rnd(x, mode):
    if mode == "bf16": return x.float().bfloat16().double()
    if mode == "fp16": return x.float().half().double()
    if mode == "hilo":   # two bf16 terms
        hi = x.float().bfloat16().float(); lo = (x.float() - hi).bfloat16().float(); return (hi + lo).double()
    if mode == "exact": return x.double()
    raise ValueError(mode)

def k1(p, cfg, pts, dirs, trunk="bf16", head_act="bf16", head_w="bf16", bias="exact"):
    P = {k: (rnd(v, bias) if (k.endswith(".bias") and (k.startswith("pts_linears") or k.startswith("feature") or k.startswith("views"))) else v.double()) for k, v in p.items()}
    gp = rnd(O.positional_encoding(pts.float(), cfg.L_pos), trunk)
    gd = rnd(O.positional_encoding(dirs.float(), cfg.L_dir), trunk)
    h = gp
    for i in range(cfg.D):
        w = rnd(p[f"pts_linears.{i}.weight"], trunk)
        acc = F.relu(F.linear(h, w, P[f"pts_linears.{i}.bias"]))
        last = i == cfg.D - 1
        h_full = acc
        h = rnd(acc, trunk)
        if i == cfg.skip: h = torch.cat([gp, h], -1)
    h7_head = rnd(h_full, head_act)
    # composed heads
    def comp(name, hl):
        ws = torch.cat([P[f"{name}.{s}.weight"] for s in ("amor_d", "amor_diag1.0", "amor_diag2.0", "amor_b")], 0)
        bs = torch.cat([P[f"{name}.{s}.bias"] for s in ("amor_d", "amor_diag1.0", "amor_diag2.0", "amor_b")], 0)
        Wc = ws @ P[f"{hl}.weight"]; bc = ws @ P[f"{hl}.bias"] + bs
        return Wc, bc
    WA, bA = comp("flows_alpha", "h_alpha_linear")
    WC, bC = comp("flows_rgb", "h_rgb_linear")
    outA = F.linear(h7_head, rnd(WA, head_w), bA)
    feat = rnd(F.linear(h, rnd(p["feature_linear.weight"], trunk), P["feature_linear.bias"]), trunk)
    vacc = F.relu(F.linear(torch.cat([feat, gd], -1), rnd(p["views_linears.0.weight"], trunk), P["views_linears.0.bias"]))
    v = rnd(vacc, head_act)
    outC = F.linear(v, rnd(WC, head_w), bC)
    return outA, outC

def finish(p, cfg, outA, outC, ea, er, z_vals, rays_d, M):
    # rebuild r1,r2,b from the composed outputs exactly as flow_conditioning does
    nF = cfg.F
    def split(out, z):
        n_d = nF * z * z; n1 = nF * z
        full_d = out[:, :n_d].reshape(M, z, z, nF)
        d1 = torch.tanh(out[:, n_d:n_d + n1]).reshape(M, z, nF)
        d2 = torch.tanh(out[:, n_d + n1:n_d + 2 * n1]).reshape(M, z, nF)
        b = out[:, n_d + 2 * n1:].reshape(M, 1, z, nF)
        mask = torch.triu(torch.ones(z, z, dtype=out.dtype), diagonal=1)[None, :, :, None]
        r1 = full_d * mask; r2 = full_d.transpose(2, 1) * mask
        idx = torch.arange(z); r1[:, idx, idx, :] = d1; r2[:, idx, idx, :] = d2
        return r1, r2, b
    K = cfg.K
    rep = lambda t: t[:, None].expand(M, K, *t.shape[1:]).reshape(M * K, *t.shape[1:])
    r1a, r2a, ba = (rep(t) for t in split(outA, 1))
    r1c, r2c, bc = (rep(t) for t in split(outC, 3))
    Pd = {k: v.double() for k, v in p.items()}
    a0 = (ea.double()[None] * Pd["alpha_std"][None, None] + Pd["alpha_mean"][None, None]).expand(M, K, 1)
    c0 = (er.double()[None] * Pd["rgb_std"][None, None] + Pd["rgb_mean"][None, None]).expand(M, K, 3)
    za, _ = O.flow_stack(a0.reshape(M * K, 1), r1a, r2a, ba, False)
    zc, _ = O.flow_stack(c0.reshape(M * K, 3), r1c, r2c, bc, False)
    raw = torch.cat([zc.reshape(M, K, 3), za.reshape(M, K, 1)], -1)
    B, N = z_vals.shape
    rgb, disp, w, depth = O.raw2outputs(raw.reshape(B, N, K, 4), z_vals.double(), rays_d.double())
    return rgb, depth

cfg = O.CfnConfig()
p = O.make_params(cfg, 0, "stressed")
sa, sr = O.make_latents(cfg, 0)
ea, er = O.test_latents(sa, sr)
rays = O.synthetic_rays(64, 1)
z = O.z_from_t(O.reference_t_schedule(), rays[:, 6:7], rays[:, 7:8], False)
pts = (rays[:, None, 0:3] + rays[:, None, 3:6] * z[..., None]).reshape(-1, 3)
dirs = rays[:, None, 8:11].expand(64, 128, 3).reshape(-1, 3)
M = pts.shape[0]
ref = None
for name, kw in [("exact", dict(trunk="exact", head_act="exact", head_w="exact")),
                 ("all bf16", dict()),
                 ("bf16 trunk, hilo head acts+weights", dict(head_act="hilo", head_w="hilo")),
                 ("bf16 trunk, hilo head weights only", dict(head_w="hilo")),
                 ("bf16 trunk, fp16 head acts+weights", dict(head_act="fp16", head_w="fp16")),
                 ("all fp16", dict(trunk="fp16", head_act="fp16", head_w="fp16")),
                 ("all fp16, fp16-rounded trunk biases", dict(trunk="fp16", head_act="fp16", head_w="fp16", bias="fp16")),
                 ("all bf16, bf16-rounded trunk biases", dict(bias="bf16"))]:
    a, c = k1(p, cfg, pts, dirs, **kw)
    rgb, depth = finish(p, cfg, a, c, ea, er, z, rays[:, 3:6], M)
    K = cfg.K
    mean = rgb.mean(-1); std = rgb.std(-1) * K / (K - 1); dm = depth.mean(-1)
    if ref is None: ref = (mean, std, dm, rgb, depth); continue
    print(f"{name:40s} mean {float((mean-ref[0]).abs().max()):.2e} std {float((std-ref[1]).abs().max()):.2e} "
          f"depth_mean {float((dm-ref[2]).abs().max()):.2e}  per-k rgb {float((rgb-ref[3]).abs().max()):.2e} per-k depth {float((depth-ref[4]).abs().max()):.2e}")
