"""Multi-GPU plumbing: one process per GPU under torchrun, rays sharded in contiguous blocks.

Rendering needs no collective (rays are independent, the latent draws are per-k, not per-ray): every rank renders
its block and the results are, bit for bit, the rows of the unsharded result.  Training averages gradients with ONE
all-reduce over a single flat fp32 bucket per step (the reference's nn.DataParallel instead re-broadcasts all
weights and scatters/gathers activations on every network call, run_nerf_uncertainty_NF.py:330).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int):
    """Contiguous block [lo, hi) of rank `rank`; the first n % world ranks get one extra ray."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(ray_batch: torch.Tensor, rank: int | None = None, world_size: int | None = None) -> torch.Tensor:
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(ray_batch.shape[0], rank, world_size)
    return ray_batch[lo:hi]


def gather_rows(t: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gather the per-rank row blocks of a sharded render back into the full (n_total, ...) tensor."""
    rank, w = world()
    if w == 1:
        return t
    sizes = [hi - lo for lo, hi in (shard_bounds(n_total, r, w) for r in range(w))]
    biggest = max(sizes)                      # all_gather wants equal shapes: pad the short (ragged) blocks
    mine = t.contiguous()
    if mine.shape[0] < biggest:
        mine = torch.cat([mine, mine.new_zeros((biggest - mine.shape[0],) + tuple(t.shape[1:]))], 0)
    parts = [torch.empty_like(mine) for _ in range(w)]
    dist.all_gather(parts, mine)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], 0)


class GradBucket:
    """One flat fp32 bucket over every parameter that receives a gradient (the two dead heads alpha_linear /
    alpha_std_linear, models.py:59-60, have grad None and are left out: 2 359 520 of 2 360 546 elements)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.flat = None

    def _live(self):
        return [p for p in self.params if p.grad is not None]

    def all_reduce_mean_(self):
        rank, w = world()
        live = self._live()
        if w == 1 or not live:
            return
        n = sum(p.grad.numel() for p in live)
        if self.flat is None or self.flat.numel() != n or self.flat.device != live[0].grad.device:
            self.flat = torch.empty(n, dtype=torch.float32, device=live[0].grad.device)
        o = 0
        for p in live:
            k = p.grad.numel()
            self.flat[o:o + k].copy_(p.grad.reshape(-1))
            o += k
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)      # NCCL over NVLink on GPU tensors, gloo on CPU tensors
        self.flat.mul_(1.0 / w)
        o = 0
        for p in live:
            k = p.grad.numel()
            p.grad.copy_(self.flat[o:o + k].view_as(p.grad))
            o += k


def train_step(network_fn, optimizer, ray_batch, target, bucket: GradBucket | None = None, beta1: float = 0.01,
               **render_kwargs):
    """One data-parallel optimisation step on this rank's shard (the trainer body of main:1014-1067 with the
    DataParallel wrapper replaced by one gradient all-reduce).  Equal shard sizes make the averaged gradient equal
    to the global-batch gradient of the mean-reduced loss."""
    from . import api

    out = api.render_rays(ray_batch, network_fn, None, 128, True, False, perturb=1., raw_noise_std=1., **render_kwargs)
    K = out["rgb_map"].shape[-1]
    losses = api.kde_nll_loss(out["rgb_map"], target, out["loss_entropy"], K, beta1)
    optimizer.zero_grad(set_to_none=True)
    losses["loss"].backward()
    if bucket is not None:
        bucket.all_reduce_mean_()
    optimizer.step()
    return {k: v.detach() for k, v in losses.items()}


class FusedTrainStep:
    """The trainer body (main:1014-1067: render in train mode -> K-mean + KDE-NLL + beta1 * entropy -> backward -> Adam)
    as a straight chain of C-ABI calls, without an autograd graph: cfn_zvals -> cfn_network_fwd(save) ->
    cfn_flow_composite_fwd -> cfn_kde_nll (loss + gradient seed) -> cfn_flow_composite_bwd_dev -> cfn_network_bwd ->
    [one all-reduce of the flat gradient] -> cfn_adam_step.  Same kernels and numbers as `train_step`; the host cost per
    step drops from ~3 ms (torch autograd engine + Python glue) to well under 1 ms, which is what bounds the reference's
    own batch size (N_rand = 512 rays).  Gradients live in ONE flat fp32 buffer (the all-reduce bucket itself)."""

    def __init__(self, network_fn, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8, beta1: float = 0.01,
                 precision: str | None = None, N_samples: int = 128, white_bkgd: bool = False, lindisp: bool = False):
        from . import api
        from .engine import _unwrap, engine_for
        self.module = _unwrap(network_fn)
        self.dev = next(self.module.parameters()).device
        self.eng = engine_for(network_fn, self.dev, precision or api.DEFAULT_TRAIN_PRECISION)
        self.lr, self.betas, self.eps, self.beta1 = lr, betas, eps, beta1
        self.N, self.white_bkgd, self.lindisp = N_samples, white_bkgd, lindisp
        ps = self.eng.params
        n = sum(p.numel() for p in ps)
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.flat_grad, self.exp_avg, self.exp_avg_sq = (torch.zeros(n, **f32) for _ in range(3))
        self.grads, self._m, self._v, o = [], [], [], 0
        for p in ps:
            k = p.numel()
            self.grads.append(self.flat_grad[o:o + k].view(p.shape))
            self._m.append(self.exp_avg[o:o + k])
            self._v.append(self.exp_avg_sq[o:o + k])
            o += k
        import ctypes as C
        arr = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])  # noqa: E731
        self._arrs = (arr(ps), arr(self.grads), arr(self._m), arr(self._v), (C.c_int64 * len(ps))(*[p.numel() for p in ps]))
        self.step_count = 0
        self._g_ld = {}

    @torch.no_grad()
    def step(self, ray_batch, target, t_rand=None, eps_alpha=None, eps_rgb=None, want_loss: bool = True):
        from . import _lib, api
        from ._lib import check
        from .engine import _f32c, _ptr, _stream, bump_weights_epoch
        import ctypes as C
        eng, dev, N, K = self.eng, self.dev, self.N, self.eng.K
        rays, target = _f32c(ray_batch, dev), _f32c(target, dev)
        B = rays.shape[0]
        with torch.cuda.device(dev):
            t_vals = api.reference_t_schedule(N, dev)
            if t_rand is None:
                t_rand = torch.rand(B, N, device=dev)                                         # main:524
            z = eng.zvals(rays, t_vals, _f32c(t_rand, dev), self.lindisp)
            if eps_alpha is None:
                eps_alpha = torch.empty([K, 1], device=dev).normal_()                         # models.py:234
                eps_rgb = torch.empty([K, 3], device=dev).normal_()                           # models.py:246
            ea, ec = _f32c(eps_alpha.reshape(-1), dev), _f32c(eps_rgb, dev)
            fp, ws = eng.network(B, N, rays=rays, z_vals=z, save=True)
            out = eng.flow_composite(fp, z, rays[:, 3:6], 11, ea, ec, self.white_bkgd, train=True)
            rgb = out["rgb_map"]
            partial = torch.empty(B, 2, dtype=torch.float32, device=dev)
            g_rgb = torch.empty_like(rgb)
            check(eng.lib.cfn_kde_nll_f32(_ptr(rgb), _ptr(target), B, K, 1.0 / (3.0 * B), _ptr(partial), _ptr(g_rgb),
                                          _stream()), "cfn_kde_nll_f32")
            cnt = float(B * N * K)
            g_ld = self._g_ld.get(B)
            if g_ld is None:        # d loss / d (sum of log-dets) = -beta1 / (B N K) for both stacks (models.py:286)
                g_ld = self._g_ld[B] = torch.full((2,), -self.beta1 / cnt, dtype=torch.float32, device=dev)
            g_fp, g_glob = eng.flow_composite_bwd(fp, z, rays[:, 3:6], 11, ea, ec, self.white_bkgd, g_rgb, None, g_ld)
            eng.network_bwd(g_fp, B, N, ws, grads=self.grads)
            # the four global latent parameters: through z0 = eps * std + mean (K4's per-ray partials) plus the base
            # log-density of the entropy term, -log(std) per latent dimension (models.py:268/283; its eps^2 part is constant)
            m = self.module
            gg = g_glob.sum(0)
            self.grads[0].copy_(gg[0:1])
            self.grads[1].copy_(gg[1:2] - self.beta1 / m.alpha_std)
            self.grads[2].copy_(gg[2:5])
            self.grads[3].copy_(gg[5:8] - self.beta1 / (3.0 * m.rgb_std))
            rank, w = world()
            if w > 1:
                dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
            self.step_count += 1
            p_arr, g_arr, m_arr, v_arr, numels = self._arrs
            check(eng.lib.cfn_adam_step_f32(len(eng.params), p_arr, g_arr, m_arr, v_arr, numels, float(self.lr),
                                            float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                            int(self.step_count), 1.0 / w, _stream()), "cfn_adam_step_f32")
            bump_weights_epoch()
            if not want_loss:
                return {}
            tot = partial.sum(0) / (3.0 * B)
            nll, mse = tot[0], tot[1]
            base_a, base_c = api._entropy_base_terms(m, ea, ec)
            ld = out["logdet_sums"].sum(0)
            ent = base_a - ld[0] / cnt + base_c - ld[1] / cnt
            import math
            return {"loss": nll + self.beta1 * ent, "loss_nll": nll, "mse": mse,
                    "psnr": -10. * torch.log(mse) / math.log(10.)}
