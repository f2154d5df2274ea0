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
