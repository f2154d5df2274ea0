"""TEST INFRASTRUCTURE ONLY — a stand-in for the reference MODULE's caller side: `batchify_rays` and `render`
(run_nerf_uncertainty_NF.py:88-170), restated so that the drop-in (`cfnerf_b200.install`) can be exercised on the GPU
box where /root/reference does not exist.  Like the reference, `batchify_rays` resolves `render_rays` through the
module namespace at call time, which is exactly the hook `install` uses."""
from __future__ import annotations

import types

import torch

from . import cfnerf_oracle as O


def make_module() -> types.ModuleType:
    m = types.ModuleType("run_nerf_uncertainty_NF_standin")

    def render_rays(*a, **k):  # replaced by install(); the stand-in has no implementation of its own
        raise RuntimeError("render_rays was not installed")

    def raw2outputs(*a, **k):
        raise RuntimeError("raw2outputs was not installed")

    def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):                                   # main:88-100
        all_ret = {}
        for i in range(0, rays_flat.shape[0], chunk):
            ret = m.render_rays(rays_flat[i:i + chunk], **kwargs)      # module-global lookup, like main:93
            for k in ret:
                all_ret.setdefault(k, []).append(ret[k])
        return {k: torch.cat(v, 0) for k, v in all_ret.items()}

    def render(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
               c2w_staticcam=None, **kwargs):                                                  # main:103-170
        if c2w is not None:
            rays_o, rays_d = O.get_rays(H, W, focal, c2w.cpu())
            rays_o, rays_d = rays_o.to(c2w.device), rays_d.to(c2w.device)
        else:
            rays_o, rays_d = rays
        viewdirs = None
        if use_viewdirs:
            viewdirs = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
            viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
        sh = rays_d.shape
        if ndc:
            rays_o, rays_d = O.ndc_rays(H, W, focal, 1., rays_o, rays_d)
        rays_o = torch.reshape(rays_o, [-1, 3]).float()
        rays_d = torch.reshape(rays_d, [-1, 3]).float()
        near_t, far_t = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
        rays_cat = torch.cat([rays_o, rays_d, near_t, far_t], -1)
        if use_viewdirs:
            rays_cat = torch.cat([rays_cat, viewdirs], -1)
        all_ret = batchify_rays(rays_cat, chunk, **kwargs)
        for k in all_ret:
            if k not in ("loss_entropy", "loss_entropy_uniformsample"):
                all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
        k_extract = ["rgb_map", "disp_map", "depth_map"]
        return [all_ret[k] for k in k_extract] + [{k: all_ret[k] for k in all_ret if k not in k_extract}]

    m.render_rays, m.raw2outputs, m.batchify_rays, m.render = render_rays, raw2outputs, batchify_rays, render
    return m
