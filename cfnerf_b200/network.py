"""Parameter container with the reference's `NeRF_Flows` surface (model/models.py:13-67, 294-350).

It holds exactly the tensors of the reference `state_dict()` (same keys and shapes, so released checkpoints
load with `load_state_dict`) plus the attributes the hot path reads off the module (`sample_alpha`,
`sample_rgb`, `K_samples`, ...).  It has no PyTorch arithmetic of its own: `forward` (the `network_fn(embedded,
is_val, is_test)` surface of models.py:188) routes through the CUDA library like every other entry of this package
(no fallback).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn


class _Amortised(nn.Module):
    """TriangularSylvesterNeRF's amortisation heads (model/models.py:339-350)."""

    def __init__(self, h: int, z: int, n_flows: int):
        super().__init__()
        self.amor_d = nn.Linear(h, n_flows * z * z)
        self.amor_diag1 = nn.Sequential(nn.Linear(h, n_flows * z))   # tanh applied by the kernel epilogue
        self.amor_diag2 = nn.Sequential(nn.Linear(h, n_flows * z))
        self.amor_b = nn.Linear(h, n_flows * z)


class NeRFFlowsParams(nn.Module):
    """Same constructor contract as the reference: `NeRFFlowsParams(args)` with the `create_nerf` namespace
    (run_nerf_uncertainty_NF.py:317-329) or keyword overrides."""

    def __init__(self, args=None, **kw):
        super().__init__()
        a = SimpleNamespace(netdepth=8, netwidth=512, input_ch=63, input_ch_views=27, K_samples=32, h_alpha_size=64,
                            h_rgb_size=64, n_flows=4, use_viewdirs=True)
        if args is not None:
            for k in vars(a):
                if hasattr(args, k):
                    setattr(a, k, getattr(args, k))
        for k, v in kw.items():
            setattr(a, k, v)
        if not a.use_viewdirs:
            raise ValueError("the reference model is only constructible with use_viewdirs (model/models.py:63-64)")
        self.D, self.W = a.netdepth, a.netwidth
        self.input_ch, self.input_ch_views = a.input_ch, a.input_ch_views
        self.K_samples = self.sample_size = a.K_samples
        self.h_alpha_size, self.h_rgb_size, self.n_flows = a.h_alpha_size, a.h_rgb_size, a.n_flows
        self.skips = [self.D / 2]  # main:327 (true division, as in the reference)
        W, D = self.W, self.D
        self.pts_linears = nn.ModuleList(
            [nn.Linear(self.input_ch, W)] +
            [nn.Linear(W, W) if i not in self.skips else nn.Linear(W + self.input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(self.input_ch_views + W, W // 2)])
        self.alpha_mean = nn.Parameter(torch.zeros(1))
        self.alpha_std = nn.Parameter(torch.ones(1))
        self.rgb_mean = nn.Parameter(torch.zeros(3))
        self.rgb_std = nn.Parameter(torch.ones(3))
        # constructor-time test latents: plain attributes in the reference, NOT in state_dict (models.py:53-55)
        self.sample_alpha = torch.empty([self.K_samples, 1]).normal_()
        self.sample_rgb = torch.empty([self.K_samples, 3]).normal_()
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)          # dead in the reference (models.py:59), kept for checkpoints
        self.alpha_std_linear = nn.Linear(W, 1)      # dead (models.py:60)
        self.h_alpha_linear = nn.Linear(W, self.h_alpha_size)
        self.h_rgb_linear = nn.Linear(W // 2, self.h_rgb_size)
        self.flows_rgb = _Amortised(self.h_rgb_size, 3, self.n_flows)
        self.flows_alpha = _Amortised(self.h_alpha_size, 1, self.n_flows)

    def forward(self, embedded, is_val=False, is_test=False, *, eps_alpha=None, eps_rgb=None, precision=None):
        """`network_fn(embedded, is_val, is_test)` (models.py:188), the call `batchify` makes (main:55).

        embedded (M, input_ch + input_ch_views) = [gamma(p) | gamma(d)] as `run_network` assembles it (main:70-80).
        `get_embedder` keeps the raw input in front of the sin/cos blocks (`include_input: True`, helpers:59), so the
        3-D point is columns 0:3 and the view direction columns input_ch : input_ch+3; the fused kernel re-encodes
        them itself (the other columns are redundant).  Returns (raw (M,K,4) [rgb|sigma], zeros_like(raw)) in test mode
        (models.py:223) and (raw, entropy scalar broadcast to (M,K,1)) otherwise (models.py:291).  Not differentiable:
        training goes through cfnerf_b200.api.render_rays (one autograd node around network + flows + compositing)."""
        from . import api
        if embedded.dim() != 2 or embedded.shape[-1] != self.input_ch + self.input_ch_views:
            raise ValueError(f"embedded must be (M,{self.input_ch + self.input_ch_views}) = [gamma(p)|gamma(d)] "
                             f"(main:70-80), got {tuple(embedded.shape)}")
        pts = embedded[:, 0:3]
        dirs = embedded[:, self.input_ch:self.input_ch + 3]
        # every row carries its own direction: M rays of one sample each
        raw, ent = api.run_network(pts[:, None, :], dirs, self, is_val, is_test, eps_alpha=eps_alpha, eps_rgb=eps_rgb,
                                   precision=precision)
        return raw[:, 0], (ent[:, 0] if is_test else ent)

    @staticmethod
    def from_oracle_params(cfg, params: dict, sample_alpha=None, sample_rgb=None) -> "NeRFFlowsParams":
        """Build from a dict keyed like `state_dict()` (e.g. oracle.make_params) — test/bench convenience."""
        m = NeRFFlowsParams(netdepth=cfg.D, netwidth=cfg.W, input_ch=3 + 6 * cfg.L_pos,
                            input_ch_views=3 + 6 * cfg.L_dir, K_samples=cfg.K, h_alpha_size=cfg.h_alpha,
                            h_rgb_size=cfg.h_rgb, n_flows=cfg.F)
        sd = m.state_dict()
        for k, v in params.items():
            assert k in sd and tuple(sd[k].shape) == tuple(v.shape), k
            sd[k] = v.detach().clone().float()
        m.load_state_dict(sd)
        if sample_alpha is not None:
            m.sample_alpha = sample_alpha.detach().clone().float()
            m.sample_rgb = sample_rgb.detach().clone().float()
        return m
