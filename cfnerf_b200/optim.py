"""F3 — `FusedAdam`: torch.optim.Adam semantics (the reference's optimiser, run_nerf_uncertainty_NF.py:339) with the whole
step in one CUDA launch (cfn_adam_step_f32).  It keeps the `torch.optim.Optimizer` surface the trainer touches:
`zero_grad()`, `step()`, `param_groups[i]['lr']` (the lr decay loop, main:1073-1077) and `state_dict()` (main:1091)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check
from .engine import bump_weights_epoch


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, grad_scale: float = 1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.grad_scale = grad_scale      # e.g. 1/world when gradients were SUM all-reduced

    @torch.no_grad()
    def step(self, closure=None):
        lib = _lib.load()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("FusedAdam needs contiguous fp32 CUDA parameters (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
            steps = {self.state[p]["step"] for p in ps}
            if len(steps) != 1:
                raise RuntimeError("FusedAdam: parameters of one group must share the step count")
            grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
            n = len(ps)
            arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
            numels = (C.c_int64 * n)(*[p.numel() for p in ps])
            b1, b2 = group["betas"]
            with torch.cuda.device(ps[0].device):
                check(lib.cfn_adam_step_f32(n, arr(ps), arr(grads), arr([self.state[p]["exp_avg"] for p in ps]),
                                            arr([self.state[p]["exp_avg_sq"] for p in ps]), numels, float(group["lr"]),
                                            float(b1), float(b2), float(group["eps"]), int(steps.pop()),
                                            float(self.grad_scale), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                      "cfn_adam_step_f32")
        # the kernel wrote the parameters behind autograd's back (no version-counter bump): tell the engines to re-pack
        bump_weights_epoch()
        return None


def decayed_lr(lrate: float, lrate_decay: int, global_step: int) -> float:
    """main:1073-1077: lrate * 0.1 ** (global_step / (lrate_decay * 1000))."""
    return lrate * (0.1 ** (global_step / (lrate_decay * 1000)))
