"""F4 — checkpoint compatibility and uncertainty-map export (host-side; no kernels).

Checkpoints use the reference layout (run_nerf_uncertainty_NF.py:1085-1100): a `torch.save`d dict with
`global_step`, `network_fn_state_dict` (keys carry the `module.` prefix because the reference saves the
`nn.DataParallel` wrapper), optionally `network_fine_state_dict`, and `optimizer_state_dict`.  Loading filters keys
like the reference does (main:360-378) and accepts either prefix convention, so released CF-NeRF weights load into
`NeRFFlowsParams` and files written here load into the reference.  The constructor-time test latents
`sample_alpha / sample_rgb` are NOT part of the reference state_dict (models.py:53-55: plain attributes, different in
every process); they are stored under an extra key the reference ignores, so a render is reproducible.
"""
from __future__ import annotations

import numpy as np
import torch

LATENT_KEY = "cfnerf_b200_latents"


def _bare(module):
    return module.module if isinstance(module, torch.nn.DataParallel) else module


def _prefixed_state_dict(module):
    return {"module." + k: v.detach().cpu() for k, v in _bare(module).state_dict().items()}


def save_checkpoint(path, global_step: int, network_fn, optimizer=None, network_fine=None):
    m = _bare(network_fn)
    ckpt = {"global_step": int(global_step), "network_fn_state_dict": _prefixed_state_dict(network_fn),
            LATENT_KEY: {"sample_alpha": m.sample_alpha.detach().cpu(), "sample_rgb": m.sample_rgb.detach().cpu()}}
    if network_fine is not None:
        ckpt["network_fine_state_dict"] = _prefixed_state_dict(network_fine)
    if optimizer is not None:
        ckpt["optimizer_state_dict"] = optimizer.state_dict()
    torch.save(ckpt, path)
    return path


def _load_into(module, pretrained: dict):
    m = _bare(module)
    model_dict = m.state_dict()
    stripped = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in pretrained.items()}
    picked = {k: v for k, v in stripped.items() if k in model_dict and tuple(v.shape) == tuple(model_dict[k].shape)}
    model_dict.update(picked)                    # main:371-375: filter, overwrite, load
    m.load_state_dict(model_dict)
    return sorted(picked)


def load_checkpoint(path, network_fn, network_fine=None, map_location="cpu", trust: bool = False):
    """-> (global_step, keys loaded).  The optimizer state is deliberately not restored (main:360).

    The file is read with `weights_only=True`: a checkpoint holds only tensors, ints and plain dicts (including the
    optimizer state_dict and the latent key), so the restricted unpickler is enough and a third-party file cannot run
    code.  `trust=True` falls back to the full unpickler for files that carry other Python objects (e.g. an argparse
    namespace someone added) — only for files you produced yourself."""
    try:
        ckpt = torch.load(path, map_location=map_location, weights_only=True)
    except Exception:
        if not trust:
            raise
        ckpt = torch.load(path, map_location=map_location, weights_only=False)
    keys = _load_into(network_fn, ckpt["network_fn_state_dict"])
    if network_fine is not None and "network_fine_state_dict" in ckpt:
        _load_into(network_fine, ckpt["network_fine_state_dict"])
    lat = ckpt.get(LATENT_KEY)
    if lat is not None:
        m = _bare(network_fn)
        m.sample_alpha = lat["sample_alpha"].clone().float()
        m.sample_rgb = lat["sample_rgb"].clone().float()
    return int(ckpt.get("global_step", 0)), keys


def to8b(x):
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)          # run_nerf_helpers.py:17


def uncertainty_maps(rgb_map, disp_map):
    """The tensorboard reduction of main:1122-1136 on a full-image K-sample render: rgb_map (H,W,3,K), disp_map
    (H,W,K) -> dict of float arrays: mean colour, 'uncertainty' std (population std * n/(n-1), main:1130) and the
    mean disparity normalised by its 90th percentile."""
    rgbs = rgb_map.detach().float().cpu().numpy() if isinstance(rgb_map, torch.Tensor) else np.asarray(rgb_map)
    disps = disp_map.detach().float().cpu().numpy() if isinstance(disp_map, torch.Tensor) else np.asarray(disp_map)
    n = rgbs.shape[-1]
    mean = rgbs.mean(-1)
    std = rgbs.std(-1) * n / (n - 1)
    dmean = disps.mean(-1)
    p90 = np.percentile(dmean, 90)
    return {"rgb_mean": mean, "rgb_std": std, "disp_mean": dmean / p90 if p90 > 0 else dmean}


def export_uncertainty_maps(rgb_map, disp_map, prefix: str):
    """Write <prefix>_rgb.png, <prefix>_uncertainty.png (JET) and <prefix>_disp.png (MAGMA) like the reference's
    tensorboard images (main:1125-1147).  Needs OpenCV for the colour maps; returns the uint8 arrays either way."""
    m = uncertainty_maps(rgb_map, disp_map)
    out = {"rgb": to8b(m["rgb_mean"]), "uncertainty": to8b(m["rgb_std"]), "disp": to8b(m["disp_mean"])}
    try:
        import cv2
        out["uncertainty"] = cv2.cvtColor(cv2.applyColorMap(out["uncertainty"], cv2.COLORMAP_JET), cv2.COLOR_BGR2RGB)
        out["disp"] = cv2.cvtColor(cv2.applyColorMap(out["disp"][..., None], cv2.COLORMAP_MAGMA), cv2.COLOR_BGR2RGB)
        for k, v in out.items():
            cv2.imwrite(f"{prefix}_{k}.png", cv2.cvtColor(v, cv2.COLOR_RGB2BGR))
    except ImportError:
        pass
    return out
