"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference (container-only).

Imports ``/root/reference/run_nerf_uncertainty_NF.py`` without touching it by pre-seeding
``sys.modules`` with empty stand-ins for the third-party modules the reference imports at
top level but never uses on the render/train hot path:
``imageio, kornia, skimage.metrics, matplotlib.{pyplot,colors}, configargparse``.

``/root/reference`` does not exist on the GPU box; there the loader falls back to ``oracle/_ref/``, the
byte-identical copy of the reference's ``*.py`` files staged by ``oracle/build_ref.py`` (git-ignored,
travels with the snapshot).  Users: ``oracle/make_golden.py`` (fixture generation), the container-side
tests that pin ``oracle/cfnerf_oracle.py`` to the live reference, the ``-m gpu`` tests that put
``cfnerf_b200.install()`` behind the real module, and the CPU legs of ``bench.py`` (kind "reference").
Never imported by the product package.
"""
from __future__ import annotations

import os
import sys
import types

def _find_root() -> str:
    """$CFNERF_REFERENCE_ROOT, else the read-only tree of the build container, else the byte-identical copy that
    oracle/build_ref.py staged under oracle/_ref/ (git-ignored; it travels to the GPU box)."""
    env = os.environ.get("CFNERF_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/run_nerf_uncertainty_NF.py"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "run_nerf_uncertainty_NF.py"))


def _stand_in(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


_CACHED = None


def load_reference():
    """Return (main_module, NeRF_Flows class).  Raises FileNotFoundError off-container."""
    global _CACHED
    if _CACHED is not None:
        return _CACHED
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    _stand_in("imageio")
    _stand_in("kornia", create_meshgrid=None)
    sk = _stand_in("skimage")
    sk.metrics = _stand_in("skimage.metrics", structural_similarity=None)
    mpl = _stand_in("matplotlib")
    mpl.pyplot = _stand_in("matplotlib.pyplot")
    mpl.colors = _stand_in("matplotlib.colors", Normalize=object)
    _stand_in("configargparse")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings

    import torch

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import run_nerf_uncertainty_NF as main  # noqa: E402
        from model.models import NeRF_Flows  # noqa: E402
    # the reference switches autograd anomaly mode on at import (models.py:5, helpers:2);
    # leave it off unless a caller wants the as-shipped timing.
    torch.autograd.set_detect_anomaly(False)
    _CACHED = (main, NeRF_Flows)
    return _CACHED


def make_reference_args(cfg, n_gpus: int = 1):
    """Namespace that ``create_nerf``/``NeRF_Flows.__init__`` read (main:317-336, models:20-36)."""
    import torch

    return types.SimpleNamespace(
        netdepth=cfg.D, netwidth=cfg.W, multires=cfg.L_pos, multires_views=cfg.L_dir, i_embed=0,
        input_ch=cfg.in_pos, input_ch_views=cfg.in_dir, K_samples=cfg.K, skips=[cfg.D / 2],
        use_viewdirs=True, h_alpha_size=cfg.h_alpha, h_rgb_size=cfg.h_rgb, n_flows=cfg.F,
        type_flows="triangular", n_hidden=128, device=torch.device("cpu"), n_gpus=n_gpus,
        netchunk_per_gpu=1024 * 64, N_importance=0,
    )


def build_reference_model(cfg, params: dict, sample_alpha, sample_rgb):
    """Instantiate the reference ``NeRF_Flows`` and load OUR deterministic parameters into it."""
    import torch

    main, NeRF_Flows = load_reference()
    args = make_reference_args(cfg)
    model = NeRF_Flows(args)
    sd = model.state_dict()
    for k, v in params.items():
        assert k in sd and tuple(sd[k].shape) == tuple(v.shape), k
        sd[k] = v.detach().clone().to(torch.float32)
    model.load_state_dict(sd)
    model.sample_alpha = sample_alpha.detach().clone().float()
    model.sample_rgb = sample_rgb.detach().clone().float()
    embed_fn, in_pos = main.get_embedder(cfg.L_pos, 0)
    embeddirs_fn, in_dir = main.get_embedder(cfg.L_dir, 0)
    assert in_pos == cfg.in_pos and in_dir == cfg.in_dir

    def network_query_fn(inputs, viewdirs, network_fn, is_val, is_test):
        return main.run_network(inputs, viewdirs, network_fn, is_val, is_test, embed_fn=embed_fn,
                                embeddirs_fn=embeddirs_fn, netchunk=1024 * 64)

    return main, model, network_query_fn
