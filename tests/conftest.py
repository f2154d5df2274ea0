"""Shared fixtures.  `-m "not gpu"` runs here (no GPU); `-m gpu` runs on a B200 box."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))


def load_golden(name):
    """-> (dict of arrays, CfnConfig, params dict) with the parameter fingerprint verified."""
    from oracle import cfnerf_oracle as O

    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))
    c = [int(v) for v in g["cfg"]]
    cfg = O.CfnConfig(D=c[0], W=c[1], L_pos=c[2], L_dir=c[3], h_alpha=c[4], h_rgb=c[5], F=c[6], K=c[7])
    variant = str(g["variant"])
    params = {}
    if variant != "none":
        params = O.make_params(cfg, int(g["seed"]), variant)
        chk = O.params_checksum(params)
        assert abs(chk - float(g["params_checksum"])) <= 1e-9 * max(1.0, abs(chk)), (
            "regenerated parameters differ from the ones the golden file was made with (torch RNG drift)")
    return g, cfg, params


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))
