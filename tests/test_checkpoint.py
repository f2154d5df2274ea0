"""F4 (CPU): checkpoints in the reference layout round-trip, load with either key prefix, and load into / from the
unmodified reference model when it is present (build container)."""
import os

import numpy as np
import pytest
import torch

from oracle import cfnerf_oracle as O
from oracle import refload


def _net(seed):
    import cfnerf_b200 as cf
    cfg = O.CfnConfig(W=256, K=64, h_alpha=32)
    return cfg, cf.NeRFFlowsParams.from_oracle_params(cfg, O.make_params(cfg, seed, "lively"), *O.make_latents(cfg, seed))


def test_roundtrip_reference_layout(tmp_path):
    from cfnerf_b200 import checkpoint as ck
    cfg, a = _net(0)
    opt = torch.optim.Adam(a.parameters(), lr=5e-4)
    path = ck.save_checkpoint(os.path.join(tmp_path, "000010_01.tar"), 10, a, opt)
    raw = torch.load(path, weights_only=False)
    assert set(raw) >= {"global_step", "network_fn_state_dict", "optimizer_state_dict"}
    assert all(k.startswith("module.") for k in raw["network_fn_state_dict"])      # what the reference writes (main:1089)
    _, b = _net(1)
    step, keys = ck.load_checkpoint(path, torch.nn.DataParallel(b))
    assert step == 10 and len(keys) == len(a.state_dict())
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k]), k
    assert torch.equal(a.sample_alpha, b.sample_alpha) and torch.equal(a.sample_rgb, b.sample_rgb)
    # un-prefixed files (a bare module's state_dict) load too
    torch.save({"global_step": 3, "network_fn_state_dict": a.state_dict()}, os.path.join(tmp_path, "bare.tar"))
    _, c = _net(2)
    assert ck.load_checkpoint(os.path.join(tmp_path, "bare.tar"), c)[0] == 3
    assert torch.equal(c.state_dict()["pts_linears.5.weight"], a.state_dict()["pts_linears.5.weight"])


@pytest.mark.skipif(not refload.reference_available(), reason="reference tree not present")
def test_checkpoint_loads_into_the_unmodified_reference(tmp_path):
    from cfnerf_b200 import checkpoint as ck
    cfg, a = _net(0)
    path = ck.save_checkpoint(os.path.join(tmp_path, "000001_01.tar"), 1, a)
    _, ref_model, _ = refload.build_reference_model(cfg, O.make_params(cfg, 9, "default"), *O.make_latents(cfg, 9))
    wrapped = torch.nn.DataParallel(ref_model)
    ckpt = torch.load(path, weights_only=False)
    model_dict = wrapped.state_dict()                              # the reference's own reload code, main:366-375
    pre = {k: v for k, v in ckpt["network_fn_state_dict"].items() if k in model_dict}
    assert len(pre) == len(a.state_dict())
    model_dict.update(pre)
    wrapped.load_state_dict(model_dict)
    for k, v in a.state_dict().items():
        assert torch.equal(ref_model.state_dict()[k], v), k
    # and the other way round: what the reference saves loads here
    torch.save({"global_step": 7, "network_fn_state_dict": wrapped.state_dict()}, os.path.join(tmp_path, "ref.tar"))
    _, b = _net(3)
    step, keys = ck.load_checkpoint(os.path.join(tmp_path, "ref.tar"), b)
    assert step == 7
    for k, v in b.state_dict().items():
        assert torch.equal(v, ref_model.state_dict()[k]), k


def test_uncertainty_maps_follow_the_tensorboard_reduction(tmp_path):
    from cfnerf_b200 import checkpoint as ck
    g = np.random.default_rng(0)
    rgb = g.random((6, 5, 3, 32)).astype(np.float32)
    disp = g.random((6, 5, 32)).astype(np.float32)
    m = ck.uncertainty_maps(torch.from_numpy(rgb), torch.from_numpy(disp))
    np.testing.assert_allclose(m["rgb_mean"], rgb.mean(-1), rtol=1e-6)
    np.testing.assert_allclose(m["rgb_std"], rgb.std(-1) * 32 / 31, rtol=1e-5)          # main:1130
    np.testing.assert_allclose(m["disp_mean"], disp.mean(-1) / np.percentile(disp.mean(-1), 90), rtol=1e-5)
    out = ck.export_uncertainty_maps(torch.from_numpy(rgb), torch.from_numpy(disp), os.path.join(tmp_path, "view0"))
    assert out["rgb"].shape == (6, 5, 3) and out["rgb"].dtype == np.uint8
