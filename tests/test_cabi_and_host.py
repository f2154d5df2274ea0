"""CPU-side checks: the C-ABI library loads and exports every symbol include/cfnerf_b200.h declares; argument
validation fails cleanly without a GPU; the host mirror of the reference interface is consistent with the oracle."""
import ctypes as C
import os
import re

import pytest
import torch

from oracle import cfnerf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from cfnerf_b200 import _lib
    return _lib.load()


def test_header_symbols_all_exported(lib):
    from cfnerf_b200 import _lib
    text = open(os.path.join(ROOT, "include", "cfnerf_b200.h")).read()
    declared = set(re.findall(r"\b(cfn_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported by the library"
    assert declared == set(_lib.SYMBOLS), "ctypes table and header disagree"


def test_invalid_config_is_rejected_without_touching_cuda(lib):
    from cfnerf_b200._lib import CfnConfigC
    cfg = CfnConfigC(D=1, W=512, L_pos=10, L_dir=4, h_alpha=64, h_rgb=64, F=4, K=32, precision=0)
    h = C.c_void_p()
    assert lib.cfn_create(C.byref(cfg), C.byref(h)) == -1
    assert b"netdepth" in lib.cfn_last_error()
    cfg = CfnConfigC(D=8, W=512, L_pos=10, L_dir=4, h_alpha=64, h_rgb=64, F=40, K=32, precision=0)
    assert lib.cfn_create(C.byref(cfg), C.byref(h)) == -1
    assert lib.cfn_version() >= 100


def test_no_cpu_fallback():
    """The product path refuses to run without its CUDA device instead of silently computing elsewhere."""
    import cfnerf_b200 as cf
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    net = cf.NeRFFlowsParams(netwidth=64, K_samples=4)
    with pytest.raises(RuntimeError):
        cf.render_rays(O.synthetic_rays(4), net, None, 128, False, False)
    with pytest.raises(RuntimeError):
        cf.raw2outputs(torch.zeros(1, 4, 2, 4), torch.zeros(1, 4), torch.zeros(1, 3))
    with pytest.raises(RuntimeError):
        net(torch.zeros(2, 90))


def test_param_container_matches_reference_state_dict_layout():
    import cfnerf_b200 as cf
    for cfg in (O.CfnConfig(), O.CfnConfig(W=256, K=64, h_alpha=32)):
        p = O.make_params(cfg, 0)
        net = cf.NeRFFlowsParams.from_oracle_params(cfg, p, *O.make_latents(cfg, 0))
        sd = net.state_dict()
        assert set(sd) == set(p)
        for k in p:
            assert tuple(sd[k].shape) == tuple(p[k].shape), k
            assert torch.equal(sd[k], p[k])
        assert sum(v.numel() for v in sd.values()) == cfg.n_params()


def test_schedule_matches_oracle():
    import cfnerf_b200 as cf
    assert torch.equal(cf.reference_t_schedule(128, "cpu"), O.reference_t_schedule())
    assert torch.equal(cf.reference_t_schedule(64, "cpu"), O.coarse_t_schedule(64))


def test_kde_nll_loss_matches_oracle():
    import cfnerf_b200 as cf
    g = torch.Generator().manual_seed(0)
    rgb = torch.rand(16, 3, 32, generator=g)
    tgt = torch.rand(16, 3, generator=g)
    ent = torch.tensor(0.37)
    a = cf.kde_nll_loss(rgb, tgt, ent.expand(16 * 128, 32, 1), 32, 0.01)
    b = O.kde_nll_loss(rgb, tgt, ent, 32, 0.01)
    for k in a:
        assert abs(float(a[k]) - float(b[k])) <= 1e-6 * max(1.0, abs(float(b[k]))), k


def test_tf32_gemm_layout_contract_is_checked_on_the_host(lib):
    """cfn_gemm_f32 engine 1 (TMA-fed tcgen05 GEMM) refuses operand layouts the TMA cannot address — unit stride along
    one axis, 16-byte aligned bases and row strides — before any CUDA call is made (so this runs without a GPU)."""
    import ctypes as C

    def call(a_ptr, a_rs, a_cs, b_ptr, b_rs, b_cs, M=64, N=32, K=63, epi=0, split=1):
        return lib.cfn_gemm_f32(1, C.c_void_p(a_ptr), a_rs, a_cs, C.c_void_p(b_ptr), b_rs, b_cs, C.c_void_p(4096), N,
                                None, None, 0, M, N, K, epi, 0, split, 0, None)

    assert call(1024, 63, 1, 2048, 32, 1) == -1          # A row stride 63 floats: not a multiple of 16 bytes
    assert b"not supported" in lib.cfn_last_error()
    assert call(1028, 64, 1, 2048, 32, 1) == -1          # A base not 16-byte aligned
    assert call(1024, 64, 2, 2048, 32, 1) == -1          # no unit stride in A
    assert call(1024, 64, 1, 2048, 32, 1, epi=1) == -1   # ReLU epilogue is only instantiated for K-major B
    assert call(1024, 64, 1, 2048, 1, 64, split=4) == -1  # split-K (wgrad) needs M-major A and N-major B
    assert lib.cfn_gemm_f32(2, None, 0, 0, None, 0, 0, None, 0, None, None, 0, 4, 4, 4, 0, 0, 1, 0, None) == -1   # engine id
