"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every symbol
include/nca_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

import nca_b200
from nca_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    return nca_b200.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "nca_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(nca_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None, name


def test_abi_version_and_workspace(lib):
    assert lib.nca_abi_version() == 6
    d = _lib.DyncaDesc(8, 16, 256, 256, 128, _lib.NCA_COND_CPE, 2, 1, 2, 0, _lib.NCA_MASK_PHILOX, 0.5)
    fwd = lib.nca_dynca_workspace_bytes(C.byref(d), 0)
    bwd = lib.nca_dynca_workspace_bytes(C.byref(d), 1)
    assert 0 < fwd < bwd
    assert bwd >= 2 * 8 * 16 * 256 * 256 * 4
    e = _lib.EncDesc(4, 20, 64, 64, 64, 3, _lib.NCA_MASK_PHILOX, 0.1, 0.5, 10.0, 0)
    assert 0 < lib.nca_enc_workspace_bytes(C.byref(e), 0) < lib.nca_enc_workspace_bytes(C.byref(e), 1)
    e.hid = 32
    assert lib.nca_enc_workspace_bytes(C.byref(e), 0) == 0


def test_bad_arguments_are_reported(lib):
    d = _lib.DyncaDesc(1, 64, 8, 8, 96, 0, 0, 1, 1, 0, _lib.NCA_MASK_PHILOX, 0.5)   # C too large
    assert lib.nca_dynca_workspace_bytes(C.byref(d), 0) == 0
    rc = lib.nca_dynca_perceive(C.byref(d), None, None, None, None)
    assert rc == -1 and b"C=64" in lib.nca_last_error()
    d = _lib.DyncaDesc(1, 12, 9, 8, 96, 0, 0, 1, 2, 0, _lib.NCA_MASK_PHILOX, 0.5)   # odd H with two scales
    assert lib.nca_dynca_perceive(C.byref(d), None, None, None, None) == -1
    assert b"even" in lib.nca_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    d = _lib.DyncaDesc(1, 12, 8, 8, 96, 0, 0, 1, 1, 0, _lib.NCA_MASK_PHILOX, 0.5)
    x = torch.zeros(1, 12, 8, 8)
    z = torch.zeros(1, 48, 8, 8)
    rc = lib.nca_dynca_perceive(C.byref(d), x.data_ptr(), None, z.data_ptr(), None)
    assert rc == -3 and b"no CPU fallback" in lib.nca_last_error()
    m = nca_b200.DyNCA_EC(12, 3, device=torch.device("cpu"))
    with pytest.raises(nca_b200.NcaError):
        m.forward_nsteps(x, 2)
    with pytest.raises(nca_b200.NcaError):
        m(x)


def test_module_state_dict_layout():
    """Parameter names / shapes are the checkpoint contract."""
    ec = nca_b200.DyNCA_EC(13, 3, fc_dim=96, pos_emb=None, padding_mode="circular", device=torch.device("cpu"))
    sd = ec.state_dict()
    assert {k: tuple(v.shape) for k, v in sd.items()} == {
        "w1.weight": (96, 52, 1, 1), "w1.bias": (96,), "w2.weight": (13, 96, 1, 1), "w2.bias": (13,)}
    assert float(sd["w2.bias"].abs().max()) == 0.0
    cd = nca_b200.DyNCA_CD(12, 3, conditioning="edges", device=torch.device("cpu"))
    sd = cd.state_dict()
    assert tuple(sd["w1.weight"].shape) == (96, 51, 1, 1)
    assert {"cond_layer.sobel_x.weight", "cond_layer.sobel_y.weight", "cond_layer.laplacian.weight"} <= set(sd)
    assert tuple(ec.seed(2, 16).shape) == (2, 12, 16, 16)        # EC seeds c_in - 1 channels (dynca.py:140)
    assert tuple(cd.seed(2, (8, 16)).shape) == (2, 12, 16, 8)
    c16 = nca_b200.DyNCA_EC(16, 3, fc_dim=128, perception_scales=[0, 1], device=torch.device("cpu"))
    assert tuple(c16.state_dict()["w1.weight"].shape) == (128, 66, 1, 1)


def test_callers_argument_checks_and_no_cpu_fallback(lib):
    """The callers' entry points validate their arguments before touching the device and, like the step, have no
    CPU path: with host tensors and no GPU they return NCA_ERR_CUDA; the host mirror refuses CPU tensors."""
    from nca_b200 import trainer as Tr, video as V
    pool = torch.zeros(4, 3, 8, 8)
    idx = torch.tensor([0, 1], dtype=torch.int64)
    out = torch.zeros(2, 3, 8, 8)
    st = None
    assert lib.nca_pool_gather(4, 3, 8, 8, pool.data_ptr(), idx.data_ptr(), 2, None, 1, None, 0, None, out.data_ptr(), st) == -1
    assert b"extra" in lib.nca_last_error()
    assert lib.nca_pool_gather(4, 3, 8, 8, pool.data_ptr(), idx.data_ptr(), 2, None, 0, None, 3, None, out.data_ptr(), st) == -1
    assert lib.nca_pool_scatter(4, 3, 8, 8, pool.data_ptr(), idx.data_ptr(), 2, out.data_ptr(), 2, st) == -1      # C < Cp
    assert lib.nca_pool_dead_flags(4, 3, 8, 8, pool.data_ptr(), idx.data_ptr(), 2, 3, 0.1, out.data_ptr(), st) == -1   # living_dim
    assert lib.nca_state_to_rgb8(1, 2, 8, 8, pool.data_ptr(), 2.0, out.data_ptr(), st) == -1                    # C < 3
    assert lib.nca_frame_to_cond_channel(1, 3, 8, 8, pool.data_ptr(), out.data_ptr(), 3, st) == -1              # channel
    assert lib.nca_overflow_loss(pool.data_ptr(), 16, out.data_ptr(), None, 1.0, None, 0, None, 0, st) == -4    # workspace
    assert lib.nca_overflow_workspace_bytes() >= 1024
    ptrs = (C.c_void_p * 1)(pool.data_ptr())
    numel = (C.c_int64 * 1)(16)
    assert lib.nca_normalized_adam_step(0, ptrs, ptrs, ptrs, ptrs, numel, 1, 1e-3, 0.9, 0.999, 1e-8, 1e-8, 1, 0, st) == -1
    assert lib.nca_normalized_adam_step(1, ptrs, ptrs, ptrs, ptrs, numel, 0, 1e-3, 0.9, 0.999, 1e-8, 1e-8, 1, 0, st) == -1   # step
    assert lib.nca_normalized_adam_step(17, ptrs, ptrs, ptrs, ptrs, numel, 1, 1e-3, 0.9, 0.999, 1e-8, 1e-8, 1, 0, st) == -1
    if not torch.cuda.is_available():
        assert lib.nca_pool_gather(4, 3, 8, 8, pool.data_ptr(), idx.data_ptr(), 2, None, 0, None, 0, None, out.data_ptr(), st) == -3
        assert b"no CPU fallback" in lib.nca_last_error()
        assert lib.nca_state_to_rgb8(1, 3, 8, 8, pool.data_ptr(), 2.0, out.data_ptr(), st) == -3
        assert lib.nca_normalized_adam_step(1, ptrs, ptrs, ptrs, ptrs, numel, 1, 1e-3, 0.9, 0.999, 1e-8, 1e-8, 1, 0, st) == -3
    for call in (lambda: Tr.pool_gather(pool, [0, 1]), lambda: Tr.pool_scatter(pool, [0, 1], out), lambda: Tr.overflow_loss(pool),
                 lambda: V.state_to_rgb8(pool), lambda: V.rgb_to_grayscale(pool[:, :3])):
        with pytest.raises(nca_b200.NcaError):
            call()
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises(nca_b200.NcaError):
        nca_b200.NormalizedAdam([p]).step()
