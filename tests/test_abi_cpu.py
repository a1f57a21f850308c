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
    assert lib.nca_abi_version() == 5
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
