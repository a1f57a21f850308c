"""Parity of the CUDA DyNCA path (through the C ABI / drop-in modules) against the golden vectors produced by
the unmodified reference and against the CPU oracle on seeded inputs.

Tolerances (fp32 MLP): per-step / rollout state 1e-5 relative (max-abs / max-abs), gradients 1e-4 relative
(BASELINE.json north_star); trained-weight 24-step cases use 2e-3 on gradients because fp32 summation order
alone moves them by 5e-4 between two CPU evaluations (see tests/test_oracle_golden.py).

Every parity case runs twice with the SAME tolerances: precision "fp32" (CUDA-core FFMA kernels) and "f16x3" (the
update MLP of forward and BPTT on tcgen05 with split-precision bf16 hi + lo operands, NCA_PREC_F16X3) - the
tensor-core mode that meets the fp32-grade bars."""
import numpy as np
import pytest
import torch

import nca_b200
from nca_b200 import functional as Fn, _lib
from oracle import nca_oracle as O
from oracle import philox
from helpers import DYNCA_CASES, load_case, rel_err, cond_for

pytestmark = pytest.mark.gpu
DEV = "cuda"


def build_model(m, t, precision="fp32"):
    dev = torch.device(DEV)
    kw = dict(c_in=m["C"], c_out=3, fc_dim=m["fc"], padding_mode=m["pad"], perception_scales=list(m["scales"]),
              device=dev, precision=precision)
    if m["flavour"] == "ec":
        model = nca_b200.DyNCA_EC(pos_emb=("CPE" if m["cond"] == "cpe" else None), **kw)
    else:
        model = nca_b200.DyNCA_CD(conditioning={"cpe": "pos_emb", "edges": "edges", None: None}[m["cond"]],
                                  edge_transform=m["edge_transform"], **kw)
    with torch.no_grad():
        model.w1.weight.copy_(t["w1"].reshape(model.w1.weight.shape))
        model.w1.bias.copy_(t["b1"])
        model.w2.weight.copy_(t["w2"].reshape(model.w2.weight.shape))
        model.w2.bias.copy_(t["b2"])
    return model


FP32_GRADE = ["fp32", "f16x3"]


def _expect_variant(cfg, B, H, W, precision):
    """f16x3 must really dispatch to the tensor-core kernels (variant 3), fp32 to the CUDA-core ones (0)"""
    want = 3 if precision == "f16x3" else 0
    assert Fn.dynca_kernel_variant(cfg, B, H, W) == want and Fn.dynca_kernel_variant(cfg, B, H, W, backward=True) == want


@pytest.mark.parametrize("precision", FP32_GRADE)
@pytest.mark.parametrize("name", DYNCA_CASES)
def test_golden_case(name, precision):
    t, m = load_case(name)
    model = build_model(m, t, precision=precision)
    kind, cc = {"cpe": (_lib.NCA_COND_CPE, 2), "edges": (_lib.NCA_COND_TENSOR, 3), None: (_lib.NCA_COND_NONE, 0)}[m["cond"]]
    _expect_variant(Fn.DyncaConfig(m["C"], m["fc"], m["pad"], m["scales"], kind, cc, precision=precision), m["B"], m["H"], m["W"], precision)
    x0 = t["x0"].to(DEV).requires_grad_(True)
    masks = t["masks"].to(DEV)
    kwargs = dict(cond_img=t["cond_img"].to(DEV)) if m["flavour"] == "cd" and m["cond"] == "edges" else {}
    if m["flavour"] == "cd" and m["cond"] != "edges":
        kwargs = dict(cond_img=None)
    # single-step perception (return_perception=True)
    with torch.no_grad():
        _, _, percept = model(x0.detach(), update_rate=m["rate"], return_perception=True, masks=masks[:1], **kwargs)
    assert rel_err(percept.cpu(), t["percept"]) < 1e-5
    if "cond_mat" in t:
        assert rel_err(model.cond_layer(t["cond_img"].to(DEV)).cpu(), t["cond_mat"]) < 1e-5
    # rollout with the reference's own fire masks
    state, rgb, mids = model.forward_nsteps(x0, m["T"], update_rate=m["rate"], return_middle_feature=True,
                                            masks=masks, **kwargs)
    assert len(mids) == m["T"]
    assert rel_err(state.detach().cpu(), t["final"]) < 1e-5
    assert rel_err(rgb.detach().cpu(), t["rgb_last"]) < 1e-5
    loss = (state * t["coef_final"].to(DEV)).sum()
    for tap in m["taps"]:
        r = mids[tap - 1]
        assert rel_err(r.detach().cpu(), t[f"rgb_tap{tap}"]) < 1e-5
        loss = loss + (r * t[f"coef_tap{tap}"].to(DEV)).sum()
    loss.backward()
    gtol = 2e-3 if name.startswith("trained_") else 1e-4
    got = dict(g_w1=model.w1.weight.grad.reshape(m["fc"], -1), g_b1=model.w1.bias.grad,
               g_w2=model.w2.weight.grad.reshape(m["C"], m["fc"]), g_b2=model.w2.bias.grad, g_x0=x0.grad)
    for k, v in got.items():
        assert rel_err(v.cpu(), t[k]) < gtol, (k, rel_err(v.cpu(), t[k]))


def test_no_grad_pingpong_matches_history():
    t, m = load_case("ec_c16_cpe_ms_circular")
    model = build_model(m, t)
    x0, masks = t["x0"].to(DEV), t["masks"].to(DEV)
    with torch.no_grad():
        a, _ = model.forward_nsteps(x0, m["T"], masks=masks)
        b, _, mids = model.forward_nsteps(x0, m["T"], masks=masks, return_middle_feature=True)
    assert torch.equal(a, b)
    assert rel_err(a.cpu(), t["final"]) < 1e-5
    assert rel_err(mids[0].cpu(), t["rgb_tap1"]) < 1e-5


def test_philox_mask_matches_restatement_and_drives_rollout():
    B, H, W, T, seed = 2, 20, 36, 3, 0x1234_5678_9ABC_DEF1
    got = Fn.philox_mask(B, H, W, 0.5, seed, T, t0=0).cpu().numpy()
    want = philox.fire_mask(seed, 0, T, B, H, W, 0.5)
    assert np.array_equal(got, want)
    got = Fn.philox_mask(B, H, W, 0.3, seed, T, t0=5, enc=True).cpu().numpy()
    assert np.array_equal(got, philox.fire_mask(seed, 5, T, B, H, W, 0.3, enc=True))
    # a rollout driven by the in-kernel generator == the same rollout with that mask supplied
    torch.manual_seed(0)
    model = nca_b200.DyNCA_EC(12, 3, fc_dim=96, padding_mode="circular", pos_emb="CPE", device=torch.device(DEV))
    with torch.no_grad():
        model.w2.weight.mul_(5.0)
    x0 = (torch.rand(B, 12, H, W, device=DEV) - 0.5)
    masks = Fn.philox_mask(B, H, W, 0.5, seed, T)
    with torch.no_grad():
        a, _ = model.forward_nsteps(x0, T, seed=seed)
        b, _ = model.forward_nsteps(x0, T, masks=masks)
    assert torch.equal(a, b)
    # gradients too (backward regenerates the mask)
    ga = torch.autograd.grad(model.forward_nsteps(x0, T, seed=seed)[0].square().sum(), model.w1.weight)[0]
    gb = torch.autograd.grad(model.forward_nsteps(x0, T, masks=masks)[0].square().sum(), model.w1.weight)[0]
    assert rel_err(ga.cpu(), gb.cpu()) < 1e-5


ORACLE_CASES = [
    # B, C, fc, H, W, T, pad, scales, cond
    (2, 12, 96, 40, 72, 3, "circular", (0,), "cpe"),
    (1, 13, 96, 37, 45, 3, "replicate", (0,), None),          # ragged: W not a multiple of 4, partial tiles
    (2, 16, 128, 36, 64, 3, "circular", (0, 1), "cpe"),
    (1, 16, 128, 22, 34, 2, "reflect", (0, 1), "cpe"),         # partial tiles in both axes, two scales
    (1, 12, 96, 10, 6, 2, "constant", (0, 1), "tensor"),
    (1, 8, 32, 4, 4, 2, "circular", (0, 1), None),             # tiny
    (1, 12, 96, 2, 66, 2, "replicate", (0, 1), "cpe"),         # one coarse row
    (1, 16, 128, 18, 40, 2, "replicate", (0, 1), "cpe"),       # 8x16 tiles: coarse ring leaves the image on a non-edge tile
    (2, 12, 96, 26, 48, 2, "circular", (0, 1), None),
    (1, 14, 64, 20, 24, 2, "reflect", (0,), "tensor"),
]


@pytest.mark.parametrize("precision", FP32_GRADE)
@pytest.mark.parametrize("case", ORACLE_CASES, ids=lambda c: "B%d_C%d_fc%d_%dx%d_T%d_%s_s%d_%s" % (c[0], c[1], c[2], c[3], c[4], c[5], c[6], len(c[7]), c[8]))
def test_against_oracle_seeded(case, precision):
    B, C, fc, H, W, T, pad, scales, cond = case
    g = torch.Generator().manual_seed(1234 + ORACLE_CASES.index(case))   # fixed inputs (hash() of a tuple with str is per-process)
    cc = {"cpe": 2, None: 0, "tensor": 3}[cond]
    P = 4 * C + cc
    w1 = torch.randn(fc, P, generator=g) * 0.15
    b1 = torch.randn(fc, generator=g) * 0.1
    w2 = torch.randn(C, fc, generator=g) * 0.1
    b2 = torch.randn(C, generator=g) * 0.02
    x0 = torch.rand(B, C, H, W, generator=g) - 0.5
    masks = (torch.rand(T, B, 1, H, W, generator=g) + 0.5).floor()
    cf = torch.randn(B, C, H, W, generator=g)
    ct = torch.randn(B, 3, H, W, generator=g)
    cond_t = O.cpe2d(B, H, W) if cond == "cpe" else (torch.randn(B, 3, H, W, generator=g) if cond == "tensor" else None)
    # oracle (CPU, fp32)
    po = [p.clone().requires_grad_(True) for p in (x0, w1, b1, w2, b2)]
    fo, hist = O.dynca_rollout(po[0], po[1], po[2], po[3], po[4], masks, scales, pad, cond_t, keep=True)
    lo = (fo * cf).sum() + (hist[1][:, :3] * 2.0 * ct).sum()
    lo.backward()
    # CUDA path through the functional API
    kind = {"cpe": _lib.NCA_COND_CPE, None: _lib.NCA_COND_NONE, "tensor": _lib.NCA_COND_TENSOR}[cond]
    cfg = Fn.DyncaConfig(C, fc, pad, scales, kind, cc, precision=precision)
    _expect_variant(cfg, B, H, W, precision)
    pg = [p.clone().to(DEV).requires_grad_(True) for p in (x0, w1, b1, w2, b2)]
    fg, taps = Fn.dynca_rollout(cfg, pg[0], pg[1], pg[2], pg[3], pg[4], T, 0.5,
                                cond=cond_t.to(DEV) if cond == "tensor" else None, masks=masks.to(DEV), return_taps=True)
    assert rel_err(fg.detach().cpu(), fo.detach()) < 1e-5
    z = Fn.dynca_perceive(cfg, pg[0].detach(), cond_t.to(DEV) if cond == "tensor" else None)
    assert rel_err(z.cpu(), O.perceive_multiscale(x0, scales, pad, cond_t)) < 1e-5
    lg = (fg * cf.to(DEV)).sum() + (taps[0] * ct.to(DEV)).sum()
    lg.backward()
    for a, b, n in zip(pg, po, ("x0", "w1", "b1", "w2", "b2")):
        assert rel_err(a.grad.cpu(), b.grad) < 1e-4, (n, rel_err(a.grad.cpu(), b.grad))


def test_linearity_of_backward_at_full_size():
    """Size-independent property at config-2 scale (256x256, C=16, fc=128, two scales): BPTT is linear in the
    incoming gradient, and with zero incoming gradient every output gradient is exactly zero."""
    torch.manual_seed(1)
    B, C, fc, H, W, T = 2, 16, 128, 256, 256, 4
    model = nca_b200.DyNCA_EC(C, 3, fc_dim=fc, padding_mode="circular", pos_emb="CPE", perception_scales=[0, 1],
                              device=torch.device(DEV))
    with torch.no_grad():
        model.w2.weight.mul_(3.0)
    x0 = (torch.rand(B, C, H, W, device=DEV) - 0.5).requires_grad_(True)
    seed = 99
    g1 = torch.randn(B, C, H, W, device=DEV)
    g2 = torch.randn(B, C, H, W, device=DEV)

    def grads(g):
        s, _ = model.forward_nsteps(x0, T, seed=seed)
        return torch.autograd.grad((s * g).sum(), [x0, model.w1.weight, model.w2.bias])

    a, b, ab, z = grads(g1), grads(g2), grads(g1 + 2.0 * g2), grads(torch.zeros_like(g1))
    for i in range(3):
        assert rel_err(ab[i].cpu(), (a[i] + 2.0 * b[i]).cpu()) < 1e-4
        assert float(z[i].abs().max()) == 0.0
    # and the state itself: a cell that never fires keeps its value
    masks = torch.zeros(T, B, 1, H, W, device=DEV)
    with torch.no_grad():
        s, _ = model.forward_nsteps(x0.detach(), T, masks=masks)
    assert torch.equal(s, x0.detach())
