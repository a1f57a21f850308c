"""Parity of the ConditionedNCA CUDA path (csrc/enc_f32.cu) with the reference (golden vectors made by the unmodified
EncoderConditioning/nca.py) and with the oracle on seeded inputs.  Tolerances (BASELINE.json north_star): state 1e-5
relative, gradients 1e-4 relative (fp32)."""
import pytest
import torch

import nca_b200
from nca_b200 import functional as Fn
from oracle import nca_oracle as O
from helpers import ENC_CASES, load_case, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
STATE_TOL, GRAD_TOL = 1e-5, 1e-4
NAMES = ("wp", "wa", "ba", "wb", "bb", "wc")


def _cuda_run(t, m, T, grads=True):
    cfg = Fn.EncConfig(m["C"], m["living_dim"], m["thr"], m["rate"])
    ps = [t[k].to(DEV).requires_grad_(grads) for k in NAMES]
    x0 = t["x0"].to(DEV).requires_grad_(grads)
    goal = t["goal_enc"].to(DEV).requires_grad_(grads)
    final = Fn.enc_rollout(cfg, x0, goal, *ps, T, masks=t["fires"][:T].to(DEV))
    return final, ps, x0, goal


@pytest.mark.parametrize("name", ENC_CASES)
def test_enc_golden_rollout_and_gradients(name):
    t, m = load_case(name)
    final, ps, x0, goal = _cuda_run(t, m, m["T"])
    assert rel_err(final.detach().cpu(), t["final"]) < STATE_TOL
    (final * t["coef_final"].to(DEV)).sum().backward()
    for p, k in zip(ps + [x0, goal], ["g_" + n for n in NAMES] + ["g_x0", "g_goal"]):
        assert rel_err(p.grad.cpu(), t[k]) < GRAD_TOL, k


@pytest.mark.parametrize("name", ENC_CASES)
def test_enc_single_steps_match_oracle(name):
    t, m = load_case(name)
    with torch.no_grad():
        x = t["x0"]
        for s in range(min(3, m["T"])):
            want = O.enc_step(x, t["goal_enc"], *[t[k] for k in NAMES], t["fires"][s], m["living_dim"], m["thr"])
            cfg = Fn.EncConfig(m["C"], m["living_dim"], m["thr"], m["rate"])
            got = Fn.enc_rollout(cfg, x.to(DEV), t["goal_enc"].to(DEV), *[t[k].to(DEV) for k in NAMES], 1,
                                 masks=t["fires"][s:s + 1].to(DEV))
            assert rel_err(got.cpu(), want) < STATE_TOL
            x = want


CASES = [  # B, C, H, W, T, living_dim
    (2, 20, 9, 37, 4, 3),       # ragged in both directions, two tiles in x
    (1, 20, 64, 64, 3, 3),      # c4 frame size
    (3, 12, 5, 70, 3, 3),       # fewer channels, three tiles in x
    (1, 20, 16, 16, 3, -1),     # use_living_channel=False
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d_C%d_%dx%d_T%d_l%d" % c)
def test_enc_against_oracle_seeded(case):
    B, C, H, W, T, liv = case
    g = torch.Generator().manual_seed(31)
    K = 3 * C
    wp = torch.randn(K, 1, 3, 3, generator=g) * 0.3
    wa = torch.randn(64, K, generator=g) * 0.2
    ba = torch.randn(64, generator=g) * 0.1
    wb = torch.randn(64, 64, generator=g) * 0.15
    bb = torch.randn(64, generator=g) * 0.1
    wc = torch.randn(C, 64, generator=g) * 0.1
    x0 = torch.randn(B, C, H, W, generator=g) * 0.4
    if liv >= 0:
        x0[:, liv] = torch.rand(B, H, W, generator=g) * 0.5 - 0.15
    x0[:, 0] *= 30.0                                              # some cells hit the +-10 clamp
    goal = torch.rand(B, C, H, W, generator=g)
    fires = (torch.rand(T, B, 1, H, W, generator=g) < 0.5).float()
    coef = torch.randn(B, C, H, W, generator=g)
    po = [p.clone().requires_grad_(True) for p in (x0, goal, wp, wa, ba, wb, bb, wc)]
    if liv >= 0:
        fo = O.enc_rollout(*po, fires, liv, 0.1)
    else:   # no living channel: alive == all ones
        x = po[0]
        for s in range(T):
            p = O.enc_perception_fast(x + po[1], po[2])
            h1 = torch.relu(torch.einsum("jk,bkhw->bjhw", po[3], p) + po[4][None, :, None, None])
            h2 = torch.relu(torch.einsum("jk,bkhw->bjhw", po[5], h1) + po[6][None, :, None, None])
            x = torch.clamp(x + fires[s] * torch.einsum("cj,bjhw->bchw", po[7], h2), -10.0, 10.0)
        fo = x
    (fo * coef).sum().backward()
    cfg = Fn.EncConfig(C, liv, 0.1, 0.5)
    pg = [p.clone().to(DEV).requires_grad_(True) for p in (x0, goal, wp, wa, ba, wb, bb, wc)]
    fg = Fn.enc_rollout(cfg, *pg, T, masks=fires.to(DEV))
    (fg * coef.to(DEV)).sum().backward()
    assert rel_err(fg.detach().cpu(), fo.detach()) < STATE_TOL
    for a, b, n in zip(pg, po, ("x0", "goal") + NAMES):
        assert rel_err(a.grad.cpu(), b.grad) < GRAD_TOL, n


def test_enc_module_grow_with_encoder_gradient():
    """drop-in module: grow() = encoder (PyTorch) + CUDA rollout; gradients reach the encoder; Philox == supplied"""
    torch.manual_seed(5)
    H = W = 32
    nca = nca_b200.ConditionedNCA(target_shape=(3, H, W), num_hidden_channels=16, living_channel_dim=3).to(DEV)
    sd = nca.state_dict()
    assert tuple(sd["perception_net.weight"].shape) == (60, 1, 3, 3)
    assert tuple(sd["update_net.out.0.weight"].shape) == (64, 60, 1, 1) and tuple(sd["update_net.out.4.weight"].shape) == (20, 64, 1, 1)
    assert "update_net.out.4.bias" not in sd and "encoder.embed.0.weight" in sd
    x = nca.generate_seed(4).to(DEV) + 0.05 * torch.randn(4, 20, H, W, device=DEV)
    goal = torch.rand(4, 3, H, W, device=DEV)
    T = 6
    out = nca.grow(x, T, goal, seed=11)
    masks = Fn.philox_mask(4, H, W, 0.5, 11, T, enc=True)
    out2 = nca.grow(x, T, goal, masks=masks)
    assert torch.equal(out, out2)
    out.square().mean().backward()
    for n, p in nca.named_parameters():
        if p.requires_grad:
            assert p.grad is not None and torch.isfinite(p.grad).all(), n
    assert float(nca.encoder.embed[0].weight.grad.abs().max()) > 0
    # oracle on the same masks, encoder run on CPU
    cpu = nca_b200.ConditionedNCA(target_shape=(3, H, W), num_hidden_channels=16, living_channel_dim=3)
    cpu.load_state_dict({k: v.cpu() for k, v in nca.state_dict().items()})
    ge = cpu._pad_goal(cpu.encoder(goal.cpu())).detach()
    wp, *mlp = [w.detach() for w in cpu._w()]
    mlp = [w.reshape(w.shape[0], -1) if w.dim() == 4 else w for w in mlp]
    want = O.enc_rollout(x.cpu(), ge, wp, *mlp, masks.cpu(), 3, 0.1)
    assert rel_err(out.detach().cpu(), want) < 5e-5
    with torch.no_grad():
        y, _ = nca((x, ge.to(DEV)), masks=masks[:1])
    assert rel_err(y.cpu(), O.enc_step(x.cpu(), ge, wp, *mlp, masks[0].cpu(), 3, 0.1)) < STATE_TOL
    assert nca.alive(out).dtype == torch.bool


# ---- tcgen05 forward path (NCA_PREC_BF16): update MLP with bf16 operands, fp32 accumulate -----------------------
BF16_STEP_TOL = 1e-2


def _life_agree(got, want):
    """cells whose alive / clamp decision agrees between the two results (a bf16-sized change of the living channel
    next to the 0.1 threshold may flip a cell; such cells are excluded and must be rare)"""
    dead_g = got.abs().sum(1, keepdim=True) == 0
    dead_w = want.abs().sum(1, keepdim=True) == 0
    return dead_g == dead_w


@pytest.mark.parametrize("name", ENC_CASES)
def test_enc_bf16_single_steps(name):
    t, m = load_case(name)
    cfg = Fn.EncConfig(m["C"], m["living_dim"], m["thr"], m["rate"], precision="bf16")
    ws = [t[k].to(DEV) for k in NAMES]
    with torch.no_grad():
        x = t["x0"]
        for s in range(min(3, m["T"])):
            want = O.enc_step(x, t["goal_enc"], *[t[k] for k in NAMES], t["fires"][s], m["living_dim"], m["thr"])
            got = Fn.enc_rollout(cfg, x.to(DEV), t["goal_enc"].to(DEV), *ws, 1, masks=t["fires"][s:s + 1].to(DEV)).cpu()
            ok = _life_agree(got, want)
            assert float(ok.float().mean()) > 0.99
            upd_g, upd_w = (got - x) * ok, (want - x) * ok
            assert rel_err(upd_g, upd_w) < BF16_STEP_TOL
            x = want


def test_enc_bf16_c4_frame_properties():
    """c4 frame size, batch 32: never-firing cells with a full-alive state keep their value bit-exactly; Philox ==
    supplied mask; bf16 update close to the fp32 kernels; gradients flow (fp32 BPTT kernels)"""
    torch.manual_seed(2)
    B, H = 32, 64
    nb = nca_b200.ConditionedNCA(target_shape=(3, H, H), num_hidden_channels=16, living_channel_dim=3, precision="bf16").to(DEV)
    nf = nca_b200.ConditionedNCA(target_shape=(3, H, H), num_hidden_channels=16, living_channel_dim=3).to(DEV)
    nf.load_state_dict(nb.state_dict())
    x = 0.3 * torch.randn(B, 20, H, H, device=DEV)
    x[:, 3] = 0.5 + 0.2 * torch.rand(B, H, H, device=DEV)          # everything alive, far from the threshold
    goal = torch.rand(B, 3, H, H, device=DEV)
    with torch.no_grad():
        ge = nb._pad_goal(nb.encoder(goal))
        cfgb, cfgf = nb._cfg(), nf._cfg()
        w = [p.detach() for p in nb._w()]
        same = Fn.enc_rollout(cfgb, x, ge, *w, 2, masks=torch.zeros(2, B, 1, H, H, device=DEV))
        assert torch.equal(same, x)
        a = Fn.enc_rollout(cfgb, x, ge, *w, 3, seed=21)
        b = Fn.enc_rollout(cfgb, x, ge, *w, 3, masks=Fn.philox_mask(B, H, H, 0.5, 21, 3, enc=True))
        assert torch.equal(a, b)
        f = Fn.enc_rollout(cfgf, x, ge, *w, 1, seed=21)
        a1 = Fn.enc_rollout(cfgb, x, ge, *w, 1, seed=21)
        assert rel_err((a1 - x).cpu(), (f - x).cpu()) < BF16_STEP_TOL
    out = nb.grow(x, 4, goal, seed=5)
    out.square().mean().backward()
    assert all(torch.isfinite(p.grad).all() for p in nb.parameters() if p.requires_grad)


def _enc_bf16_vs_emu(t, m, T):
    cfg = Fn.EncConfig(m["C"], m["living_dim"], m["thr"], m["rate"], precision="bf16")
    ps = [t[k].to(DEV).requires_grad_(True) for k in NAMES]
    x0 = t["x0"].to(DEV).requires_grad_(True)
    goal = t["goal_enc"].to(DEV).requires_grad_(True)
    final = Fn.enc_rollout(cfg, x0, goal, *ps, T, masks=t["fires"][:T].to(DEV))
    (final * t["coef_final"].to(DEV)).sum().backward()
    fe, ge = O.enc_bf16emu_rollout_grads(t["x0"], t["goal_enc"], *[t[k] for k in NAMES], t["fires"][:T], t["coef_final"],
                                         m["living_dim"], m["thr"])
    errs = {"state": rel_err(final.detach().cpu(), fe)}
    for p, k in zip(ps + [x0, goal], list(NAMES) + ["x0", "goal"]):
        errs[k] = float((p.grad.cpu() - ge[k]).norm() / (ge[k].norm() + 1e-30))
    return errs


@pytest.mark.parametrize("name", ENC_CASES)
def test_enc_bf16_bptt_vs_emulated_oracle(name):
    """tcgen05 forward + BPTT against the oracle that rounds the GEMM operands to bf16 at the same points"""
    t, m = load_case(name)
    errs = _enc_bf16_vs_emu(t, m, min(m["T"], 4))
    print(name, {k: "%.1e" % v for k, v in errs.items()})
    assert errs["state"] < 3e-3, errs
    assert max(v for k, v in errs.items() if k != "state") < 1e-2, errs


TC_CASES = [  # B, C, H, W, T, living_dim  (W % 4 == 0: tcgen05 kernels)
    (2, 20, 9, 36, 3, 3),        # ragged: partial tiles in both directions
    (1, 20, 20, 24, 3, 3),
    (1, 13, 16, 16, 2, 3),       # odd channel count: half-empty perception chunk
    (1, 20, 16, 32, 2, -1),      # use_living_channel=False
]


@pytest.mark.parametrize("case", TC_CASES, ids=lambda c: "B%d_C%d_%dx%d_T%d_l%d" % c)
def test_enc_bf16_ragged_shapes_vs_emulated_oracle(case):
    B, C, H, W, T, liv = case
    g = torch.Generator().manual_seed(41)
    K = 3 * C
    t = dict(wp=torch.randn(K, 1, 3, 3, generator=g) * 0.3, wa=torch.randn(64, K, generator=g) * 0.2,
             ba=torch.randn(64, generator=g) * 0.1, wb=torch.randn(64, 64, generator=g) * 0.15,
             bb=torch.randn(64, generator=g) * 0.1, wc=torch.randn(C, 64, generator=g) * 0.1)
    x0 = torch.randn(B, C, H, W, generator=g) * 0.4
    if liv >= 0:
        x0[:, liv] = torch.rand(B, H, W, generator=g) * 0.5 - 0.15
    x0[:, 0] *= 30.0                                              # some cells hit the +-10 clamp
    t.update(x0=x0, goal_enc=torch.rand(B, C, H, W, generator=g), fires=(torch.rand(T, B, 1, H, W, generator=g) < 0.5).float(),
             coef_final=torch.randn(B, C, H, W, generator=g))
    m = dict(C=C, living_dim=liv, thr=0.1, rate=0.5)
    errs = _enc_bf16_vs_emu(t, m, T)
    assert errs["state"] < 3e-3, errs
    assert max(v for k, v in errs.items() if k != "state") < 1e-2, errs
