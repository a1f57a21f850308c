"""Parity of every precision mode of the CUDA path against the CPU oracle at the BASELINE.json shapes (batch reduced so
that the CPU side finishes in seconds), over long rollouts, and over a training run.

Yardsticks (all CPU, all independent of the kernels):
  * ``O.dynca_rollout_aten``      - the fp32 restatement of dynca.py:71-128 (pinned to the reference goldens);
  * ``O.dynca_rollout_bf16ops``   - "the MLP runs in BF16" stated from the math: GEMM operands rounded to bf16, fp32
                                     accumulation, autograd with straight-through rounding (oracle/nca_oracle.py).

Tolerances (DESIGN.md section 4 states and justifies them):
  fp32 / f16x3 : per-step and 4-step state 1e-5 (max-abs / max-abs) vs the fp32 oracle                         [north_star]
                 gradients 1e-3 in L2, isolated relu flips below 5e-3 of the maximum (see GRAD_RMS_TOL_FP32_FULL below; the
                 1e-4 max-relative bar is enforced on the goldens and small seeded shapes in test_dynca_gpu.py)
  bf16         : per-step update 1e-2 vs the fp32 oracle                                                       [north_star]
                 gradients: within GRAD_RMS_VS_BF16OPS (rms) of the bf16-operand oracle, and no further from the fp32 oracle
                 than BF16_EXCESS x the bf16-operand oracle itself is (rounding the MLP operands to bf16 is what moves the
                 gradients; the kernels add nothing of their own beyond that)
  rollouts     : state error vs the fp32 oracle over T = 128 steps below a stated curve, and below 2x the bf16-operand
                 oracle's own curve
  loss curves  : 40 optimiser iterations (MSE + overflow loss, per-parameter gradient normalisation, Adam - the reference's
                 training step, experiments.py:236-255) with the CUDA path in each precision against the same loop on the CPU
                 oracle.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import nca_b200
from nca_b200 import functional as Fn, _lib
from oracle import nca_oracle as O
from helpers import load_case

pytestmark = pytest.mark.gpu
DEV = "cuda"

STATE_TOL_FP32 = 1e-5
# gradients at these sizes: 3e7 relu units per rollout, so a few pre-activations sit within fp32 summation noise of zero and
# their relu derivative differs between ANY two fp32 evaluations - the two CPU phrasings of the oracle (dynca_step vs
# dynca_step_aten) differ by 5.9e-4 (dL/dx0), 2.5e-4 (dL/dw1) max-relative at the config-2 shape.  The 1e-4 max-relative bar
# of north_star is therefore enforced where it is meaningful (tests/test_dynca_gpu.py: the reference goldens and seeded
# small shapes) and the full shapes use an L2 bar plus a bound on isolated flips.
GRAD_RMS_TOL_FP32_FULL, GRAD_MAX_TOL_FP32_FULL = 1e-3, 5e-3
STEP_TOL_BF16 = 1e-2
ROLLOUT4_TOL_BF16 = 1e-2
GRAD_RMS_VS_BF16OPS = {1: 1e-2, 2: 4e-2}       # by number of perception scales (two scales: the coarse operand is rounded separately)
BF16_EXCESS = 1.5


def _rel_max(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _rel_rms(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _ref_init(C, fc, cc, seed=0):
    """the reference's initialisation (dynca.py:56-61): xavier_normal gains 0.2 / 0.1, default conv bias, zero w2 bias"""
    torch.manual_seed(seed)
    w1 = torch.nn.Conv2d(4 * C + cc, fc, 1)
    torch.nn.init.xavier_normal_(w1.weight, gain=0.2)
    w2 = torch.nn.Conv2d(fc, C, 1)
    torch.nn.init.xavier_normal_(w2.weight, gain=0.1)
    torch.nn.init.zeros_(w2.bias)
    return [w1.weight.detach()[:, :, 0, 0].clone(), w1.bias.detach().clone(), w2.weight.detach()[:, :, 0, 0].clone(), w2.bias.detach().clone()]


# name -> B, C, fc, H, W, T, pad, scales, cond  (BASELINE.json configs 1, 2, 3, 5 with the batch reduced)
SHAPES = {
    "c1_128x128_C12_fc96_cpe_B4": (4, 12, 96, 128, 128, 4, "replicate", (0,), "cpe"),
    "c2_256x256_C16_fc128_cpe_ms_B1": (1, 16, 128, 256, 256, 4, "replicate", (0, 1), "cpe"),
    "c3_256x256_C12_fc96_edges_B1": (1, 12, 96, 256, 256, 4, "circular", (0,), "edges"),
    "c5_1080x1920_C13_fc96_B1_one_step": (1, 13, 96, 1080, 1920, 1, "circular", (0,), None),
}


def _inputs(name):
    B, C, fc, H, W, T, pad, scales, cond = SHAPES[name]
    cc = {"cpe": 2, "edges": 3, None: 0}[cond]
    w = _ref_init(C, fc, cc, seed=0)
    g = torch.Generator().manual_seed(42)
    x0 = torch.rand(B, C, H, W, generator=g) - 0.5
    masks = (torch.rand(T, B, 1, H, W, generator=torch.Generator().manual_seed(424)) + 0.5).floor()
    cf = torch.randn(B, C, H, W, generator=g)
    cond_t = None
    if cond == "cpe":
        cond_t = O.cpe2d(B, H, W)
    elif cond == "edges":
        cond_t = O.edge_extract(torch.rand(B, 1, H, W, generator=g) * 2 - 1, "None")
    return x0, w, masks, cf, cond_t


def _cpu(name, step_fn, grads):
    B, C, fc, H, W, T, pad, scales, cond = SHAPES[name]
    x0, w, masks, cf, cond_t = _inputs(name)
    ps = [p.clone().requires_grad_(grads) for p in [x0] + w]
    x = ps[0]
    one = None
    with torch.set_grad_enabled(grads):
        for t in range(T):
            x = step_fn(x, ps[1], ps[2], ps[3], ps[4], masks[t], scales, pad, cond_t)
            if t == 0:
                one = x.detach().clone()
        if grads:
            (x * cf).sum().backward()
    return one, x.detach(), [p.grad for p in ps] if grads else None


def _gpu(name, precision, grads):
    B, C, fc, H, W, T, pad, scales, cond = SHAPES[name]
    x0, w, masks, cf, cond_t = _inputs(name)
    kind, cc = {"cpe": (_lib.NCA_COND_CPE, 2), "edges": (_lib.NCA_COND_TENSOR, 3), None: (_lib.NCA_COND_NONE, 0)}[cond]
    cfg = Fn.DyncaConfig(C, fc, pad, scales, kind, cc, precision=precision)
    ps = [p.clone().to(DEV).requires_grad_(grads) for p in [x0] + w]
    cd = cond_t.to(DEV) if cond == "edges" else None
    m = masks.to(DEV)
    with torch.no_grad():
        one, _ = Fn.dynca_rollout(cfg, ps[0].detach(), *[p.detach() for p in ps[1:]], 1, 0.5, cond=cd, masks=m[:1])
    with torch.set_grad_enabled(grads):
        fin, _ = Fn.dynca_rollout(cfg, *ps, T, 0.5, cond=cd, masks=m)
        if grads:
            (fin * cf.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    return one.cpu(), fin.detach().cpu(), [p.grad.cpu() for p in ps] if grads else None, cfg


_CPU_CACHE = {}


def _cpu_cached(name, which, grads):
    key = (name, which, grads)
    if key not in _CPU_CACHE:
        fn = O.dynca_step_aten if which == "fp32" else (lambda *a: O.dynca_step_bf16ops(*a, fast=True))
        _CPU_CACHE[key] = _cpu(name, fn, grads)
    return _CPU_CACHE[key]


@pytest.mark.parametrize("precision", ["fp32", "f16x3", "bf16"])
@pytest.mark.parametrize("name", list(SHAPES))
def test_baseline_shape_against_cpu_oracle(name, precision, record_property):
    grads = SHAPES[name][5] > 1          # the 1080p case is the inference shape: one no-grad step
    x0 = _inputs(name)[0]
    one_o, fin_o, g_o = _cpu_cached(name, "fp32", grads)
    one, fin, g, cfg = _gpu(name, precision, grads)
    B, C, fc, H, W = SHAPES[name][:5]
    want_variant = {"fp32": 0, "f16x3": 3, "bf16": 2}[precision]
    assert Fn.dynca_kernel_variant(cfg, B, H, W) == want_variant
    e_step = _rel_max(one - x0, one_o - x0)
    e_fin = _rel_max(fin, fin_o)
    record_property("update_err_step1", e_step)
    record_property("state_err_final", e_fin)
    print(f"\n[{name} {precision}] step-1 update err {e_step:.2e}, final state err {e_fin:.2e}")
    if precision != "bf16":
        assert e_step < 10 * STATE_TOL_FP32      # the update is ~10x smaller than the state it is added to
        assert _rel_max(one, one_o) < STATE_TOL_FP32 and e_fin < STATE_TOL_FP32
        if grads:
            # 3e7 relu units per rollout at these sizes: a pre-activation within fp32 summation noise of zero flips its relu
            # derivative between any two fp32 evaluations (CPU vs CPU included) and moves one 3x3 patch of dL/dx0 by ~1e-3 of the
            # maximum, so the gradient bar is applied in L2 and the maximum is bounded separately
            for a, b, n in zip(g, g_o, ("x0", "w1", "b1", "w2", "b2")):
                e, em = _rel_rms(a, b), _rel_max(a, b)
                print(f"    grad {n}: rms {e:.2e} max {em:.2e}")
                assert e < GRAD_RMS_TOL_FP32_FULL and em < GRAD_MAX_TOL_FP32_FULL, (n, e, em)
        return
    assert e_step < STEP_TOL_BF16
    assert e_fin < ROLLOUT4_TOL_BF16
    if grads:
        _, _, g_q = _cpu_cached(name, "bf16ops", True)
        for a, bq, bo, n in zip(g, g_q, g_o, ("x0", "w1", "b1", "w2", "b2")):
            e_q, e_o, q_o = _rel_rms(a, bq), _rel_rms(a, bo), _rel_rms(bq, bo)
            print(f"    grad {n}: rms vs bf16-operand oracle {e_q:.2e}, vs fp32 oracle {e_o:.2e} (bf16-operand oracle vs fp32 oracle {q_o:.2e})")
            assert e_q < GRAD_RMS_VS_BF16OPS[len(SHAPES[name][7])], (n, e_q)
            assert e_o < BF16_EXCESS * q_o + 2e-3, (n, e_o, q_o)


# ---- ConditionedNCA at the config-4 shape (batch 4) ---------------------------------------------------------------------
def test_c4_conditioned_nca_against_cpu_oracle():
    torch.manual_seed(0)
    B, C, H, T = 4, 20, 64, 4
    g = torch.Generator().manual_seed(5)
    wp = torch.randn(60, 1, 3, 3, generator=g) * 0.3
    wa = torch.randn(64, 60, generator=g) * 0.1
    ba = torch.randn(64, generator=g) * 0.05
    wb = torch.randn(64, 64, generator=g) * 0.1
    bb = torch.randn(64, generator=g) * 0.05
    wc = torch.randn(20, 64, generator=g) * 0.05
    x0 = torch.zeros(B, C, H, H)
    x0[:, 3:, H // 2 - 3:H // 2 + 3, H // 2 - 3:H // 2 + 3] = 1.0
    x0 = x0 + 0.2 * torch.rand(B, C, H, H, generator=g) * (x0[:, 3:4] > 0)
    goal = torch.zeros(B, C, H, H)
    goal[:, 4:] = torch.randn(B, 16, H, H, generator=g) * 0.3
    fires = (torch.rand(T, B, 1, H, H, generator=g) < 0.5).float()
    cf = torch.randn(B, C, H, H, generator=g)
    ws = [wp, wa, ba, wb, bb, wc]
    po = [p.clone().requires_grad_(True) for p in [x0, goal] + ws]
    fo = O.enc_rollout(po[0], po[1], *po[2:], fires)
    (fo * cf).sum().backward()
    pq = [p.clone().requires_grad_(True) for p in [x0, goal] + ws]
    fq = O.enc_rollout_bf16ops(pq[0], pq[1], *pq[2:], fires)
    (fq * cf).sum().backward()
    for precision, st, gt in (("fp32", 1e-5, 1e-4), ("bf16", 3e-2, None)):
        cfg = Fn.EncConfig(C, 3, precision=precision)
        pg = [p.clone().to(DEV).requires_grad_(True) for p in [x0, goal] + ws]
        fg = Fn.enc_rollout(cfg, *pg, T, masks=fires.to(DEV))
        (fg * cf.to(DEV)).sum().backward()
        e = _rel_max(fg.detach().cpu(), fo.detach())
        print(f"\n[c4 {precision}] final state err {e:.2e}")
        assert e < st
        for a, b, bq, n in zip(pg, po, pq, ("x0", "goal", "wp", "wa", "ba", "wb", "bb", "wc")):
            em, er = _rel_max(a.grad.cpu(), b.grad), _rel_rms(a.grad.cpu(), b.grad)
            if gt is not None:
                print(f"    grad {n}: max {em:.2e} rms {er:.2e}")
                assert em < gt, (n, em)
            else:       # bf16 operands: close to the bf16-operand oracle, and no further from the fp32 oracle than that model is
                e_q, q_o = _rel_rms(a.grad.cpu(), bq.grad), _rel_rms(bq.grad, b.grad)
                print(f"    grad {n}: rms vs bf16-operand oracle {e_q:.2e}, vs fp32 oracle {er:.2e} (bf16-operand oracle vs fp32 oracle {q_o:.2e})")
                assert e_q < 1e-2, (n, e_q)
                assert er < BF16_EXCESS * q_o + 2e-3, (n, er, q_o)


# ---- long rollouts: state error against the fp32 oracle as a function of t ------------------------------------------------
CHECKPOINTS = (1, 2, 4, 8, 16, 32, 64, 96, 128)
# stated curve for the bf16 path: relative max error <= BF16_CURVE_A * sqrt(t) (capped), measured ~3x below it
BF16_CURVE_A, BF16_CURVE_CAP = 1.5e-3, 3e-2
BF16_CURVE_A_TRAINED = 8e-3              # trained weights are ~4x larger than the initialisation: so is the per-step update
FP32_CURVE_A = 2e-6                      # fp32 / f16x3: summation-order noise amplified by the dynamics


def _curve_cpu(step_fn, x0, w, masks, scales, pad, cond_t):
    out, x = {}, x0
    with torch.no_grad():
        for t in range(masks.shape[0]):
            x = step_fn(x, *w, masks[t], scales, pad, cond_t)
            if t + 1 in CHECKPOINTS:
                out[t + 1] = x.clone()
    return out


def _curve_gpu(cfg, x0, w, masks, cond_dev):
    out, x, t0 = {}, x0.to(DEV), 0
    wd = [p.to(DEV) for p in w]
    m = masks.to(DEV)
    with torch.no_grad():
        for tc in CHECKPOINTS:
            if tc > masks.shape[0]:
                break
            x, _ = Fn.dynca_rollout(cfg, x, *wd, tc - t0, 0.5, cond=cond_dev, masks=m[t0:tc])
            out[tc] = x.cpu()
            t0 = tc
    return out


def _check_curves(tag, ref, q, runs, bf16_a):
    bad = []
    for t in sorted(ref):
        eq = _rel_max(q[t], ref[t])
        line = f"  t={t:3d} |x|max {float(ref[t].abs().max()):7.3f}  bf16-operand oracle {eq:.2e}"
        for prec, cur in runs.items():
            e = _rel_max(cur[t], ref[t])
            line += f"  {prec} {e:.2e}"
            if prec == "bf16":
                if not e < min(bf16_a * math.sqrt(t), BF16_CURVE_CAP):
                    bad.append((tag, prec, t, e, "stated curve"))
                if not e < 2.0 * eq + 5e-4:
                    bad.append((tag, prec, t, e, eq, "2x the bf16-operand oracle"))
            elif not e < FP32_CURVE_A * t + 1e-5:
                bad.append((tag, prec, t, e, "fp32-grade curve"))
        print(line)
    assert not bad, bad


def test_c2_rollout_divergence_curve_T128():
    """config 2 (256x256, C=16, fc=128, two scales, CPE, replicate padding), one sample, reference initialisation, T = 128
    with the fire masks supplied: state error of every precision against the fp32 oracle at t = 1 .. 128"""
    B, C, fc, H, W, T = 1, 16, 128, 256, 256, 128
    w = _ref_init(C, fc, 2, seed=0)
    x0 = torch.rand(B, C, H, W, generator=torch.Generator().manual_seed(42)) - 0.5
    masks = (torch.rand(T, B, 1, H, W, generator=torch.Generator().manual_seed(424)) + 0.5).floor()
    cpe = O.cpe2d(B, H, W)
    ref = _curve_cpu(O.dynca_step_aten, x0, w, masks, (0, 1), "replicate", cpe)
    q = _curve_cpu(lambda *a: O.dynca_step_bf16ops(*a, fast=True), x0, w, masks, (0, 1), "replicate", cpe)
    runs = {}
    for prec in ("bf16", "f16x3", "fp32"):
        cfg = Fn.DyncaConfig(C, fc, "replicate", (0, 1), _lib.NCA_COND_CPE, 2, precision=prec)
        runs[prec] = _curve_gpu(cfg, x0, w, masks, None)
    print("\n[c2 T=128 divergence, relative max error vs the fp32 oracle]")
    _check_curves("c2", ref, q, runs, BF16_CURVE_A)


def test_trained_model_rollout_divergence_curve_T128():
    """trained weights of the reference's own demo model (docs/data/vec_field_models/large: CD flavour, edge conditioning) at
    128x128: the regime the step is used in (bounded states over hundreds of steps)"""
    t, m = load_case("trained_starry_edges")
    B, C, fc, H, W, T = 2, m["C"], m["fc"], 128, 128, 128
    w = [t["w1"], t["b1"], t["w2"], t["b2"]]
    g = torch.Generator().manual_seed(7)
    x0 = torch.rand(B, C, H, W, generator=g) - 0.5
    masks = (torch.rand(T, B, 1, H, W, generator=g) + 0.5).floor()
    img = F.interpolate(t["cond_img"], size=(H, W), mode="bilinear", align_corners=False).repeat(B, 1, 1, 1)
    cond_t = O.edge_extract(img, m["edge_transform"])
    ref = _curve_cpu(O.dynca_step_aten, x0, w, masks, tuple(m["scales"]), m["pad"], cond_t)
    q = _curve_cpu(lambda *a: O.dynca_step_bf16ops(*a, fast=True), x0, w, masks, tuple(m["scales"]), m["pad"], cond_t)
    runs = {}
    for prec in ("bf16", "f16x3", "fp32"):
        cfg = Fn.DyncaConfig(C, fc, m["pad"], m["scales"], _lib.NCA_COND_TENSOR, 3, precision=prec)
        runs[prec] = _curve_gpu(cfg, x0, w, masks, cond_t.to(DEV))
    print("\n[trained CD model T=128 divergence, relative max error vs the fp32 oracle]")
    _check_curves("trained", ref, q, runs, BF16_CURVE_A_TRAINED)


# ---- loss curves -----------------------------------------------------------------------------------------------------------
LOSS_ITERS = 40
# relative deviation of the loss from the CPU-oracle run, at every one of the 40 iterations.  The optimiser (per-parameter gradient
# normalisation + Adam) turns any gradient difference into a parameter difference of the same relative size at every step, so
# the curves drift apart as the run goes on: measured maxima 7e-4 (fp32), 1e-3 (f16x3), 3e-2 .. 5e-2 (bf16; the bf16 BPTT is not
# bit-reproducible run to run), all reached in the last ten iterations where the loss has fallen 60x.
LOSS_TOL = {"fp32": 5e-3, "f16x3": 5e-3, "bf16": 1e-1}


def _train_setup():
    B, C, fc, H, W = 4, 12, 96, 32, 32
    g = torch.Generator().manual_seed(3)
    w = _ref_init(C, fc, 2, seed=1)
    seed_state = torch.rand(1, C, H, W, generator=g) - 0.5
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    target = torch.stack([0.5 + 0.4 * torch.sin(3 * xx), 0.5 + 0.4 * torch.cos(2 * yy), 0.5 + 0.3 * xx * yy])[None].repeat(B, 1, 1, 1)
    Ts = [int(v) for v in torch.randint(8, 17, (LOSS_ITERS,), generator=g)]
    masks = [(torch.rand(T, B, 1, H, W, generator=g) + 0.5).floor() for T in Ts]
    return B, C, fc, H, W, w, seed_state.repeat(B, 1, 1, 1), target, Ts, masks


def _loss_terms(state, target):
    rgb = state[:, :3] * 2.0                                       # DyNCA.to_rgb (dynca.py:130-131)
    if state.is_cuda:
        overflow = nca_b200.overflow_loss(state)                   # the product's fused kernel
    else:
        overflow = (state - state.clamp(-1.0, 1.0)).abs().mean()   # utils/loss/loss.py:33-36
    return F.mse_loss(rgb, target) + overflow


def _train(rollout, w, x0, target, Ts, masks, device, opt_factory):
    ps = [p.clone().to(device).requires_grad_(True) for p in w]
    opt = opt_factory(ps)
    losses = []
    for it in range(LOSS_ITERS):
        state = rollout(x0.to(device), ps, Ts[it], masks[it].to(device))
        loss = _loss_terms(state, target.to(device))
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    return np.array(losses)


def test_loss_curves_against_cpu_oracle():
    """the reference's training step (forward_nsteps, loss, per-parameter gradient normalisation, Adam; experiments.py:213-255)
    for 40 iterations: the CUDA path in each precision against the same loop run on the CPU oracle with torch.optim.Adam"""
    B, C, fc, H, W, w, x0, target, Ts, masks = _train_setup()
    cpe = O.cpe2d(B, H, W)

    class CpuNormAdam(torch.optim.Adam):       # experiments.py:252-255: p.grad /= p.grad.norm() + 1e-8, then Adam
        def step(self):
            for grp in self.param_groups:
                for p in grp["params"]:
                    p.grad /= (p.grad.norm() + 1e-8)
            super().step()

    ref = _train(lambda x, ps, T, m: O.dynca_rollout(x, *ps, m, (0,), "circular", cpe), w, x0, target, Ts, masks, "cpu",
                 lambda ps: CpuNormAdam(ps, lr=2e-3))
    assert ref[-1] < 0.7 * ref[0], "the oracle run itself must be learning"
    print(f"\n[loss curve] CPU oracle: {ref[0]:.4f} -> {ref[-1]:.4f}")
    for prec in ("fp32", "f16x3", "bf16"):
        cfg = Fn.DyncaConfig(C, fc, "circular", (0,), _lib.NCA_COND_CPE, 2, precision=prec)
        got = _train(lambda x, ps, T, m: Fn.dynca_rollout(cfg, x, *ps, T, 0.5, masks=m)[0], w, x0, target, Ts, masks, DEV,
                     lambda ps: nca_b200.NormalizedAdam(ps, lr=2e-3))
        dev = np.abs(got - ref) / ref
        print(f"[loss curve] {prec}: {got[0]:.4f} -> {got[-1]:.4f}, max relative deviation {dev.max():.2e} (iteration {int(dev.argmax())}), "
              f"at iterations 0/9/19/29/39: " + " ".join(f"{dev[i]:.1e}" for i in (0, 9, 19, 29, 39)))
        assert dev.max() < LOSS_TOL[prec], (prec, dev)
        assert dev[:20].max() < 0.1 * LOSS_TOL[prec], (prec, dev[:20])      # first half of the run: ten times tighter
        assert got[-1] < 0.7 * got[0]
