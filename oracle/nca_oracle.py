"""CPU oracle for the NCA step hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in explicit tensor arithmetic, the algorithm of the reference's
NCA step so the CUDA path can be checked against it.  It is NOT a product path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it.  The product (the package under
``video-stylization-with-nca_b200/``) never imports anything from ``oracle/`` and
fails loudly when its CUDA library is missing.

Parity status: the reference ships no tests, golden vectors or fixtures for this path
 so parity is pinned against the reference ITSELF: ``oracle/make_golden.py``
imports the unmodified reference modules from /root/reference in the build container and
commits their outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
restatement against those vectors (and, when /root/reference is present, live).

Reference lines followed (all fp32, NCHW):
  * DyNCA filters / perception / multi-scale / step / rollout:
      ExtraChannels/models/dynca.py:63-69, 71-96, 98-111, 113-128, 130-131, 158-167
      ConditioneDyNCA/models/dynca.py:117-138 (cond_img path), 182-213 (EdgeExtractor)
  * CPE2D positional encoding: ExtraChannels/models/dynca.py:180-207
  * ConditionedNCA (encoder-conditioned): EncoderConditioning/nca.py:29-58, 99-110, 152-209
  * ImageEncoder: EncoderConditioning/encoder.py:5-64
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

PAD_MODES = ("constant", "circular", "replicate", "reflect")


# --------------------------------------------------------------------------------------
# index maps for a 1-px border, per axis (F.pad semantics; dynca.py:81)
# --------------------------------------------------------------------------------------
def _shift(x: torch.Tensor, dy: int, dx: int, mode: str) -> torch.Tensor:
    """Return s with s[..., i, j] = x_padded[..., i+dy, j+dx] for dy, dx in {-1,0,1}."""
    H, W = x.shape[-2:]

    def idx(n: int, d: int):
        i = torch.arange(n) + d
        if mode == "circular":
            return i % n, None
        if mode == "replicate":
            return i.clamp(0, n - 1), None
        if mode == "reflect":
            i = torch.where(i < 0, -i, i)
            i = torch.where(i > n - 1, 2 * (n - 1) - i, i)
            return i, None
        if mode == "constant":
            valid = (i >= 0) & (i < n)
            return i.clamp(0, n - 1), valid
        raise ValueError(mode)

    iy, vy = idx(H, dy)
    ix, vx = idx(W, dx)
    s = x.index_select(-2, iy).index_select(-1, ix)
    if vy is not None:
        s = s * vy.to(x.dtype)[:, None] * vx.to(x.dtype)[None, :]
    return s


# unnormalised filters, cross-correlation taps w[a][b] applied to x[i+a-1, j+b-1]
SOBEL_X = ((-1.0, 0.0, 1.0), (-2.0, 0.0, 2.0), (-1.0, 0.0, 1.0))   # d/dW  (dynca.py:63)
SOBEL_Y = ((-1.0, -2.0, -1.0), (0.0, 0.0, 0.0), (1.0, 2.0, 1.0))   # d/dH  (dynca.py:65)
LAPLACE = ((1.0, 2.0, 1.0), (2.0, -12.0, 2.0), (1.0, 2.0, 1.0))    # dynca.py:68


def _stencil(x: torch.Tensor, taps, mode: str) -> torch.Tensor:
    out = torch.zeros_like(x)
    for a in range(3):
        for b in range(3):
            w = taps[a][b]
            if w != 0.0:
                out = out + w * _shift(x, a - 1, b - 1, mode)
    return out


def down2(x: torch.Tensor) -> torch.Tensor:
    """bilinear, align_corners=False, exact /2 on even dims == 2x2 mean (dynca.py:73-77)."""
    H, W = x.shape[-2:]
    assert H % 2 == 0 and W % 2 == 0
    return 0.25 * (x[..., 0::2, 0::2] + x[..., 0::2, 1::2] + x[..., 1::2, 0::2] + x[..., 1::2, 1::2])


def _up_axis(x: torch.Tensor, dim: int) -> torch.Tensor:
    n = x.shape[dim]
    i = torch.arange(n)
    lo = x.index_select(dim, (i - 1).clamp(0, n - 1))
    hi = x.index_select(dim, (i + 1).clamp(0, n - 1))
    even = 0.25 * lo + 0.75 * x          # dst 2Q   <- src Q-0.25
    odd = 0.75 * x + 0.25 * hi           # dst 2Q+1 <- src Q+0.25
    st = torch.stack([even, odd], dim=dim + 1 if dim >= 0 else dim)
    shp = list(x.shape)
    shp[dim] = 2 * n
    return st.reshape(shp)


def up2(x: torch.Tensor) -> torch.Tensor:
    """bilinear x2 upsample, align_corners=False, edge clamped (dynca.py:93-94)."""
    x = _up_axis(x, x.dim() - 2)
    x = _up_axis(x, x.dim() - 1)
    return x


def perceive(x: torch.Tensor, scale: int, mode: str) -> torch.Tensor:
    """[B,C,H,W] -> [B,4C,H,W], channel blocks [id | sobel_x | sobel_y | lap] (dynca.py:71-96)."""
    H, W = x.shape[-2:]
    if scale == 1:
        x = down2(x)
    elif scale > 1:
        x = F.interpolate(x, size=(H // 2 ** scale, W // 2 ** scale), mode="bilinear", align_corners=False)
    y = torch.cat([x, _stencil(x, SOBEL_X, mode), _stencil(x, SOBEL_Y, mode), _stencil(x, LAPLACE, mode)], dim=1)
    if scale == 1:
        y = up2(y)
    elif scale > 1:
        y = F.interpolate(y, size=(H, W), mode="bilinear", align_corners=False)
    return y


def cpe2d(B: int, H: int, W: int) -> torch.Tensor:
    """Cartesian positional encoding [B,2,H,W] (dynca.py:180-207)."""
    xs = torch.arange(H) / H
    ys = torch.arange(W) / W
    xs = 2.0 * (xs - 0.5 + 0.5 / H)
    ys = 2.0 * (ys - 0.5 + 0.5 / W)
    emb = torch.zeros(2, H, W)
    emb[0] = xs[:, None]
    emb[1] = ys[None, :]
    return emb[None].repeat(B, 1, 1, 1)


def edge_extract(img: torch.Tensor, transform: str = "tanh") -> torch.Tensor:
    """CD EdgeExtractor: zero-padded sobel_x, sobel_y, laplacian of a 1-ch image
    (ConditioneDyNCA/models/dynca.py:182-213)."""
    e = torch.cat([_stencil(img, SOBEL_X, "constant"), _stencil(img, SOBEL_Y, "constant"),
                   _stencil(img, LAPLACE, "constant")], dim=1)
    return torch.tanh(e) if transform == "tanh" else e


def perceive_multiscale(x, scales: Sequence[int], mode: str, cond: Optional[torch.Tensor]):
    y = sum(perceive(x, s, mode) for s in scales) / len(scales)          # dynca.py:98-106
    if cond is not None:
        y = torch.cat([y, cond], dim=1)                                  # dynca.py:108-109
    return y


def dynca_mask_from_uniform(u: torch.Tensor, rate: float) -> torch.Tensor:
    """floor(u + rate), u in [0,1) (dynca.py:121)."""
    return (u + rate).floor()


def dynca_step(x, w1, b1, w2, b2, mask, scales=(0,), mode="circular", cond=None):
    """One DyNCA step with a SUPPLIED fire mask [B,1,H,W] (dynca.py:113-123).

    w1: [fc, 4C+cc], b1: [fc], w2: [C, fc], b2: [C]."""
    z = perceive_multiscale(x, scales, mode, cond)
    h = torch.relu(torch.einsum("jk,bkhw->bjhw", w1, z) + b1[None, :, None, None])
    y = torch.einsum("cj,bjhw->bchw", w2, h) + b2[None, :, None, None]
    return x + y * mask


def dynca_rollout(x, w1, b1, w2, b2, masks, scales=(0,), mode="circular", cond=None, keep=False):
    """T steps (dynca.py:158-167). masks: [T,B,1,H,W]. Returns final state (and all states if keep)."""
    hist = [x]
    for t in range(masks.shape[0]):
        x = dynca_step(x, w1, b1, w2, b2, masks[t], scales, mode, cond)
        if keep:
            hist.append(x)
    return (x, hist) if keep else x


# --------------------------------------------------------------------------------------
# The same DyNCA step phrased with the ATen ops the reference dispatches on CPU (F.pad + depthwise
# F.conv2d + F.interpolate + 1x1 conv2d, dynca.py:71-123).  Used ONLY as the CPU baseline that bench.py
# times on the host cores (cpu_baseline / --impl reference); checked against the golden vectors too.
# --------------------------------------------------------------------------------------
def _filters(C, dtype):
    f = torch.tensor([SOBEL_X, SOBEL_Y, LAPLACE], dtype=dtype)          # [3,3,3]
    return [f[i][None, None].repeat(C, 1, 1, 1) for i in range(3)]


def perceive_aten(x, scale, mode):
    H, W = x.shape[-2:]
    C = x.shape[1]
    if scale != 0:
        x = F.interpolate(x, size=(H // 2 ** scale, W // 2 ** scale), mode="bilinear", align_corners=False)
    xp = F.pad(x, [1, 1, 1, 1], mode)
    ys = [x] + [F.conv2d(xp, w, groups=C) for w in _filters(C, x.dtype)]
    y = torch.cat(ys, dim=1)
    if scale != 0:
        y = F.interpolate(y, size=(H, W), mode="bilinear", align_corners=False)
    return y


def dynca_step_aten(x, w1, b1, w2, b2, mask, scales=(0,), mode="circular", cond=None):
    z = sum(perceive_aten(x, s, mode) for s in scales) / len(scales)
    if cond is not None:
        z = torch.cat([z, cond], dim=1)
    y = F.conv2d(torch.relu(F.conv2d(z, w1[:, :, None, None], b1)), w2[:, :, None, None], b2)
    return x + y * mask


def dynca_rollout_aten(x, w1, b1, w2, b2, masks, scales=(0,), mode="circular", cond=None):
    for t in range(masks.shape[0]):
        x = dynca_step_aten(x, w1, b1, w2, b2, masks[t], scales, mode, cond)
    return x


# --------------------------------------------------------------------------------------
# BF16-operand emulation of the DyNCA step and its BPTT (checker for the tcgen05 path).
# The tensor-core kernels round the GEMM operands to bfloat16 (round-to-nearest-even) and accumulate in
# fp32; everything else (perception, bias b2, fire mask, residual, transposed perception) stays fp32.
# This restatement rounds at exactly those points so the CUDA path can be checked to accumulation-order
# accuracy instead of to bf16 accuracy.  Rounding points (csrc/dynca_bf16.cu):
#   forward : z -> bf16 (cond inputs and b1 as bf16 hi + lo pairs, i.e. ~fp32), W1 -> bf16, h = relu(a) -> bf16,
#             W2 -> bf16
#   backward: g_y = mask * g_next -> bf16, g_a = g_h * [a > 0] -> bf16; gb2 sums the unrounded g_y
# --------------------------------------------------------------------------------------
def bf16r(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _hilo(t):
    hi = bf16r(t)
    return hi + bf16r(t - hi)


# --------------------------------------------------------------------------------------
# "MLP in BF16" stated from the MATH of dynca.py:113-123, not from any kernel: the operands of the two 1x1-conv GEMMs
# (perception vector, W1, hidden layer, W2) are rounded to bfloat16, products accumulate in fp32, everything else
# (perception, cond inputs, biases, mask, residual) is fp32; the backward is torch autograd with the roundings as
# straight-through estimators (so gradients are the fp32 gradients of the rounded forward, with no rounding of their own).
# This is the independent yardstick for the tcgen05 path (BASELINE.json north_star: "1e-2 when the MLP runs in BF16"):
# it knows nothing about tiles, operand layouts or where the kernels round in the backward pass.  The kernel-mirroring
# emulation further below (dynca_bf16emu_*) is a regression tool only.
# --------------------------------------------------------------------------------------
class _RoundBf16STE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def dynca_step_bf16ops(x, w1, b1, w2, b2, mask, scales=(0,), mode="circular", cond=None, fast=False):
    """fast=True phrases the same arithmetic with the ATen ops of dynca_step_aten (4x quicker on big grids)."""
    q = _RoundBf16STE.apply
    C = x.shape[1]
    if fast:
        z = sum(perceive_aten(x, s, mode) for s in scales) / len(scales)
        pre = F.conv2d(q(z), q(w1[:, :4 * C])[:, :, None, None], b1)
        if cond is not None:
            pre = pre + F.conv2d(cond, q(w1[:, 4 * C:])[:, :, None, None])
        y = F.conv2d(q(torch.relu(pre)), q(w2)[:, :, None, None], b2)
        return x + y * mask
    z = perceive_multiscale(x, scales, mode, None)
    pre = torch.einsum("jk,bkhw->bjhw", q(w1[:, :4 * C]), q(z)) + b1[None, :, None, None]
    if cond is not None:      # cond inputs stay fp32 (the reference feeds them unrounded); their weights are MLP weights
        pre = pre + torch.einsum("jk,bkhw->bjhw", q(w1[:, 4 * C:]), cond)
    h = torch.relu(pre)
    y = torch.einsum("cj,bjhw->bchw", q(w2), q(h)) + b2[None, :, None, None]
    return x + y * mask


def dynca_rollout_bf16ops(x, w1, b1, w2, b2, masks, scales=(0,), mode="circular", cond=None, keep=False, fast=False):
    hist = [x]
    for t in range(masks.shape[0]):
        x = dynca_step_bf16ops(x, w1, b1, w2, b2, masks[t], scales, mode, cond, fast=fast)
        if keep:
            hist.append(x)
    return (x, hist) if keep else x


def perceive_coarse(x, mode):
    """[id | sobel_x | sobel_y | lap] of the 2x2-mean coarse state, NOT upsampled: [B,4C,H/2,W/2]."""
    xc = down2(x)
    return torch.cat([xc, _stencil(xc, SOBEL_X, mode), _stencil(xc, SOBEL_Y, mode), _stencil(xc, LAPLACE, mode)], dim=1)


def _emu_preact(xr, w1q, b1q, scales, mode, cond, variant):
    """pre-activation a of one step with the rounding points of kernel `variant` (1: dynca_bf16.cu, 2: dynca_tc2.cu).
    Returns (a, zp, zq): zp = fp32 perception that autograd can differentiate, zq = the rounded GEMM operand the
    weight gradient pairs with (variant 1 layout)."""
    C = xr.shape[1]
    with torch.enable_grad():
        zp = perceive_multiscale(xr, scales, mode, None)
    zq = bf16r(zp.detach())
    if cond is not None:
        zq = torch.cat([zq, _hilo(cond)], dim=1)
    if variant == 2 and len(scales) == 2:
        # tcgen05 v2 (dynca_tc2.cu): fine and coarse perception are rounded separately, their sum Z = z_fine + up(z_coarse) is
        # accumulated in fp32 (constant-matrix upsample on the tensor cores) and rounded once more: Z is the ONE perception
        # operand of the step (forward GEMM, BPTT recompute and weight gradient); W1 / n_scales is exact
        xd = xr.detach()
        zf = bf16r(perceive(xd, 0, mode))
        zc = bf16r(perceive_coarse(xd, mode))
        zsum = bf16r(zf + up2(zc))
        w1h = w1q[:, :4 * C] * 0.5
        a = torch.einsum("jk,bkhw->bjhw", w1h, zsum)
        zq = zsum
        if cond is not None:
            a = a + torch.einsum("jk,bkhw->bjhw", w1q[:, 4 * C:], _hilo(cond))
            zq = torch.cat([zsum, _hilo(cond)], dim=1)
        a = a + b1q[None, :, None, None]
    else:
        a = torch.einsum("jk,bkhw->bjhw", w1q, zq) + b1q[None, :, None, None]
    return a, zp, zq


def dynca_bf16emu_rollout_grads(x0, w1, b1, w2, b2, masks, scales, mode, cond, g_final, taps=None, fwd_variant=1,
                                bwd_variant=1):
    """T steps forward with bf16-rounded GEMM operands, then manual BPTT with the kernel's rounding points.
    taps: {step t in 1..T: dL/d(2 * states[t][:, :3])}.  fwd_variant / bwd_variant: which tensor-core kernel generation
    ran the forward rollout / recomputes the step inside the BPTT (nca_dynca_kernel_variant).
    Returns (final, dict of gradients, states)."""
    taps = taps or {}
    T = masks.shape[0]
    C = x0.shape[1]
    w1q = bf16r(w1)                             # cond columns: the bf16 weight multiplies the hi and the lo input slot
    w2q, b1q = bf16r(w2), _hilo(b1)
    xs, saved = [x0], []
    x = x0
    for t in range(T):
        xr = x.detach().clone().requires_grad_(True)
        a, zp, zq = _emu_preact(xr, w1q, b1q, scales, mode, cond, fwd_variant)
        hq = bf16r(torch.relu(a))
        y = torch.einsum("cj,bjhw->bchw", w2q, hq) + b2[None, :, None, None]
        x = x.detach() + y * masks[t]
        if bwd_variant != fwd_variant:          # the BPTT kernel recomputes the step with its own rounding points
            a, zp, zq = _emu_preact(xr, w1q, b1q, scales, mode, cond, bwd_variant)
            hq = bf16r(torch.relu(a))
        saved.append((xr, zp, zq, a, hq))
        xs.append(x)
    g = g_final.clone() if g_final is not None else torch.zeros_like(x0)
    gw1 = torch.zeros_like(w1); gb1 = torch.zeros_like(b1); gw2 = torch.zeros_like(w2); gb2 = torch.zeros_like(b2)
    two = bwd_variant == 2 and len(scales) == 2
    for t in range(T - 1, -1, -1):
        if (t + 1) in taps:
            g = g.clone()
            g[:, :3] += 2.0 * taps[t + 1]
        xr, zp, zq, a, hq = saved[t]
        gy32 = masks[t] * g
        gy = bf16r(gy32)
        gb2 += gy32.sum(dim=(0, 2, 3))
        gw2 += torch.einsum("bchw,bjhw->cj", gy, hq)
        gh = torch.einsum("bchw,cj->bjhw", gy, w2q)
        ga = bf16r(gh * (a > 0).to(gh.dtype))
        gb1 += ga.sum(dim=(0, 2, 3))
        if two:
            # dynca_tc3_bwd.cu: the weight gradient pairs g_a with the recorded operand Z; g_z = g_a . W1h is fp32, its fine part goes
            # through the transposed stencils as it is, its coarse part through U^T applied to bf16(g_z)
            xd = xr.detach().clone().requires_grad_(True)
            with torch.enable_grad():
                zf = perceive(xd, 0, mode)
                zc = perceive_coarse(xd, mode)
                zc_leaf = zc.detach().clone().requires_grad_(True)
                zu = up2(zc_leaf)
            w1h = w1q[:, :4 * C] * 0.5
            gw1[:, :4 * C] += 0.5 * torch.einsum("bjhw,bkhw->jk", ga, zq[:, :4 * C])
            if cond is not None:
                gw1[:, 4 * C:] += torch.einsum("bjhw,bkhw->jk", ga, zq[:, 4 * C:])
            gzf = torch.einsum("bjhw,jk->bkhw", ga, w1h)
            (gzc,) = torch.autograd.grad(zu, zc_leaf, bf16r(gzf))
            (gx,) = torch.autograd.grad([zf, zc], xd, [gzf, gzc])
        else:
            gw1 += torch.einsum("bjhw,bkhw->jk", ga, zq)
            gz = torch.einsum("bjhw,jk->bkhw", ga, w1q[:, :4 * C])
            (gx,) = torch.autograd.grad(zp, xr, gz)
        g = g + gx
    return xs[-1], dict(x0=g, w1=gw1, b1=gb1, w2=gw2, b2=gb2), xs


# --------------------------------------------------------------------------------------
# EncoderConditioning/nca.py
# --------------------------------------------------------------------------------------
def enc_alive(x, living_dim: int, thr: float = 0.1):
    """nca.py:152-163 : 3x3 max-pool (-inf padded) of the living channel > thr."""
    a = x[:, living_dim:living_dim + 1]
    B, _, H, W = a.shape
    p = torch.full((B, 1, H + 2, W + 2), -math.inf, dtype=a.dtype)
    p[:, :, 1:-1, 1:-1] = a
    m = p[:, :, 1:-1, 1:-1]
    for dy in range(3):
        for dx in range(3):
            m = torch.maximum(m, p[:, :, dy:dy + H, dx:dx + W])
    return m > thr


def gaussian_kernel5():
    """encoder.py:59-63: normalised 5x5 gaussian, sigma 1, built in float32 from python floats"""
    k = torch.tensor([[(1 / (2 * math.pi)) * math.exp(-((i - 2) ** 2 + (j - 2) ** 2) / 2.0) for j in range(5)] for i in range(5)])
    return (k / torch.sum(k)).float()


def image_encoder(x, w1, b1, w2):
    """ImageEncoder.forward (EncoderConditioning/encoder.py:37-57): x [B,ch,H,W] -> [B,E,H,W].
    w1 [E,ch+3,3,3], b1 [E], w2 [E,E,3,3]; every convolution zero padded."""
    ch = x.shape[1]
    gray = torch.mean(x, dim=1, keepdim=True)
    feats = [_stencil(gray, SOBEL_X, "constant"), _stencil(gray, SOBEL_Y, "constant"), _stencil(gray, LAPLACE, "constant")]
    gk = gaussian_kernel5().to(x.dtype)[None, None]
    feats += [F.conv2d(x[:, i:i + 1], gk, padding=2) for i in range(ch)]
    h = torch.relu(F.conv2d(torch.cat(feats, dim=1), w1, b1, padding=1))
    return F.conv2d(h, w2, None, padding=1)


def enc_perception(x, wp):
    """Learned depthwise 3x3, zero pad, out channel j reads in channel j//3 (nca.py:99-107).
    wp: [3C,1,3,3] (cross-correlation)."""
    B, C, H, W = x.shape
    outs = []
    for j in range(3 * C):
        c = j // 3
        taps = [[float(wp[j, 0, a, b]) for b in range(3)] for a in range(3)]
        outs.append(_stencil(x[:, c:c + 1], taps, "constant"))
    return torch.cat(outs, dim=1)


def enc_perception_fast(x, wp):
    C = x.shape[1]
    return F.conv2d(x, wp, padding=1, groups=C)


def enc_step(x, goal, wp, wa, ba, wb, bb, wc, fire, living_dim=3, thr=0.1, fast=True):
    """nca.py:176-195 with a SUPPLIED fire mask (float [B,1,H,W], = (u < rate))."""
    pre = enc_alive(x, living_dim, thr)
    xin = x + goal * pre
    p = enc_perception_fast(xin, wp) if fast else enc_perception(xin, wp)
    h1 = torch.relu(torch.einsum("jk,bkhw->bjhw", wa, p) + ba[None, :, None, None])
    h2 = torch.relu(torch.einsum("jk,bkhw->bjhw", wb, h1) + bb[None, :, None, None])
    out = torch.einsum("cj,bjhw->bchw", wc, h2)
    x = x + fire * out
    post = enc_alive(x, living_dim, thr)
    x = x * (pre & post).to(x.dtype)
    return torch.clamp(x, -10.0, 10.0)


def enc_step_bf16ops(x, goal, wp, wa, ba, wb, bb, wc, fire, living_dim=3, thr=0.1):
    """enc_step with "the update MLP runs in BF16" stated from the math (see dynca_step_bf16ops): the operands of the three 1x1-conv
    GEMMs (perception vector, Wa, h1, Wb, h2, Wc) rounded to bfloat16, fp32 accumulation, everything else fp32; autograd with
    straight-through rounding.  Knows nothing about the kernels."""
    q = _RoundBf16STE.apply
    pre = enc_alive(x, living_dim, thr)
    xin = x + goal * pre
    p = enc_perception_fast(xin, wp)
    h1 = torch.relu(torch.einsum("jk,bkhw->bjhw", q(wa), q(p)) + ba[None, :, None, None])
    h2 = torch.relu(torch.einsum("jk,bkhw->bjhw", q(wb), q(h1)) + bb[None, :, None, None])
    out = torch.einsum("cj,bjhw->bchw", q(wc), q(h2))
    x = x + fire * out
    post = enc_alive(x, living_dim, thr)
    x = x * (pre & post).to(x.dtype)
    return torch.clamp(x, -10.0, 10.0)


def enc_rollout_bf16ops(x, goal, wp, wa, ba, wb, bb, wc, fires, living_dim=3, thr=0.1):
    for t in range(fires.shape[0]):
        x = enc_step_bf16ops(x, goal, wp, wa, ba, wb, bb, wc, fires[t], living_dim, thr)
    return x


def enc_rollout(x, goal, wp, wa, ba, wb, bb, wc, fires, living_dim=3, thr=0.1, keep=False):
    hist = [x]
    for t in range(fires.shape[0]):
        x = enc_step(x, goal, wp, wa, ba, wb, bb, wc, fires[t], living_dim, thr)
        if keep:
            hist.append(x)
    return (x, hist) if keep else x


# --------------------------------------------------------------------------------------
# BF16-operand emulation of the ConditionedNCA step and its BPTT (checker for csrc/enc_tc.cu).
# Rounding points: p -> bf16, Wa / Wb / Wc -> bf16, ba as bf16 hi + lo, h1 / h2 -> bf16 (bb added in fp32 before the relu);
# backward: g_o = fire * g1 -> bf16, g_a2 / g_a1 -> bf16; gbb / gba sum the ROUNDED g_a2 / g_a1 (they ride the weight-gradient
# MMAs); g_p, the transposed depthwise conv, gwp and the pass-through stay fp32.
# --------------------------------------------------------------------------------------
def enc_bf16emu_rollout_grads(x0, goal, wp, wa, ba, wb, bb, wc, fires, g_final, living_dim=3, thr=0.1, clampv=10.0):
    T = fires.shape[0]
    waq, wbq, wcq, baq = bf16r(wa), bf16r(wb), bf16r(wc), _hilo(ba)
    xs, saved = [x0], []
    x = x0
    for t in range(T):
        pre = enc_alive(x, living_dim, thr).to(x.dtype) if living_dim >= 0 else torch.ones_like(x[:, :1])
        xin = (x + goal * pre).detach().clone().requires_grad_(True)
        with torch.enable_grad():
            p = enc_perception_fast(xin, wp)
        pq = bf16r(p.detach())
        a1 = torch.einsum("jk,bkhw->bjhw", waq, pq) + baq[None, :, None, None]
        h1 = bf16r(torch.relu(a1))
        a2 = torch.einsum("jk,bkhw->bjhw", wbq, h1) + bb[None, :, None, None]
        h2 = bf16r(torch.relu(a2))
        out = torch.einsum("cj,bjhw->bchw", wcq, h2)
        x1 = x + fires[t] * out
        post = enc_alive(x1, living_dim, thr).to(x.dtype) if living_dim >= 0 else torch.ones_like(pre)
        life = pre * post
        v = x1 * life
        saved.append((xin, p, pq, h1, h2, pre, life, (v >= -clampv) & (v <= clampv)))
        x = torch.clamp(v, -clampv, clampv)
        xs.append(x)
    g = g_final.clone()
    G = {k: torch.zeros_like(t_) for k, t_ in dict(wp=wp, wa=wa, ba=ba, wb=wb, bb=bb, wc=wc, goal=goal).items()}
    for t in range(T - 1, -1, -1):
        xin, p, pq, h1, h2, pre, life, cm = saved[t]
        g1 = g * cm.to(g.dtype) * life
        go = bf16r(fires[t] * g1)
        G["wc"] += torch.einsum("bchw,bjhw->cj", go, h2)
        ga2 = bf16r(torch.einsum("bchw,cj->bjhw", go, wcq) * (h2 > 0).to(g.dtype))
        G["wb"] += torch.einsum("bjhw,bkhw->jk", ga2, h1)
        G["bb"] += ga2.sum(dim=(0, 2, 3))
        ga1 = bf16r(torch.einsum("bjhw,jk->bkhw", ga2, wbq) * (h1 > 0).to(g.dtype))
        G["wa"] += torch.einsum("bjhw,bkhw->jk", ga1, pq)
        G["ba"] += ga1.sum(dim=(0, 2, 3))
        gp = torch.einsum("bjhw,jk->bkhw", ga1, waq)
        wpl = wp.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            p2 = enc_perception_fast(xin, wpl)
        gxin, gwp = torch.autograd.grad(p2, [xin, wpl], gp)
        G["wp"] += gwp
        G["goal"] += gxin * pre
        g = g1 + gxin
    G["x0"] = g
    return xs[-1], G
