"""Functional layer over the C ABI: rollouts as autograd Functions, perception, edge maps, Philox masks.

Everything here runs on CUDA through libnca_b200.so; tensors are fp32, contiguous, NCHW.  torch is used only
for device memory, streams and autograd bookkeeping.
"""
import collections.abc
import os
import warnings
import ctypes as C

import torch

from . import _lib
from ._lib import NcaError, check, load_library


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _need_cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise NcaError("the NCA step runs on CUDA only (got a CPU tensor); there is no CPU fallback")
        if t.dtype != torch.float32:
            raise NcaError(f"expected float32, got {t.dtype}")


def _c(t):
    return None if t is None else t.detach().contiguous()


def new_seed():
    """Philox key drawn from torch's default generator, so torch.manual_seed controls the fire masks."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


# ------------------------------------------------------------------------------------------------
# DyNCA
# ------------------------------------------------------------------------------------------------
class DyncaConfig:
    """Static description of a DyNCA model (ctor arguments of the reference's DyNCA, dynca.py:30-69)."""

    def __init__(self, C_, fc, pad, scales, cond_kind, cc, precision="fp32"):
        scales = list(scales)
        if scales not in ([0], [0, 1]):
            raise NcaError(f"perception_scales {scales} not supported (only [0] and [0, 1])")
        if pad not in _lib.NCA_PAD:
            raise NcaError(f"padding_mode {pad!r} not supported")
        self.C, self.fc, self.pad, self.ns = C_, fc, _lib.NCA_PAD[pad], len(scales)
        self.cond_kind, self.cc, self.precision = cond_kind, cc, _lib.NCA_PREC[precision]

    def desc(self, B, H, W, rate, supplied):
        return _lib.DyncaDesc(B, self.C, H, W, self.fc, self.cond_kind, self.cc, self.pad, self.ns, self.precision,
                              _lib.NCA_MASK_SUPPLIED if supplied else _lib.NCA_MASK_PHILOX, float(rate))


def _weights_struct(w1, b1, w2, b2):
    return _lib.DyncaWeights(w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr())


def dynca_kernel_variant(cfg, B, H, W, backward=False):
    """Which step kernel the library dispatches this shape to (0 fp32, 1 tcgen05 4x32 tiles, 2 tcgen05 8x16 + TMA)."""
    d = cfg.desc(B, H, W, 0.5, False)
    return int(load_library().nca_dynca_kernel_variant(C.byref(d), int(backward)))


def dynca_perceive(cfg, x, cond=None):
    """DyNCA.perceive_multiscale (dynca.py:98-111) -> [B, 4C+cc, H, W]."""
    _need_cuda(x, cond)
    if torch.is_grad_enabled() and (x.requires_grad or (cond is not None and cond.requires_grad)):
        # the reference returns a differentiable tensor here; this entry point has no backward, so refuse instead of
        # silently cutting the graph (the rollouts, which do have a BPTT, never come through here)
        raise NcaError("perceive_* is forward-only on the CUDA path: call it under torch.no_grad() or on detached tensors")
    lib = load_library()
    x, cond = _c(x), _c(cond)
    B, Cc, H, W = x.shape
    z = torch.empty(B, 4 * Cc + cfg.cc, H, W, device=x.device, dtype=torch.float32)
    d = cfg.desc(B, H, W, 0.5, False)
    with torch.cuda.device(x.device):
        check(lib.nca_dynca_perceive(C.byref(d), _ptr(x), _ptr(cond), _ptr(z), _stream()))
    return z


def edge_extract(img, tanh):
    """EdgeExtractor.forward (ConditioneDyNCA/models/dynca.py:204-213): [B,1,H,W] -> [B,3,H,W]."""
    _need_cuda(img)
    lib = load_library()
    img = _c(img)
    B, one, H, W = img.shape
    if one != 1:
        raise NcaError("edge_extract expects a one-channel image")
    out = torch.empty(B, 3, H, W, device=img.device, dtype=torch.float32)
    with torch.cuda.device(img.device):
        check(lib.nca_edge_extract(B, H, W, _ptr(img), int(bool(tanh)), _ptr(out), _stream()))
    return out


def philox_mask(B, H, W, rate, seed, T, t0=0, enc=False, device="cuda"):
    """The fire masks the kernels generate for (seed, t0..t0+T), as float [T,B,1,H,W]."""
    lib = load_library()
    out = torch.empty(T, B, 1, H, W, device=device, dtype=torch.float32)
    with torch.cuda.device(out.device):
        check(lib.nca_philox_mask(B, H, W, float(rate), int(enc), C.c_uint64(seed), t0, T, _ptr(out), _stream()))
    return out


# Operand history (nca_b200.h: op_hist): the forward records the bf16 perception operands of every step so that the BPTT
# loads them instead of recomputing the perception.  128-224 B per cell and step (DESIGN.md section 2 lists every configuration);
# kept within NCA_OP_HIST_MAX_GB (default 48; 0 disables).  A rollout whose history would exceed the cap keeps it for the LAST
# steps that fit and recomputes the perception for the earlier ones (the rollout is split into two C calls at that step); a
# one-time warning says so.  Results are bit-identical either way.
def _op_hist_limit():
    return int(float(os.environ.get("NCA_OP_HIST_MAX_GB", "48")) * (1 << 30))


_OP_HIST_WARNED = False


def _op_hist_plan(lib, d, T):
    """(steps without history, bytes of the history for the remaining steps)"""
    global _OP_HIST_WARNED
    per_step = lib.nca_dynca_op_hist_bytes(C.byref(d), 1)
    if per_step == 0 or T <= 0:
        return T, 0
    k = min(T, _op_hist_limit() // per_step)
    if k < T and not _OP_HIST_WARNED and _op_hist_limit() > 0:       # (a cap of 0 is an explicit opt-out, not a surprise)
        _OP_HIST_WARNED = True
        warnings.warn(f"NCA operand history: {T} steps need {T * per_step / 2**30:.1f} GiB, the cap (NCA_OP_HIST_MAX_GB) allows "
                      f"{k}: the first {T - k} steps of the BPTT recompute the perception (slower, same results)", stacklevel=3)
    return T - k, k * per_step


def _dynca_forward_raw(cfg, x0, w1, b1, w2, b2, cond, masks, seed, T, rate, keep_history, want_ops=False):
    lib = load_library()
    B, Cc, H, W = x0.shape
    d = cfg.desc(B, H, W, rate, masks is not None)
    n_slots = T + 1 if keep_history else 2
    states = torch.empty(n_slots, B, Cc, H, W, device=x0.device, dtype=torch.float32)
    states[0].copy_(x0)
    coarse = None
    if keep_history and cfg.ns == 2:     # coarse (2x2-mean) state history, reused by the BPTT
        coarse = torch.empty(n_slots, B, Cc, H // 2, W // 2, device=x0.device, dtype=torch.float32)
    ops, t1 = None, 0
    with torch.cuda.device(x0.device):
        if keep_history and want_ops and T > 0:
            t1, ob = _op_hist_plan(lib, d, T)
            if ob > 0:
                ops = torch.empty(ob, device=x0.device, dtype=torch.uint8)
            else:
                t1 = 0
        nbytes = lib.nca_dynca_workspace_bytes(C.byref(d), 0)
        ws = torch.empty(max(nbytes, 16), device=x0.device, dtype=torch.uint8)
        wst = _weights_struct(w1, b1, w2, b2)
        if t1 > 0:      # steps [0, t1) without the operand history, steps [t1, T) with it: two calls on the same buffers
            check(lib.nca_dynca_forward(C.byref(d), C.byref(wst), _ptr(cond), _ptr(masks), C.c_uint64(seed), 0, t1,
                                        1, _ptr(states), _ptr(coarse), None, _ptr(ws), nbytes, _stream()))
            check(lib.nca_dynca_forward(C.byref(d), C.byref(wst), _ptr(cond), _ptr(masks[t1:] if masks is not None else None), C.c_uint64(seed),
                                        t1, T - t1, 1, _ptr(states[t1:]), _ptr(coarse[t1:] if coarse is not None else None), _ptr(ops),
                                        _ptr(ws), nbytes, _stream()))
        else:
            check(lib.nca_dynca_forward(C.byref(d), C.byref(wst), _ptr(cond), _ptr(masks), C.c_uint64(seed), 0, T,
                                        int(keep_history), _ptr(states), _ptr(coarse), _ptr(ops), _ptr(ws), nbytes, _stream()))
    if keep_history and want_ops:
        return states, coarse, (ops, t1)
    if keep_history:
        return states, coarse
    return states


class _RolloutHandle:
    """Shared between the rollout Function and the lazy rgb taps: gradients of taps land here."""

    def __init__(self):
        self.tap_grads = {}
        self.hist = None
        self.c_out = 0


class _DyncaRollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, w1, b1, w2, b2, cond, masks, cfg, T, rate, seed, handle):
        x0c, w1c, b1c, w2c, b2c = _c(x0), _c(w1), _c(b1), _c(w2), _c(b2)
        cond, masks = _c(cond), _c(masks)
        hist, coarse, (ops, ops_t1) = _dynca_forward_raw(cfg, x0c, w1c, b1c, w2c, b2c, cond, masks, seed, T, rate, True, want_ops=True)
        ctx.coarse = coarse
        ctx.ops, ctx.ops_t1 = ops, ops_t1
        ctx.cfg, ctx.T, ctx.rate, ctx.seed, ctx.handle = cfg, T, rate, seed, handle
        ctx.w_shapes = (w1.shape, b1.shape, w2.shape, b2.shape)
        ctx.save_for_backward(w1c, b1c, w2c, b2c, cond, masks)
        ctx.hist = hist
        handle.hist = hist
        ctx.set_materialize_grads(False)
        token = x0.new_zeros(())
        # a copy, not a view: the BPTT recomputes from `hist`, which autograd's version counters do not cover, so an in-place
        # edit of the returned state (clamp_, overwriting the conditioning channel) must not reach it (1/T of the history)
        return hist[T].clone(), token

    @staticmethod
    def backward(ctx, g_final, _g_token):
        lib = load_library()
        w1, b1, w2, b2, cond, masks = ctx.saved_tensors
        hist, cfg, T = ctx.hist, ctx.cfg, ctx.T
        _, B, Cc, H, W = hist.shape
        d = cfg.desc(B, H, W, ctx.rate, masks is not None)
        g_final = _c(g_final)
        taps = sorted(ctx.handle.tap_grads.items())
        tap_c = ctx.handle.c_out
        gx0 = torch.empty(B, Cc, H, W, device=hist.device, dtype=torch.float32)
        # the four weight gradients are written back to back into ONE buffer (in parameter order): the data-parallel
        # all-reduce then runs on it as it is (parallel.flat_view), without a flatten / concatenate launch
        sizes = [t.numel() for t in (w1, b1, w2, b2)]
        flat = torch.empty(sum(sizes), device=hist.device, dtype=torch.float32)
        gw1, gb1, gw2, gb2 = flat.split(sizes)
        t1 = ctx.ops_t1 if ctx.ops is not None else 0
        with torch.cuda.device(hist.device):
            nbytes = lib.nca_dynca_workspace_bytes(C.byref(d), 1)
            ws = torch.empty(nbytes, device=hist.device, dtype=torch.uint8)
            wst = _weights_struct(w1, b1, w2, b2)

            def call(ta, tb, ops, g_in, g_out, gws):
                """BPTT through steps [ta, tb): taps at states[ta+1 .. tb] belong to it"""
                mine = [(s - ta, t) for s, t in taps if ta < s <= tb]
                n = len(mine)
                tap_ptrs = (C.c_void_p * max(n, 1))(*[t.data_ptr() for _, t in mine])
                tap_steps = (C.c_int32 * max(n, 1))(*[s for s, _ in mine])
                gst = _weights_struct(*gws)
                check(lib.nca_dynca_backward(C.byref(d), C.byref(wst), _ptr(cond), _ptr(masks[ta:] if masks is not None else None),
                                             C.c_uint64(ctx.seed), ta, tb - ta, _ptr(hist[ta:]),
                                             _ptr(ctx.coarse[ta:] if ctx.coarse is not None else None), _ptr(ops), _ptr(g_in), tap_ptrs, tap_steps,
                                             n, max(tap_c, 1), 2.0, _ptr(g_out), C.byref(gst), _ptr(ws), nbytes, _stream()))

            if t1 > 0:      # [t1, T) with the operand history, then [0, t1) recomputing the perception (see _op_hist_plan)
                g_mid = torch.empty_like(gx0)
                part = torch.empty_like(flat)
                call(t1, T, ctx.ops, g_final, g_mid, part.split(sizes))
                call(0, t1, None, g_mid, gx0, (gw1, gb1, gw2, gb2))
                flat += part
            else:
                call(0, T, ctx.ops, g_final, gx0, (gw1, gb1, gw2, gb2))
        ctx.handle.tap_grads = {}
        ctx.ops = None
        s1, sb1, s2, sb2 = ctx.w_shapes
        return (gx0, gw1.view(s1), gb1.view(sb1), gw2.view(s2), gb2.view(sb2), None, None, None, None, None, None, None)


class _RgbTap(torch.autograd.Function):
    """rgb_t = 2 * states[t][:, :c_out] (DyNCA.to_rgb, dynca.py:130-131) whose gradient is routed into the
    rollout's BPTT as a tap instead of through a dense gradient of the whole history."""

    @staticmethod
    def forward(ctx, token, handle, t):
        ctx.handle, ctx.t = handle, t
        return handle.hist[t][:, :handle.c_out] * 2.0

    @staticmethod
    def backward(ctx, g):
        h = ctx.handle
        g = g.contiguous()
        h.tap_grads[ctx.t] = g if ctx.t not in h.tap_grads else h.tap_grads[ctx.t] + g
        return torch.zeros((), device=g.device, dtype=g.dtype), None, None


class RgbTaps(collections.abc.Sequence):
    """List-like view of the per-step rgb outputs of forward_nsteps(return_middle_feature=True): entry i is the
    rgb after step i+1 (dynca.py:161-165).  Entries are materialised on access."""

    def __init__(self, handle, token, T):
        self._h, self._token, self._T = handle, token, T

    def __len__(self):
        return self._T

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._T))]
        if i < 0:
            i += self._T
        if not 0 <= i < self._T:
            raise IndexError(i)
        if self._token is None:
            return self._h.hist[i + 1][:, :self._h.c_out] * 2.0
        return _RgbTap.apply(self._token, self._h, i + 1)


def dynca_rollout(cfg, x0, w1, b1, w2, b2, T, rate=0.5, cond=None, masks=None, seed=None, c_out=3,
                  return_taps=False):
    """T DyNCA steps (DyNCA.forward_nsteps, dynca.py:158-167).

    masks: optional supplied fire masks [T,B,1,H,W] (1 = fire); otherwise in-kernel Philox keyed on `seed`.
    Returns (final_state, taps) with taps an RgbTaps sequence when return_taps else None."""
    _need_cuda(x0, w1, b1, w2, b2, cond, masks)
    if seed is None:
        seed = new_seed() if masks is None else 0
    if masks is not None and tuple(masks.shape) != (T, x0.shape[0], 1, x0.shape[2], x0.shape[3]):
        raise NcaError(f"masks must be [T,B,1,H,W] = {(T, x0.shape[0], 1, x0.shape[2], x0.shape[3])}, got {tuple(masks.shape)}")
    needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (x0, w1, b1, w2, b2))
    handle = _RolloutHandle()
    handle.c_out = c_out
    if T == 0:
        return x0, (RgbTaps(handle, None, 0) if return_taps else None)
    if needs_grad:
        final, token = _DyncaRollout.apply(x0, w1, b1, w2, b2, cond, masks, cfg, T, rate, seed, handle)
        return final, (RgbTaps(handle, token, T) if return_taps else None)
    states = _dynca_forward_raw(cfg, _c(x0), _c(w1), _c(b1), _c(w2), _c(b2), _c(cond), _c(masks), seed, T, rate,
                                return_taps)
    if return_taps:
        states = states[0]
        handle.hist = states
        return states[T], RgbTaps(handle, None, T)
    return states[T & 1], None


# ------------------------------------------------------------------------------------------------
# ConditionedNCA (EncoderConditioning/nca.py)
# ------------------------------------------------------------------------------------------------
class EncConfig:
    """Static description of a ConditionedNCA (ctor arguments of nca.py:62-74)."""

    def __init__(self, C_, living_dim, alive_thr=0.1, fire_rate=0.5, hid=64, clamp=10.0, precision="fp32"):
        self.C, self.living_dim, self.alive_thr, self.fire_rate, self.hid, self.clamp = C_, living_dim, alive_thr, fire_rate, hid, clamp
        self.precision = _lib.NCA_PREC[precision]

    def desc(self, B, H, W, supplied):
        return _lib.EncDesc(B, self.C, H, W, self.hid, self.living_dim,
                            _lib.NCA_MASK_SUPPLIED if supplied else _lib.NCA_MASK_PHILOX,
                            float(self.alive_thr), float(self.fire_rate), float(self.clamp), self.precision)


def _enc_weights_struct(ws):
    return _lib.EncWeights(*[t.data_ptr() for t in ws])


def _enc_forward_raw(cfg, x0, goal, ws, masks, seed, T, keep_history):
    lib = load_library()
    B, Cc, H, W = x0.shape
    d = cfg.desc(B, H, W, masks is not None)
    states = torch.empty(T + 1 if keep_history else 2, B, Cc, H, W, device=x0.device, dtype=torch.float32)
    states[0].copy_(x0)
    life = torch.empty(T, B, H, W, device=x0.device, dtype=torch.uint8) if keep_history else None
    with torch.cuda.device(x0.device):
        nbytes = lib.nca_enc_workspace_bytes(C.byref(d), 0)
        wsb = torch.empty(max(nbytes, 16), device=x0.device, dtype=torch.uint8)
        wst = _enc_weights_struct(ws)
        check(lib.nca_enc_forward(C.byref(d), C.byref(wst), _ptr(goal), _ptr(masks), C.c_uint64(seed), 0, T,
                                  int(keep_history), _ptr(states), _ptr(life), _ptr(wsb), nbytes, _stream()))
    return states, life


class _EncRollout(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, goal, wp, wa, ba, wb, bb, wc, masks, cfg, T, seed):
        ws = [_c(t) for t in (wp, wa, ba, wb, bb, wc)]
        x0c, goalc, masks = _c(x0), _c(goal), _c(masks)
        hist, life = _enc_forward_raw(cfg, x0c, goalc, ws, masks, seed, T, True)
        ctx.cfg, ctx.T, ctx.seed = cfg, T, seed
        ctx.w_shapes = tuple(t.shape for t in (wp, wa, ba, wb, bb, wc))
        ctx.save_for_backward(goalc, masks, *ws)
        ctx.hist, ctx.life = hist, life
        return hist[T].clone()      # a copy: see _DyncaRollout.forward

    @staticmethod
    def backward(ctx, g_final):
        lib = load_library()
        goal, masks, *ws = ctx.saved_tensors
        hist, life, cfg, T = ctx.hist, ctx.life, ctx.cfg, ctx.T
        _, B, Cc, H, W = hist.shape
        d = cfg.desc(B, H, W, masks is not None)
        g_final = _c(g_final)
        gx0 = torch.empty(B, Cc, H, W, device=hist.device, dtype=torch.float32)
        ggoal = torch.empty_like(gx0)
        sizes = [t.numel() for t in ws]
        gws = list(torch.empty(sum(sizes), device=hist.device, dtype=torch.float32).split(sizes))      # one flat buffer, parameter order
        with torch.cuda.device(hist.device):
            nbytes = lib.nca_enc_workspace_bytes(C.byref(d), 1)
            wsb = torch.empty(nbytes, device=hist.device, dtype=torch.uint8)
            wst, gst = _enc_weights_struct(ws), _enc_weights_struct(gws)
            check(lib.nca_enc_backward(C.byref(d), C.byref(wst), _ptr(goal), _ptr(masks), C.c_uint64(ctx.seed), 0, T,
                                       _ptr(hist), _ptr(life), _ptr(g_final), _ptr(gx0), _ptr(ggoal), C.byref(gst),
                                       _ptr(wsb), nbytes, _stream()))
        return (gx0, ggoal, *[g.view(s) for g, s in zip(gws, ctx.w_shapes)], None, None, None, None)


class _ImageEncoderFn(torch.autograd.Function):
    """ImageEncoder.forward fused (EncoderConditioning/encoder.py:37-57) -> the zero-padded goal tensor [B,C,H,W] the rollout reads;
    backward = the weight gradients from the BPTT's d(goal) in one kernel (the image is data: no gradient)."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, goal_channels):
        lib = load_library()
        xc, w1c, b1c, w2c = _c(x), _c(w1), _c(b1), _c(w2)
        B, ch, H, W = xc.shape
        E = w2c.shape[0]
        keep = any(ctx.needs_input_grad[1:4])
        feats = torch.empty(B, ch + 3, H, W, device=x.device, dtype=torch.float32) if keep else None
        hidden = torch.empty(B, E, H, W, device=x.device, dtype=torch.float32) if keep else None
        goal = torch.empty(B, goal_channels, H, W, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            check(lib.nca_encoder_forward(B, ch, H, W, E, _ptr(xc), _ptr(w1c), _ptr(b1c), _ptr(w2c), _ptr(feats), _ptr(hidden), _ptr(goal),
                                          goal_channels, _stream()))
        ctx.save_for_backward(feats, hidden, w2c)
        ctx.shapes = (w1.shape, b1.shape, w2.shape, ch, E, goal_channels)
        return goal

    @staticmethod
    def backward(ctx, g_goal):
        lib = load_library()
        feats, hidden, w2c = ctx.saved_tensors
        s1, sb, s2, ch, E, gc = ctx.shapes
        g = _c(g_goal)
        B, _, H, W = g.shape
        sizes = [E * (ch + 3) * 9, E, E * E * 9]
        gw1, gb1, gw2 = torch.empty(sum(sizes), device=g.device, dtype=torch.float32).split(sizes)
        with torch.cuda.device(g.device):
            check(lib.nca_encoder_backward(B, ch, H, W, E, _ptr(feats), _ptr(hidden), _ptr(w2c), _ptr(g), gc, _ptr(gw1), _ptr(gb1), _ptr(gw2),
                                           _stream()))
        return None, gw1.view(s1), gb1.view(sb), gw2.view(s2), None


def image_encoder_supported(channels, embedding_dim):
    return channels == 3 and embedding_dim == 16


def image_encoder(x, w1, b1, w2, goal_channels):
    """ImageEncoder (encoder.py:5-64) on the CUDA path: x [B,3,H,W] -> zero-padded goal encoding [B,goal_channels,H,W]."""
    _need_cuda(x, w1, b1, w2)
    if x.requires_grad and torch.is_grad_enabled():
        raise NcaError("the fused ImageEncoder does not differentiate its input image (it is data in the reference's trainer)")
    return _ImageEncoderFn.apply(x, w1, b1, w2, int(goal_channels))


def enc_rollout(cfg, x0, goal, wp, wa, ba, wb, bb, wc, T, masks=None, seed=None):
    """T ConditionedNCA steps (the loop of ConditionedNCA.grow, nca.py:207-208) on an already encoded, zero-padded
    goal [B,C,H,W].  masks: optional supplied fire masks [T,B,1,H,W] (1 = fire, i.e. u < rate)."""
    _need_cuda(x0, goal, wp, wa, ba, wb, bb, wc, masks)
    if tuple(goal.shape) != tuple(x0.shape):
        raise NcaError(f"goal encoding must have the state's shape {tuple(x0.shape)}, got {tuple(goal.shape)}")
    if seed is None:
        seed = new_seed() if masks is None else 0
    if masks is not None and tuple(masks.shape) != (T, x0.shape[0], 1, x0.shape[2], x0.shape[3]):
        raise NcaError(f"masks must be [T,B,1,H,W], got {tuple(masks.shape)}")
    if T == 0:
        return x0
    ts = (x0, goal, wp, wa, ba, wb, bb, wc)
    if torch.is_grad_enabled() and any(t.requires_grad for t in ts):
        return _EncRollout.apply(*ts, masks, cfg, T, seed)
    states, _ = _enc_forward_raw(cfg, _c(x0), _c(goal), [_c(t) for t in ts[2:]], _c(masks), seed, T, False)
    return states[T & 1]
