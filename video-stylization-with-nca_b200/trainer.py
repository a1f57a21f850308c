"""Host mirror of the step's immediate callers in the reference's training loops:
sample-pool batch assembly / write-back, the per-parameter gradient normalisation fused with Adam, and the overflow loss.

The reference has these inline (ExtraChannels/experiments.py:203-211,252-259; ConditioneDyNCA/experiments.py:210-218,259-266;
EncoderConditioning/conditioned_trainer.py:107-113,134-137,155-156,167) as ~10 + ~45 + ~5 ATen launches per iteration; here each
is one launch of libnca_b200.so through the C ABI.  CUDA only: like the step itself there is no CPU / PyTorch fallback.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import NcaError, check, load_library
from .functional import _need_cuda, _ptr, _stream

ADAM_MAX_TENSORS = 16        # NCA_ADAM_MAX_TENSORS


def _idx_tensor(idx, device):
    if isinstance(idx, torch.Tensor):
        t = idx.to(device=device, dtype=torch.int64)
    else:
        t = torch.as_tensor(np.asarray(idx, dtype=np.int64)).to(device, non_blocking=True)
    if t.dim() != 1:
        raise NcaError("pool indices must be one-dimensional")
    return t.contiguous()


def pool_gather(pool, idx, extra=None, seed_state=None, inject_n=0, reseed_dead=None):
    """`nca_pool[batch_idx]` + seed injection into the first `inject_n` samples + `torch.cat((states, extra), 1)`
    (experiments.py:203-211) in one pass.  pool [N,Cp,H,W]; idx: numpy / list / tensor of B slots; extra [B,Cx,H,W] or None;
    seed_state [Cp,H,W] or None (zeros).  reseed_dead=(living_channel_dim, alpha_living_threshold) also replaces every sample
    without a living cell by the seed (conditioned_trainer.py:107-112), tested on the device.  Returns a new [B,Cp+Cx,H,W] tensor."""
    _need_cuda(pool, extra, seed_state)
    if not pool.is_contiguous():
        raise NcaError("the pool tensor must be contiguous")
    N, Cp, H, W = pool.shape
    it = _idx_tensor(idx, pool.device)
    B = it.numel()
    Cx = 0
    if extra is not None:
        extra = extra.detach().contiguous()
        if extra.dim() != 4 or extra.shape[0] != B or tuple(extra.shape[2:]) != (H, W):
            raise NcaError(f"extra must be [B={B},Cx,{H},{W}], got {tuple(extra.shape)}")
        Cx = extra.shape[1]
    if seed_state is not None:
        seed_state = seed_state.detach().contiguous()
        if seed_state.numel() != Cp * H * W:
            raise NcaError(f"seed_state must have {Cp}x{H}x{W} elements, got {tuple(seed_state.shape)}")
    out = torch.empty(B, Cp + Cx, H, W, device=pool.device, dtype=torch.float32)
    lib = load_library()
    with torch.cuda.device(pool.device):
        flags = None
        if reseed_dead is not None:
            flags = torch.empty(B, device=pool.device, dtype=torch.uint8)
            check(lib.nca_pool_dead_flags(N, Cp, H, W, _ptr(pool), _ptr(it), B, int(reseed_dead[0]), float(reseed_dead[1]),
                                          _ptr(flags), _stream()))
        check(lib.nca_pool_gather(N, Cp, H, W, _ptr(pool), _ptr(it), B, _ptr(extra), Cx, _ptr(seed_state),
                                  int(inject_n), _ptr(flags), _ptr(out), _stream()))
    return out


def pool_scatter(pool, idx, states):
    """`nca_pool[batch_idx] = nca_states_after[:, :Cp]` (experiments.py:259), in place on `pool`.  idx must not repeat."""
    _need_cuda(pool, states)
    if not pool.is_contiguous():
        raise NcaError("the pool tensor must be contiguous")
    N, Cp, H, W = pool.shape
    states = states.detach().contiguous()
    it = _idx_tensor(idx, pool.device)
    B = it.numel()
    if states.dim() != 4 or states.shape[0] != B or states.shape[1] < Cp or tuple(states.shape[2:]) != (H, W):
        raise NcaError(f"states must be [B={B},C>={Cp},{H},{W}], got {tuple(states.shape)}")
    with torch.cuda.device(pool.device):
        check(load_library().nca_pool_scatter(N, Cp, H, W, _ptr(pool), _ptr(it), B, _ptr(states), states.shape[1], _stream()))
    return pool


class TensorSamplePool:
    """Tensor-backed stand-in for EncoderConditioning/sample_pool.py:14-33 (`SamplePool`, a Python list of per-sample tensors):
    same `len()` / `pool[idxs]` / `pool[idxs] = outputs` surface, but one contiguous device tensor [pool_size,C,H,W], so that
    `sample_batch` (conditioned_trainer.py:100-113: stack the sampled slots, reseed the None / dead ones, then `batch[:2] = seed`
    :167) is two launches instead of B host round trips.  Slots never written are all-zero, which the dead test reseeds exactly
    like the reference's `None`."""

    def __init__(self, pool_size, shape, device):
        self.pool_size = int(pool_size)
        self.data = torch.zeros(self.pool_size, *shape, device=device, dtype=torch.float32)
        self._written = [False] * self.pool_size

    def __len__(self):
        return self.pool_size

    def __getitem__(self, idx):
        if isinstance(idx, int):
            return self.data[idx] if self._written[idx] else None
        return [self[int(i)] for i in idx]

    def __setitem__(self, idx, value):
        if isinstance(idx, int):
            self.data[idx].copy_(value)
            self._written[idx] = True
            return
        pool_scatter(self.data, idx, value)
        for i in idx:
            self._written[int(i)] = True

    def sample_batch(self, idx, seed_state, living_dim, alive_thr=0.1, inject_n=2):
        return pool_gather(self.data, idx, None, seed_state, inject_n, reseed_dead=(living_dim, alive_thr))


class NormalizedAdam(torch.optim.Optimizer):
    """`for p: p.grad /= (p.grad.norm() + norm_eps)` followed by `torch.optim.Adam.step()` (experiments.py:252-255,
    conditioned_trainer.py:134-137) as ONE kernel launch per <= 16 parameter tensors.

    Drop-in for the `torch.optim.Adam` object of those scripts: same constructor arguments (`lr`, `betas`, `eps`; amsgrad and
    weight decay are not used by the reference and not supported), same `param_groups` / `state` layout (`step`, `exp_avg`,
    `exp_avg_sq`), so `torch.optim.lr_scheduler.MultiStepLR` and `state_dict()` work unchanged.  The caller drops its own
    normalisation loop (or passes normalize=False to keep it).  `zero_grads=True` also folds `optimizer.zero_grad()` in
    (gradients are zeroed in place instead of set to None)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, norm_eps=1e-8, normalize=True, zero_grads=False):
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, norm_eps=norm_eps, normalize=normalize, zero_grads=zero_grads))
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = load_library()
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                _need_cuda(p, p.grad)
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise NcaError("NormalizedAdam needs contiguous parameters and gradients")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] = int(st["step"]) + 1
            # parameters that share a step count go out together (they always do unless parameters were added later)
            by_step = {}
            for p in ps:
                by_step.setdefault(self.state[p]["step"], []).append(p)
            for step, plist in by_step.items():
                for o in range(0, len(plist), ADAM_MAX_TENSORS):
                    chunk = plist[o:o + ADAM_MAX_TENSORS]
                    n = len(chunk)
                    # the pointer tables are rebuilt only when a storage moved (autograd allocates new .grad tensors after
                    # zero_grad(set_to_none=True); with zero_grads=True they stay put)
                    key = tuple(t.data_ptr() for p in chunk for t in (p, p.grad, self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"]))
                    tables = self.__dict__.setdefault("_tables", {})       # not part of the pickled state
                    cached = tables.get(id(chunk[0]))
                    if cached is None or cached[0] != key:
                        arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])      # noqa: E731
                        cached = (key, arr(chunk), arr([p.grad for p in chunk]), arr([self.state[p]["exp_avg"] for p in chunk]),
                                  arr([self.state[p]["exp_avg_sq"] for p in chunk]), (C.c_int64 * n)(*[p.numel() for p in chunk]))
                        tables[id(chunk[0])] = cached
                    with torch.cuda.device(chunk[0].device):
                        check(lib.nca_normalized_adam_step(
                            n, cached[1], cached[2], cached[3], cached[4], cached[5], step, float(group["lr"]),
                            float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]), float(group["norm_eps"]),
                            int(bool(group["normalize"])), int(bool(group["zero_grads"])), _stream()))
        return loss


def _overflow_call(x, grad, scale, scale_dev, accumulate):
    lib = load_library()
    loss = torch.empty(1, device=x.device, dtype=torch.float32)
    nws = lib.nca_overflow_workspace_bytes()
    ws = torch.empty(nws, device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        check(lib.nca_overflow_loss(_ptr(x), x.numel(), _ptr(loss), _ptr(grad), float(scale), _ptr(scale_dev), int(accumulate),
                                    _ptr(ws), nws, _stream()))
    return loss.reshape(())


class _OverflowLoss(torch.autograd.Function):
    """forward: one read of the state -> loss.  backward: one read of the state + one write of the gradient, scaled by the incoming
    dL/dloss read on the device (12 B per element in total; the reference's expression moves ~64 B per element)."""

    @staticmethod
    def forward(ctx, x):
        xc = x.detach().contiguous()
        ctx.save_for_backward(xc)
        ctx.shape = x.shape
        return _overflow_call(xc, None, 1.0, None, 0)

    @staticmethod
    def backward(ctx, gout):
        (xc,) = ctx.saved_tensors
        grad = torch.empty_like(xc)
        gout = gout.detach().to(device=xc.device, dtype=torch.float32).reshape(1).contiguous()
        _overflow_call(xc, grad, 1.0, gout, 0)
        return grad.view(ctx.shape)


def overflow_loss(nca_state):
    """`(nca_state - nca_state.clamp(-1.0, 1.0)).abs().mean()` (utils/loss/loss.py:33-36): the loss and its gradient come out of
    one pass over the state (the reference runs 4 elementwise passes forward and 4 backward)."""
    _need_cuda(nca_state)
    return _OverflowLoss.apply(nca_state)


def overflow_loss_into(nca_state, g_final, weight=1.0):
    """The same loss, with `weight * dloss/dstate` ADDED into an existing gradient buffer `g_final` (what the BPTT consumes):
    returns the unweighted loss as a 0-d tensor."""
    _need_cuda(nca_state, g_final)
    if not (nca_state.is_contiguous() and g_final.is_contiguous()) or nca_state.shape != g_final.shape:
        raise NcaError("overflow_loss_into needs contiguous state / gradient tensors of the same shape")
    return _overflow_call(nca_state.detach(), g_final, weight, None, 1)
