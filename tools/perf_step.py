"""Per-launch timings of the step kernels (CUDA events around nca_dynca_forward / nca_dynca_backward calls, divided by T).
usage: python tools/perf_step.py [config ...] [--prec bf16] [--T 32]      configs: c1 c2 c3 c5"""
import argparse
import ctypes as Ct
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nca_b200
from nca_b200 import functional as Fn, _lib

CONFIGS = {
    "c1": dict(B=4, C=12, fc=96, H=128, W=128, pad="replicate", scales=[0], cond=_lib.NCA_COND_CPE, cc=2),
    "c2": dict(B=8, C=16, fc=128, H=256, W=256, pad="replicate", scales=[0, 1], cond=_lib.NCA_COND_CPE, cc=2),
    "c3": dict(B=8, C=12, fc=96, H=256, W=256, pad="circular", scales=[0], cond=_lib.NCA_COND_TENSOR, cc=3),
    "c5": dict(B=1, C=13, fc=96, H=1080, W=1920, pad="circular", scales=[0], cond=_lib.NCA_COND_NONE, cc=0, nograd=True),
}


def timed(fn, n=3, w=2):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run(name, prec, T):
    c = CONFIGS[name]
    dev = torch.device("cuda:0")
    lib = nca_b200.load_library()
    cfg = Fn.DyncaConfig(c["C"], c["fc"], c["pad"], c["scales"], c["cond"], c["cc"], precision=prec)
    B, C, H, W, fc = c["B"], c["C"], c["H"], c["W"], c["fc"]
    g = torch.Generator().manual_seed(0)
    P = 4 * C + c["cc"]
    w1 = (torch.randn(fc, P, generator=g) * (0.2 * (2.0 / (fc + P)) ** 0.5)).to(dev)
    b1 = ((torch.rand(fc, generator=g) - 0.5) * 0.2).to(dev)
    w2 = (torch.randn(C, fc, generator=g) * (0.1 * (2.0 / (fc + C)) ** 0.5)).to(dev)
    b2 = torch.zeros(C, device=dev)
    x0 = (torch.rand(B, C, H, W, generator=g) - 0.5).to(dev)
    cond = (torch.rand(B, 3, H, W, generator=g) * 2 - 1).to(dev) if c["cond"] == _lib.NCA_COND_TENSOR else None
    out = {"config": name, "precision": prec, "cells_per_launch": B * H * W,
           "variant_fwd": Fn.dynca_kernel_variant(cfg, B, H, W), "variant_bwd": Fn.dynca_kernel_variant(cfg, B, H, W, backward=True)}
    ms = timed(lambda: Fn._dynca_forward_raw(cfg, x0, w1, b1, w2, b2, cond, None, 7, T, 0.5, False))
    out["fwd_us"] = ms / T * 1e3
    out["fwd_G_per_s"] = B * H * W / (ms / T * 1e-3) / 1e9
    if not c.get("nograd"):
        res = Fn._dynca_forward_raw(cfg, x0, w1, b1, w2, b2, cond, None, 7, T, 0.5, True, want_ops=True)
        ms = timed(lambda: Fn._dynca_forward_raw(cfg, x0, w1, b1, w2, b2, cond, None, 7, T, 0.5, True, want_ops=True))
        out["fwd_hist_us"] = ms / T * 1e3
        hist, coarse, (ops, _t1) = res
        d = cfg.desc(B, H, W, 0.5, False)
        nbytes = lib.nca_dynca_workspace_bytes(Ct.byref(d), 1)
        ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        wst = Fn._weights_struct(w1, b1, w2, b2)
        gouts = [torch.empty_like(p) for p in (w1, b1, w2, b2)]
        gst = Fn._weights_struct(*gouts)
        gx0 = torch.empty_like(x0)
        g_final = torch.randn(B, C, H, W, device=dev) / (B * C * H * W)
        stream = Ct.c_void_p(torch.cuda.current_stream().cuda_stream)

        def bwd():
            Fn.check(lib.nca_dynca_backward(Ct.byref(d), Ct.byref(wst), cond.data_ptr() if cond is not None else None, None, Ct.c_uint64(7), 0, T,
                                            hist.data_ptr(), coarse.data_ptr() if coarse is not None else None,
                                            ops.data_ptr() if ops is not None else None, g_final.data_ptr(), (Ct.c_void_p * 1)(),
                                            (Ct.c_int32 * 1)(), 0, 1, 2.0, gx0.data_ptr(), Ct.byref(gst), ws.data_ptr(), nbytes, stream))
        ms = timed(bwd)
        out["bwd_us"] = ms / T * 1e3
        out["fwd_bptt_G_per_s"] = B * H * W / ((out["fwd_hist_us"] + out["bwd_us"]) * 1e-6) / 1e9
        out["gw1_sum"] = float(gouts[0].double().abs().sum())
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c2"])
    ap.add_argument("--prec", default="bf16")
    ap.add_argument("--T", type=int, default=32)
    a = ap.parse_args()
    for n in a.configs:
        for p in a.prec.split(","):
            run(n, p, a.T)
