"""Run-to-run determinism of the BPTT (first runs of a process vs later ones) over a few variants."""
import os, sys, torch
sys.path.insert(0, '.')
import nca_b200
from nca_b200 import functional as Fn, _lib
DEV = "cuda"
def rel(a, b): return float((a - b).abs().max() / (b.abs().max() + 1e-30))
def case(B, C, fc, H, W, T, pad, scales, limit):
    g = torch.Generator().manual_seed(5)
    cfg = Fn.DyncaConfig(C, fc, pad, list(scales), _lib.NCA_COND_CPE, 2, precision="bf16")
    params = [torch.randn(fc, 4 * C + 2, generator=g) * 0.15, torch.randn(fc, generator=g) * 0.1,
              torch.randn(C, fc, generator=g) * 0.1, torch.randn(C, generator=g) * 0.02]
    x0 = torch.rand(B, C, H, W, generator=g) - 0.5
    masks = (torch.rand(T, B, 1, H, W, generator=g) + 0.5).floor().to(DEV)
    cf = torch.randn(B, C, H, W, generator=g).to(DEV)
    os.environ["NCA_OP_HIST_MAX_GB"] = limit
    rs = []
    for _ in range(5):
        pg = [p.clone().to(DEV).requires_grad_(True) for p in [x0] + params]
        fg, _ = Fn.dynca_rollout(cfg, *pg, T, 0.5, masks=masks)
        (fg * cf).sum().backward()
        rs.append([p.grad.cpu() for p in pg])
    d = [rel(r[0], rs[-1][0]) for r in rs[:-1]]
    bad = [i for i, v in enumerate(d) if v > 1e-6]
    where = ""
    if bad:
        dd = (rs[bad[0]][0] - rs[-1][0]).abs()
        idx = (dd > 1e-5 * rs[-1][0].abs().max()).nonzero()
        where = "cells %d, y %d..%d x %d..%d b %s" % (idx.shape[0], int(idx[:, 2].min()), int(idx[:, 2].max()), int(idx[:, 3].min()), int(idx[:, 3].max()), sorted(set(idx[:, 0].tolist())))
    print((B, C, fc, H, W, T, pad, scales, limit), ["%.0e" % v for v in d], where, flush=True)
which = sys.argv[1]
if which == "a": case(2, 16, 128, 64, 96, 3, "replicate", (0, 1), "48")
if which == "b": case(2, 16, 128, 64, 96, 1, "replicate", (0, 1), "48")
if which == "c": case(2, 16, 128, 64, 96, 3, "constant", (0, 1), "48")
if which == "d": case(2, 16, 128, 64, 96, 3, "replicate", (0,), "48")
if which == "e": case(2, 16, 128, 64, 64, 3, "replicate", (0, 1), "48")
if which == "f": case(1, 16, 128, 64, 96, 2, "replicate", (0, 1), "48")
