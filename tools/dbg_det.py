"""Run-to-run determinism of the forward history / operand history and of the BPTT on the 64x96 two-scale case."""
import os, sys, torch
sys.path.insert(0, '.')
import nca_b200
from nca_b200 import functional as Fn, _lib
DEV = "cuda"
B, C, fc, H, W, T, pad, scales = 2, 16, 128, 64, 96, 3, "replicate", (0, 1)
g = torch.Generator().manual_seed(5)
cfg = Fn.DyncaConfig(C, fc, pad, list(scales), _lib.NCA_COND_CPE, 2, precision="bf16")
params = [torch.randn(fc, 4 * C + 2, generator=g) * 0.15, torch.randn(fc, generator=g) * 0.1,
          torch.randn(C, fc, generator=g) * 0.1, torch.randn(C, generator=g) * 0.02]
x0 = torch.rand(B, C, H, W, generator=g) - 0.5
masks = (torch.rand(T, B, 1, H, W, generator=g) + 0.5).floor().to(DEV)
pd = [p.to(DEV) for p in params]
outs = []
for i in range(6):
    hist, coarse, ops = Fn._dynca_forward_raw(cfg, x0.to(DEV), *pd, None, masks, 0, T, 0.5, True, want_ops=True)
    torch.cuda.synchronize()
    outs.append((hist.clone(), coarse.clone(), ops.clone()))
for i in range(1, 6):
    dh = (outs[i][0] - outs[0][0]).abs()
    dc = (outs[i][1] - outs[0][1]).abs()
    do = (outs[i][2] != outs[0][2])
    print(i, "hist max diff %.2e at steps %s | coarse %.2e | op bytes differing %d" % (
        float(dh.max()), [int(t) for t in range(T + 1) if float(dh[t].max()) > 0], float(dc.max()), int(do.sum())))
    if int(do.sum()):
        idx = do.nonzero().flatten()
        per_step = ops.numel() // T
        tile_b = per_step // 96
        print("   first differing bytes: step,tile,offset", [(int(j) // per_step, (int(j) % per_step) // tile_b, (int(j) % per_step) % tile_b) for j in idx[:6]])
