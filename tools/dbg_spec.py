import os, sys; sys.path.insert(0, '.')
import torch
import nca_b200
from nca_b200 import functional as Fn, _lib
dev = "cuda"
g = torch.Generator().manual_seed(3)
B, C, fc, H, W = 1, 16, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 32, int(sys.argv[2]) if len(sys.argv) > 2 else 64
w1 = (torch.randn(fc, 4 * C + 2, generator=g) * 0.1).to(dev); b1 = (torch.randn(fc, generator=g) * 0.1).to(dev)
w2 = (torch.randn(C, fc, generator=g) * 0.1).to(dev); b2 = (torch.randn(C, generator=g) * 0.02).to(dev)
x0 = (torch.rand(B, C, H, W, generator=g) - 0.5).to(dev)
cf = torch.randn(B, C, H, W, generator=g).to(dev)
SC = [0, 1] if os.environ.get("DBG_NS", "2") == "2" else [0]
cfg = Fn.DyncaConfig(C, fc, "circular", SC, _lib.NCA_COND_CPE, 2, precision="bf16")
def run(T):
    ps = [p.clone().requires_grad_(True) for p in (x0, w1, b1, w2, b2)]
    masks = (torch.rand(T, B, 1, H, W, generator=torch.Generator().manual_seed(9)) + 0.5).floor().to(dev)
    fin, _ = Fn.dynca_rollout(cfg, *ps, T, 0.5, masks=masks)
    (fin * cf).sum().backward()
    torch.cuda.synchronize()
    return [p.grad.clone() for p in ps]
for T in (1, 2, 3):
    os.environ["NCA_T2_NOSPEC_BWD"] = "0"; a = run(T)
    os.environ["NCA_T2_NOSPEC_BWD"] = "1"; b = run(T)
    print("T", T, [f"{float((x - y).abs().max() / (y.abs().max() + 1e-30)):.2e}" for x, y in zip(a, b)])
    if T == 1:
        d = (a[0] - b[0]).abs()[0].amax(0)
        bad = (d > 1e-3 * b[0].abs().max()).nonzero()
        print("  bad x0-grad cells:", bad.shape[0], "rows", sorted(set(bad[:, 0].tolist()))[:40], "cols", sorted(set(bad[:, 1].tolist()))[:40])
        dw = (a[1] - b[1]).abs()
        print("  gw1 bad rows (hidden):", (dw.amax(1) > 1e-3 * b[1].abs().max()).nonzero().flatten().tolist()[:20], "cols:", (dw.amax(0) > 1e-3 * b[1].abs().max()).nonzero().flatten().tolist()[:70])
