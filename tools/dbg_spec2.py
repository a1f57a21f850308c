import os, sys; sys.path.insert(0, '.')
import torch
import nca_b200
from nca_b200 import functional as Fn, _lib
dev = "cuda"
g = torch.Generator().manual_seed(3)
B, C, fc, H, W = 1, 16, 128, 8, 16
w1 = (torch.randn(fc, 4 * C + 2, generator=g) * 0.1).to(dev); b1 = (torch.randn(fc, generator=g) * 0.1).to(dev)
w2 = (torch.randn(C, fc, generator=g) * 0.1).to(dev); b2 = (torch.randn(C, generator=g) * 0.02).to(dev)
x0 = (torch.rand(B, C, H, W, generator=g) - 0.5).to(dev)
cf = torch.randn(B, C, H, W, generator=g).to(dev)
cfg = Fn.DyncaConfig(C, fc, "circular", [0, 1], _lib.NCA_COND_CPE, 2, precision="bf16")
def run(T, w2v):
    ps = [p.clone().requires_grad_(True) for p in (x0, w1, b1, w2v, b2)]
    masks = torch.ones(T, B, 1, H, W, device=dev)
    fin, _ = Fn.dynca_rollout(cfg, *ps, T, 0.5, masks=masks)
    (fin * cf).sum().backward()
    torch.cuda.synchronize()
    return [p.grad.clone() for p in ps]
for name, w2v in (("w2=0", torch.zeros_like(w2)), ("w2", w2)):
    os.environ["NCA_T2_NOSPEC_BWD"] = "0"; a = run(1, w2v)
    os.environ["NCA_T2_NOSPEC_BWD"] = "1"; b = run(1, w2v)
    print(name, [f"{float((x - y).abs().max() / (y.abs().max() + 1e-30)):.2e}" for x, y in zip(a, b)])
    d = (a[3] - b[3]).abs()
    print("   gw2 err by hidden unit (max over c), first 32:", [f"{v:.1e}" for v in d.amax(0)[:32].tolist()])
    print("   gw2 ref magnitude:", float(b[3].abs().max()))
