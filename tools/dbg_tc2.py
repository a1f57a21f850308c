import sys, torch
sys.path.insert(0, '.')
import nca_b200
from nca_b200 import functional as Fn, _lib
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1
B, C, fc, H, W, T = 1, 16, 128, 16, 32, 2
g = torch.Generator().manual_seed(1)
w1 = torch.randn(fc, 4 * C + 2, generator=g) * 0.15
b1 = torch.randn(fc, generator=g) * 0.1
w2 = torch.randn(C, fc, generator=g) * 0.1
b2 = torch.randn(C, generator=g) * 0.02
x0 = torch.rand(B, C, H, W, generator=g) - 0.5
masks = (torch.rand(T, B, 1, H, W, generator=g) + 0.5).floor()
cfg = Fn.DyncaConfig(C, fc, "circular", [0, 1] if ns == 2 else [0], _lib.NCA_COND_CPE, 2, precision="bf16")
print("variant", Fn.dynca_kernel_variant(cfg, B, H, W))
with torch.no_grad():
    out, _ = Fn.dynca_rollout(cfg, *[t.cuda() for t in (x0, w1, b1, w2, b2)], T, 0.5, masks=masks.cuda())
torch.cuda.synchronize()
from oracle import nca_oracle as O
want = O.dynca_rollout(x0, w1, b1, w2, b2, masks, (0, 1) if ns == 2 else (0,), "circular", O.cpe2d(B, H, W))
print("rel err", float((out.cpu() - want).abs().max() / want.abs().max()))
