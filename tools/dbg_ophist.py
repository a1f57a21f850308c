"""Which BPTT mode (operand history vs recompute) deviates?  Runs the 64x96 two-scale case several times in both modes and
compares run to run and against the bf16-emulating oracle."""
import os, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import nca_b200
from nca_b200 import functional as Fn, _lib
from oracle import nca_oracle as O
DEV = "cuda"
B, C, fc, H, W, T, pad, scales = 2, 16, 128, 64, 96, 3, "replicate", (0, 1)
g = torch.Generator().manual_seed(5)
cfg = Fn.DyncaConfig(C, fc, pad, list(scales), _lib.NCA_COND_CPE, 2, precision="bf16")
params = [torch.randn(fc, 4 * C + 2, generator=g) * 0.15, torch.randn(fc, generator=g) * 0.1,
          torch.randn(C, fc, generator=g) * 0.1, torch.randn(C, generator=g) * 0.02]
x0 = torch.rand(B, C, H, W, generator=g) - 0.5
masks = (torch.rand(T, B, 1, H, W, generator=g) + 0.5).floor()
cf = torch.randn(B, C, H, W, generator=g)
fe, ge, _ = O.dynca_bf16emu_rollout_grads(x0, *params, masks, scales, pad, O.cpe2d(B, H, W), cf, {}, 2, 2)
def run(limit):
    os.environ["NCA_OP_HIST_MAX_GB"] = limit
    pg = [p.clone().to(DEV).requires_grad_(True) for p in [x0] + params]
    fg, _ = Fn.dynca_rollout(cfg, *pg, T, 0.5, masks=masks.to(DEV))
    (fg * cf.to(DEV)).sum().backward()
    return [p.grad.cpu() for p in pg]
def rel(a, b): return float((a - b).abs().max() / (b.abs().max() + 1e-30))
ref = {}
order = sys.argv[1].split(",") if len(sys.argv) > 1 else ["48", "0"]
for limit in order:
    rs = [run(limit) for _ in range(4)]
    ref[limit] = rs[0]
    print("mode", "ophist" if limit == "48" else "recompute",
          "run-to-run gx0", ["%.1e" % rel(r[0], rs[0][0]) for r in rs[1:]],
          "vs emu gx0 %.2e w1 %.2e" % (rel(rs[0][0], ge["x0"]), rel(rs[0][1].reshape(fc, -1), ge["w1"])))
print("ophist vs recompute gx0 %.2e" % rel(ref["48"][0], ref["0"][0]))
d = (ref["48"][0] - ref["0"][0]).abs()
idx = (d > 1e-4 * ref["0"][0].abs().max()).nonzero()
print("cells differing:", idx.shape[0], idx[:12].tolist())
