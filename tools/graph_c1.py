"""c1-sized rollouts (4 x 12 x 128 x 128, one scale): eager launches vs one captured CUDA graph per T-step rollout (is the small-grid
step host bound?).  usage: python tools/graph_c1.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nca_b200
from nca_b200 import functional as Fn, _lib

dev = torch.device("cuda:0")
B, C, fc, H, W, T = 4, 12, 96, 128, 128, 64
cfg = Fn.DyncaConfig(C, fc, "replicate", [0], _lib.NCA_COND_CPE, 2, precision="bf16")
g = torch.Generator().manual_seed(0)
P = 4 * C + 2
w1 = (torch.randn(fc, P, generator=g) * 0.05).to(dev); b1 = torch.zeros(fc, device=dev)
w2 = (torch.randn(C, fc, generator=g) * 0.02).to(dev); b2 = torch.zeros(C, device=dev)
x0 = (torch.rand(B, C, H, W, generator=g) - 0.5).to(dev)


def timed(fn, n=20, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def roll():
    return Fn._dynca_forward_raw(cfg, x0, w1, b1, w2, b2, None, None, 7, T, 0.5, False)

ms_eager = timed(roll)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    roll(); roll()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        out = roll()
torch.cuda.synchronize()
ms_graph = timed(gr.replay)
print(json.dumps({"config": "c1 forward", "T": T, "eager_us_per_step": ms_eager / T * 1e3, "graph_us_per_step": ms_graph / T * 1e3,
                  "eager_G_per_s": B * H * W * T / ms_eager / 1e6, "graph_G_per_s": B * H * W * T / ms_graph / 1e6}))
