"""Timings of the other BASELINE.json configurations (not bench lines; recorded in profiles/ for DESIGN.md)."""
import sys, time, json
sys.path.insert(0, '.')
import torch
import nca_b200
from nca_b200 import functional as Fn

dev = torch.device("cuda:0")

def timed(fn, n=3, w=2):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def dynca(name, cls, B, C, fc, H, W, T, prec, grad=True, **kw):
    model = cls(C, 3, fc_dim=fc, device=dev, precision=prec, **kw)
    x0 = torch.rand(B, C, H, W, device=dev) - 0.5
    extra = {}
    if cls is nca_b200.DyNCA_CD and kw.get("conditioning") == "edges":
        extra["cond_img"] = torch.rand(B, 1, H, W, device=dev) * 2 - 1
    def fwd():
        with torch.no_grad():
            model.forward_nsteps(x0, T, seed=1, **extra)
    def both():
        s, _ = model.forward_nsteps(x0, T, seed=1, **extra)
        torch.autograd.grad(s.square().mean(), list(p for p in model.parameters() if p.requires_grad))
    cells = B * H * W * T
    out = {"config": name, "precision": prec, "fwd_G_per_s": cells / timed(fwd) / 1e6}
    if grad:
        out["fwd_bptt_G_per_s"] = cells / timed(both) / 1e6
    print(json.dumps(out), flush=True)

def enc(name, B, H, T, prec="fp32"):
    nca = nca_b200.ConditionedNCA(target_shape=(3, H, H), num_hidden_channels=16, living_channel_dim=3, precision=prec).to(dev)
    with torch.no_grad():
        for p in nca.update_net.parameters(): p.mul_(0.5)
    x0 = nca.generate_seed(B).to(dev) + 0.2 * torch.rand(B, 20, H, H, device=dev)
    goal = torch.rand(B, 3, H, H, device=dev)
    def fwd():
        with torch.no_grad():
            nca.grow(x0, T, goal, seed=3)
    def both():
        s = nca.grow(x0, T, goal, seed=3)
        torch.autograd.grad(s.square().mean(), [p for p in nca.parameters() if p.requires_grad])
    cells = B * H * H * T
    print(json.dumps({"config": name, "precision": prec, "fwd_G_per_s": cells / timed(fwd) / 1e6,
                      "fwd_bptt_G_per_s": cells / timed(both) / 1e6}), flush=True)

import sys as _s0
for prec in (() if (len(_s0.argv) > 1 and _s0.argv[1] == "enc") else ("bf16", "fp32")):
    dynca("c1 EC 128x128 C12 fc96 CPE B4 T64", nca_b200.DyNCA_EC, 4, 12, 96, 128, 128, 64, prec, padding_mode="replicate", pos_emb="CPE")
    dynca("c2 EC 256x256 C16 fc128 CPE ms B8 T128", nca_b200.DyNCA_EC, 8, 16, 128, 256, 256, 128, prec, padding_mode="replicate", pos_emb="CPE", perception_scales=[0, 1])
    dynca("c3 CD 256x256 C12 fc96 edges circular B64 T80", nca_b200.DyNCA_CD, 64, 12, 96, 256, 256, 80, prec, padding_mode="circular", conditioning="edges", edge_transform="None")
    dynca("c5 EC 1920x1080 C13 fc96 none circular B1 T64 (no grad)", nca_b200.DyNCA_EC, 1, 13, 96, 1080, 1920, 64, prec, grad=False, padding_mode="circular", pos_emb=None)
import sys as _s
if len(_s.argv) > 1 and _s.argv[1] == "enc":
    pass
enc("c4 ENC 64x64 C20 B256 T72", 256, 64, 72, "fp32")
enc("c4 ENC 64x64 C20 B256 T72", 256, 64, 72, "bf16")
