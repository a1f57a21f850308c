"""Short runs of the other named shapes for an ncu launch list: c3 (CD edges, B=64, 256x256, circular) fwd + BPTT and
c5 (EC C=13, 1920x1080, circular) fwd.  usage: python profiles/prof_other.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nca_b200
dev = torch.device("cuda:0")
torch.manual_seed(0)
m3 = nca_b200.DyNCA_CD(12, 3, fc_dim=96, padding_mode="circular", conditioning="edges", edge_transform="None", device=dev, precision="bf16")
x3 = torch.rand(64, 12, 256, 256, device=dev) - 0.5
c3 = torch.rand(64, 1, 256, 256, device=dev) * 2 - 1
for it in range(2):
    s, _ = m3.forward_nsteps(x3, 4, seed=it, cond_img=c3)
    torch.autograd.grad(s.square().mean(), [p for p in m3.parameters() if p.requires_grad])
m5 = nca_b200.DyNCA_EC(13, 3, fc_dim=96, padding_mode="circular", pos_emb=None, device=dev, precision="bf16")
x5 = torch.rand(1, 13, 1080, 1920, device=dev) - 0.5
with torch.no_grad():
    for it in range(2):
        m5.forward_nsteps(x5, 4, seed=it)
torch.cuda.synchronize()
print("ok")
