"""Short c2-shaped run for ncu: B=8, 256x256, C=16, fc=128, CPE, scales [0,1]; T steps forward (history) + BPTT.
Same per-launch shapes as bench.py's workload, fewer launches.  usage: python profiles/prof_step.py [T] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nca_b200

T = int(sys.argv[1]) if len(sys.argv) > 1 else 6
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = nca_b200.DyNCA_EC(16, 3, fc_dim=128, padding_mode="replicate", pos_emb="CPE", perception_scales=[0, 1],
                          device=dev, precision=prec)
x0 = torch.rand(8, 16, 256, 256, device=dev) - 0.5
for it in range(2):
    state, _, mids = model.forward_nsteps(x0, T, return_middle_feature=True, seed=it)
    loss = state.square().mean() + mids[0].square().mean()
    torch.autograd.grad(loss, list(model.parameters()))
    with torch.no_grad():
        model.forward_nsteps(x0, T, seed=it)
torch.cuda.synchronize()
print("ok")
