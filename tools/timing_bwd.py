import os, sys, torch
sys.path.insert(0, '.')
import nca_b200
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = nca_b200.DyNCA_EC(16, 3, fc_dim=128, padding_mode="replicate", pos_emb="CPE", perception_scales=[0, 1], device=dev, precision="bf16")
x0 = torch.rand(8, 16, 256, 256, device=dev) - 0.5
s, _ = model.forward_nsteps(x0, 3, seed=1)
s.square().mean().backward()
torch.cuda.synchronize()
s, _ = model.forward_nsteps(x0, 2, seed=1)
os.environ["NCA_T2_TDBG"] = "1"
s.square().mean().backward()
torch.cuda.synchronize()
