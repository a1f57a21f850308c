import os, sys, torch
sys.path.insert(0, '.')
import nca_b200
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = nca_b200.DyNCA_EC(13, 3, fc_dim=96, padding_mode="circular", pos_emb=None, device=dev, precision="bf16")
x0 = torch.rand(1, 13, 1080, 1920, device=dev) - 0.5
with torch.no_grad():
    model.forward_nsteps(x0, 3, seed=1)
    torch.cuda.synchronize()
    os.environ["NCA_T2_TDBG"] = "1"
    model.forward_nsteps(x0, 2, seed=1)
torch.cuda.synchronize()
