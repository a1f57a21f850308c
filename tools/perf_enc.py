"""c4 timings: ConditionedNCA forward / forward + BPTT (B=256, 64x64, T=72), per-launch and G cell-updates/s."""
import sys, os, json; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, nca_b200
dev = torch.device("cuda:0")
B, H, T = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 64, 72
def timed(fn, n=3, w=2):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for prec in ("bf16", "fp32"):
    torch.manual_seed(0)
    nca = nca_b200.ConditionedNCA(target_shape=(3, H, H), num_hidden_channels=16, living_channel_dim=3, precision=prec).to(dev)
    with torch.no_grad():
        for p in nca.update_net.parameters(): p.mul_(0.5)
    x0 = nca.generate_seed(B).to(dev) + 0.2 * torch.rand(B, 20, H, H, device=dev)
    goal = torch.rand(B, 3, H, H, device=dev)
    ps = [p for p in nca.parameters() if p.requires_grad]
    def fwd():
        with torch.no_grad(): nca.grow(x0, T, goal, seed=3)
    def both():
        s = nca.grow(x0, T, goal, seed=3)
        torch.autograd.grad(s.square().mean(), ps)
    cells = B * H * H * T
    mf, mb = timed(fwd), timed(both)
    print(json.dumps({"config": "c4", "precision": prec, "B": B, "fwd_us_per_step": mf / T * 1e3, "fwd_G_per_s": cells / mf / 1e6,
                      "both_us_per_step": mb / T * 1e3, "fwd_bptt_G_per_s": cells / mb / 1e6}), flush=True)
    if prec == "bf16" and len(sys.argv) > 2: break
