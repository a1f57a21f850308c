#!/usr/bin/env python
"""One whole training iteration of ExtraChannels/experiments.py:187-259 at config-3 shapes on one GPU (batch 8 = 64 / 8 ranks, pool
256, 256x256, C = 12 + the conditioning channel, fc = 96, circular padding, T = 80), twice: (a) the reference's own lines around the
drop-in module (tensor indexing, cat, `.abs().mean()` overflow term, per-parameter normalisation loop, torch.optim.Adam, indexed
write-back) and (b) the callers of SURVEY.md §8f (pool_gather / overflow_loss / NormalizedAdam / pool_scatter).  The appearance /
motion losses need downloaded networks, so a mean-square term on the rgb output stands in for them in both arms.
CUDA events around 10 iterations after 3 warm-up; prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nca_b200  # noqa: E402

DEV = torch.device("cuda:0")
N, Cp, H, W, B, T = 256, 12, 256, 256, 8, 80


def make():
    torch.manual_seed(0)
    m = nca_b200.DyNCA_EC(Cp + 1, 3, fc_dim=96, padding_mode="circular", pos_emb=None, perception_scales=[0], device=DEV, precision="bf16")
    pool = m.seed(N, size=(W, H))
    return m, pool


def run(arm, iters=10, warm=3):
    m, pool = make()
    gs = torch.rand(B, 1, H, W, device=DEV) * 2 - 1
    target = torch.rand(B, 3, H, W, device=DEV) * 2 - 1
    opt = torch.optim.Adam(m.parameters(), lr=1e-3) if arm == "reference" else nca_b200.NormalizedAdam(m.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [1000, 2000], 0.5)
    rs = np.random.RandomState(0)

    def iteration(i):
        batch_idx = rs.choice(N, B, replace=False)
        inject = i % 8 == 0
        with torch.no_grad():
            if arm == "reference":
                input_states = pool[batch_idx]
                if inject:
                    input_states[:1] = m.seed(1, size=(W, H))[:1]
                input_states = torch.cat((input_states, gs), 1)
            else:
                input_states = nca_b200.pool_gather(pool, batch_idx, extra=gs, inject_n=int(inject))
            m.forward_nsteps(input_states, 1)                                   # the "before" image (experiments.py:213)
        state, rgb = m.forward_nsteps(input_states, T)
        if arm == "reference":
            overflow = (state - state.clamp(-1.0, 1.0)).abs().mean()
        else:
            overflow = nca_b200.overflow_loss(state)
        loss = overflow + (rgb - target).square().mean()
        loss.backward()
        with torch.no_grad():
            if arm == "reference":
                for p in m.parameters():
                    p.grad /= (p.grad.norm() + 1e-8)
            opt.step()
            opt.zero_grad()
            sched.step()
            if arm == "reference":
                pool[batch_idx] = state[:, :Cp, :, :]
            else:
                nca_b200.pool_scatter(pool, batch_idx, state)
        return loss

    for i in range(warm):
        iteration(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        loss = iteration(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, float(loss)


def main():
    ms_ref, l_ref = run("reference")
    ms_new, l_new = run("callers")
    cells = B * H * W * (T + 1)
    print(json.dumps({"item": "training iteration, config-3 shapes per GPU (B=8, 256x256, C=13, fc=96, T=80 + the 1-step probe)",
                      "ms_reference_phrasing": ms_ref, "ms_callers": ms_new, "speedup": ms_ref / ms_new,
                      "cell_updates_per_s_callers": cells / (ms_new * 1e-3), "loss_reference": l_ref, "loss_callers": l_new}), flush=True)


if __name__ == "__main__":
    main()
