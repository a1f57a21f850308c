#!/usr/bin/env python
"""Measures the callers' kernels on one B200: each against its HBM roofline (algorithmic bytes /
CUDA-event time, inputs larger than L2 or rotated through > L2 of buffers) and against the reference's own phrasing of the same
lines in PyTorch on the same GPU.  Prints one JSON line per item; `python tools/bench_callers.py > profiles/callers_r1.jsonl`."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nca_b200  # noqa: E402
from nca_b200 import trainer as Tr, video as V  # noqa: E402

DEV = torch.device("cuda:0")
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


QUICK = "--quick" in sys.argv      # a few launches of every kernel for an ncu launch list (numbers printed in this mode are not bench values)


def timed(fn, iters=20, warm=3):
    if QUICK:
        iters, warm = 2, 1
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def line(name, secs, nbytes, ref_secs, **kw):
    d = {"item": name, "us": secs * 1e6, "algorithmic_bytes": nbytes, "GBps": nbytes / secs / 1e9 if nbytes else None,
         "hbm_frac": nbytes / secs / 1e9 / HBM if nbytes else None, "torch_reference_us": ref_secs * 1e6,
         "speedup_vs_torch_phrasing": ref_secs / secs}
    d.update(kw)
    print(json.dumps(d), flush=True)


def main():
    torch.manual_seed(0)
    # ---- N1: pool gather / scatter at c3 (pool 256 x 12 x 256 x 256 = 805 MB, batch 64, EC conditioning channel) ----
    N, Cp, H, W, B = 256, 12, 256, 256, 64
    pool = torch.rand(N, Cp, H, W, device=DEV) - 0.5
    extra = torch.rand(B, 1, H, W, device=DEV)
    rs = np.random.RandomState(0)
    idxs = [torch.as_tensor(rs.choice(N, B, replace=False)).to(DEV) for _ in range(8)]
    k = [0]

    def ours_gather():
        k[0] += 1
        return Tr.pool_gather(pool, idxs[k[0] % 8], extra, None, 1)

    def ref_gather():
        k[0] += 1
        s = pool[idxs[k[0] % 8]]
        s[:1] = 0.0
        return torch.cat((s, extra), 1)
    nb = B * (Cp + 1) * H * W * 4 * 2 - H * W * 4 * Cp          # read + write, the injected sample is not read
    line("pool_gather c3 (B=64 of 256, +1 conditioning channel, seed injection)", timed(ours_gather), nb, timed(ref_gather), launches=1, torch_launches=3)
    after = torch.rand(B, Cp + 1, H, W, device=DEV)

    def ours_scatter():
        k[0] += 1
        Tr.pool_scatter(pool, idxs[k[0] % 8], after)

    def ref_scatter():
        k[0] += 1
        pool[idxs[k[0] % 8]] = after[:, :Cp]
    line("pool_scatter c3", timed(ours_scatter), B * Cp * H * W * 4 * 2, timed(ref_scatter), launches=1, torch_launches=2)

    # ---- N4: overflow loss + gradient on the c3 final state (201 MB > L2) ----
    x = (torch.randn(B, Cp, H, W, device=DEV) * 0.8).requires_grad_(True)

    def ours_overflow():
        x.grad = None
        nca_b200.overflow_loss(x).backward()

    def ref_overflow():
        x.grad = None
        (x - x.clamp(-1.0, 1.0)).abs().mean().backward()
    # ours: forward reads x; backward reads x and writes the gradient (scaled by the incoming dL/dloss read on the device)
    line("overflow_loss fwd+bwd c3 state (autograd.Function)", timed(ours_overflow), x.numel() * 4 * 3, timed(ref_overflow), launches=4)
    gf = torch.zeros_like(x)
    xd = x.detach()
    line("overflow_loss_into g_final (loss + gradient accumulate, one pass)", timed(lambda: Tr.overflow_loss_into(xd, gf, 0.5)), x.numel() * 4 * 3, timed(ref_overflow), launches=2)

    # ---- N1: normalise + Adam, DyNCA c2 parameter set and the ConditionedNCA + encoder set ----
    for name, shapes in (("DyNCA c2 (4 tensors, 10 640 params)", [(128, 66, 1, 1), (128,), (16, 128, 1, 1), (16,)]),
                         ("ConditionedNCA + encoder (10 tensors, 13 068 params)", [(60, 1, 3, 3), (64, 60, 1, 1), (64,), (64, 64, 1, 1), (64,), (20, 64, 1, 1), (16, 6, 3, 3), (16,), (16, 16, 3, 3), (16,)])):
        mine = [torch.nn.Parameter(torch.randn(s, device=DEV) * 0.1) for s in shapes]
        ref = [torch.nn.Parameter(p.detach().clone()) for p in mine]
        grads = [torch.randn(s, device=DEV) for s in shapes]
        o_mine = nca_b200.NormalizedAdam(mine, lr=1e-3)
        o_ref = torch.optim.Adam(ref, lr=1e-3)

        def ours_adam():
            for p, g in zip(mine, grads):
                p.grad = g
            o_mine.step()

        def ref_adam():
            for p, g in zip(ref, grads):
                p.grad = g.clone()
            for p in ref:
                p.grad /= (p.grad.norm() + 1e-8)
            o_ref.step()
        t_o, t_r = timed(ours_adam, 200, 20), timed(ref_adam, 200, 20)
        # wall clock per call too: these are launch-bound, the host cost is what the training loop sees
        NW = 2 if QUICK else 200
        t0 = time.perf_counter(); [ours_adam() for _ in range(NW)]; torch.cuda.synchronize(); w_o = (time.perf_counter() - t0) / NW
        t0 = time.perf_counter(); [ref_adam() for _ in range(NW)]; torch.cuda.synchronize(); w_r = (time.perf_counter() - t0) / NW
        line("normalise + Adam step, " + name, t_o, None, t_r, launches=1, wall_us=w_o * 1e6, torch_wall_us=w_r * 1e6)

    # ---- N2: frame kernels and the frame stream at c5 (1920x1080, C=13 EC flavour) ----
    del pool, extra, after, x, gf, xd
    torch.cuda.empty_cache()
    Hf, Wf, C = 1080, 1920, 13
    states = [torch.rand(1, C, Hf, Wf, device=DEV) - 0.5 for _ in range(4)]      # 4 x 108 MB: rotated, > L2
    frames = [torch.rand(1, 3, Hf, Wf, device=DEV) * 2 - 1 for _ in range(8)]
    outs = [torch.empty(1, Hf, Wf, 3, device=DEV, dtype=torch.uint8) for _ in range(4)]

    def ours_rgb8():
        k[0] += 1
        V.state_to_rgb8(states[k[0] % 4], 2.0, outs[k[0] % 4])

    def ref_rgb8():
        k[0] += 1
        z = states[k[0] % 4][:, :3] * 2.0
        img = z[0].permute(1, 2, 0).clamp(-1.0, 1.0)
        img = (img + 1.0) / 2.0
        return (img.clamp(0, 1) * 255).to(torch.uint8)
    line("state_to_rgb8 1080p (device; the reference does this on the host after a 25 MB float D2H)", timed(ours_rgb8, 50), Hf * Wf * (12 + 3), timed(ref_rgb8, 50), launches=1)

    def ours_gray():
        k[0] += 1
        V.frame_to_cond_channel(states[k[0] % 4], frames[k[0] % 8], C - 1)

    def ref_gray():
        k[0] += 1
        return torch.cat((states[k[0] % 4][:, :-1], torch.mean(frames[k[0] % 8], dim=1, keepdim=True)), 1)
    line("frame_to_cond_channel 1080p (in place) vs cat((h, gray), 1)", timed(ours_gray, 50), Hf * Wf * 16, timed(ref_gray, 50), launches=1)

    # the reference's own video size (nca_size 256, video_utils.py:50-52): host bound without the per-frame CUDA graph
    m = nca_b200.DyNCA_EC(C, 3, fc_dim=96, padding_mode="circular", pos_emb=None, perception_scales=[0], device=DEV, precision="bf16")
    with torch.no_grad():
        m.w2.weight.mul_(0.1)          # random-init weights: keep the 2400-step stream bounded (a trained model is)
    F = 4 if QUICK else 300
    clip = (torch.rand(F, 3, 256, 256) * 2 - 1).pin_memory()
    res = {}
    for mode in ("eager", "graph"):
        st = V.FrameStylizer(m, (256, 256), step_n=8, seed=1, graph=(mode == "graph"))
        st.run(clip)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); st.reset(); st.run(clip); torch.cuda.synchronize(); res[mode] = time.perf_counter() - t0

    def ref_stream_small():
        with torch.no_grad():
            h = m.seed(1, size=(256, 256))
            out = []
            for f in range(F):
                fr = clip[f].unsqueeze(0).to(DEV)
                h = torch.cat((h, torch.mean(fr, dim=1, keepdim=True)), 1)
                nca_state, z = m.forward_nsteps(h, 8, seed=1)
                h = nca_state[:, :-1, :, :]
                img = z.detach().cpu().numpy()[0].transpose(1, 2, 0)
                img = np.clip(img, -1.0, 1.0)
                img = (img + 1.0) / 2.0
                out.append(np.uint8(img.clip(0, 1) * 255))
            return out
    ref_stream_small()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); ref_stream_small(); torch.cuda.synchronize(); t_r = time.perf_counter() - t0
    print(json.dumps({"item": f"frame stream 256x256 EC C=13 bf16, step_n=8, {F} frames (wall clock)",
                      "frames_per_s_eager": F / res["eager"], "frames_per_s_graph": F / res["graph"],
                      "cell_updates_per_s_graph": F * 8 * 65536 / res["graph"], "reference_loop_frames_per_s": F / t_r,
                      "speedup_graph_vs_reference_loop": t_r / res["graph"], "speedup_graph_vs_eager": res["eager"] / res["graph"]}), flush=True)

    for step_n, F in (((8, 2),) if QUICK else ((8, 24), (256, 3))):
        m = nca_b200.DyNCA_EC(C, 3, fc_dim=96, padding_mode="circular", pos_emb=None, perception_scales=[0], device=DEV, precision="bf16")
        with torch.no_grad():
            m.w2.weight.mul_(0.1)
        clip = (torch.rand(F, 3, Hf, Wf) * 2 - 1).pin_memory()
        st = V.FrameStylizer(m, (Hf, Wf), step_n=step_n, seed=1)

        def ours_stream():
            st.reset()
            return st.run(clip)

        def ref_stream():          # video_utils.py:65-82 with the drop-in module (same step kernels), reference host handling
            with torch.no_grad():
                h = m.seed(1, size=(Wf, Hf))
                res = []
                for f in range(F):
                    fr = clip[f].unsqueeze(0).to(DEV)
                    h = torch.cat((h, torch.mean(fr, dim=1, keepdim=True)), 1)
                    nca_state, z = m.forward_nsteps(h, step_n, seed=1)
                    h = nca_state[:, :-1, :, :]
                    img = z.detach().cpu().numpy()[0].transpose(1, 2, 0)
                    img = np.clip(img, -1.0, 1.0)
                    img = (img + 1.0) / 2.0
                    res.append(np.uint8(img.clip(0, 1) * 255))
                return res
        for fn in (ours_stream, ref_stream):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter(); ours_stream(); torch.cuda.synchronize(); t_o = time.perf_counter() - t0
        t0 = time.perf_counter(); ref_stream(); torch.cuda.synchronize(); t_r = time.perf_counter() - t0
        cells = F * step_n * Hf * Wf
        print(json.dumps({"item": f"frame stream c5 1080p EC C=13 bf16, step_n={step_n}, {F} frames from pinned host memory, uint8 frames back on the host (wall clock)",
                          "frames_per_s": F / t_o, "cell_updates_per_s": cells / t_o, "reference_loop_frames_per_s": F / t_r,
                          "reference_loop_cell_updates_per_s": cells / t_r, "speedup_vs_reference_loop": t_r / t_o}), flush=True)


if __name__ == "__main__":
    main()
