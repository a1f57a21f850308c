"""Generates tests/golden/callers.npz for the callers' parity tests FROM THE REFERENCE's own code and the torch
objects it constructs.  Run in the build container (needs /root/reference):  python oracle/make_golden_callers.py

  gray          RGBToGrayscale of ExtraChannels/utils/misc/preprocess_texture.py:178-179 (imported by file path)
  rgb8          the lines of save_video / VideoWriter.add (video_utils.py:78-82,20-27) applied with numpy, as there
                (moviepy is not installed, so the module itself cannot be imported)
  p_final*      4 parameters after n_steps of: p.grad /= p.grad.norm()+1e-8; torch.optim.Adam(lr=1e-3).step();
                MultiStepLR([2, 4], 0.5).step()  (experiments.py:157,170-172,252-257)
  overflow*     (x - x.clamp(-1, 1)).abs().mean() and its autograd gradient (utils/loss/loss.py:33-36)
"""
import importlib.util
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    spec = importlib.util.spec_from_file_location("ref_pt", "/root/reference/ExtraChannels/utils/misc/preprocess_texture.py")
    pt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pt)
    g = torch.Generator().manual_seed(11)
    out = {}
    frame = torch.rand(2, 3, 24, 40, generator=g) * 2 - 1
    out["frame"] = frame.numpy()
    out["gray"] = pt.RGBToGrayscale(frame).numpy()
    state = torch.randn(2, 13, 24, 40, generator=g) * 0.6
    out["state"] = state.numpy()
    z = (state[:, :3] * 2.0).detach().cpu().numpy()
    frames = []
    for b in range(2):
        img = z[b].transpose(1, 2, 0)
        img = np.clip(img, -1.0, 1.0)
        img = (img + 1.0) / 2.0
        frames.append(np.uint8(img.clip(0, 1) * 255))
    out["rgb8"] = np.stack(frames)
    x = state.clone().requires_grad_(True)
    loss = (x - x.clamp(-1.0, 1.0)).abs().mean()
    loss.backward()
    out["overflow"] = np.float32(loss.item())
    out["overflow_grad"] = x.grad.numpy()
    shapes = [(96, 52, 1, 1), (96,), (13, 96, 1, 1), (13,)]
    ps = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.1) for s in shapes]
    for i, p in enumerate(ps):
        out[f"p{i}"] = p.detach().numpy().copy()
    opt = torch.optim.Adam(ps, lr=1e-3)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, [2, 4], 0.5)
    n_steps = 6
    for it in range(n_steps):
        for i, p in enumerate(ps):
            gr = torch.randn(p.shape, generator=g) * (3.0 ** (it - 2))
            out[f"g{it}_{i}"] = gr.numpy().copy()
            p.grad = gr.clone()
        for p in ps:
            p.grad /= (p.grad.norm() + 1e-8)
        opt.step(); opt.zero_grad(); sched.step()
    out["n_steps"] = np.int64(n_steps)
    for i, p in enumerate(ps):
        out[f"p_final{i}"] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "callers.npz"), **out)
    print("wrote tests/golden/callers.npz")


if __name__ == "__main__":
    main()
