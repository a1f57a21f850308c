#!/usr/bin/env python
"""Where does a 1080p frame of the stream go at 8 steps per frame?  (diagnostic, not a bench line)"""
import os, sys, time, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nca_b200
from nca_b200 import video as V
DEV = torch.device("cuda:0")
H, W, C, F = 1080, 1920, 13, 24
m = nca_b200.DyNCA_EC(C, 3, fc_dim=96, padding_mode="circular", pos_emb=None, perception_scales=[0], device=DEV, precision="bf16")
with torch.no_grad():
    m.w2.weight.mul_(0.1)
clip = (torch.rand(F, 3, H, W) * 2 - 1).pin_memory()
clip_d = clip.to(DEV)
res = {}
for name, frames, graph in (("host_frames", clip, False), ("device_frames", clip_d, False), ("device_frames_graph", clip_d, True), ("host_frames_graph", clip, True)):
    st = V.FrameStylizer(m, (H, W), step_n=8, seed=1, graph=graph)
    st.run(frames); torch.cuda.synchronize()
    t0 = time.perf_counter(); st.run(frames); torch.cuda.synchronize(); res[name + "_ms_per_frame"] = (time.perf_counter() - t0) / F * 1e3
st = V.FrameStylizer(m, (H, W), step_n=8, seed=1)
fr = clip_d[0:1]
for _ in range(3): st.push(fr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): st.push(fr)
e1.record(); torch.cuda.synchronize()
res["push_only_ms"] = e0.elapsed_time(e1) / 20
buf = torch.empty(1, 3, H, W, device=DEV)
e0.record()
for f in range(20): buf.copy_(clip[f % F:f % F + 1], non_blocking=True)
e1.record(); torch.cuda.synchronize()
res["h2d_25MB_ms"] = e0.elapsed_time(e1) / 20
o8 = torch.empty(1, H, W, 3, device=DEV, dtype=torch.uint8); h8 = torch.empty(1, H, W, 3, dtype=torch.uint8).pin_memory()
e0.record()
for f in range(20): h8.copy_(o8, non_blocking=True)
e1.record(); torch.cuda.synchronize()
res["d2h_6MB_ms"] = e0.elapsed_time(e1) / 20
print(json.dumps(res))
